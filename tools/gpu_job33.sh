#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_gpu_wide.py -q -p no:cacheprovider -k "128-128-128 or long_x_axis or 64-64-768 or 64-64-512" 2>&1 | tail -2
timeout 60 python tools/single_grid.py 472 6 v2 2>/dev/null | cut -c 1-200
