#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_wide.py tests/test_gpu_slab.py -q -p no:cacheprovider -k "long_x_axis or heterogeneous_absorbing or fused_passes" 2>&1 | tail -2
