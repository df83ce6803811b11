#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
LIFU_WIDE_ZSPLIT=1 timeout 200 python -m pytest tests/test_gpu_wide.py tests/test_gpu_slab.py -q -p no:cacheprovider -k "128-128-128 or long_x_axis or heterogeneous_absorbing or fused_passes or 64-64-768" 2>&1 | tail -3
LIFU_WIDE_ZSPLIT=1 timeout 120 python tools/single_grid.py 472 6 v2 > gpurun_out/r2_wide_512_zsplit.jsonl 2> gpurun_out/r2_wide_zsplit.err; cut -c 1-800 gpurun_out/r2_wide_512_zsplit.jsonl; tail -2 gpurun_out/r2_wide_zsplit.err
LIFU_WIDE_ZSPLIT=1 timeout 150 python tools/single_grid.py 728 4 v2 > gpurun_out/r2_wide_768_zsplit.jsonl 2> gpurun_out/r2_wide_zsplit.err; cut -c 1-800 gpurun_out/r2_wide_768_zsplit.jsonl; tail -2 gpurun_out/r2_wide_zsplit.err
