#!/bin/bash
# Round-2 GPU call 32: last check of the final tree -- the full GPU suite and smoke()
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_g32_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2_g32_smoke.log
