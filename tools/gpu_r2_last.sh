#!/bin/bash
# round 2, last GPU job: the whole -m gpu suite and smoke() on the committed build
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/r2q_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2q_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
