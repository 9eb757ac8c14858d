#!/bin/bash
# calibration of the synthetic full-size checkpoint, full-size parity, first bench + launch list
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g | head -2 >> gpurun_out/nproc.txt
echo "=== calibrate"; timeout -k 10 900 python scripts/calibrate_blank.py --config full --out gpurun_out/calib_full.json 2>&1 | tail -5
echo "=== full-model tests"; timeout -k 10 1500 python -m pytest tests/test_gpu_full_model.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/full_model.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/full_model.log
echo "=== bench"; timeout -k 10 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
