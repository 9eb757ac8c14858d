#!/bin/bash
# Full GPU pass: calibration of the synthetic checkpoint, every -m gpu test, smoke(), bench, reference arm.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "=== calibrate full"; timeout -k 10 900 python scripts/calibrate_synthetic.py --config full --out gpurun_out/synth_calib_full.json 2>&1 | grep -v "fine scan\|rate curve" | tail -2 | cut -c1-260
cp gpurun_out/synth_calib_full.json reazonspeech_b200/data/synth_calib_24x1024_v3000_p640_j640_seed0.json
echo "=== pytest -m gpu"; timeout -k 10 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
echo "=== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout -k 10 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
echo "=== bench reference arm"; timeout -k 10 1200 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json; tail -n 3 gpurun_out/bench_ref.err
