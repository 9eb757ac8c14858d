#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "=== decode tests (tiny, both kernels)"; timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "teacher or end_to_end" -p no:cacheprovider 2>&1 | tail -12
echo "=== calibrate full"; timeout -k 10 900 python scripts/calibrate_synthetic.py --config full --out gpurun_out/synth_calib_full.json 2>&1 | grep -v "fine scan" | tail -3
cp gpurun_out/synth_calib_full.json reazonspeech_b200/data/synth_calib_24x1024_v3000_p640_j640_seed0.json
echo "=== full-model tests"; timeout -k 10 1500 python -m pytest tests/test_gpu_full_model.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/full_model.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/full_model.log
echo "=== bench (batched decode)"; timeout -k 10 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
echo "=== bench (per-utterance decode)"; RS_DECODE_MODE=1 timeout -k 10 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stage_ms'], d['config']['tokens_per_clip'])"
