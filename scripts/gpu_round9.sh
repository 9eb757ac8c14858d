#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "=== gemm + e2e tests (new epilogue)"; timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm or end_to_end or encoder" -p no:cacheprovider 2>&1 | tail -5
echo "=== calibrate full"; timeout -k 10 900 python scripts/calibrate_synthetic.py --config full --out gpurun_out/synth_calib_full.json 2>&1 | grep -v "fine scan\|rate curve" | tail -2 | cut -c1-300
cp gpurun_out/synth_calib_full.json reazonspeech_b200/data/synth_calib_24x1024_v3000_p640_j640_seed0.json
echo "=== bench (batched decode)"; timeout -k 10 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], d['config']['tokens_per_clip'], d['roofline']['achieved'], d['roofline']['gemm_ms_per_step'], d['decode_cycles_cta0'])"; tail -3 gpurun_out/bench.err
