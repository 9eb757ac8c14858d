#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "=== encoder/e2e tests"; timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "encoder or end_to_end or teacher" -p no:cacheprovider 2>&1 | tail -4
echo "=== bench"; timeout -k 10 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['config']['tokens_per_clip'], d['roofline']['achieved'], d['roofline']['gemm_ms_per_step'], d['decode_cycles_cta0'])"; tail -3 gpurun_out/bench.err
