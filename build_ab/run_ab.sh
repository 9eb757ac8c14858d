#!/bin/bash
# ONE box: M0 = shipped LayerNorm (one row prefetched in registers, 2 CTAs / SM), M1 = same at 3 CTAs / SM (80 registers, spills),
# M2 = rows through per-thread cp.async FIFOs in shared memory (3 rows in flight per warp)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
cp reazonspeech_b200/librs_engine.so build_ab/librs_engine_SHIPPED.so
for v in M1 M2; do
  cp build_ab/librs_engine_$v.so reazonspeech_b200/librs_engine.so
  echo "== unit tests $v"; timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k layernorm -p no:cacheprovider 2>&1 | tail -1
done
for round in 1 2 3; do for v in M0 M1 M2; do
  cp build_ab/librs_engine_$v.so reazonspeech_b200/librs_engine.so
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=j['kernel_ms']
print('$v round $round  ms/step %.3f  layers %.3f ln %.3f conv_dw %.3f gemm_ms %.3f clocks %s tok %.3f' % (j['ms_per_step'], j['stage_ms']['layers'], k['launch_layernorm']['ms'], k['launch_conv_dw']['ms'], j['roofline']['gemm_ms_per_step'], j['clocks']['sm_mhz'], j['config']['tokens_per_clip']))"
done; done
cp build_ab/librs_engine_SHIPPED.so reazonspeech_b200/librs_engine.so
