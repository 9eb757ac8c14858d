#!/bin/bash
# A = shipped (timeline stamps + first ring between the set-up barriers), B = no stamps, C = no stamps, no early ring
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for round in 1 2 3; do for v in A B C; do
  cp build_ab/librs_engine_$v.so reazonspeech_b200/librs_engine.so
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=j['kernel_ms']
print('$v round $round  ms/step %.3f  gemm_ms %.3f  frac %.4f  clocks %s  N1024K1024 %.3f N4096 %.3f K4096 %.3f qkv %.3f glu %.3f' % (j['ms_per_step'], j['roofline']['gemm_ms_per_step'], j['roofline']['frac'], j['clocks']['sm_mhz'], k['gemm N=1024 K=1024 epi=4']['ms'], k['gemm N=4096 K=1024 epi=2']['ms'], k['gemm N=1024 K=4096 epi=4']['ms'], k['gemm N=3072 K=1024 epi=7']['ms'], k['gemm N=2048 K=1024 epi=3']['ms']))"
done; done
cp build_ab/librs_engine_A.so reazonspeech_b200/librs_engine.so
