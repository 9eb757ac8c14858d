#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) --set full captures of the dominant
# kernel (2-CTA tcgen05 GEMM) and of the other hot kernels.  Numbers printed by runs under ncu are never bench values.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "=== full-model tests"; timeout -k 10 900 python -m pytest tests/test_gpu_full_model.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
SKIP=${SKIP:-1250}
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 420 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_2cta_kernel -s 40 -c 8 \
    -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_gemm.log 2>&1
echo "gemm capture exit $?"
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:"local_attention_kernel|global_row|logmel_kernel|sub_conv0_dw1_kernel|conv_dw_kernel|layernorm_kernel|gemm_bf16_tn_kernel" -s 10 -c 12 \
    -o gpurun_out/prof_other -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_other.log 2>&1
echo "other capture exit $?"
timeout -k 10 1500 ncu --set full --clock-control none -k regex:rnnt_greedy_batched_kernel -c 1 \
    -o gpurun_out/prof_decode -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_decode.log 2>&1
echo "decode capture exit $?"
ls -la gpurun_out/ | head -30
