#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step (device time per launch),
# (2) one --set full capture of the dominant kernel (tcgen05 GEMM) and of the decode / attention kernels.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
SKIP=${SKIP:-1200}
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 420 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 40 -c 6 \
    -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_gemm.log 2>&1
echo "gemm capture exit $?"
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:"local_attention_kernel|rnnt_greedy_kernel|logmel_kernel|sub_conv0_dw1_kernel|conv_dw_kernel|layernorm_kernel" -c 8 \
    -o gpurun_out/prof_other -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_other.log 2>&1
echo "other capture exit $?"
ls -la gpurun_out/
