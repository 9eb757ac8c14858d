#!/bin/bash
# Round evidence on one B200: tests, bench, and the ncu captures summarised under profiles/
# (1) launch list of one bench step, (2) --set full captures of the dominant kernel (2-CTA tcgen05 GEMM), the decode
# kernel and the attention kernels.  Numbers printed by runs under ncu are never bench values.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${TAG:-r01_v4}
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/${TAG}_tests.log
timeout -k 10 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
SKIP=${SKIP:-1250}
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 420 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_2cta_kernel -s 40 -c 6 \
    -o gpurun_out/${TAG}_prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_gemm.log 2>&1
echo "gemm capture exit $?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:rnnt_greedy_spec_kernel -c 1 \
    -o gpurun_out/${TAG}_prof_decode -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_decode.log 2>&1
echo "decode capture exit $?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"local_attention_tc_kernel|global_row_attention_tc|layernorm_kernel|conv_dw_kernel|sub_dw_kernel|gemm_bf16_tn_kernel" -s 4 -c 8 \
    -o gpurun_out/${TAG}_prof_other -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_other.log 2>&1
echo "other capture exit $?"
ls -la gpurun_out/ | tail -12
