#!/bin/bash
# Run the GPU parity tests group by group, each in its own process under a timeout, so a
# hung kernel in one group cannot take the others (or the box) with it.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
i=0
for k in "test_gemm_bias_f32" "test_gemm_epilogues or test_gemm_resid" "test_layernorm" "test_logmel" \
         "test_encoder_vs_oracle" "test_greedy_teacher" "test_end_to_end"; do
  i=$((i+1))
  echo "=== group $i: $k"
  timeout -k 10 ${GROUP_TIMEOUT:-240} python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "$k" -p no:cacheprovider \
      > gpurun_out/group_$i.log 2>&1
  echo "exit $?"; tail -n 25 gpurun_out/group_$i.log
done
