cd "${GRAFT_REPO_ROOT:-.}"
timeout -k 5 120 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "teacher_forced or windowed_decode" 2>&1 | grep -vE "^\s*$" | tail -15 | cut -c1-300
rc=${PIPESTATUS[0]}; echo "decode tests rc=$rc"
if [ "$rc" != "0" ]; then exit 0; fi
RS_DECODE_MODE=4 timeout -k 10 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r1q_bench_mode4.json 2> gpurun_out/r1q_bench_mode4.err; echo "bench mode4 exit $?"; tail -3 gpurun_out/r1q_bench_mode4.err | cut -c1-300
RS_DECODE_MODE=3 timeout -k 10 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r1q_bench_mode3.json 2> gpurun_out/r1q_bench_mode3.err; echo "bench mode3 exit $?"
