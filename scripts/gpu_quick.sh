cd "${GRAFT_REPO_ROOT:-.}"
timeout -k 5 150 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "gemm" 2>&1 | grep -vE "^\s*$" | tail -8 | cut -c1-250
rc=${PIPESTATUS[0]}; echo "gemm tests rc=$rc"
if [ "$rc" != "0" ]; then exit 0; fi
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -8 | cut -c1-250 | tee gpurun_out/r1l_tests.log
timeout -k 10 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r1l_bench.err
RS_GEMM_MODE=1 timeout -k 10 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r1l_bench_mode1.json 2> gpurun_out/r1l_bench_mode1.err; echo "bench mode1 exit $?"
