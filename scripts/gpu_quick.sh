cd "${GRAFT_REPO_ROOT:-.}"
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "long_form" -s 2>&1 | grep -E "utt|passed|failed|Error|error" | head -12 | cut -c1-250
timeout -k 10 600 python bench.py --no-cpu-baseline --steps 5 --batch 128 > gpurun_out/r1n_bench_b128.json 2> gpurun_out/r1n_bench_b128.err; echo "bench b128 exit $?"; tail -3 gpurun_out/r1n_bench_b128.err | cut -c1-300
