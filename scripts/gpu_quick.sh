cd "${GRAFT_REPO_ROOT:-.}"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | cut -c1-300
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -6 | cut -c1-250
timeout -k 10 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r1o_bench.err
