cd "${GRAFT_REPO_ROOT:-.}"
CAL=reazonspeech_b200/data/synth_calib_24x1024_v3000_p640_j640_seed0.json
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -12 | cut -c1-250 | tee gpurun_out/r1j_tests.log
timeout 600 python scripts/calibrate_synthetic.py --config full --out gpurun_out/calib_full.json > gpurun_out/r1j_calib.log 2>&1 && cp gpurun_out/calib_full.json $CAL
tail -1 gpurun_out/r1j_calib.log | cut -c1-400
timeout -k 10 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r1j_bench.err
