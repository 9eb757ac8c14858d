cd "${GRAFT_REPO_ROOT:-.}"
CAL=reazonspeech_b200/data/synth_calib_24x1024_v3000_p640_j640_seed0.json
timeout -k 10 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -vE "^\s*$" | tail -12 | cut -c1-250 | tee gpurun_out/r1m_tests.log
timeout 600 python scripts/calibrate_synthetic.py --config full --out gpurun_out/calib_full.json > gpurun_out/r1m_calib.log 2>&1 && cp gpurun_out/calib_full.json $CAL
tail -1 gpurun_out/r1m_calib.log | cut -c1-200
timeout -k 10 600 python bench.py --steps 20 > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r1m_bench.err
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 420 --csv --log-file gpurun_out/r01_v5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list exit $?"
