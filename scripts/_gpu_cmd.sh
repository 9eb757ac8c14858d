set -x
CAL=reazonspeech_b200/data/synth_calib_24x1024_v3000_p640_j640_seed0.json
RS_DECODE_MODE=2 timeout 600 python scripts/calibrate_synthetic.py --config full --out gpurun_out/calib_full.json > gpurun_out/r1c_calib.log 2>&1 && cp gpurun_out/calib_full.json $CAL
tail -2 gpurun_out/r1c_calib.log | cut -c1-600
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r1c_tests.log
tail -25 gpurun_out/r1c_tests.log
RS_DECODE_MODE=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1c_bench_mode2.json 2> gpurun_out/r1c_bench_mode2.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1c_bench_mode3.json 2> gpurun_out/r1c_bench_mode3.err
python - <<'PY'
import json
for m in ("mode2","mode3"):
    try:
        d=json.load(open(f"gpurun_out/r1c_bench_{m}.json"))
        print(m, d["value"], d["e2e"]["value"], d["stage_ms"], d["decode_cycles_cta0"], d["config"].get("tokens_per_clip"))
    except Exception as e:
        print(m, "failed", e); print(open(f"gpurun_out/r1c_bench_{m}.err").read()[-1500:])
PY
