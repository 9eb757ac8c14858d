set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r1e_tests.log
tail -5 gpurun_out/r1e_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r1e_bench.json"))
print(d["value"], d["e2e"]["value"], d["stage_ms"], d["decode_cycles_cta0"], d["roofline"]["frac"])
for k,v in d["kernel_ms"].items(): print(f"{k:45s} {v['launches']:4d} {v['ms']:8.3f}")
PY
