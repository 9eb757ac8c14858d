#!/bin/bash
# One GPU call of the development cycle:  gpurun --timeout 1800 -- 'bash scripts/gpu_cycle.sh <tag> [tests] [calib] [bench] [sanitize] [ncu:<regex>[:<skip>[:<count>]]] [launches]'
# Everything lands in gpurun_out/<tag>_*.  Sections run in the order calib, tests, bench, sanitize, profiles regardless of argument order.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
tag=$1; shift
mkdir -p gpurun_out
has() { for a in "$@"; do [ "$a" = "$want" ] && return 0; done; return 1; }
args=("$@")
want=calib; if has "${args[@]}"; then
  echo "=== calibrate the synthetic checkpoints (tiny, then full config); the files come back as gpurun_out/${tag}_calib_*.json"
  timeout -k 10 600 python scripts/calibrate_synthetic.py --config tiny > gpurun_out/${tag}_calib_tiny.log 2>&1; tail -2 gpurun_out/${tag}_calib_tiny.log | cut -c1-300
  timeout -k 10 900 python scripts/calibrate_synthetic.py --config full > gpurun_out/${tag}_calib.log 2>&1; tail -2 gpurun_out/${tag}_calib.log | cut -c1-300
  cp reazonspeech_b200/data/synth_calib_24x1024_v3000_p640_j640_seed0.json gpurun_out/${tag}_calib_full.json
  cp reazonspeech_b200/data/synth_calib_2x256_v127_p128_j128_seed0.json gpurun_out/${tag}_calib_tiny.json
fi
want=tests; if has "${args[@]}"; then
  echo "=== pytest -m gpu"; timeout -k 10 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
  grep -E "FAIL|whole path:|decode kernel alone|near-tie|Error" gpurun_out/${tag}_tests.log | cut -c1-260 | head -60
fi
want=bench; if has "${args[@]}"; then
  echo "=== bench"; timeout -k 10 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -5 gpurun_out/${tag}_bench.err
  python - <<PY
import json
j = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("RTFx", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "clocks", j["clocks"])
print("stage_ms", j["stage_ms"], "tokens/clip", j["config"].get("tokens_per_clip"), "iterations", (j.get("decode_cycles_cta0") or {}).get("iterations"))
print("roofline", {k: j["roofline"][k] for k in ("achieved", "frac", "gemm_ms_per_step")}, [(r["kernel"], round(r["frac"], 3)) for r in j["roofline_hbm"]])
for k, v in list(j["kernel_ms"].items())[:18]: print("   ", k, v)
for k in ("cpu_baseline", "python_api", "python_api_multi_gpu", "decode_sensitivity", "alsd", "config2"):
    if j.get(k): print(k, j[k])
PY
fi
want=sanitize; if has "${args[@]}"; then
  for tool in memcheck racecheck; do
    echo "=== compute-sanitizer $tool"
    timeout -k 10 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
    echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/${tag}_sanitizer_$tool.log | sort | uniq -c | head -12
  done
fi
for a in "${args[@]}"; do
  case "$a" in ncu:*)
    spec=${a#ncu:}; IFS=: read -r pat skip count <<< "$spec"; skip=${skip:-2}; count=${count:-3}     # ncu:<regex>[:<skip>[:<count>]]
    name=$(echo "$pat" | tr -c 'A-Za-z0-9' '_')
    echo "=== ncu --set full on $pat (skip $skip, capture $count)"
    timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $count -f -o gpurun_out/${tag}_ncu_$name python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_$name.log 2>&1; tail -2 gpurun_out/${tag}_ncu_$name.log;;
  launches)
    echo "=== ncu launch list"
    timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches.log 2>&1; tail -1 gpurun_out/${tag}_launches.log;;
  esac
done
