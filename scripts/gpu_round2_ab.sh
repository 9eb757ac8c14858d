#!/bin/bash
# First GPU call of round 2: validate and A/B the experiments that round 1 left unmeasured.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_ab.sh'
# Writes gpurun_out/ab_<variant>.json (one bench line each) and gpurun_out/ab_summary.txt.
# An experiment changes the numerics slightly, which moves the synthetic checkpoint's token count:
# compare ms of the ENCODER stages and the per-kernel table, not only the headline.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "=== default suite"; timeout -k 10 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1; tail -3 gpurun_out/r2a_tests.log
grep -E "clip[0-9]+:|ragged|32 x 30 s|FAIL|near-tie|utt[0-9]" gpurun_out/r2a_tests.log | cut -c1-220 | head -80
for t in test_gpu_decode_edge_cases test_gpu_asymmetric_window test_gpu_logmel_b test_gpu_ln_fold test_gpu_splitk; do   # one process each, under timeout: a scheduling bug in an experiment would hang
  echo "=== experiment $t"; RS_RUN_EXPERIMENTS=1 timeout -k 10 300 python -m pytest tests/experiments/$t.py -m gpu -q -s -x -p no:cacheprovider 2>&1 | grep -E "utt|rep=|passed|failed|Error|error" | cut -c1-200
done
echo "=== frontend tests on the log-mel variant B"; RS_LOGMEL_VARIANT=B timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parakeet.py tests/test_gpu_nemo_port.py -m gpu -q -x -k "logmel or parakeet or nemo_port or end_to_end" -p no:cacheprovider 2>&1 | tail -2
echo "=== GEMM and encoder tests on the 6-stage ring"; RS_GEMM_STAGES=6 timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm or encoder" -p no:cacheprovider 2>&1 | tail -2
# compile-time variant with programmatic dependent launch (csrc/common.cuh RS_PDL): build it here BEFORE the gpurun call
# (`python -m reazonspeech_b200.build --variant pdl`, the .so travels); rebuilt on the box only if it is stale
python -m reazonspeech_b200.build --variant pdl | tail -1        # mtime-based: a no-op when it was built after the last source change
echo "=== tiny-config -m gpu suite on the PDL variant"; RS_ENGINE_VARIANT=pdl timeout -k 10 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nemo_port.py tests/test_gpu_parakeet.py -m gpu -q -x -k "not long_form" -p no:cacheprovider 2>&1 | tail -3
run() {   # name, env assignments...
  local name=$1; shift
  echo "=== bench $name"
  env "$@" timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -n 5 gpurun_out/ab_$name.err
}
run default RS_NONE=1
run stages6 RS_GEMM_STAGES=6
run lnfold RS_LN_FOLD=1
run logmelb RS_LOGMEL_VARIANT=B
run splitk RS_GEMM_SPLITK=1
run lnfold_splitk RS_LN_FOLD=1 RS_GEMM_SPLITK=1
run pdl RS_ENGINE_VARIANT=pdl
run pdl_lnfold_splitk RS_ENGINE_VARIANT=pdl RS_LN_FOLD=1 RS_GEMM_SPLITK=1
python - <<'PY' | tee gpurun_out/ab_summary.txt
import json, glob, os
rows = []
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e); continue
    k = j.get("kernel_ms", {})
    gem = sum(v["ms"] for n, v in k.items() if n.startswith("gemm"))
    ln = sum(v["ms"] for n, v in k.items() if "layernorm" in n)
    lm = sum(v["ms"] for n, v in k.items() if "logmel" in n)
    rows.append((os.path.basename(f)[3:-5], j["value"], j["e2e"]["value"], j["ms_per_step"], gem, ln, j.get("roofline", {}).get("frac"), lm))
print(f"{'variant':18s} {'RTFx':>9s} {'e2e':>9s} {'ms/step':>8s} {'gemm ms':>8s} {'LN ms':>7s} {'gemm frac':>9s} {'logmel ms':>9s}")
for r in rows:
    print(f"{r[0]:18s} {r[1]:9.0f} {r[2]:9.0f} {r[3]:8.2f} {r[4]:8.2f} {r[5]:7.2f} {r[6] if r[6] is None else round(r[6], 3)!s:>9s} {r[7]:9.3f}")
PY
# compute-sanitizer over the tiny-config smoke path with the default kernels (last: a tool hang must not cost the rest)
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer $tool"
  timeout -k 10 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/r2a_sanitizer_$tool.log | sort | uniq -c | head -20
done
