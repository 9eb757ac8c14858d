cd "${GRAFT_REPO_ROOT:-.}"
python scripts/probes/ln_probe.py 2>&1 | tail -5
