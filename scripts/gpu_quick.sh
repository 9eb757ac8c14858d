cd "${GRAFT_REPO_ROOT:-.}"
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout -k 10 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "Error|error:|passed|failed" | head -6 | cut -c1-300
