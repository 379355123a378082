#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c16; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c16] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 300 python -X faulthandler bench.py --workload c4 --c4-cells 16 --c4-block 256 --c4-batch 64 > $O/c4_small.json 2> $O/c4_small.err; lap "c4 small rc=$?"
grep -v "^$" $O/c4_small.err | tail -25 | cut -c1-200
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c16/c4_small.json").read().strip().splitlines()[-1])["c4"]
    print({k:d[k] for k in d if k not in ("workload",)})
except Exception as e: print("c4 unreadable", e)
PY
timeout -k 5 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log; lap pytest
