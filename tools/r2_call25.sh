#!/usr/bin/env bash
# Round 2, session 2: C1 (4012 small cells) against the order delay D of the processing order; full GPU suite with the new default build
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c25; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c25] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 300 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
lap "pytest rc=$? $(tail -1 $O/pytest_gpu.log | cut -c1-120)"
for D in 0 37 74 148 296 592 1184; do
  HXB200_ORDER_DELAY=$D timeout -k 5 60 python bench.py --workload c1 --quick --no-cpu > $O/c1_D$D.json 2> $O/c1_D$D.err; lap "c1 D=$D rc=$?"
done
python - <<'PY'
import json
for D in (0,37,74,148,296,592,1184):
    try:
        d=json.loads(open(f"gpurun_out/r2c25/c1_D{D}.json").read().strip().splitlines()[-1])
        print("D",D,": cell %.4f apply %.4f ms/step %.3f value %.2f"%(d["roofline"]["kernel_ms_per_launch"],d["hx_apply"]["ms"],d["ms_per_step"],d["value"]))
    except Exception as e: print(D,"unreadable",e)
PY
