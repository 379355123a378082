#!/usr/bin/env bash
# round 2, call 1: gated tests (periodic + experimental variants) and the prepared A/Bs
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
T0=$(date +%s)
lap() { echo "[r2c1] $1 at $(( $(date +%s) - T0 )) s"; }
HXB200_EXPERIMENTS=1 timeout 300 python -m pytest tests -m gpu -q -k "experimental or periodic" > $O/pytest_exp.log 2>&1
echo "pytest exp rc=$?"; tail -15 $O/pytest_exp.log
lap pytest
for m in 0 1 2; do HXB200_PRODUCER_ADDR=$m timeout 60 python bench.py --quick > $O/prod$m.json 2> $O/prod$m.err; lap "prod$m rc=$?"; done
for m in 2 3; do HXB200_CELL_MINB=$m timeout 60 python bench.py --workload c1 --quick > $O/minb$m.json 2> $O/minb$m.err; lap "minb$m rc=$?"; done
HXB200_CELL_MTW=1 timeout 60 python bench.py --workload c1 --quick > $O/mtw1.json 2> $O/mtw1.err; lap "mtw1 rc=$?"
for m in 0 1; do HXB200_SPLIT_ROWLIST=$m timeout 90 python bench.py --workload c2a --quick > $O/c2a_split$m.json 2> $O/c2a_split$m.err; lap "c2a$m rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c1/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.2f" % d["value"], "ms/step %.3f" % d["ms_per_step"],
              "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "frac %.3f" % d["roofline"]["frac"],
              "apply ms %.4f" % d["hx_apply"]["ms"], d["chebyshev_filter"]["phase_ms_per_degree"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
