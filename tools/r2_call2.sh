#!/usr/bin/env bash
# round 2, call 2: first run of the pipelined cell kernel (parity suite, then A/B against the round-1 kernel)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c2; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c2] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; rc=$?
echo "smoke rc=$rc"; tail -3 $O/smoke.log; lap smoke
if [ $rc -ne 0 ]; then exit 1; fi
HXB200_EXPERIMENTS=1 timeout -k 5 500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $O/pytest.log; lap pytest
timeout -k 5 60 python bench.py --quick > $O/new_kc2.json 2> $O/new_kc2.err; lap "new kc2 rc=$?"
HXB200_CELL_KC=4 timeout -k 5 60 python bench.py --quick > $O/new_kc4.json 2> $O/new_kc4.err; lap "new kc4 rc=$?"
HXB200_CELL_KERNEL=v1 timeout -k 5 60 python bench.py --quick > $O/v1.json 2> $O/v1.err; lap "v1 rc=$?"
timeout -k 5 60 python bench.py --workload c1 --quick > $O/new_c1.json 2> $O/new_c1.err; lap "c1 new rc=$?"
HXB200_CELL_KERNEL=v1 timeout -k 5 60 python bench.py --workload c1 --quick > $O/v1_c1.json 2> $O/v1_c1.err; lap "c1 v1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c2/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.2f" % d["value"], "ms/step %.3f" % d["ms_per_step"],
              "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "frac %.3f" % d["roofline"]["frac"],
              "apply ms %.4f" % d["hx_apply"]["ms"], d["chebyshev_filter"]["phase_ms_per_degree"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
