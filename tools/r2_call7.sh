#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c7; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c7] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; rc=$?
echo "smoke rc=$rc"; tail -3 $O/smoke.log; lap smoke
if [ $rc -ne 0 ]; then exit 1; fi
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest.log; lap pytest
for v in 0 1 2 3; do
  HXB200_CELL_VARIANT=$v timeout -k 5 60 python bench.py --quick --no-cpu > $O/var$v.json 2> $O/var$v.err; lap "var$v rc=$?"
done
HXB200_CELL_VARIANT=1 HXB200_CELL_DIAG=8 timeout -k 5 60 python bench.py --quick --no-cpu > $O/var1_diag8.json 2> $O/var1_diag8.err; lap "var1 diag8 rc=$?"
timeout -k 5 60 python bench.py --workload c1 --quick --no-cpu > $O/c1.json 2> $O/c1.err; lap "c1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c7/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"], "value %.2f" % d["value"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
