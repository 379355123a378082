#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c6; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c6] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; rc=$?
echo "smoke rc=$rc"; tail -3 $O/smoke.log; lap smoke
if [ $rc -ne 0 ]; then exit 1; fi
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $O/pytest.log; lap pytest
for d in 0 8 12 15 4; do
  HXB200_CELL_DIAG=$d timeout -k 5 60 python bench.py --quick --no-cpu > $O/diag$d.json 2> $O/diag$d.err; lap "diag$d rc=$?"
done
HXB200_CELL_KC=4 timeout -k 5 60 python bench.py --quick --no-cpu > $O/kc4.json 2> $O/kc4.err; lap "kc4 rc=$?"
timeout -k 5 60 python bench.py --workload c1 --quick --no-cpu > $O/c1.json 2> $O/c1.err; lap "c1 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c6/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"], "value %.2f" % d["value"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout -k 5 240 ncu --set full --import-source on --clock-control none -k regex:cell_apply_pipe -s 40 -c 1 -f -o $O/pipe_fuse \
   python bench.py --quick --no-cpu --steps 3 --warmup 2 > $O/ncu.log 2>&1; lap "ncu rc=$?"
python tools/ncu_summary.py $O/pipe_fuse.ncu-rep 45 > $O/pipe_fuse_summary.txt 2>&1; lap summary
head -60 $O/pipe_fuse_summary.txt
