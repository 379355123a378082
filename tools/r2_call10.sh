#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c10; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c10] $1 at $(( $(date +%s) - T0 )) s"; }
for d in 12 13 14 76 77; do
  HXB200_CELL_DIAG=$d timeout -k 5 60 python bench.py --quick --no-cpu --steps 8 --warmup 3 > $O/diag$d.json 2> $O/diag$d.err; lap "diag$d rc=$?"
done
for d in 12 13; do
  HXB200_CELL_KC=4 HXB200_CELL_DIAG=$d timeout -k 5 60 python bench.py --quick --no-cpu --steps 8 --warmup 3 > $O/kc4_diag$d.json 2> $O/kc4_diag$d.err; lap "kc4 diag$d rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c10/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "clk %.1f" % d["roofline"]["kernel_sm_clock_mhz"], "apply ms %.4f" % d["hx_apply"]["ms"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
