#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c8; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c8] $1 at $(( $(date +%s) - T0 )) s"; }
# power / clocks sampled at 50 ms during the runs
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active,temperature.gpu --format=csv -lms 50 > $O/smi.csv 2>&1 &
SMI=$!
for d in 0 12 4 8; do
  HXB200_CELL_DIAG=$d timeout -k 5 60 python bench.py --quick --no-cpu > $O/diag$d.json 2> $O/diag$d.err; lap "diag$d rc=$?"
  echo "MARK diag$d $(date +%H:%M:%S.%N)" >> $O/marks.txt
done
kill $SMI
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c8/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "clk %.1f" % d["roofline"]["kernel_sm_clock_mhz"], "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
awk -F, 'NR>1{print $2,$4,$5}' $O/smi.csv | sort | uniq -c | sort -rn | head -40
