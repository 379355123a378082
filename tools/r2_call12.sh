#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c12; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c12] $1 at $(( $(date +%s) - T0 )) s"; }
for d in 0 128 256 384; do
  HXB200_CELL_DIAG=$d timeout -k 5 60 python bench.py --quick --no-cpu > $O/diag$d.json 2> $O/diag$d.err; lap "diag$d rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c12/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "clk %.1f" % d["roofline"].get("kernel_sm_clock_mhz",0), "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
for d in 0 384; do
HXB200_CELL_DIAG=$d timeout -k 5 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cell_apply_pipe -s 40 -c 2 --csv --log-file $O/dram_diag$d.csv \
   python bench.py --quick --no-cpu --steps 3 --warmup 2 > $O/ncu$d.log 2>&1; lap "ncu diag$d rc=$?"
grep -v "^==" $O/dram_diag$d.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | tail -8
done
