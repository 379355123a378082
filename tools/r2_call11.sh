#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c11; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c11] $1 at $(( $(date +%s) - T0 )) s"; }
for D in 74 148 296 592 1184; do
  HXB200_ORDER_DELAY=$D timeout -k 5 60 python bench.py --quick --no-cpu > $O/new_D$D.json 2> $O/new_D$D.err; lap "new D$D rc=$?"
done
for D in 296 592; do
  HXB200_CELL_KERNEL=v1 HXB200_ORDER_DELAY=$D timeout -k 5 60 python bench.py --quick --no-cpu > $O/v1_D$D.json 2> $O/v1_D$D.err; lap "v1 D$D rc=$?"
done
HXB200_NO_DISCARD=1 timeout -k 5 60 python bench.py --quick --no-cpu > $O/new_nodiscard.json 2> $O/new_nodiscard.err; lap "nodiscard rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c11/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "clk %.1f" % d["roofline"].get("kernel_sm_clock_mhz",0), "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout -k 5 240 ncu --set full --import-source on --clock-control none -k regex:cell_apply_pipe -s 40 -c 1 -f -o $O/pipe_fuse \
   python bench.py --quick --no-cpu --steps 3 --warmup 2 > $O/ncu.log 2>&1; lap "ncu rc=$?"
python tools/ncu_summary.py $O/pipe_fuse.ncu-rep 30 > $O/pipe_fuse_summary.txt 2>&1; lap summary
head -40 $O/pipe_fuse_summary.txt
