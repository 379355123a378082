#!/usr/bin/env bash
# round 2, call 3: what bounds the pipelined cell kernel?  timing with parts switched off + one full ncu capture
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c3; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c3] $1 at $(( $(date +%s) - T0 )) s"; }
for d in 0 1 2 3 4 8 16 32 12 7 15; do
  HXB200_CELL_DIAG=$d timeout -k 5 60 python bench.py --quick --no-cpu --steps 5 --warmup 3 > $O/diag$d.json 2> $O/diag$d.err; lap "diag$d rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c3/diag*.json"), key=lambda s:int(s.split('diag')[-1].split('.')[0])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout -k 5 240 ncu --set full --import-source on --clock-control none -k regex:cell_apply_pipe -s 40 -c 1 -f -o $O/pipe_fuse \
   python bench.py --quick --no-cpu --steps 3 --warmup 2 > $O/ncu.log 2>&1; lap "ncu rc=$?"
python tools/ncu_summary.py $O/pipe_fuse.ncu-rep 45 > $O/pipe_fuse_summary.txt 2>&1; lap summary
head -50 $O/pipe_fuse_summary.txt
