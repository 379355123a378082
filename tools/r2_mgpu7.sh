#!/usr/bin/env bash
# round 2: what the in-kernel halo exchange buys - the C2 filter step on N GPUs with the overlapped and with the serial exchange
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2m7; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2m7] $1 at $(( $(date +%s) - T0 )) s"; }
for V in 1 0 1 0; do
  n=overlap$V; [ -e $O/$n.json ] && n=${n}_b
  HXB200_HALO_OVERLAP=$V timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29520 + V)) \
    bench.py --gpus "$N" --quick --no-cpu > $O/$n.json 2> $O/$n.err
  lap "bench N=$N HALO_OVERLAP=$V rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m7/overlap*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.2f ms/step %.3f cell %.4f apply %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["hx_apply"]["ms"]), d["chebyshev_filter"]["phase_ms_per_degree"], d["run"].get("halo_overlap"))
    except Exception as e:
        print(f, "unreadable", e)
PY
