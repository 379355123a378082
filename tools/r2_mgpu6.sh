#!/usr/bin/env bash
# round 2, session 2: fused projector exchange (single-block accumulate + update) - parity on N GPUs and A/B of the filter step
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2m6; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2m6] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/pytest_mgpu_n${N}.log 2>&1
lap "pytest multi-GPU rc=$? $(tail -1 $O/pytest_mgpu_n${N}.log | cut -c1-150)"
grep -h "halo transport" $O/pytest_mgpu_n${N}.log | cut -c1-400 | head -4
for F in 1 0 1 0; do
  n=fused$F; [ -e $O/$n.json ] && n=${n}_b
  HXB200_NL_FUSED=$F timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29510 + F)) \
    bench.py --gpus "$N" --quick --no-cpu > $O/$n.json 2> $O/$n.err
  lap "bench N=$N NL_FUSED=$F rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m6/fused*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "value %.2f ms/step %.3f cell %.4f apply %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["hx_apply"]["ms"]), d["chebyshev_filter"]["phase_ms_per_degree"], (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(f, "unreadable", e)
PY
