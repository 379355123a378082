#!/usr/bin/env bash
# round 2, session 2: the driver's multi-GPU bench line at N GPUs (parity block, overlapped halo, c3_strong block), timed
set -u
N=${1:-4}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2m5; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2m5] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29500 \
    bench.py --gpus "$N" > $O/bench_n${N}.json 2> $O/bench_n${N}.err
lap "bench N=$N rc=$?"; grep -v "^$\|^W1\|^\*\*\*\|NCCL version\|OMP_NUM" $O/bench_n${N}.err | tail -4 | cut -c1-300
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/r2m5/bench_n{n}.json").read().strip().splitlines()[-1])
print("N=%s value %.2f ms/step %.3f cell %.4f" % (n, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"]), d["chebyshev_filter"]["phase_ms_per_degree"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_batch"], "single", d["e2e"]["single_block"]["value"])
print("parity", {k: v for k, v in (d.get("parity") or {}).items() if k not in ("what", "tolerance")})
c3 = d.get("c3_strong") or {}
print("c3", {k: c3.get(k) for k in ("ms_per_step", "value", "cell_kernel_ms_per_launch", "cell_kernel_tflops_per_gpu", "phase_ms_per_degree", "host_build_s", "error", "skipped")})
PY
