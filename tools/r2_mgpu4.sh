#!/usr/bin/env bash
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2m4; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2m4] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_n${N}.log 2>&1
echo "pytest (all, $N GPUs) rc=$?"; tail -5 $O/pytest_gpu_n${N}.log | cut -c1-300; lap pytest
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29503 \
    bench.py --gpus "$N" --workload c4 --c4-cells ${2:-24} --c4-block ${3:-512} --c4-batch ${4:-128} > $O/c4_n${N}.json 2> $O/c4_n${N}.err
lap "c4 N=$N rc=$?"; grep -v "^$\|^W1\|^\*\*\*\|NCCL version\|OMP_NUM" $O/c4_n${N}.err | tail -5 | cut -c1-300
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2m4/c4_n{n}.json").read().strip().splitlines()[-1])["c4"]
    print({k:d[k] for k in d})
except Exception as e: print("c4 unreadable", e)
PY
