#!/usr/bin/env bash
# round 2, first multi-GPU call: parity of the serial halo path with the pipelined cell kernel (both launch modes), bench lines
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2m1; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2m1] $1 at $(( $(date +%s) - T0 )) s"; }
for PDL in 1 0; do
  HXB200_PDL=$PDL timeout -k 5 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > $O/pytest_mgpu_n${N}_pdl${PDL}.log 2>&1
  echo "multi-GPU parity (PDL=$PDL) rc=$?"; tail -4 $O/pytest_mgpu_n${N}_pdl${PDL}.log
  lap "pytest PDL=$PDL"
done
for PDL in 0 1; do
  HXB200_PDL=$PDL timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port $((29500 + PDL)) bench.py --gpus "$N" --quick --no-cpu > $O/bench_c2_n${N}_pdl${PDL}.json 2> $O/bench_c2_n${N}_pdl${PDL}.err
  lap "bench N=$N PDL=$PDL rc=$?"
done
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
for pdl in (0, 1):
    try:
        d = json.loads(open(f"gpurun_out/r2m1/bench_c2_n{n}_pdl{pdl}.json").read().strip().splitlines()[-1])
        print(f"N={n} PDL={pdl}: value %.2f  ms/step %.3f  cell ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"]),
              d["chebyshev_filter"]["phase_ms_per_degree"], d["run"]["halo_transport"])
    except Exception as e:
        print(f"N={n} PDL={pdl}: unreadable: {e}")
PY
