#!/usr/bin/env bash
# round 2: halo exchange overlapped with the cell kernel - first run (N GPUs)
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2m2; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2m2] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "processing_order or hx_full or chebyshev_filter or plain_mesh or golden" > $O/pytest_1gpu.log 2>&1
echo "1-GPU subset rc=$?"; tail -3 $O/pytest_1gpu.log; lap "pytest 1gpu"
timeout -k 5 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > $O/pytest_mgpu_n${N}.log 2>&1
echo "multi-GPU parity rc=$?"; tail -6 $O/pytest_mgpu_n${N}.log | cut -c1-400
lap "pytest mgpu"
for OV in 1 0; do
  HXB200_HALO_OVERLAP=$OV timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port $((29500 + OV)) bench.py --gpus "$N" --quick --no-cpu > $O/bench_c2_n${N}_ov${OV}.json 2> $O/bench_c2_n${N}_ov${OV}.err
  lap "bench N=$N overlap=$OV rc=$?"
done
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
for ov in (1, 0):
    try:
        d = json.loads(open(f"gpurun_out/r2m2/bench_c2_n{n}_ov{ov}.json").read().strip().splitlines()[-1])
        print(f"N={n} overlap={ov}: value %.2f  ms/step %.3f  cell ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"]),
              d["chebyshev_filter"]["phase_ms_per_degree"], d["run"]["halo_transport"])
    except Exception as e:
        print(f"N={n} overlap={ov}: unreadable: {e}")
PY
