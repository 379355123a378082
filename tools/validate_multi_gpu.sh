#!/usr/bin/env bash
# Multi-GPU validation, to be run with `gpurun --gpus N -- bash tools/validate_multi_gpu.sh N` (N = 2, 4 or 8; charged N x).
#   1. the multi-GPU parity cases that fit N GPUs (both halo transports), with serialised launches (the round-1 default for
#      multi-rank plans) and with programmatic dependent launch forced on (HXB200_PDL=1: not yet run on N >= 2);
#   2. the weak-scaling bench line at N for both launch modes.
# If (1) is green with HXB200_PDL=1 and (2) shows a gain, make dependent launch the default for multi-rank plans too
# (api.cu: pdl_enabled / g_multirank_plan).
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T0=$(date +%s)
lap() { echo "[validate-mgpu] $1 at $(( $(date +%s) - T0 )) s"; }
for PDL in 0 1; do
  HXB200_PDL=$PDL timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu_n${N}_pdl${PDL}.log 2>&1
  echo "multi-GPU parity (PDL=$PDL) rc=$?"; tail -3 gpurun_out/pytest_mgpu_n${N}_pdl${PDL}.log
  lap "pytest PDL=$PDL"
done
for PDL in 0 1; do
  HXB200_PDL=$PDL timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
    --master-port $((29500 + PDL)) bench.py --gpus "$N" --quick > gpurun_out/bench_c2_n${N}_pdl${PDL}.json 2> gpurun_out/bench_c2_n${N}_pdl${PDL}.err
  lap "bench N=$N PDL=$PDL rc=$?"
done
python - "$N" <<'EOF'
import json, sys
n = sys.argv[1]
for pdl in (0, 1):
    try:
        d = json.loads(open(f"gpurun_out/bench_c2_n{n}_pdl{pdl}.json").read().strip().splitlines()[-1])
        print(f"N={n} PDL={pdl}: value %.2f  ms/step %.3f  cell ms %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"]),
              d["chebyshev_filter"]["phase_ms_per_degree"], d["run"]["halo_transport"])
    except Exception as e:
        print(f"N={n} PDL={pdl}: unreadable: {e}")
EOF
