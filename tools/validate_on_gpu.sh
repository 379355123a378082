#!/usr/bin/env bash
# One-call validation of the current tree on a B200 box (run through gpurun from the repo root):
#   parity suite with programmatic dependent launch on (its own test compares both launch modes bit for bit),
#   A/B bench lines for both launch modes (c2 = BASELINE configs[1], c1 = configs[0]),
#   compute-sanitizer memcheck over smoke(), ncu launch list of the bench step.
# Everything lands in gpurun_out/; every leg has its own timeout so one slow leg cannot eat the others.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T0=$(date +%s)
lap() { echo "[validate] $1 at $(( $(date +%s) - T0 )) s"; }

export HXB200_PDL=1
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_pdl1.log 2>&1
RC=$?
tail -3 gpurun_out/pytest_gpu_pdl1.log
lap "pytest (PDL on) rc=$RC"
if [ $RC -ne 0 ]; then
  HXB200_PDL=0 timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_pdl0.log 2>&1
  echo "pytest (PDL off) rc=$?"; tail -3 gpurun_out/pytest_gpu_pdl0.log
  lap "pytest (PDL off)"
fi

HXB200_PDL=1 timeout 240 python bench.py > gpurun_out/bench_c2_pdl1.json 2> gpurun_out/bench_c2_pdl1.err
lap "bench c2 PDL on rc=$?"
HXB200_PDL=0 timeout 150 python bench.py --quick > gpurun_out/bench_c2_pdl0_quick.json 2> gpurun_out/bench_c2_pdl0.err
lap "bench c2 PDL off rc=$?"
HXB200_PDL=1 timeout 90 python bench.py --workload c1 --quick > gpurun_out/bench_c1_pdl1_quick.json 2> gpurun_out/bench_c1_pdl1.err
lap "bench c1 PDL on rc=$?"
HXB200_PDL=0 timeout 90 python bench.py --workload c1 --quick > gpurun_out/bench_c1_pdl0_quick.json 2> gpurun_out/bench_c1_pdl0.err
lap "bench c1 PDL off rc=$?"
python - <<'EOF'
import json
for f in ("bench_c2_pdl1", "bench_c2_pdl0_quick", "bench_c1_pdl1_quick", "bench_c1_pdl0_quick"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.2f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["value"],
              "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "frac %.3f" % d["roofline"]["frac"],
              "apply ms %.4f" % d["hx_apply"]["ms"], d["chebyshev_filter"]["phase_ms_per_degree"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
EOF

HXB200_PDL=1 timeout 240 compute-sanitizer --tool memcheck --error-exitcode 3 \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_smoke.log
lap "memcheck"

HXB200_PDL=1 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_filter_step.csv python bench.py --steps 2 --warmup 1 --quick > gpurun_out/ncu_bench.log 2>&1
lap "ncu launch list c2 rc=$?"
HXB200_PDL=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_filter_step_c1.csv python bench.py --workload c1 --steps 2 --warmup 1 --quick > gpurun_out/ncu_bench_c1.log 2>&1
lap "ncu launch list c1 rc=$?"
