#!/usr/bin/env bash
# Short form of validate_on_gpu.sh: parity suite, the default bench line, the c1 line, memcheck over smoke().
set -u
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T0=$(date +%s)
lap() { echo "[validate] $1 at $(( $(date +%s) - T0 )) s"; }
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
lap "pytest"
timeout 60 python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err
lap "bench c2 rc=$?"
timeout 30 python bench.py --workload c1 --quick > gpurun_out/bench_c1_final_quick.json 2> gpurun_out/bench_c1_final.err
lap "bench c1 rc=$?"
python - <<'EOF'
import json
for f in ("bench_c2_final", "bench_c1_final_quick"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.2f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["value"],
              "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "frac %.3f" % d["roofline"]["frac"],
              "apply ms %.4f" % d["hx_apply"]["ms"], d["chebyshev_filter"]["phase_ms_per_degree"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
EOF
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 3 \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_smoke.log
lap "memcheck"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_filter_step.csv python bench.py --steps 2 --warmup 1 --quick > gpurun_out/ncu_bench.log 2>&1
lap "ncu launch list c2 rc=$?"
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_filter_step_c1.csv python bench.py --workload c1 --steps 2 --warmup 1 --quick > gpurun_out/ncu_bench_c1.log 2>&1
lap "ncu launch list c1 rc=$?"
