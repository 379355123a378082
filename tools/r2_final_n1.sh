#!/usr/bin/env bash
# Round 2 final N=1 evidence: GPU test suite, the default bench line (as the driver runs it), C1 line, ncu launch list of the
# same command, one ncu --set full capture of the dominant kernel, memcheck over smoke()
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2final; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2final] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
lap "pytest rc=$? $(tail -1 $O/pytest_gpu.log | cut -c1-120)"
timeout -k 5 600 python bench.py > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err
lap "bench c2 (default) rc=$?"
timeout -k 5 60 python bench.py --workload c1 --quick > $O/bench_c1_n1_quick.json 2> $O/bench_c1.err
lap "bench c1 rc=$?"
python - <<'PY'
import json
for f in ("bench_c2_n1", "bench_c1_n1_quick"):
    try:
        d = json.loads(open(f"gpurun_out/r2final/{f}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "value %.2f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["value"],
              "cell ms %.4f" % r["kernel_ms_per_launch"], "clk", r.get("kernel_sm_clock_mhz"), "frac %.3f" % r["frac"], r["bound"],
              "apply ms %.4f" % d["hx_apply"]["ms"], d["chebyshev_filter"]["phase_ms_per_degree"], d["clocks"])
        if d.get("c3_strong"): print("  c3:", {k: d["c3_strong"].get(k) for k in ("ms_per_step", "value", "cell_kernel_ms_per_launch", "cell_kernel_tflops_per_gpu", "roofline")})
        if d.get("cpu_baseline"): print("  cpu:", d["cpu_baseline"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout -k 5 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file $O/launches_filter_step.csv python bench.py --steps 2 --warmup 1 --quick --no-cpu > $O/ncu_bench.log 2>&1
lap "ncu launch list c2 rc=$?"
timeout -k 5 240 ncu --set full --import-source on --clock-control none -k regex:cell_apply_pipe -s 40 -c 1 -f -o $O/pipe_fuse_final \
   python bench.py --quick --no-cpu --steps 3 --warmup 2 > $O/ncu_full.log 2>&1; lap "ncu full rc=$?"
python tools/ncu_summary.py $O/pipe_fuse_final.ncu-rep 30 > $O/pipe_fuse_final_summary.txt 2>&1
ncu -i $O/pipe_fuse_final.ncu-rep --page source --csv > $O/pipe_fuse_final_source.csv 2>/dev/null
python tools/ncu_roles.py $O/pipe_fuse_final_source.csv 8 > $O/pipe_fuse_final_roles.txt 2>&1; lap "summaries"
head -22 $O/pipe_fuse_final_summary.txt | cut -c1-150
timeout -k 5 120 compute-sanitizer --tool memcheck --error-exitcode 3 \
  python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck_smoke.log 2>&1
lap "memcheck rc=$?"; tail -2 $O/memcheck_smoke.log
