#!/usr/bin/env bash
# Round 2, session 2, call 3: sleep-poll wait of the scatter warps (s100 / s250) against the new default; C1 and C3 shapes with the new default
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c19; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c19] $1 at $(( $(date +%s) - T0 )) s"; }
E=$PWD/dft_efe_b200/lib/exp
M=$PWD/dft_efe_b200/lib/libhxb200.so
run() { # name lib [bench args...]
  local n=$1 l=$2; shift 2
  HXB200_LIB=$l timeout -k 5 150 python bench.py --quick --no-cpu "$@" > $O/$n.json 2> $O/$n.err; lap "$n rc=$?"
}
run main $M
run s100 $E/libhxb200_s100.so
run s250 $E/libhxb200_s250.so
run main_b $M
run s100_b $E/libhxb200_s100.so
run s250_b $E/libhxb200_s250.so
run c1_main $M --workload c1
run c1_s100 $E/libhxb200_s100.so --workload c1
run c3_main $M --workload c3 --steps 2 --warmup 1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c19/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f.split('/')[-1], "cell ms %.4f" % r["kernel_ms_per_launch"], "clk %.1f" % r.get("kernel_sm_clock_mhz",0), "cycles %.0fk" % (r["kernel_ms_per_launch"]*r.get("kernel_sm_clock_mhz",0)), "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"], "value %.2f" % d["value"], "frac %.3f" % r["frac"], r["bound"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
