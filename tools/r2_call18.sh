#!/usr/bin/env bash
# Round 2, session 2, call 2: c112 (RB=2, 112/56 registers) full GPU test suite; A prefetch variant; order-delay retune
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c18; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c18] $1 at $(( $(date +%s) - T0 )) s"; }
E=$PWD/dft_efe_b200/lib/exp
HXB200_LIB=$E/libhxb200_c112.so timeout -k 5 300 python -m pytest tests -x -q -m gpu > $O/pytest_c112.log 2>&1
lap "pytest c112 rc=$? $(tail -1 $O/pytest_c112.log | cut -c1-120)"
run() { # name lib [env...]
  local n=$1 l=$2; shift 2
  env "$@" HXB200_LIB=$l timeout -k 5 80 python bench.py --quick --no-cpu > $O/$n.json 2> $O/$n.err; lap "$n rc=$?"
}
run c112 $E/libhxb200_c112.so
run p112 $E/libhxb200_p112.so
run c112_D296 $E/libhxb200_c112.so HXB200_ORDER_DELAY=296
run c112_D1184 $E/libhxb200_c112.so HXB200_ORDER_DELAY=1184
run p112_D1184 $E/libhxb200_p112.so HXB200_ORDER_DELAY=1184
run c112_b $E/libhxb200_c112.so
run p112_b $E/libhxb200_p112.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c18/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "clk %.1f" % d["roofline"].get("kernel_sm_clock_mhz",0), "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"], "value %.2f" % d["value"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
