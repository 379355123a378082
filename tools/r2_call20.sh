#!/usr/bin/env bash
# Round 2, session 2, call 5: L2 eviction-policy variants of the scatter / gather side (HX_PIPE_EF bits) against the default
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c20; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c20] $1 at $(( $(date +%s) - T0 )) s"; }
E=$PWD/dft_efe_b200/lib/exp
M=$PWD/dft_efe_b200/lib/libhxb200.so
run() { local n=$1 l=$2; shift 2
  HXB200_LIB=$l timeout -k 5 150 python bench.py --quick --no-cpu "$@" > $O/$n.json 2> $O/$n.err; lap "$n rc=$?"; }
for v in ef5 ef13; do
  HXB200_LIB=$E/libhxb200_$v.so timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hx or cheb or determin or golden" > $O/pytest_$v.log 2>&1
  lap "pytest $v rc=$? $(tail -1 $O/pytest_$v.log | cut -c1-120)"
done
run main $M
for v in ef1 ef3 ef4 ef5 ef13; do run $v $E/libhxb200_$v.so; done
run main_b $M
for v in ef5 ef13; do run ${v}_b $E/libhxb200_$v.so; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c20/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f.split('/')[-1], "cell ms %.4f" % r["kernel_ms_per_launch"], "clk %.1f" % r.get("kernel_sm_clock_mhz",0), "cycles %.0fk" % (r["kernel_ms_per_launch"]*r.get("kernel_sm_clock_mhz",0)), "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"], "value %.2f" % d["value"], "frac %.3f" % r["frac"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
for v in main ef5 ef13; do
  L=$E/libhxb200_$v.so; [ $v = main ] && L=$M
  HXB200_LIB=$L timeout -k 5 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:cell_apply_pipe -s 40 -c 2 --csv --log-file $O/dram_$v.csv \
     python bench.py --quick --no-cpu --steps 3 --warmup 2 > $O/ncu_$v.log 2>&1; lap "ncu $v rc=$?"
  grep -v "^==" $O/dram_$v.csv | awk -F'","' '{print $(NF-3), $(NF-2), $(NF-1), $NF}' | tail -8
done
