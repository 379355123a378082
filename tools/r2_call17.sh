#!/usr/bin/env bash
# Round 2, session 2, call 1: A/B of the scatter-side variants of the pipelined cell kernel (tools/build_variants.py)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c17; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c17] $1 at $(( $(date +%s) - T0 )) s"; }
E=dft_efe_b200/lib/exp
for v in h112 h112n2; do
  HXB200_LIB=$PWD/$E/libhxb200_$v.so timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hx or cheb or determin or golden" > $O/pytest_$v.log 2>&1
  lap "pytest $v rc=$? $(tail -1 $O/pytest_$v.log | cut -c1-120)"
done
for v in base c112 h112 h112n2 n2c112 base h112; do
  L=$PWD/$E/libhxb200_$v.so; [ $v = base ] && L=$PWD/dft_efe_b200/lib/libhxb200.so
  n=$v; [ -e $O/$v.json ] && n=${v}_b
  HXB200_LIB=$L timeout -k 5 80 python bench.py --quick --no-cpu > $O/$n.json 2> $O/$n.err; lap "$n rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c17/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "cell ms %.4f" % d["roofline"]["kernel_ms_per_launch"], "clk %.1f" % d["roofline"].get("kernel_sm_clock_mhz",0), "apply ms %.4f" % d["hx_apply"]["ms"], "ms/step %.3f" % d["ms_per_step"], "value %.2f" % d["value"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout -k 5 240 env HXB200_LIB=$PWD/$E/libhxb200_h112.so ncu --set full --import-source on --clock-control none -k regex:cell_apply_pipe -s 40 -c 1 -f -o $O/pipe_h112 \
   python bench.py --quick --no-cpu --steps 3 --warmup 2 > $O/ncu.log 2>&1; lap "ncu rc=$?"
python tools/ncu_summary.py $O/pipe_h112.ncu-rep 40 > $O/pipe_h112_summary.txt 2>&1; lap summary
head -45 $O/pipe_h112_summary.txt | cut -c1-180
ls -la $O | head -30
