#!/usr/bin/env bash
# Round 2, session 2: small cells (order 2 / 3) - scatter-side variants: publisher warp (sh), + two accumulator tiles (shn2),
# four rows per batch in the plain apply (rbp4), two tiles only (n2)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c22; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c22] $1 at $(( $(date +%s) - T0 )) s"; }
E=$PWD/dft_efe_b200/lib/exp
M=$PWD/dft_efe_b200/lib/libhxb200.so
for v in sh shn2; do
  HXB200_LIB=$E/libhxb200_$v.so timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hx or cheb or determin or golden" > $O/pytest_$v.log 2>&1
  lap "pytest $v rc=$? $(tail -1 $O/pytest_$v.log | cut -c1-120)"
done
for v in main sh shn2 rbp4 n2; do
  L=$E/libhxb200_$v.so; [ $v = main ] && L=$M
  HXB200_LIB=$L timeout -k 5 300 python tools/sweep.py --points 3:32,3:128,2:128 --out $O/sweep_$v.json > $O/sweep_$v.log 2>&1; lap "sweep $v rc=$?"
  HXB200_LIB=$L timeout -k 5 60 python bench.py --workload c1 --quick --no-cpu > $O/c1_$v.json 2> $O/c1_$v.err; lap "c1 $v rc=$?"
done
python - <<'PY'
import json,glob
for v in ("main","sh","shn2","rbp4","n2"):
    try:
        d=json.load(open(f"gpurun_out/r2c22/sweep_{v}.json"))
        print(v, [(r["p"],r["B"],"apply %.3f cell %.3f filt %.3f frac %.3f"%(r["apply_ms"],r["cell_kernel_ms"],r["filter_ms_per_degree"],r["roofline_frac"])) for r in d["points"]])
    except Exception as e: print(v,"sweep unreadable",e)
    try:
        d=json.loads(open(f"gpurun_out/r2c22/c1_{v}.json").read().strip().splitlines()[-1])
        print(v,"c1: cell %.4f apply %.4f ms/step %.3f"%(d["roofline"]["kernel_ms_per_launch"],d["hx_apply"]["ms"],d["ms_per_step"]))
    except Exception as e: print(v,"c1 unreadable",e)
PY
