#!/usr/bin/env bash
# Round 2, session 2: 128 x 128-tile Gram kernel - parity, subspace bench A/B against the 64-tile kernel, DRAM bytes at B = 1024
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c21; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c21] $1 at $(( $(date +%s) - T0 )) s"; }
timeout -k 5 300 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
lap "pytest rc=$? $(tail -1 $O/pytest_gpu.log | cut -c1-120)"
timeout -k 5 200 python tools/subspace_bench.py --widths 96,128,256,512,1024 --out $O/subspace_tile128.json > $O/sub128.log 2>&1; lap "subspace 128 rc=$?"
HXB200_GRAM_TILE64=1 timeout -k 5 200 python tools/subspace_bench.py --widths 96,128,256,512,1024 --out $O/subspace_tile64.json > $O/sub64.log 2>&1; lap "subspace 64 rc=$?"
python - <<'PY'
import json
for n in ("tile128","tile64"):
    try:
        d=json.load(open(f"gpurun_out/r2c21/subspace_{n}.json"))
        for r in d["points"]:
            print(n, "B", r["B"], "gram ms %.3f frac %.3f" % (r["gram"]["ms"], r["gram"]["frac_dmma"]), "rotate ms %.3f frac %.3f" % (r["rotate"]["ms"], r["rotate"]["frac_dmma"]), "ortho %.1e" % r["orthonormality_error"])
    except Exception as e: print(n, "unreadable", e)
PY
for m in 0 1; do
  HXB200_GRAM_TILE64=$m timeout -k 5 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gram_kernel -c 3 --csv --log-file $O/dram_gram_tile64_$m.csv \
     python tools/subspace_bench.py --widths 1024 --reps 1 --out $O/ncu_sub_$m.json > $O/ncu_gram_$m.log 2>&1; lap "ncu gram TILE64=$m rc=$?"
  grep -v "^==" $O/dram_gram_tile64_$m.csv | awk -F'","' '{print $5, $(NF-3), $(NF-2), $(NF-1), $NF}' | tail -12
done
