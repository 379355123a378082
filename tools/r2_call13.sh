#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c13; mkdir -p $O
P="2:16,2:128,3:16,3:32,3:128,4:16,4:128,5:256,6:32,6:128,7:32,8:32,8:128,4:32:16:4"
timeout -k 5 400 python tools/sweep.py --points $P --out $O/sweep_new.json > $O/sweep_new.log 2>&1; echo "new rc=$?"
HXB200_CELL_KERNEL=v1 timeout -k 5 400 python tools/sweep.py --points $P --out $O/sweep_v1.json > $O/sweep_v1.log 2>&1; echo "v1 rc=$?"
python - <<'PY'
import json
a=json.load(open("gpurun_out/r2c13/sweep_new.json"))["points"]; b=json.load(open("gpurun_out/r2c13/sweep_v1.json"))["points"]
for x,y in zip(a,b):
    if "error" in x or "error" in y: print(x.get("p"),x.get("B"),x.get("error"),y.get("error")); continue
    print("p=%d B=%4d enr=%2d  new: cell %.4f ms frac %.2f filt %.4f | v1: cell %.4f frac %.2f filt %.4f | ratio cell %.2f filt %.2f ok %s %s" % (x["p"],x["B"],x["n_enr_per_cell"],x["cell_kernel_ms"],x["roofline_frac"],x["filter_ms_per_degree"],y["cell_kernel_ms"],y["roofline_frac"],y["filter_ms_per_degree"],y["cell_kernel_ms"]/x["cell_kernel_ms"],y["filter_ms_per_degree"]/x["filter_ms_per_degree"],x["ok"],y["ok"]))
PY
