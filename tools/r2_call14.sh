#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2c14; mkdir -p $O
T0=$(date +%s)
lap() { echo "[r2c14] $1 at $(( $(date +%s) - T0 )) s"; }
free -g | head -2; nproc
timeout -k 5 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log; lap pytest
timeout -k 5 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; lap "bench default rc=$?"
tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c14/bench_default.json").read().strip().splitlines()[-1])
print("value %.2f ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["e2e"].items() if k not in ("call","single_block")}, "single", d["e2e"]["single_block"]["value"])
print("roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","unit","frac","frac_hbm","frac_tensor","kernel_ms_per_launch","kernel_sm_clock_mhz")})
print("c3", json.dumps(d.get("c3_strong"))[:900])
print("cpu", d.get("cpu_baseline",{}).get("value"))
PY
