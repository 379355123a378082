"""CPU tests of bench.py's output contract: the reference arm (`--impl reference`, the oracle port on the host cores)
prints ONE JSON line with the keys the driver reads, runs exactly the requested number of steps, and the GPU arm
refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def _check_reference_line(r, kind):
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hx_throughput_fp64" and d["unit"] == "GDoF*vec/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "2 filter call" in d["config"]["sample"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    return d


def test_reference_arm_prints_one_contract_line():
    """With oracle/_ref built (the reference's own sources compiled) the arm times THAT code: kind "reference"."""
    from oracle import ref
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "small")
    d = _check_reference_line(r, "reference" if ref.available() else "port")
    if ref.available():
        assert "oracle/_ref" in d["config"]["sample"] and "CELL_BATCH_SIZE=1" in d["config"]["sample"]


def test_reference_arm_falls_back_to_the_oracle_port():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "small",
                  env={"HXB200_BENCH_REFERENCE_PORT": "1"})
    _check_reference_line(r, "port")


def test_reference_arm_nonzero_ranks_exit_without_work():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = run_bench("--steps", "1", "--warmup", "1", "--workload", "small")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
