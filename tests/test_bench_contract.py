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
    assert "workload" in d["config"] and "2 filter call" in d["reference_run"]["sample"]
    # both arms print the same `config` object (it depends on the workload and N only)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.shared_config("small", 1)
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
        assert "oracle/_ref" in d["reference_run"]["sample"] and "CELL_BATCH_SIZE=1" in d["reference_run"]["sample"]
        assert "the whole workload" in d["reference_run"]["sample"]  # the same mesh the GPU arm times, cut into one partition per core


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


def test_reference_arm_threads_reproduce_the_serial_result_and_the_port():
    """The reference-compiled filter (oracle/_ref) driven from a thread pool - the stand-in for `mpirun -n cores` - gives
    on every thread the bits of a serial run and of the oracle port (same dgemm): nothing in the reference's compiled
    routines shares hidden state, and the arm measures the same arithmetic the GPU path is checked against."""
    import copy
    import ctypes as ct
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    import pytest

    sys.path.insert(0, ROOT)
    import bench
    from dft_efe_b200 import synth
    from oracle import oracle as orc, ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    orc.use_scipy_dgemm(True)
    L = ref.lib()
    L.ref_set_dgemm.argtypes = [ct.c_void_p]
    L.ref_set_dgemm(bench._scipy_dgemm_pointer())
    try:
        base = synth.build_problem(synth.MeshSpec(ncell=(4, 4, 4), p=3, h=0.8, atoms=np.array([[1.5, 1.6, 1.7]]),
                                                  n_enr_per_atom=2, enr_cutoff=1.3, n_proj_per_atom=2, proj_cutoff=1.1,
                                                  nranks=1, boundary="dirichlet"))[0]
        B, deg, T = 8, 4, 4
        a0, a_, b_ = bench.FILTER_BOUNDS
        probs = []
        for _ in range(T):
            q = copy.copy(base)
            q.h_cell = base.h_cell.copy()
            probs.append(q)
        worlds = [orc.OracleWorld([q]) for q in probs]
        X0 = synth.make_block(base, B)

        def make_cb(q, W):
            def cb(_u, op_id, xp, yp, n_, B_, ugx, ugy):
                X = np.ctypeslib.as_array(xp, shape=(n_, B_))
                Y = np.ctypeslib.as_array(yp, shape=(n_, B_))
                if op_id == 0:
                    ref.hx_apply_serial(q, X, cell_block=1, out=Y)
                else:
                    W.minv_apply([X], [Y], bool(ugx), bool(ugy))
            return ref.APPLY_CB(cb)

        cbs = [make_cb(q, W) for q, W in zip(probs, worlds)]

        def one(i):
            x, y = X0.copy(), np.zeros_like(X0)
            L.ref_chebyshev_filter(cbs[i], None, orc._f64(x), orc._f64(y), ct.c_uint32(base.n_local), ct.c_uint32(B),
                                   ct.c_uint32(deg), ct.c_double(a0), ct.c_double(a_), ct.c_double(b_))
            return y

        serial = one(0)
        with ThreadPoolExecutor(T) as pool:
            for _ in range(2):
                assert all(np.array_equal(y, serial) for y in pool.map(one, range(T)))
        F = worlds[0].chebyshev_filter([X0.copy()], deg, a0, a_, b_)[0]
        assert np.abs(F - serial).max() <= 1e-13 * np.abs(F).max()
    finally:
        L.ref_set_dgemm(None)
        orc.use_scipy_dgemm(False)
