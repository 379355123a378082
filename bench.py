#!/usr/bin/env python
"""bench.py — H.X throughput (GDoF.vec/s, FP64) inside the Chebyshev filter + filter seconds per SCF iteration on
N B200s, beside the CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|small]

A "step" is one ChebyshevFilter call (reference src/linearAlgebra/ChebyshevFilter.t.cpp:39-134) of degree 24 over
one block of B wavefunctions: 24 x [KohnShamOperatorContextFE::apply (updateGhostX=true) + M^-1 apply + the
recurrence] - the unit of work one SCF iteration repeats per wavefunction block.  N=1 workload = BASELINE.json
configs[1]: CH4-like pseudopotential OrthoEFE, FE order 4, 25^3 cells (1.03 M DoFs), 5 atoms x 4 enrichment
functions, nonlocal projectors, B = 32.  For N > 1 the mesh is N times longer in z and cut into N slabs (one per
GPU, "weak" scaling, per-GPU work fixed); the only data-path communication is the halo exchange.

Prints ONE JSON line (rank 0).  `value` = degree*N_global*B / t_step (H.X applications per second inside the filter,
inputs resident in HBM); `e2e` = the same metric through hx_chebyshev_filter_host (pinned HOST buffers: H2D of the
block, the filter, D2H of the filtered block inside the timed region); `hx_apply` = the bare operator apply;
`roofline` is for the dominant kernel (the fused gather->DMMA->scatter cell kernel), timed with CUDA events on its
own stream inside the timed region; `cpu_baseline` = the oracle port timed on this box's host cores on a bounded
sample of the same step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hx_throughput_fp64"
UNIT = "GDoF*vec/s"


def shared_config(workload: str, nranks: int):
    """the `config` object both arms print (identical by construction: it depends on the workload and N only)"""
    spec, B = workload_spec(workload, nranks)
    n_c = (spec.p + 1) ** 3
    cells = int(np.prod(spec.ncell)) if spec.refine_mask is None else None
    return {"workload": f"{workload}: ChebyshevFilter degree {DEGREE} (= {DEGREE} H.X applies + M^-1 + recurrence per step) on a "
                        f"synthetic pseudopotential OrthoEFE problem, FE order {spec.p}, {spec.ncell[0]}x{spec.ncell[1]}x{spec.ncell[2]} "
                        f"coarse cells{'' if spec.refine_mask is None else ' refined 2:1 around the atoms'}, B={B}, "
                        f"{0 if spec.atoms is None else len(spec.atoms)} atoms x {spec.n_enr_per_atom} enrichment fns + "
                        f"{spec.n_proj_per_atom} projectors, one z-slab per GPU",
            "block": B, "degree": DEGREE, "n_gpus": nranks,
            "l2_policy": "GPU arm: inputs larger than L2 (the cell matrices alone are %s per GPU, read once per apply)"
                         % ("%.2f GB" % (8.0 * n_c * n_c * cells / nranks / 1e9) if cells else "> 1 GB")}


def workload_spec(name: str, nranks: int):
    from dft_efe_b200 import synth
    if name == "c2":
        nc, p, B, h = (25, 25, 25 * nranks), 4, 32, 0.8
        n_atoms = 5 * nranks
    elif name == "small":
        nc, p, B, h = (8, 8, 8 * nranks), 4, 32, 0.8
        n_atoms = 2 * nranks
    elif name == "c3":
        # BASELINE configs[2]: benzene-dimer-like, FE order 6, ~7 M DoFs, 128-vector block; the mesh is FIXED and cut
        # into nranks z-slabs (strong scaling; 31 GB of cell matrices in total, meant for 4 or 8 GPUs)
        nc, p, B, h = (32, 32, 32), 6, 128, 0.8
        n_atoms = 24
    elif name == "c2a":
        # C2 on an ADAPTIVE mesh (dft-efe meshes are refined around the atoms): 24^3 cells of which the ones within 2.5 h
        # of an atom are split 2:1 -> ~15.9 k cells, ~1.0 M DoFs, ~30 k hanging-node rows, parents with a few hundred
        # children.  Exercises the constraint kernels and the row-list recurrence at the size of C2.
        nc, p, B, h = (24, 24, 24 * nranks), 4, 32, 0.8
        rng = np.random.default_rng(7)
        L = np.array(nc) * h
        atoms = (0.25 + 0.5 * rng.uniform(size=(5 * nranks, 3))) * L[None, :]
        spec = synth.MeshSpec(ncell=nc, p=p, h=h, refine_mask=synth.refine_ball(nc, h, atoms, 2.5 * h), atoms=atoms,
                              n_enr_per_atom=4, enr_cutoff=1.6 * h, n_proj_per_atom=4, proj_cutoff=1.3 * h, nranks=nranks,
                              boundary="dirichlet")
        return spec, B
    elif name == "c1":
        # BASELINE configs[0]: H2 all-electron classical EFE on a small adaptive mesh, block of 8 wavefunctions (the
        # reference's own CPU-runnable test/ksdft case: 15^3 cells refined around the two nuclei, FE order 3, one
        # enrichment function per atom, no pseudopotential)
        nc, p, B, h = (15, 15, 15 * nranks), 3, 8, 1.0
        L = np.array(nc) * h
        atoms = np.concatenate([np.array([[0.5 * L[0] - 0.7, 0.5 * L[1], (k + 0.5) * 15.0 * h],
                                          [0.5 * L[0] + 0.7, 0.5 * L[1], (k + 0.5) * 15.0 * h]]) for k in range(nranks)])
        spec = synth.MeshSpec(ncell=nc, p=p, h=h, refine_mask=synth.refine_ball(nc, h, atoms, 2.5), atoms=atoms,
                              n_enr_per_atom=1, enr_cutoff=3.0 * h, n_proj_per_atom=0, nranks=nranks, boundary="dirichlet")
        return spec, B
    else:
        raise SystemExit(f"unknown workload {name}")
    rng = np.random.default_rng(7)
    L = np.array(nc) * h
    atoms = (0.25 + 0.5 * rng.uniform(size=(n_atoms, 3))) * L[None, :]
    spec = synth.MeshSpec(ncell=nc, p=p, h=h, atoms=atoms, n_enr_per_atom=4, enr_cutoff=1.6 * h,
                          n_proj_per_atom=4, proj_cutoff=1.3 * h, nranks=nranks, boundary="dirichlet")
    return spec, B


class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i):
                reasons.append(name)
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU leg ----
DEGREE = 24                       # CHEBY_ORDER_LOOKUP bucket for a <= 500 Ha spectral bound (src/ksdft/Defaults.cpp:51-58)
FILTER_BOUNDS = (-3.0, 1.0, 400.0)  # wantedLower, wantedUpper, unwantedUpper of the synthetic spectrum


def cpu_filter_throughput(sample_cells=(10, 10, 10), p=4, B=32, threads=1, seconds=12.0, warm=1, max_steps=50,
                          degree=DEGREE, workload=None):
    """Oracle port (reference-faithful: ChebyshevFilter = per degree one H.X [gather, one dgemm per cell through an
    optimised BLAS, sequential scatter, BLAS-1 constraints], one mass-lumped M^-1 apply, two axpby) on a bounded
    sample of the same cell shape; `threads` partitions run concurrently (the stand-in for `mpirun -n threads`)."""
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from concurrent.futures import ThreadPoolExecutor
    from dft_efe_b200 import synth
    from oracle import oracle as orc
    blas = orc.use_scipy_dgemm(True)
    if workload == "c1":
        # the reference's own CPU-runnable case is small enough to be timed whole: one full C1 mesh per thread
        spec, B = workload_spec("c1", threads)
        nc, p = spec.ncell, spec.p
    else:
        nc = (sample_cells[0], sample_cells[1], sample_cells[2] * threads)
        rng = np.random.default_rng(7)
        L = np.array(nc) * 0.8
        atoms = (0.25 + 0.5 * rng.uniform(size=(2 * threads, 3))) * L[None, :]
        spec = synth.MeshSpec(ncell=nc, p=p, h=0.8, atoms=atoms, n_enr_per_atom=4, enr_cutoff=1.28,
                              n_proj_per_atom=4, proj_cutoff=1.04, nranks=threads, boundary="dirichlet")
    probs = synth.build_problem(spec)
    W = orc.OracleWorld(probs)
    if threads > 1:
        W.pool = ThreadPoolExecutor(max_workers=threads)
    X0 = [synth.make_block(q, B) for q in probs]
    N = sum(q.n_owned for q in probs)

    def step():
        W.chebyshev_filter([x.copy() for x in X0], degree, *FILTER_BOUNDS)

    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > seconds or n >= max_steps:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": degree * N * B / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"ChebyshevFilter degree {degree} on {nc[0]}x{nc[1]}x{nc[2]} cells order {p} ({N} DoFs) B={B}, "
                      f"{threads} partition(s) on {threads} thread(s), {n} filter calls, per-cell dgemm via "
                      f"{'SciPy OpenBLAS' if blas else 'built-in loops'}",
            "ms_per_step": dt * 1e3}


def _scipy_dgemm_pointer():
    import ctypes as ct
    import scipy.linalg.cython_blas as cb
    cap = cb.__pyx_capi__["dgemm"]
    ct.pythonapi.PyCapsule_GetName.restype = ct.c_char_p
    ct.pythonapi.PyCapsule_GetName.argtypes = [ct.py_object]
    ct.pythonapi.PyCapsule_GetPointer.restype = ct.c_void_p
    ct.pythonapi.PyCapsule_GetPointer.argtypes = [ct.py_object, ct.c_char_p]
    return ct.pythonapi.PyCapsule_GetPointer(cap, ct.pythonapi.PyCapsule_GetName(cap))


def cpu_filter_throughput_ref(sample_cells=(8, 8, 8), p=4, B=32, threads=1, warm=1, steps=1, degree=DEGREE, workload=None):
    """The reference's OWN code on the host cores (oracle/_ref = its sources compiled where they lie): its ChebyshevFilter
    template (linearAlgebra/ChebyshevFilter.t.cpp:39-134) drives KohnShamOperatorContextFE::apply assembled from its
    compiled gather / gemmStridedVarBatched / scaleStridedVarBatched / scatter-add / constraint routines with the
    reference's default CELL_BATCH_SIZE = 1 (ksdft/Defaults.cpp:84; ref_hx_apply_serial) - the one piece that cannot be
    compiled here, the mass-lumped M^-1 apply (deal.II-dependent class), is the oracle port.  dgemm_ is SciPy's OpenBLAS,
    the stand-in for the MKL/BLIS a site links.  One independent single-rank problem per thread (no halo traffic at all:
    this favours the reference over an MPI run of the same cells)."""
    import copy
    import ctypes as ct
    from concurrent.futures import ThreadPoolExecutor
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from dft_efe_b200 import synth
    from oracle import oracle as orc, ref
    assert ref.available(), "oracle/_ref/libdftefe_ref.so not built"
    orc.use_scipy_dgemm(True)
    L = ref.lib()
    L.ref_set_dgemm.argtypes = [ct.c_void_p]
    L.ref_set_dgemm(_scipy_dgemm_pointer())
    if workload == "c1":
        spec, B = workload_spec("c1", 1)
        nc, p = spec.ncell, spec.p
    else:
        nc = tuple(sample_cells)
        rng = np.random.default_rng(7)
        Lbox = np.array(nc) * 0.8
        atoms = (0.25 + 0.5 * rng.uniform(size=(2, 3))) * Lbox[None, :]
        spec = synth.MeshSpec(ncell=nc, p=p, h=0.8, atoms=atoms, n_enr_per_atom=4, enr_cutoff=1.28,
                              n_proj_per_atom=4, proj_cutoff=1.04, nranks=1, boundary="dirichlet")
    base = synth.build_problem(spec)[0]
    probs = []
    for _ in range(threads):  # same cells, private copies of the big arrays (no cache sharing between the "ranks")
        q = copy.copy(base)
        q.h_cell = base.h_cell.copy()
        if getattr(base, "cell_c", None) is not None:
            q.cell_c = base.cell_c.copy()
        probs.append(q)
    worlds = [orc.OracleWorld([q]) for q in probs]
    X0 = synth.make_block(base, B)
    a0, a_, b_ = FILTER_BOUNDS

    def make_cb(q, W):
        def cb(_user, op_id, xp, yp, n_, B_, ugx, ugy):
            X = np.ctypeslib.as_array(xp, shape=(n_, B_))
            Y = np.ctypeslib.as_array(yp, shape=(n_, B_))
            if op_id == 0:
                ref.hx_apply_serial(q, X, cell_block=1, out=Y)  # the C call releases the GIL; no Python-side copy
            else:
                W.minv_apply([X], [Y], bool(ugx), bool(ugy))
        return ref.APPLY_CB(cb)

    cbs = [make_cb(q, W) for q, W in zip(probs, worlds)]

    def one(i):
        x, y = X0.copy(), np.zeros_like(X0)
        L.ref_chebyshev_filter(cbs[i], None, orc._f64(x), orc._f64(y), ct.c_uint32(base.n_local), ct.c_uint32(B),
                               ct.c_uint32(degree), ct.c_double(a0), ct.c_double(a_), ct.c_double(b_))
        return float(np.abs(y).max())

    pool = ThreadPoolExecutor(max_workers=threads)

    def step():
        return list(pool.map(one, range(threads)))

    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        chk = step()
    dt = (time.perf_counter() - t0) / steps
    assert all(np.isfinite(chk)), "reference filter produced non-finite values"
    N = base.n_owned * threads
    return {"value": degree * N * B / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"ChebyshevFilter degree {degree} through oracle/_ref (the reference's compiled ChebyshevFilter template + "
                      f"H.X apply assembled from its compiled gather / gemmStridedVarBatched / scatter / constraint routines, "
                      f"CELL_BATCH_SIZE=1; M^-1 apply = oracle port) on {threads} independent single-rank problem(s) of "
                      f"{nc[0]}x{nc[1]}x{nc[2]} cells order {p} ({base.n_owned} DoFs each) B={B}, one per thread, {steps} "
                      f"filter call(s) per problem, dgemm_ = SciPy OpenBLAS",
            "ms_per_step": dt * 1e3}


def cpu_filter_throughput_ref_partitioned(workload, threads, warm=1, steps=1, degree=DEGREE):
    """The reference's own code on the host cores over the SAME problem the GPU arm times: the workload's mesh cut into one
    partition per core (the stand-in for `mpirun -n cores`), KohnShamOperatorContextFE::apply on every partition through the
    reference's compiled routines (oracle/_ref: ref_hx_phase_a / ref_hx_phase_b = its gather, gemmStridedVarBatched with
    CELL_BATCH_SIZE = 1, scaleStridedVarBatched, scatter-add, constraint distribute), halo exchanges / projector all-reduce /
    M^-1 / the recurrence's axpbys through the oracle port's C routines (in-process copies instead of MPI messages, which
    favours the reference).  Rank-local phases run on a thread pool; the C routines release the GIL.  dgemm_ = SciPy OpenBLAS."""
    import ctypes as ct
    from concurrent.futures import ThreadPoolExecutor
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from dft_efe_b200 import synth
    from oracle import oracle as orc, ref
    assert ref.available(), "oracle/_ref/libdftefe_ref.so not built"
    orc.use_scipy_dgemm(True)
    L = ref.lib()
    L.ref_set_dgemm.argtypes = [ct.c_void_p]
    L.ref_set_dgemm(_scipy_dgemm_pointer())
    spec, B = workload_spec(workload, 1)
    spec.nranks = threads
    spec.with_k_cell = False
    probs = synth.build_problem(spec)
    W = orc.OracleWorld(probs)
    if threads > 1:
        W.pool = ThreadPoolExecutor(max_workers=threads)
    PA = ref.PartitionedApply(W, cell_block=1)
    W.hx_apply = lambda Xs, Ys, gx=False, gy=False, **kw: PA(Xs, Ys, gx, gy)
    X0 = [synth.make_block(q, B) for q in probs]
    N = sum(q.n_owned for q in probs)

    def step():
        return W.chebyshev_filter([x.copy() for x in X0], degree, *FILTER_BOUNDS)

    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        F = step()
    dt = (time.perf_counter() - t0) / steps
    assert all(np.isfinite(f).all() for f in F), "reference filter produced non-finite values"
    nc = spec.ncell
    return {"value": degree * N * B / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"the whole workload: ChebyshevFilter degree {degree} on {nc[0]}x{nc[1]}x{nc[2]} cells order {spec.p} ({N} DoFs) "
                      f"B={B} cut into {threads} partitions, one per host core; H.X through oracle/_ref (the reference's compiled "
                      f"gather / gemmStridedVarBatched (CELL_BATCH_SIZE=1) / scatter-add / constraint routines), halo exchange + "
                      f"M^-1 + axpby through the oracle port, {steps} filter call(s), dgemm_ = SciPy OpenBLAS",
            "ms_per_step": dt * 1e3, "global_dofs": N}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    spec, B = workload_spec(args.workload, 1)
    # bounded sample: each step = one filter call over `cores` problems / partitions of 8^3 cells; W warm-up and exactly
    # K timed steps, as on the GPU arm (a step is 1.5-4 s on 8-64 cores)
    res = None
    try:
        from oracle import ref as _ref
        if _ref.available() and not os.environ.get("HXB200_BENCH_REFERENCE_PORT"):
            if args.workload in ("c2", "small", "c2a", "c1") and not os.environ.get("HXB200_BENCH_REFERENCE_SAMPLE"):
                res = cpu_filter_throughput_ref_partitioned(args.workload, cores, warm=max(0, min(args.warmup, 2)),
                                                            steps=max(1, min(args.steps, 8)))
            else:
                res = cpu_filter_throughput_ref(sample_cells=(8, 8, 8), p=spec.p, B=B, threads=cores, warm=max(0, args.warmup),
                                                steps=max(1, args.steps), workload=args.workload)
    except Exception as e:  # noqa: BLE001 - the reference-compiled arm must not take the baseline down with it
        sys.stderr.write(f"[bench] oracle/_ref arm failed ({e}); timing the oracle port instead\n")
        res = None
    if res is None:
        res = cpu_filter_throughput(sample_cells=(8, 8, 8), p=spec.p, B=B, threads=cores, seconds=1e9,
                                    warm=max(0, args.warmup), max_steps=max(1, args.steps), workload=args.workload)
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": shared_config(args.workload, args.gpus),
            "reference_run": {"sample": res["sample"], "timed_steps": max(1, min(args.steps, 8)),
                              "note": "steps/warmup echo the command line; the CPU arm times at most 8 filter calls of the whole "
                                      "workload (a call takes seconds on the host cores)"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def multi_gpu_parity(torch, dist, capi, synth, rank, world):
    """N > 1 only, before anything is timed: the checks of tests/mgpu_worker.py on a small mesh (order 4, enrichment,
    projectors, 3 cell layers per rank) through the same plan / communicator machinery the timed run uses - H.X, the fused
    Chebyshev filter and X^T H X on every rank against the CPU oracle of the whole rank set, and the halo exchange
    overlapped with the cell kernel against the serial exchange (bitwise).  The oracle is the checker here, never the
    thing measured."""
    from oracle import oracle as orc
    nc = (4, 4, 3 * world)
    L = np.array(nc, float)
    atoms = np.array([[0.5 * L[0], 0.5 * L[1], 0.5 * L[2]], [0.3 * L[0], 0.7 * L[1], 0.26 * L[2]]])
    spec = synth.MeshSpec(ncell=nc, p=4, atoms=atoms, n_enr_per_atom=2, enr_cutoff=1.2, n_proj_per_atom=2, proj_cutoff=1.0,
                          nranks=world)
    probs = synth.build_problem(spec)
    q = probs[rank]
    B = 16
    plan = capi.Plan(q, max_block=B)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    plan.attach_comm(bytes(uid.cpu().numpy().tobytes()))
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, q.diag_inv, q.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    W = orc.OracleWorld(probs)
    Xs = [synth.make_block(p_, B) for p_ in probs]
    for p_, x in zip(probs, Xs):
        x[p_.n_owned:] = 7.0  # stale ghosts: the update must fix them

    def rel(a, b):
        den = np.linalg.norm(b, axis=0)
        den[den == 0] = 1.0
        return float((np.linalg.norm(a - b, axis=0) / den).max())

    out = {}
    dX, dY = plan.block(B, Xs[rank]), plan.block(B)
    H.apply(dX, dY, True, True)
    Yo = [np.zeros_like(x) for x in Xs]
    W.hx_apply([x.copy() for x in Xs], Yo, True, True)
    out["hx"] = rel(dY.download(), Yo[rank])
    os.environ["HXB200_HALO_OVERLAP"] = "0"
    dXs, dYs = plan.block(B, Xs[rank]), plan.block(B)
    H.apply(dXs, dYs, True, True)
    os.environ.pop("HXB200_HALO_OVERLAP")
    out["overlap_vs_serial_max_abs_diff"] = float(np.abs(dY.download() - dYs.download()).max())
    dX, dF = plan.block(B, Xs[rank]), plan.block(B)
    capi.chebyshev_filter(H, minv, dX, dF, 6, -3.0, 1.0, 60.0)
    F = W.chebyshev_filter([x.copy() for x in Xs], 6, -3.0, 1.0, 60.0)
    out["cheb"] = rel(dF.download()[:q.n_owned], F[rank][:q.n_owned])
    dX = plan.block(B, Xs[rank])
    S = H.xtopx(dX, 8)
    So = W.xtopx([x.copy() for x in Xs], lambda a, b, c, d_: W.hx_apply(a, b, c, d_), 8)
    out["xtopx"] = float(np.abs(S - So).max() / np.abs(So).max())
    plan.synchronize()
    transport = plan.halo_transport()
    t = torch.tensor([out["hx"], out["cheb"], out["xtopx"], out["overlap_vs_serial_max_abs_diff"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    for o in (H, minv):
        o.destroy()
    plan.destroy()
    dist.barrier()
    v = [float(x) for x in t.cpu()]
    tol = {"hx": 1e-12, "cheb": 1e-11, "xtopx": 1e-12}
    return {"hx": v[0], "cheb": v[1], "xtopx": v[2], "overlap_vs_serial_max_abs_diff": v[3], "tolerance": tol,
            "ok": bool(v[0] <= tol["hx"] and v[1] <= tol["cheb"] and v[2] <= tol["xtopx"] and v[3] == 0.0),
            "what": f"max over {world} ranks, order-4 mesh {nc[0]}x{nc[1]}x{nc[2]} with enrichment and projectors, B={B}: H.X and "
                    f"fused Chebyshev filter rel. L2 per vector, X^T H X max rel., against the CPU oracle of the whole rank set; "
                    f"overlapped vs serial halo exchange bitwise", "halo_transport": transport}


def c3_strong_block(torch, dist, capi, synth, rank, nranks, local_rank, micro, steps=2):
    """BASELINE configs[2] on N GPUs: the FIXED benzene-dimer-like mesh (order 6, 32^3 cells, 7.1 M DoFs, B = 128) cut
    into N z-slabs - strong scaling; its cell kernel is bound by the FP64 tensor pipe (arithmetic intensity ~ 30 flop/B)."""
    import psutil
    spec, B = workload_spec("c3", nranks)
    spec.with_k_cell = False
    need = 8.0 * 343 * 343 * (32 ** 3) / nranks   # per rank; all ranks of the node generate theirs at the same time
    ok_mem = psutil.virtual_memory().available > 2.3 * need * nranks
    flag = torch.tensor([1 if ok_mem else 0], device="cuda")
    if nranks > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not int(flag.item()):
        return {"skipped": "not enough host memory to generate %.1f GB of synthetic cell matrices per rank" % (need / 1e9)}
    t0 = time.perf_counter()
    prob = synth.build_problem(spec, only_rank=rank)[0]
    t_build = time.perf_counter() - t0
    stream = torch.cuda.Stream()
    plan = capi.Plan(prob, max_block=B, stream=stream.cuda_stream)
    if nranks > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        plan.attach_comm(bytes(uid.cpu().numpy().tobytes()))
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, prob.diag_inv, prob.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(prob, B)
    N_global = prob.n_owned
    S2 = prob.S2 + (int(np.sum(prob.num_cell_proj.astype(np.int64) * prob.num_cell_dofs.astype(np.int64)))
                    if prob.num_cell_proj is not None else 0)
    del prob.h_cell
    if nranks > 1:
        t = torch.tensor([N_global], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        N_global = int(t.item())

    class Blk:
        def __init__(self, host=None):
            self.B = B
            self.t = torch.zeros(X.shape[0] * B, dtype=torch.float64, device="cuda") if host is None else \
                torch.from_numpy(np.ascontiguousarray(host)).reshape(-1).cuda()
            self.p = C.cast(self.t.data_ptr(), capi.f64p)

    a0, a_, b_ = FILTER_BOUNDS
    with torch.cuda.stream(stream):
        dX0, dX, dF = Blk(X), Blk(X), Blk()

        def step():
            dX.t.copy_(dX0.t, non_blocking=True)
            capi.chebyshev_filter(H, minv, dX, dF, DEGREE, a0, a_, b_)

        step()
        plan.synchronize()
        if nranks > 1:
            dist.barrier()
        torch.cuda.synchronize()
        plan.enable_kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        plan.synchronize()
        if nranks > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1) / steps
        cell_ms, nl = plan.cell_kernel_time_ms()
        clk = plan.cell_kernel_sm_clock_mhz()
        plan.enable_kernel_timing(False)
        phases = {}
        try:
            plan.trace(True)
            step()
            rep = plan.trace_report()
            plan.trace(False)
            phases = {k: round(v["ms"] / DEGREE, 5) for k, v in rep.items()}
        except Exception as e:  # noqa: BLE001
            phases = {"error": str(e)[:200]}
        finite = bool(torch.isfinite(dF.t).all())
    tm = torch.tensor([ms, cell_ms / max(nl, 1)], dtype=torch.float64, device="cuda")
    if nranks > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, cell_ms1 = [float(v) for v in tm.cpu()]
    flops = 2.0 * B * S2
    dmma = micro["dmma_tflops"] if micro else None
    res = {"workload": "c3: ChebyshevFilter degree %d, order 6, 32x32x32 cells, %d DoFs, B=%d, fixed mesh cut into %d z-slab(s)"
                       % (DEGREE, N_global, B, nranks),
           "scaling": "strong", "n_gpus": nranks, "ms_per_step": ms, "ms_per_degree": ms / DEGREE,
           "value": DEGREE * N_global * B / (ms * 1e-3) / 1e9, "unit": UNIT, "steps": steps, "result_finite": finite,
           "cell_kernel_ms_per_launch": cell_ms1, "cell_kernel_sm_clock_mhz": round(clk, 1),
           "cell_kernel_tflops_per_gpu": flops / (cell_ms1 * 1e-3) / 1e12,
           "roofline": {"bound": "tensor", "achieved": flops / (cell_ms1 * 1e-3) / 1e12, "peak": dmma, "unit": "TFLOP/s",
                        "frac": (flops / (cell_ms1 * 1e-3) / 1e12 / dmma) if dmma else None,
                        "peak_source": "hx_microbench: mma.sync.m8n8k4.f64 (DMMA.8x8x4) issue rate measured on this GPU in this run"},
           "phase_ms_per_degree": phases, "halo_transport": plan.halo_transport(), "host_build_s": round(t_build, 1),
           "efficiency_note": "strong-scaling efficiency = ms_per_step(N=1) / (N * ms_per_step(N)) over the c3_strong blocks of "
                              "the per-N runs"}
    for o in (H, minv):
        o.destroy()
    del dX0, dX, dF
    if nranks > 1:
        dist.barrier()
    plan.destroy()
    torch.cuda.empty_cache()
    return res


def c4_block(torch, dist, capi, synth, rank, nranks, micro, cells=78, B_total=1024, batch=256):
    """BASELINE configs[3]: synthetic PERIODIC cluster, FE order 5, cells^3 cells (78^3: 59.3 M DoFs), a block of B_total
    wavefunctions resident on the GPUs and filtered `batch` columns at a time (ChebyshevFilteredEigenSolver's column batches),
    then the subspace steps over the WHOLE block across the ranks: X^T M X (Gram, NCCL all-reduce) + Cholesky + rotation,
    X^T H X + symmetric eigensolve + rotation.  The cell matrices (22 GB per GPU at 8 GPUs) are assembled on the device from
    the basis values and a synthetic local potential (FEBasisOperations::computeFEMatrices straight into the kernel's packed
    stream), never on the host.  The subspace steps are timed on the UNFILTERED block (a degree-24 filter of random vectors
    is too ill-conditioned for one Cholesky pass; their cost does not depend on the values)."""
    t0 = time.perf_counter()
    spec = synth.MeshSpec(ncell=(cells, cells, cells), p=5, h=0.8, boundary="periodic", nranks=nranks, with_k_cell=False,
                          with_h_cell=False)
    prob = synth.build_problem(spec, only_rank=rank)[0]
    t_build = time.perf_counter() - t0
    stream = torch.cuda.Stream()
    plan = capi.Plan(prob, max_block=B_total, stream=stream.cuda_stream)
    if nranks > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        plan.attach_comm(bytes(uid.cpu().numpy().tobytes()))
    t0 = time.perf_counter()
    H = capi.CellOp(plan, with_nonlocal=False, matrices=False)
    fe = synth.fe_basis_data(prob)
    nq = int(fe["num_cell_quad"][0])
    feb = capi.FeBasis(plan, fe["num_cell_quad"], fe["basis"], fe["jxw"], fe["same_basis"])
    vq = 1.05 + 0.05 * (synth.potential_at_quad_points(prob, nq) + 0.8) / 0.35   # in [1.0, 1.1]: a positive definite operator
    feb.assemble_into(H, vq)
    plan.synchronize()
    t_asm = time.perf_counter() - t0
    minv = capi.DiagOp(plan, prob.diag_inv, None, capi.DIAG_CFE)
    Mop = capi.DiagOp(plan, prob.diag, None, capi.DIAG_OEFE_MASS)
    n_local, N_global = prob.n_local, prob.n_owned
    S2 = prob.S2
    if nranks > 1:
        t = torch.tensor([N_global], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        N_global = int(t.item())

    class Blk:
        def __init__(self, width, tensor=None):
            self.B = width
            self.t = torch.zeros(n_local * width, dtype=torch.float64, device="cuda") if tensor is None else tensor
            self.p = C.cast(self.t.data_ptr(), capi.f64p)

    a0, a_, b_ = 0.99, 1.04, 1.11
    out = {}
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda")
        g.manual_seed(1234 + rank)
        Xt = (torch.rand(n_local * B_total, dtype=torch.float64, device="cuda", generator=g) - 0.5)
        X2 = Xt.view(n_local, B_total)
        X = Blk(B_total, Xt)
        xin, xout = Blk(batch), Blk(batch)

        def filter_all():
            for j0 in range(0, B_total, batch):
                xin.t.view(n_local, batch).copy_(X2[:, j0:j0 + batch])
                capi.chebyshev_filter(H, minv, xin, xout, DEGREE, a0, a_, b_)

        # warm-up: one batch (peer transport set-up, scratch allocation)
        xin.t.view(n_local, batch).copy_(X2[:, :batch])
        capi.chebyshev_filter(H, minv, xin, xout, 2, a0, a_, b_)
        plan.synchronize()
        if nranks > 1:
            dist.barrier()
        torch.cuda.synchronize()
        plan.enable_kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        filter_all()
        e1.record(stream)
        plan.synchronize()
        if nranks > 1:
            dist.barrier()
        t_filter = e0.elapsed_time(e1)
        cell_ms, nl = plan.cell_kernel_time_ms()
        clk = plan.cell_kernel_sm_clock_mhz()
        plan.enable_kernel_timing(False)
        finite = bool(torch.isfinite(xout.t).all())
        phases = {}
        try:
            plan.trace(True)
            xin.t.view(n_local, batch).copy_(X2[:, :batch])
            capi.chebyshev_filter(H, minv, xin, xout, 4, a0, a_, b_)
            rep = plan.trace_report()
            plan.trace(False)
            phases = {k: round(v["ms"] / 4, 4) for k, v in rep.items()}
        except Exception as e:  # noqa: BLE001
            phases = {"error": str(e)[:200]}
        del xin, xout
        torch.cuda.empty_cache()
        # subspace steps over the whole block (in place)
        plan.synchronize()
        if nranks > 1:
            dist.barrier()
        plan.trace(True)
        t0 = time.perf_counter()
        st_o = capi.cholesky_gram_schmidt(Mop, X, X, batch)
        plan.synchronize()
        t_cgs = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        w, st_r = capi.rayleigh_ritz(H, X, X, batch, True)
        plan.synchronize()
        t_rr = (time.perf_counter() - t0) * 1e3
        rep = plan.trace_report()
        plan.trace(False)
        sub_phases = {k: round(v["ms"], 2) for k, v in rep.items()}
        # orthonormality of the first columns after CholGS + rotation: X^T M X = I
        chk = Blk(32, X2[:, :32].contiguous().view(-1))
        Sx = Mop.xtopx(chk, 32) if hasattr(Mop, "xtopx") else None
    tm = torch.tensor([t_filter, cell_ms / max(nl, 1), t_cgs, t_rr], dtype=torch.float64, device="cuda")
    if nranks > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    t_filter, cell1, t_cgs, t_rr = [float(v) for v in tm.cpu()]
    flops = 2.0 * batch * S2
    dmma = micro["dmma_tflops"] if micro else None
    res = {"workload": "c4: periodic FE order 5, %dx%dx%d cells, %d DoFs, block of %d wavefunctions in column batches of %d, %d GPU(s)"
                       % (cells, cells, cells, N_global, B_total, batch, nranks),
           "n_gpus": nranks, "filter_ms": t_filter, "filter_degree": DEGREE,
           "value": DEGREE * N_global * B_total / (t_filter * 1e-3) / 1e9, "unit": UNIT,
           "cell_kernel_ms_per_launch": cell1, "cell_kernel_sm_clock_mhz": round(clk, 1),
           "cell_kernel_tflops_per_gpu": flops / (cell1 * 1e-3) / 1e12,
           "roofline": {"bound": "tensor", "achieved": flops / (cell1 * 1e-3) / 1e12, "peak": dmma, "unit": "TFLOP/s",
                        "frac": (flops / (cell1 * 1e-3) / 1e12 / dmma) if dmma else None},
           "phase_ms_per_degree": phases, "result_finite": finite,
           "cholesky_gram_schmidt_ms": t_cgs, "cholesky_gram_schmidt_status": int(st_o),
           "rayleigh_ritz_ms": t_rr, "rayleigh_ritz_status": int(st_r), "subspace_phase_ms": sub_phases,
           "ritz_values_lowest": [float(v) for v in w[:4]], "ritz_values_ascending": bool(np.all(np.diff(w) >= -1e-9)),
           "orthonormality_max_dev_first_32": (float(np.abs((Sx + Sx.T - np.diag(np.diag(Sx))) - np.eye(32)).max()) if Sx is not None else None),
           "halo_transport": plan.halo_transport(), "host_build_s": round(t_build, 1), "device_assembly_s": round(t_asm, 1),
           "cell_matrices_gb_per_gpu": 8.0 * S2 / 1e9, "block_gb_per_gpu": 8.0 * n_local * B_total / 1e9}
    feb.destroy()
    for o in (H, minv, Mop):
        o.destroy()
    del X, Xt, X2
    if nranks > 1:
        dist.barrier()
    plan.destroy()
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------- GPU leg ----
def run_ours(args):
    # keep stdout for the ONE JSON line: libraries (NCCL's version banner) write to fd 1 during initialisation
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from dft_efe_b200 import capi, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    nranks = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    capi.check(capi.lib().hx_set_device(local_rank))
    affinity = None
    try:  # run (and first-touch the pinned host buffers) on the CPUs next to this rank's GPU
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(hnd)
        affinity = sorted(os.sched_getaffinity(0))
        affinity = f"{affinity[0]}-{affinity[-1]} ({len(affinity)} cpus)"
    except Exception as e:  # noqa: BLE001
        affinity = f"unchanged ({str(e)[:60]})"
    if nranks > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "c4":
        micro = capi.microbench()
        res = c4_block(torch, dist, capi, synth, rank, nranks, micro, cells=args.c4_cells, B_total=args.c4_block, batch=args.c4_batch)
        if rank == 0:
            line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": nranks, "steps": 1, "warmup": 1,
                    "ms_per_step": res["filter_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "f64", "data": "synthetic", "config": {"workload": res["workload"], "n_gpus": nranks}, "c4": res}
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(line), flush=True)
            os.dup2(2, 1)
        if nranks > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    parity = None
    if nranks > 1:
        try:
            parity = multi_gpu_parity(torch, dist, capi, synth, rank, nranks)
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": str(e)[:300]}
    spec, B = workload_spec(args.workload, nranks)
    if args.workload in ("c3",):
        spec.with_k_cell = False
    prob = synth.build_problem(spec, only_rank=rank)[0]
    stream = torch.cuda.Stream()
    plan = capi.Plan(prob, max_block=B, stream=stream.cuda_stream)
    if nranks > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        plan.attach_comm(bytes(uid.cpu().numpy().tobytes()))
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, prob.diag_inv, prob.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(prob, B)
    N_local = prob.n_owned
    N_global = N_local
    if nranks > 1:
        t = torch.tensor([N_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        N_global = int(t.item())

    class Block:  # a block vector in a torch allocation (torch = device memory + streams only)
        def __init__(self, host=None):
            self.B = B
            self.t = torch.zeros(prob.n_local * B, dtype=torch.float64, device="cuda") if host is None else \
                torch.from_numpy(np.ascontiguousarray(host)).reshape(-1).cuda()
            self.p = C.cast(self.t.data_ptr(), capi.f64p)

    def barrier():
        if nranks > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        plan.synchronize()
        barrier()
        return e0.elapsed_time(e1) / reps

    a0, a_, b_ = FILTER_BOUNDS
    warm = max(args.warmup, 3)
    with torch.cuda.stream(stream):
        dX0, dX, dY, dF = Block(X), Block(X), Block(), Block()
        torch.cuda.synchronize()

        # ---- bare operator apply (KohnShamOperatorContextFE::apply, updateGhostX = true) ----
        for _ in range(3):
            H.apply(dX, dY, True, False)
        apply_ms = timed(lambda: H.apply(dX, dY, True, False), 20)

        # ---- the step: one ChebyshevFilter call of degree DEGREE; the block is reset from a pristine device copy
        # first (one D2D copy per step, inside the timed region) so values stay finite over many steps ----
        def step():
            dX.t.copy_(dX0.t, non_blocking=True)
            capi.chebyshev_filter(H, minv, dX, dF, DEGREE, a0, a_, b_)

        for _ in range(warm):
            step()
        plan.synchronize()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        plan.enable_kernel_timing(True)
        l0 = plan.launch_count()
        ms_per_step = timed(step, args.steps)
        launches = plan.launch_count() - l0
        cell_ms, cell_launches = plan.cell_kernel_time_ms()
        cell_clock_mhz = plan.cell_kernel_sm_clock_mhz()
        plan.enable_kernel_timing(False)

        # ---- phase breakdown of one filter call (untimed pass with CUDA events at the phase boundaries) ----
        phases = {}
        try:
            plan.trace(True)
            step()
            rep = plan.trace_report()
            plan.trace(False)
            phases = {k: round(v["ms"] / DEGREE, 5) for k, v in rep.items()}
        except Exception as e:  # noqa: BLE001
            phases = {"error": str(e)[:200]}

        # ---- subspace projections: X^T H X (one column batch, Op.apply + Gram GEMM) and the rotation X <- X Q ----
        sub = {}
        try:
            if args.quick:
                raise RuntimeError("skipped (--quick)")
            dXs = Block(X)
            H.xtopx(dXs, B)
            sub["xtopx_ms"] = timed(lambda: H.xtopx(dXs, B), 3)
            Q = np.linalg.qr(np.random.default_rng(3).standard_normal((B, B)))[0]
            plan.subspace_rotation(dXs, Q, True, False)
            sub["rotation_ms"] = timed(lambda: plan.subspace_rotation(dXs, Q, True, False), 3)
        except Exception as e:  # noqa: BLE001
            sub["error"] = str(e)[:200]

        # ---- one full ChebyshevFilteredEigenSolver pass as KohnShamEigenSolver::solve drives it: Lanczos bounds (20 B=1
        # applies), Chebyshev degree from the reference's lookup table, column-batched filter, Cholesky-Gram-Schmidt,
        # Rayleigh-Ritz - everything on the device, only the B Ritz values come back ----
        chfsi = {}
        try:
            if args.quick:
                raise RuntimeError("skipped (--quick)")
            Mop = capi.DiagOp(plan, prob.diag, prob.enr_block, capi.DIAG_OEFE_MASS)
            lg = np.random.default_rng(11).uniform(-0.5, 0.5, (prob.n_local, 1))
            dlg = capi.DeviceBlock(prob.n_local, 1, lg)
            capi.lanczos_extreme(H, Mop, minv, dlg, 20)
            plan.synchronize()
            t0 = time.perf_counter()
            ev_l, ldiag, lsub, lst = capi.lanczos_extreme(H, Mop, minv, dlg, 20)
            chfsi["lanczos_ms"] = (time.perf_counter() - t0) * 1e3
            unwanted = float(ev_l[1] + lsub[-1] / 10.0)
            lower = float(ev_l[0])
            upper = (unwanted - lower) * (B * 200.0 / N_global) + lower
            if upper >= unwanted:
                upper = 0.5 * (unwanted + lower)
            deg_l = capi.chebyshev_polynomial_degree(unwanted)
            chfsi.update({"lanczos_bounds": [lower, upper, unwanted], "degree": deg_l, "lanczos_status": int(lst)})
            dG, dV = Block(X), Block()
            evs, st = capi.chfsi_solve(H, Mop, minv, dG, dV, B, deg_l, lower, upper, unwanted)  # warm-up pass
            chfsi["status_first_pass"] = int(st)
            if st == 0:
                plan.synchronize()
                plan.trace(True)
                t0 = time.perf_counter()
                evs, st = capi.chfsi_solve(H, Mop, minv, dG, dV, B, deg_l, float(evs[0]), float(evs[-1]), unwanted)
                chfsi["pass_ms"] = (time.perf_counter() - t0) * 1e3
                rep = plan.trace_report()
                plan.trace(False)
                chfsi["status"] = int(st)
                chfsi["phase_ms"] = {k: round(v["ms"], 4) for k, v in rep.items()}
                chfsi["ritz_values_lowest"] = [float(v) for v in evs[:4]]
                resn = capi.eigen_residual_norms(H, Mop, dV, evs, B)
                chfsi["residual_norms_lowest"] = [float(v) for v in resn[:4]]
            del Mop, dG, dV
        except Exception as e:  # noqa: BLE001
            chfsi["error"] = str(e)[:300]

        # ---- the steps either side of the path in one SCF iteration: cell matrices of the local
        # potential (computeFEMatrices + reinit's component sum + re-tiling for the cell kernel) and the density ----
        scf = {}
        try:
            if args.quick:
                raise RuntimeError("skipped (--quick)")
            if nranks == 1:
                fe = synth.fe_basis_data(prob)
                nqc = int(fe["num_cell_quad"][0])
                feb = capi.FeBasis(plan, fe["num_cell_quad"], fe["basis"], fe["jxw"], fe["same_basis"])
                vq = synth.potential_at_quad_points(prob, nqc)
                d_kin = capi.DeviceBlock(prob.S2, 1, 0.5 * prob.k_cell)
                d_h = capi.DeviceBlock(prob.S2, 1)
                d_v = capi.DeviceBlock(vq.size, 1, vq)
                feb.compute_fe_matrices(None, d_h, add_to=d_kin, f_device=d_v)
                scf["compute_fe_matrices_ms"] = timed(lambda: feb.compute_fe_matrices(None, d_h, add_to=d_kin, f_device=d_v), 5)
                H2 = capi.CellOp(plan, with_nonlocal=False)
                H2.set_matrices_device(d_h.ptr)
                scf["reinit_retile_ms"] = timed(lambda: H2.set_matrices_device(d_h.ptr), 3)
                feb.assemble_into(H2, None, add_to=d_kin, f_device=d_v)
                scf["assemble_into_operator_ms"] = timed(lambda: feb.assemble_into(H2, None, add_to=d_kin, f_device=d_v), 5)
                ncd64 = prob.num_cell_dofs.astype(np.int64)
                tiles = (ncd64 + 63) // 64
                fl_done = float(np.sum(2.0 * 64 * 64 * nqc * tiles * (tiles + 1) / 2))
                fl_useful = float(np.sum(2.0 * ncd64 * ncd64 * nqc)) / 2.0
                t_a = scf["compute_fe_matrices_ms"] * 1e-3
                scf["compute_fe_matrices"] = {"quad_points_per_cell": nqc, "same_basis_in_all_cells": bool(fe["same_basis"]),
                                              "dmma_tflops_issued": fl_done / t_a / 1e12,
                                              "tflops_symmetric_minimum": fl_useful / t_a / 1e12,
                                              "bytes_written_read": 16.0 * prob.S2, "gbs": 16.0 * prob.S2 / t_a / 1e9}
                occ = np.clip(np.linspace(1.5, -0.5, B), 0.0, 1.0)
                rho = feb.compute_rho(dX0, occ)
                d_rho = capi.DeviceBlock(rho.size, 1)
                scf["compute_rho_ms"] = timed(lambda: feb.compute_rho_device(dX0, occ, d_rho), 5)
                fl_r = float(np.sum(2.0 * nqc * ncd64 * B))
                scf["compute_rho"] = {"tflops": fl_r / (scf["compute_rho_ms"] * 1e-3) / 1e12, "rho_finite": bool(np.isfinite(rho).all()),
                                      "electrons": float(np.dot(rho, fe["jxw"]))}
                del feb, d_kin, d_h, d_v, H2, fe
        except Exception as e:  # noqa: BLE001
            scf["error"] = str(e)[:300]

        # ---- electrostatics: Laplace apply + Jacobi-preconditioned CG iterations on one right-hand side ----
        poisson = {}
        try:
            if args.quick:
                raise RuntimeError("skipped (--quick)")
            if nranks == 1:
                A_u = capi.CellOp(plan, h_cell=prob.k_cell, with_nonlocal=False)
                A_l = capi.CellOp(plan, h_cell=prob.k_cell, with_nonlocal=False, share_identical=True)
                poisson["unique_cell_matrices"] = [A_l.num_unique_matrices(), prob.n_cells]
                pc_l = capi.DiagOp(plan, 1.0 / prob.k_diag, None, capi.DIAG_JACOBI)
                rhs = np.zeros((prob.n_local, 1)); rhs[:prob.n_owned, 0] = X[:prob.n_owned, 0]
                rhs[prob.row_ids.astype(np.int64)] = 0.0
                db1, dx1, dy1 = capi.DeviceBlock(prob.n_local, 1, rhs), capi.DeviceBlock(prob.n_local, 1), capi.DeviceBlock(prob.n_local, 1)
                A_l.apply(db1, dy1, True, True)
                A_u.apply(db1, dy1, True, True)
                poisson["laplace_apply_ms_B1_every_cell_its_own_matrix"] = timed(lambda: A_u.apply(db1, dy1, True, True), 10)
                del A_u
                poisson["laplace_apply_ms_B1"] = timed(lambda: A_l.apply(db1, dy1, True, True), 10)
                n_it = 40
                plan.synchronize()
                t0 = time.perf_counter()
                it_done, st, rn = capi.cg_solve(A_l, pc_l, db1, dx1, n_it, 1e-30, 1e-30, 1e300)
                poisson["cg_ms_per_iteration_B1"] = (time.perf_counter() - t0) * 1e3 / max(it_done, 1)
                poisson["cg_iterations_timed"] = int(it_done)
                poisson["cg_residual_reduction"] = float(rn[0] / np.linalg.norm(rhs[:prob.n_owned, 0]))
                poisson["gdof_per_s_apply"] = N_global / (poisson["laplace_apply_ms_B1"] * 1e-3) / 1e9
                del A_l, pc_l
        except Exception as e:  # noqa: BLE001
            poisson["error"] = str(e)[:200]

        # ---- BASELINE configs[2] (the largest config that north_star's 80 % parallel-efficiency target is stated on) ----
        c3 = None
        if not args.quick and not args.no_c3 and args.workload == "c2":
            try:
                c3 = c3_strong_block(torch, dist, capi, synth, rank, nranks, local_rank, capi.microbench())
            except Exception as e:  # noqa: BLE001
                c3 = {"error": str(e)[:300]}

        # ---- end to end through the host-buffer entry points (pinned host memory, H2D + D2H every step) ----
        # (a) the call a ChebyshevFilteredEigenSolver with HOST wavefunctions makes: E2E_BATCHES column batches of B vectors,
        #     hx_chebyshev_filter_host_batches pipelines copy-in / filter / copy-out of neighbouring batches;
        # (b) a single block (nothing to overlap with: H2D, filter, D2H in sequence) for comparison.
        E2E_BATCHES = 6
        xb = [torch.from_numpy(X).pin_memory() for _ in range(E2E_BATCHES)]
        yb = [torch.zeros_like(xb[0]).pin_memory() for _ in range(E2E_BATCHES)]
        xp_, yp_ = [t_.data_ptr() for t_ in xb], [t_.data_ptr() for t_ in yb]

        def e2e_step():
            capi.chebyshev_filter_host_batches(H, minv, xp_, yp_, B, DEGREE, a0, a_, b_)

        e2e_step()
        barrier()
        e2e_steps = max(2, min(args.steps, 4))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_finite = bool(all(torch.isfinite(t_).all() for t_ in yb))
        e2e_same = bool(torch.equal(yb[0], yb[-1]))  # every batch holds the same input here: same output, bit for bit

        def e2e1_step():
            capi.chebyshev_filter_host_ptr(H, minv, xp_[0], yp_[0], B, DEGREE, a0, a_, b_, False)

        e2e1_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e1_step()
        barrier()
        e2e1_ms = (time.perf_counter() - t0) * 1e3 / 3
        sampler.stop_flag = True
        sampler.join(timeout=2)
        del xb, yb

    tms = torch.tensor([ms_per_step, apply_ms, e2e_ms, cell_ms, e2e1_ms], dtype=torch.float64, device="cuda")
    if nranks > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_per_step, apply_ms, e2e_ms, cell_ms, e2e1_ms = [float(v) for v in tms.cpu()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        # algorithmic bytes per H.X apply on one rank: stream the cell matrices once, read X once,
        # write Y once, the cell->DoF map and the constraint CSR
        S2 = prob.S2 + int(np.sum(prob.num_cell_proj.astype(np.int64) * prob.num_cell_dofs.astype(np.int64))) \
            if prob.num_cell_proj is not None else prob.S2
        alg_apply = 8 * S2 + 16 * B * prob.n_local + 4 * prob.S + 12 * prob.col_vals.size + 16 * len(prob.row_ids)
        # inside the filter the kernel also applies the recurrence on the fusable rows: + xprev (8B) and dinv per
        # such row; `out` replaces the Y write of those rows
        n_fus, n_other = plan.fusable_rows()
        alg_bytes = alg_apply + (8 * B + 8) * n_fus
        flops = 2.0 * B * S2
        cell_ms_per_launch = cell_ms / max(cell_launches, 1)
        achieved = alg_bytes / (cell_ms_per_launch * 1e-3) / 1e9
        micro = None
        try:
            micro = capi.microbench()
        except Exception:
            pass
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "cell_kernel_traffic.json")))
            if tj.get("workload") == args.workload and nranks == 1:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        blk_bytes = 8 * B * prob.n_local
        # which roof bounds the kernel: the larger of the two minimum times (algorithmic bytes at the HBM peak, flops at the
        # FP64 tensor peak); `frac` is that minimum time over the measured time, both single fractions are reported as well
        dmma_peak = micro["dmma_tflops"] if micro else None
        t_k = cell_ms_per_launch * 1e-3
        t_bytes = alg_bytes / (hbm_peak * 1e9)
        t_flops = flops / (dmma_peak * 1e12) if dmma_peak else 0.0
        tensor_bound = t_flops > t_bytes
        roof = {"bound": "tensor" if tensor_bound else "hbm",
                "achieved": (flops / t_k / 1e12) if tensor_bound else achieved,
                "peak": dmma_peak if tensor_bound else hbm_peak, "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": max(t_bytes, t_flops) / t_k, "frac_hbm": t_bytes / t_k, "frac_tensor": (t_flops / t_k) if dmma_peak else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src + "; FP64 tensor peak = hx_microbench (mma.sync.m8n8k4.f64 / SASS DMMA.8x8x4 issue rate measured on "
                                          "this GPU in this run; MEASURED_PEAKS.json has no FP64 entry)",
                "kernel": "cell_apply_pipe_kernel<FUSE> (one persistent launch per H.X apply: 4 DMMA warps + 8 scatter warps + A-stream "
                          "and gather warps per CTA; the Chebyshev update of %d of %d owned rows is applied in its scatter "
                          "epilogue)" % (n_fus, n_fus + n_other),
                "algorithmic_bytes_bare_apply": alg_apply,
                "kernel_ms_per_launch": cell_ms_per_launch, "kernel_launches_timed": int(cell_launches),
                "kernel_sm_clock_mhz": round(cell_clock_mhz, 1),
                "kernel_share_of_step": cell_ms / args.steps / ms_per_step,
                "algorithmic_bytes_per_launch": alg_bytes, "flops_per_launch": flops,
                "tensor": {"achieved_tflops": flops / t_k / 1e12,
                           "dmma_peak_tflops_measured": micro["dmma_tflops"] if micro else None,
                           "dfma_peak_tflops_measured": micro["dfma_tflops"] if micro else None,
                           "copy_gbs_measured": micro["copy_gbs"] if micro else None}}
        line = {
            "metric": METRIC, "value": DEGREE * N_global * B / (ms_per_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": nranks,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c3" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(args.workload, nranks),
            "run": {"global_dofs": N_global, "cells_per_gpu": prob.n_cells, "parallelism": f"cells/{nranks}",
                    "halo_transport": plan.halo_transport(), "halo_overlap": os.environ.get("HXB200_HALO_OVERLAP", "1") != "0",
                    "programmatic_dependent_launch": capi.pdl_enabled(), "cpu_affinity_rank0": affinity,
                    "l2_policy": "inputs larger than L2 (cell matrices %.2f GB + 4 block vectors %.2f GB per GPU)"
                                 % (8 * S2 / 1e9, 4 * blk_bytes / 1e9)},
            "e2e": {"value": DEGREE * N_global * B * E2E_BATCHES / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": blk_bytes * E2E_BATCHES, "d2h_bytes_per_step": blk_bytes * E2E_BATCHES,
                    "ms_per_step": e2e_ms, "ms_per_batch": e2e_ms / E2E_BATCHES, "column_batches": E2E_BATCHES,
                    "result_finite": e2e_finite, "batches_bitwise_equal": e2e_same,
                    "call": "hx_chebyshev_filter_host_batches (pinned host wavefunctions in %d column batches of B=%d, as "
                            "ChebyshevFilteredEigenSolver filters them: per batch H2D, %d H.X applies + M^-1 + recurrence, D2H; "
                            "the copies of neighbouring batches overlap the filter)" % (E2E_BATCHES, B, DEGREE),
                    "single_block": {"value": DEGREE * N_global * B / (e2e1_ms * 1e-3) / 1e9, "ms_per_step": e2e1_ms,
                                     "call": "hx_chebyshev_filter_host (one block: H2D, filter, D2H in sequence)"}},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof,
            "hx_apply": {"ms": apply_ms, "value": N_global * B / (apply_ms * 1e-3) / 1e9, "unit": UNIT,
                         "what": "bare KohnShamOperatorContextFE::apply (updateGhostX=true), block resident in HBM"},
            "subspace": sub,
            "chfsi_pass": chfsi,
            "scf_neighbours": scf,
            "poisson": poisson,
            "c3_strong": c3,
            "parity": parity,
            "chebyshev_filter": {"degree": DEGREE, "seconds_per_scf_iter": ms_per_step * 1e-3,
                                 "ms_per_degree": ms_per_step / DEGREE, "fused_recurrence": True,
                                 "phase_ms_per_degree": phases},
        }
        if not args.no_cpu and not args.quick and nranks == 1:
            try:
                line["cpu_baseline"] = {k: v for k, v in cpu_filter_throughput(threads=1, seconds=10.0, p=spec.p, B=B,
                                                                                 workload=args.workload).items()
                                        if k != "ms_per_step"}
            except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {e}"}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if nranks > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "small", "c3", "c1", "c2a", "c4"])
    ap.add_argument("--c4-cells", type=int, default=78, help="c4: cells per direction (78 = BASELINE configs[3], needs 8 GPUs)")
    ap.add_argument("--c4-block", type=int, default=1024)
    ap.add_argument("--c4-batch", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c3", action="store_true", help="skip the c3_strong block (BASELINE configs[2] on the same GPUs)")
    ap.add_argument("--quick", action="store_true",
                    help="only the step, its phase trace, the bare apply and e2e (skips the subspace / ChFSI pass / SCF-neighbour / "
                         "Poisson extras and the cpu_baseline leg): for A/B runs of one setting")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
