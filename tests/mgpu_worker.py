"""One rank of the multi-GPU parity run (launched by tests/test_multi_gpu.py through torch.distributed.run).

Every rank builds its own partition of the same synthetic mesh (synth.build_problem(only_rank=r)), creates a
plan on its GPU, joins the NCCL communicator (unique id broadcast over torch.distributed), and runs the H.X
apply, the Chebyshev filter, X^T H X and the column norms through the C ABI.  The checker is the CPU oracle of
the WHOLE rank set computed redundantly on every rank (OracleWorld over all partitions, in-process "MPI").
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    den = np.linalg.norm(b, axis=0)
    den[den == 0] = 1.0
    return float((np.linalg.norm(a - b, axis=0) / den).max())


def main():
    import torch
    import torch.distributed as dist
    from dft_efe_b200 import capi, synth
    from oracle import oracle as orc

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    capi.check(capi.lib().hx_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nc = (4, 4, 3 * world)
    L = np.array(nc, float)
    atoms = np.array([[0.5 * L[0], 0.5 * L[1], 0.5 * L[2]], [0.3 * L[0], 0.7 * L[1], 0.26 * L[2]]])
    spec = synth.MeshSpec(ncell=nc, p=3, refine_mask=synth.refine_ball(nc, 1.0, [atoms[0]], 0.9), atoms=atoms,
                          n_enr_per_atom=3, enr_cutoff=1.2, n_proj_per_atom=2, proj_cutoff=1.0, nranks=world)
    probs = synth.build_problem(spec)          # all partitions: needed by the oracle world
    mine = synth.build_problem(spec, only_rank=rank)[0]
    ref = probs[rank]
    for name in ("cell_local_ids", "row_ids", "col_ids"):
        assert np.array_equal(getattr(mine, name), getattr(ref, name)), f"only_rank build differs in {name}"

    B = 16
    plan = capi.Plan(mine, max_block=B)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    plan.attach_comm(bytes(uid.cpu().numpy().tobytes()))

    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, mine.diag_inv, mine.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    W = orc.OracleWorld(probs)
    Xs = [synth.make_block(q, B) for q in probs]
    for q, x in zip(probs, Xs):          # start from stale ghosts: the update must fix them
        x[q.n_owned:] = 7.0
    errs = {}

    # ---- ghost communicator ----
    d = plan.block(B, Xs[rank])
    plan.update_ghost_values(d)
    Xo = [x.copy() for x in Xs]
    W.update_ghost_values(Xo)
    errs["update_ghost"] = float(np.abs(d.download() - Xo[rank]).max())
    d = plan.block(B, Xs[rank])
    plan.accumulate_add_locally_owned(d)
    Yo = [x.copy() for x in Xs]
    W.accumulate_add_locally_owned(Yo)
    errs["accumulate_add"] = rel(d.download()[:mine.n_owned], Yo[rank][:mine.n_owned])

    # ---- H.X with ghost update on both sides ----
    dX, dY = plan.block(B, Xs[rank]), plan.block(B)
    H.apply(dX, dY, True, True)
    Xo = [x.copy() for x in Xs]
    Yo = [np.zeros_like(x) for x in Xs]
    W.hx_apply(Xo, Yo, True, True)
    errs["hx"] = rel(dY.download(), Yo[rank])
    errs["hx_x_modified"] = rel(dX.download(), Xo[rank])

    # ---- the halo exchange overlapped with the cell kernel (default with the peer-memory transport) against the serial
    # exchange: same processing order, same arithmetic -> bitwise the same H.X and X (ghosts included) ----
    os.environ["HXB200_HALO_OVERLAP"] = "0"
    dXs, dYs = plan.block(B, Xs[rank]), plan.block(B)
    H.apply(dXs, dYs, True, True)
    os.environ.pop("HXB200_HALO_OVERLAP")
    errs["overlap_vs_serial"] = float(np.abs(dY.download() - dYs.download()).max() + np.abs(dX.download() - dXs.download()).max())
    for rep in range(3):                      # repeated applies: sequence numbers, acknowledgements, stamp reuse
        dXr, dYr = plan.block(B, Xs[rank]), plan.block(B)
        H.apply(dXr, dYr, True, False)
        H.apply(dYr, dXr, True, True)
        if rep == 0:
            first = dXr.download()
        else:
            errs["overlap_vs_serial"] += float(np.abs(dXr.download() - first).max())

    # ---- Chebyshev filter ----
    dX, dF = plan.block(B, Xs[rank]), plan.block(B)
    capi.chebyshev_filter(H, minv, dX, dF, 6, -3.0, 1.0, 60.0)
    F = W.chebyshev_filter([x.copy() for x in Xs], 6, -3.0, 1.0, 60.0)
    errs["cheb"] = rel(dF.download()[:mine.n_owned], F[rank][:mine.n_owned])

    # ---- OrthoEFEOverlapInverseOpContextGLL: one dense block over the enrichment functions of ALL ranks (all-reduce) ----
    nEs = [q.n_owned - q.n_owned_classical for q in probs]
    nEg = int(sum(nEs))
    if nEg:
        rng = np.random.default_rng(23)
        Rm = rng.standard_normal((nEg, nEg))
        blk = np.asfortranarray(Rm @ Rm.T / nEg + np.eye(nEg) + 0.05 * rng.standard_normal((nEg, nEg)))
        MIg = capi.DiagOpGlobalEnrichment(plan, mine.diag_inv, blk.ravel(order="F"), nEg, int(sum(nEs[:rank])))
        dX, dY = plan.block(B, Xs[rank]), plan.block(B)
        MIg.apply(dX, dY, True, True)
        Xo = [x.copy() for x in Xs]
        Yo = [np.zeros_like(x) for x in Xs]
        W.minv_apply_global_enrichment(Xo, Yo, blk.ravel(order="F"), True, True)
        errs["minv_global_enrichment"] = rel(dY.download(), Yo[rank])
        MIg.destroy()
    else:
        errs["minv_global_enrichment"] = 0.0

    # ---- X^T H X (NCCL all-reduce of the Gram blocks) and column norms ----
    dX = plan.block(B, Xs[rank])
    S = H.xtopx(dX, 8)
    So = W.xtopx([x.copy() for x in Xs], lambda a, b, c, d_: W.hx_apply(a, b, c, d_), 8)
    errs["xtopx"] = float(np.abs(S - So).max() / np.abs(So).max())
    dX = plan.block(B, Xs[rank])
    nr = plan.l2_norms(dX)
    errs["l2"] = float(np.abs(nr - W.l2_norms(Xs)).max() / np.abs(nr).max())

    # ---- the eigensolve around the path across ranks: Lanczos bounds, one ChebyshevFilteredEigenSolver pass (filter in
    # column batches, Cholesky-Gram-Schmidt, Rayleigh-Ritz; Gram blocks all-reduced on the device), eigen-residuals ----
    from oracle import eigensolver as es
    Mop = capi.DiagOp(plan, mine.diag, mine.enr_block, capi.DIAG_OEFE_MASS)
    A_ = lambda X_, Y_, gx, gy: W.hx_apply(X_, Y_, gx, gy)  # noqa: E731
    M_ = lambda X_, Y_, gx, gy: W.m_apply(X_, Y_, gx, gy)  # noqa: E731
    MI_ = lambda X_, Y_, gx, gy: W.minv_apply(X_, Y_, gx, gy)  # noqa: E731
    lg = [np.random.default_rng(3 + q.rank).uniform(-0.5, 0.5, (q.n_local, 1)) for q in probs]
    evo, do_, so_, sto = es.lanczos_extreme(W, A_, M_, MI_, [g.copy() for g in lg], 12)
    ev, dg, sg, st = capi.lanczos_extreme(H, Mop, minv, plan.block(1, lg[rank]), 12)
    assert st == 0 and sto == 0
    errs["lanczos"] = float(max(np.abs(dg[:4] - do_[:4]).max() / np.abs(do_).max(), np.abs(sg[:4] - so_[:4]).max() / np.abs(so_).max()))
    unwanted = float(evo[1] + so_[-1])
    guesses = [x.copy() for x in Xs]
    wo, sto, vo = es.chfsi_solve(W, guesses, np.zeros(B), 8, 10, -1.0, 6.0, unwanted)
    dG, dV = plan.block(B, Xs[rank]), plan.block(B)
    w, st = capi.chfsi_solve(H, Mop, minv, dG, dV, 8, 10, -1.0, 6.0, unwanted)
    assert st == 0 and sto == 0, (st, sto)
    errs["chfsi_ritz"] = float(np.abs(w - wo).max() / np.abs(wo).max())
    rn = capi.eigen_residual_norms(H, Mop, dV, w, 8)
    rno = es.eigen_residual_norms(W, vo, wo, 8)
    errs["eig_residuals"] = float(np.abs(rn - rno).max() / np.abs(rno).max())
    Mop.destroy()

    # ---- mesh without hanging nodes: the Chebyshev filter runs its fused path across ranks (recurrence applied in
    # the cell kernel's scatter for interior rows, row-list pass for the partition-face rows) ----
    spec2 = synth.MeshSpec(ncell=nc, p=4, atoms=atoms, n_enr_per_atom=2, enr_cutoff=1.2, n_proj_per_atom=2,
                           proj_cutoff=1.0, nranks=world)
    probs2 = synth.build_problem(spec2)
    plan2 = capi.Plan(probs2[rank], max_block=B)
    uid2 = torch.zeros(128, dtype=torch.uint8, device="cuda")   # a communicator needs its own unique id
    if rank == 0:
        uid2.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid2, 0)
    plan2.attach_comm(bytes(uid2.cpu().numpy().tobytes()))
    H2 = capi.CellOp(plan2)
    minv2 = capi.DiagOp(plan2, probs2[rank].diag_inv, probs2[rank].enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X2 = [synth.make_block(q, B) for q in probs2]
    dX, dF = plan2.block(B, X2[rank]), plan2.block(B)
    capi.chebyshev_filter(H2, minv2, dX, dF, 7, -3.0, 1.0, 60.0)
    F2 = orc.OracleWorld(probs2).chebyshev_filter([x.copy() for x in X2], 7, -3.0, 1.0, 60.0)
    errs["cheb_fused"] = rel(dF.download()[:probs2[rank].n_owned], F2[rank][:probs2[rank].n_owned])
    os.environ["HXB200_HALO_OVERLAP"] = "0"
    dXs, dFs = plan2.block(B, X2[rank]), plan2.block(B)
    capi.chebyshev_filter(H2, minv2, dXs, dFs, 7, -3.0, 1.0, 60.0)
    os.environ.pop("HXB200_HALO_OVERLAP")
    errs["overlap_vs_serial"] += float(np.abs(dF.download()[:probs2[rank].n_owned] - dFs.download()[:probs2[rank].n_owned]).max())
    plan2.synchronize()

    # ---- small cells (order 2, no enrichment): the one-m-tile-per-warp kernels with three gather warps, each of which
    # waits for the in-kernel ghost unpack on its own; H.X and the fused filter, overlapped against serial exchange ----
    spec3 = synth.MeshSpec(ncell=(5, 4, 3 * world), p=2, nranks=world)
    probs3 = synth.build_problem(spec3)
    plan3 = capi.Plan(probs3[rank], max_block=B)
    uid3 = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid3.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid3, 0)
    plan3.attach_comm(bytes(uid3.cpu().numpy().tobytes()))
    H3 = capi.CellOp(plan3)
    minv3 = capi.DiagOp(plan3, probs3[rank].diag_inv, probs3[rank].enr_block_inv, capi.DIAG_CFE)
    X3 = [synth.make_block(q, B) for q in probs3]
    W3 = orc.OracleWorld(probs3)
    n3 = probs3[rank].n_owned
    dX, dY3 = plan3.block(B, X3[rank]), plan3.block(B)
    H3.apply(dX, dY3, True, False)
    Y3 = [np.zeros_like(x) for x in X3]
    W3.hx_apply([x.copy() for x in X3], Y3, True, False)
    errs["hx_small_cells"] = rel(dY3.download()[:n3], Y3[rank][:n3])
    dX, dF3 = plan3.block(B, X3[rank]), plan3.block(B)
    capi.chebyshev_filter(H3, minv3, dX, dF3, 5, -3.0, 1.0, 60.0)
    F3 = W3.chebyshev_filter([x.copy() for x in X3], 5, -3.0, 1.0, 60.0, minv_variant="cfe")
    errs["cheb_small_cells"] = rel(dF3.download()[:n3], F3[rank][:n3])
    os.environ["HXB200_HALO_OVERLAP"] = "0"
    dXs, dFs = plan3.block(B, X3[rank]), plan3.block(B)
    capi.chebyshev_filter(H3, minv3, dXs, dFs, 5, -3.0, 1.0, 60.0)
    os.environ.pop("HXB200_HALO_OVERLAP")
    errs["overlap_vs_serial"] += float(np.abs(dF3.download()[:n3] - dFs.download()[:n3]).max())
    plan3.synchronize()
    want = os.environ.get("HXB200_EXPECT_TRANSPORT")
    if want:
        assert plan.halo_transport() == want and plan2.halo_transport() == want, (plan.halo_transport(), want)

    tol = {"hx_small_cells": 1e-12, "cheb_small_cells": 1e-11, "minv_global_enrichment": 1e-13, "overlap_vs_serial": 0.0, "update_ghost": 0.0, "cheb_fused": 1e-11, "accumulate_add": 1e-14, "hx": 1e-12, "hx_x_modified": 1e-14, "cheb": 1e-11,
           "xtopx": 1e-12, "l2": 1e-13, "lanczos": 1e-10, "chfsi_ritz": 1e-9, "eig_residuals": 1e-6}
    bad = {k: v for k, v in errs.items() if not v <= tol[k]}
    print(f"[rank {rank}/{world}] halo transport {plan.halo_transport()} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
    t = torch.tensor([len(bad)], device="cuda")
    dist.all_reduce(t)
    dist.barrier()
    # collective teardown (peer transports synchronise across ranks): operators first, then plans
    for o in (H, minv, H2, minv2, H3, minv3):
        o.destroy()
    plan.destroy()
    plan2.destroy()
    plan3.destroy()
    dist.barrier()
    dist.destroy_process_group()
    if int(t.item()):
        print(f"[rank {rank}] FAILED: {bad}", flush=True)
        sys.exit(1)


if __name__ == "__main__":
    main()
