"""debug of the overlapped halo (N ranks): H.X in serial and overlapped mode against the oracle, row classes of the differences"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    import torch, torch.distributed as dist
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
    for label, spec in (("p3 hanging enr+proj", synth.MeshSpec(ncell=nc, p=3, refine_mask=synth.refine_ball(nc, 1.0, [atoms[0]], 0.9), atoms=atoms,
                          n_enr_per_atom=3, enr_cutoff=1.2, n_proj_per_atom=2, proj_cutoff=1.0, nranks=world)),
                        ("p3 hanging plain", synth.MeshSpec(ncell=nc, p=3, refine_mask=synth.refine_ball(nc, 1.0, [atoms[0]], 0.9), nranks=world)),
                        ("p4 enr+proj", synth.MeshSpec(ncell=nc, p=4, atoms=atoms, n_enr_per_atom=2, enr_cutoff=1.2, n_proj_per_atom=2, proj_cutoff=1.0, nranks=world))):
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
        W = orc.OracleWorld(probs)
        Xs = [synth.make_block(p_, B) for p_ in probs]
        for p_, x_ in zip(probs, Xs):
            x_[p_.n_owned:] = 7.0   # stale ghosts: the update must fix them
        Xo = [x.copy() for x in Xs]
        Yo = [np.zeros_like(x) for x in Xs]
        W.hx_apply(Xo, Yo, True, False)
        ids = q.cell_local_ids.astype(np.int64)
        inc = np.bincount(ids, minlength=q.n_local)
        minv = capi.DiagOp(plan, q.diag_inv, q.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
        F = W.chebyshev_filter([x.copy() for x in Xs], 5, -3.0, 1.0, 60.0)
        Yg = [np.zeros_like(x) for x in Xs]
        W.hx_apply([x.copy() for x in Xs], Yg, True, True)
        for mode in ("0", "1"):
            os.environ["HXB200_HALO_OVERLAP"] = mode
            dX, dF = plan.block(B, Xs[rank]), plan.block(B)
            capi.chebyshev_filter(H, minv, dX, dF, 5, -3.0, 1.0, 60.0)
            eF = np.abs(dF.download()[:q.n_owned] - F[rank][:q.n_owned]).max(axis=1)
            dX, dY = plan.block(B, Xs[rank]), plan.block(B)
            H.apply(dX, dY, True, True)
            eG = np.abs(dY.download() - Yg[rank]).max(axis=1)
            print(f"[{label}] rank {rank} overlap={mode}: filter rows wrong {int((eF > 1e-9).sum())} max {eF.max():.2e}; apply(ugy) owned wrong {int((eG[:q.n_owned] > 1e-10).sum())} ghost wrong {int((eG[q.n_owned:] > 1e-10).sum())} of {q.n_ghost}", flush=True)
        for mode in ("0", "1", "1"):
            os.environ["HXB200_HALO_OVERLAP"] = mode
            dX, dY = plan.block(B, Xs[rank]), plan.block(B)
            H.apply(dX, dY, True, False)
            Y = dY.download()
            err = np.abs(Y - Yo[rank]).max(axis=1)
            bad = np.nonzero(err[:q.n_owned] > 1e-10)[0]
            acc = set(q.halo.owned_local_ids_for_targets.tolist()) if hasattr(q, "halo") else set()
            nb_acc = sum(1 for r in bad if int(r) in acc)
            print(f"[{label}] rank {rank} overlap={mode}: owned rows wrong {len(bad)} of {q.n_owned} (of which halo targets {nb_acc} of {len(acc)}), "
                  f"max err {err[:q.n_owned].max():.2e}; staged(>8) among bad {int((inc[bad] > 8).sum())}; X err {np.abs(dX.download() - Xo[rank]).max():.2e}", flush=True)
            if len(bad):
                r = int(bad[0])
                print(f"   first bad row {r}: got {Y[r, :3]} want {Yo[rank][r, :3]} diff {Y[r,:3]-Yo[rank][r,:3]}", flush=True)
        os.environ.pop("HXB200_HALO_OVERLAP")
        plan.synchronize()
        dist.barrier()
        minv.destroy(); H.destroy(); plan.destroy()
        dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
