"""World-size-2 CPU check (gloo) of the host-side N>1 logic: independently built partitions + halo lists.

Each process builds only its own partition (synth.build_problem(only_rank=r)), runs the per-rank oracle
kernels, and exchanges ghost rows with torch.distributed send/recv driven by the same MPIPatternP2P lists
the CUDA library feeds to NCCL (reference src/utils/MPICommunicatorP2P.t.cpp:77-470).  The result must equal
the in-process OracleWorld over all partitions.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def exchange(dist, torch, send_rows, send_procs, send_counts, recv_procs, recv_counts, B):
    """segment i of send_rows -> send_procs[i]; returns the concatenated receive buffer."""
    reqs, off = [], 0
    recv = [torch.empty((int(c), B), dtype=torch.float64) for c in recv_counts]
    for pr, buf in zip(recv_procs, recv):
        if buf.numel():
            reqs.append(dist.irecv(buf, src=int(pr)))
    for pr, c in zip(send_procs, send_counts):
        c = int(c)
        if c:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_rows[off:off + c])), dst=int(pr)))
        off += c
    for r in reqs:
        r.wait()
    return np.concatenate([b.numpy() for b in recv]) if recv else np.zeros((0, B))


def update_ghosts(dist, torch, h, X):
    B = X.shape[1]
    counts = h.ghost_ranges[1::2] - h.ghost_ranges[0::2]
    recv = exchange(dist, torch, X[h.owned_local_ids_for_targets.astype(np.int64)], h.target_proc_ids,
                    h.num_owned_for_target, h.ghost_proc_ids, counts, B)
    X[h.n_owned + h.ghost_local_ids.astype(np.int64)] = recv


def accumulate(dist, torch, h, Y):
    B = Y.shape[1]
    counts = h.ghost_ranges[1::2] - h.ghost_ranges[0::2]
    recv = exchange(dist, torch, Y[h.n_owned + h.ghost_local_ids.astype(np.int64)], h.ghost_proc_ids, counts,
                    h.target_proc_ids, h.num_owned_for_target, B)
    np.add.at(Y, h.owned_local_ids_for_targets.astype(np.int64), recv)


def main():
    import torch
    import torch.distributed as dist
    from dft_efe_b200 import synth
    from oracle import oracle as orc

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    nc = (4, 3, 3 * world)
    L = np.array(nc, float)
    atoms = np.array([[0.5 * L[0], 0.5 * L[1], 0.5 * L[2]]])
    spec = synth.MeshSpec(ncell=nc, p=2, refine_mask=synth.refine_ball(nc, 1.0, atoms, 0.9), atoms=atoms,
                          n_enr_per_atom=2, enr_cutoff=1.2, n_proj_per_atom=2, proj_cutoff=1.0, nranks=world)
    mine = synth.build_problem(spec, only_rank=rank)[0]
    B = 5
    X = synth.make_block(mine, B)
    X[mine.n_owned:] = -3.0  # stale ghosts
    R = orc.OracleRank(mine)

    # KohnShamOperatorContextFE::apply, one rank of it (src/ksdft/KohnShamOperatorContextFE.t.cpp:1313-1443)
    update_ghosts(dist, torch, mine.halo, X)
    R.p2c(X)
    Y = np.zeros_like(X)
    CX = np.zeros((R.n_proj_local, B))
    R.loop_a(X, CX)
    accumulate(dist, torch, mine.proj_halo, CX)
    update_ghosts(dist, torch, mine.proj_halo, CX)
    CX *= R.proj_v[:, None]
    R.loop_b(Y, CX)
    R.c2p(Y)
    accumulate(dist, torch, mine.halo, Y)

    # checker: the whole world in one process
    probs = synth.build_problem(spec)
    W = orc.OracleWorld(probs)
    Xs = [synth.make_block(q, B) for q in probs]
    for q, x in zip(probs, Xs):
        x[q.n_owned:] = -3.0
    Ys = [np.zeros_like(x) for x in Xs]
    W.hx_apply(Xs, Ys, True, False)
    own = mine.n_owned
    err = np.abs(Y[:own] - Ys[rank][:own]).max() / np.abs(Ys[rank][:own]).max()
    errx = np.abs(X - Xs[rank]).max()
    print(f"[gloo rank {rank}] hx err {err:.2e}, x err {errx:.2e}", flush=True)
    ok = torch.tensor([int(err < 1e-13 and errx == 0.0)])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) else 1)


if __name__ == "__main__":
    main()
