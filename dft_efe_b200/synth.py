"""Synthetic FE problem generator (host side, numpy).

dft-efe gets its mesh, DoF numbering and hanging-node constraints from deal.II
(reference: src/basis/CFEBasisDofHandlerDealii.t.cpp:133-360,
src/basis/EFEBasisDofHandlerDealii.t.cpp:944-948,1159-1188) which is not
available here.  This module produces, for a structured hex mesh of order p
with an optional one-level 2:1 refined region, optional enrichment DoFs and
optional nonlocal projectors, *exactly the flat arrays* the reference hands to
its hot path, in the reference's types and conventions:

  * cell -> local DoF map and per-cell DoF counts      (FEBasisManager.t.cpp:110-146,262-287)
  * local ids = owned range 0 (classical), owned range 1 (enrichment), then
    ghosts in ascending global id                         (utils/MPIPatternP2P.h:107-116)
  * constraint CSR (rows owned-first then ghost, ascending global id)
                                                          (CFEConstraintsLocalDealii.t.cpp:286-462)
  * halo pattern lists of MPIPatternP2P                   (utils/MPIPatternP2P.h:425-459)
  * flat cell Hamiltonian, each cell n_c x n_c row-major  (KohnShamOperatorContextFE.t.cpp:488-507)
  * nonlocal projector cell matrices, column-major nProj_c x n_c
                                                          (AtomCenterNonLocalOpContextFE.t.cpp:478-494)

size_type = uint32, global_size_type = uint64 (utils/TypeConfig.h:8-9).
Nothing here runs on the GPU and nothing here is part of the oracle.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

U32 = np.uint32
U64 = np.uint64


# --------------------------------------------------------------------------
# 1-D GLL machinery
# --------------------------------------------------------------------------
def gll_nodes_weights(p: int) -> Tuple[np.ndarray, np.ndarray]:
    """Gauss-Lobatto-Legendre nodes / weights on [-1, 1] for order p (p+1 nodes)."""
    from numpy.polynomial import legendre as L

    if p < 1:
        raise ValueError("order must be >= 1")
    cp = np.zeros(p + 1)
    cp[p] = 1.0
    interior = np.sort(np.real(L.legroots(L.legder(cp)))) if p > 1 else np.zeros(0)
    x = np.concatenate(([-1.0], interior, [1.0]))
    # a couple of Newton steps on (1-x^2) P_p'(x) to tidy the roots
    for _ in range(3):
        if p > 1:
            d1 = L.legval(x[1:-1], L.legder(cp))
            d2 = L.legval(x[1:-1], L.legder(cp, 2))
            x[1:-1] -= d1 / d2
    w = 2.0 / (p * (p + 1) * L.legval(x, cp) ** 2)
    x = 0.5 * (x - x[::-1])  # symmetrise
    w = 0.5 * (w + w[::-1])
    return x, w


def lagrange_eval(nodes: np.ndarray, x: np.ndarray) -> np.ndarray:
    """L[i, b] = value of the b-th Lagrange polynomial on `nodes` at x[i]."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    n = len(nodes)
    out = np.ones((len(x), n))
    for b in range(n):
        for m in range(n):
            if m != b:
                out[:, b] *= (x - nodes[m]) / (nodes[b] - nodes[m])
    return out


def lagrange_deriv_at_nodes(nodes: np.ndarray, x: np.ndarray) -> np.ndarray:
    """D[i, b] = d/dx L_b at x[i]."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    n = len(nodes)
    out = np.zeros((len(x), n))
    for b in range(n):
        for k in range(n):
            if k == b:
                continue
            term = np.full(len(x), 1.0 / (nodes[b] - nodes[k]))
            for m in range(n):
                if m != b and m != k:
                    term *= (x - nodes[m]) / (nodes[b] - nodes[m])
            out[:, b] += term
    return out


def ref_matrices_1d(p: int):
    """Exact 1-D stiffness K1, consistent mass M1 and lumped (GLL) mass w on [-1,1]."""
    from numpy.polynomial import legendre as L

    x, w = gll_nodes_weights(p)
    q, qw = L.leggauss(p + 2)
    N = lagrange_eval(x, q)
    D = lagrange_deriv_at_nodes(x, q)
    K1 = (D * qw[:, None]).T @ D
    M1 = (N * qw[:, None]).T @ N
    K1 = 0.5 * (K1 + K1.T)
    M1 = 0.5 * (M1 + M1.T)
    return x, w, K1, M1


# --------------------------------------------------------------------------
# problem containers
# --------------------------------------------------------------------------
@dataclass
class MeshSpec:
    ncell: Tuple[int, int, int]
    p: int
    h: float = 1.0                      # coarse cell edge (bohr)
    refine_mask: Optional[np.ndarray] = None   # bool [nx,ny,nz] (x fastest index 0)
    boundary: str = "dirichlet"         # 'dirichlet' | 'none' | 'periodic'
    atoms: Optional[np.ndarray] = None  # [nA,3] positions in bohr
    n_enr_per_atom: int = 0
    enr_cutoff: float = 0.0
    n_proj_per_atom: int = 0
    proj_cutoff: float = 0.0
    nranks: int = 1
    seed: int = 1234
    with_k_cell: bool = True            # also build the Laplace cell matrices (Poisson tests); off for the big bench meshes
    with_h_cell: bool = True            # build the Hamiltonian cell matrices on the host (off: they are assembled on the device)


@dataclass
class HaloPattern:
    """The MPIPatternP2P getters the ghost communicator consumes
    (reference: src/utils/MPIPatternP2P.h:425-459)."""
    n_owned: int
    n_ghost: int
    owned_ranges: np.ndarray            # u64 [K,2] global [a,b)
    ghost_global_ids: np.ndarray        # u64 [n_ghost] ascending
    ghost_proc_ids: np.ndarray          # u32 [nGP]
    ghost_ranges: np.ndarray            # u32 [2*nGP] into ghost_local_ids
    ghost_local_ids: np.ndarray         # u32 [n_ghost] position in ghost_global_ids
    target_proc_ids: np.ndarray         # u32 [nTP]
    num_owned_for_target: np.ndarray    # u32 [nTP]
    owned_local_ids_for_targets: np.ndarray  # u32 [sum]

    @property
    def n_local(self) -> int:
        return self.n_owned + self.n_ghost


@dataclass
class RankProblem:
    rank: int
    nranks: int
    p: int
    halo: HaloPattern
    n_owned_classical: int
    n_cells: int
    num_cell_dofs: np.ndarray           # u32 [C]
    cell_local_ids: np.ndarray          # u32 [S]
    cell_global_index: np.ndarray       # i64 [C] position in the global cell list
    # constraints (CSR-like, reference a3)
    row_ids: np.ndarray                 # u32 [nR]
    row_sizes: np.ndarray               # u32 [nR]
    row_offsets: np.ndarray             # u32 [nR]
    col_ids: np.ndarray                 # u32 [nnz]
    col_vals: np.ndarray                # f64 [nnz]
    inhom: np.ndarray                   # f64 [nR]
    # operators
    h_cell: np.ndarray                  # f64 [S2]
    diag: np.ndarray                    # f64 [n_local]   lumped mass
    diag_inv: np.ndarray                # f64 [n_local]
    enr_block: np.ndarray               # f64 [nE_own^2]  atom-block overlap (col-major, symmetric)
    enr_block_inv: np.ndarray           # f64 [nE_own^2]
    # electrostatics: grad N_i . grad N_j cell matrices of LaplaceOperatorContextFE (same layout as
    # h_cell), the assembled diagonal on the local rows (Jacobi preconditioner) and the constraint set of the
    # "X" basis manager: the mesh's constraints with inhomogeneous Dirichlet values on the boundary rows
    k_cell: Optional[np.ndarray] = None
    k_diag: Optional[np.ndarray] = None
    inhom_dirichlet: Optional[np.ndarray] = None    # f64 [nR]: inhomogeneities of the X constraint set
    cell_edge: Optional[np.ndarray] = None          # f64 [C]: edge length of each local cell
    # nonlocal projectors
    proj_halo: Optional[HaloPattern] = None
    num_cell_proj: Optional[np.ndarray] = None      # u32 [C]
    cell_proj_local_ids: Optional[np.ndarray] = None  # u32 [sum nProj_c]
    cell_c: Optional[np.ndarray] = None             # f64 [sum n_c*nProj_c] col-major nProj_c x n_c
    proj_v: Optional[np.ndarray] = None             # f64 [nProjLocal]
    # bookkeeping for tests
    local_to_global: np.ndarray = field(default_factory=lambda: np.zeros(0, U64))
    natural_ids: np.ndarray = field(default_factory=lambda: np.zeros(0, U64))  # partition-independent dof id
    node_coords: Optional[np.ndarray] = None        # f64 [n_local_classical?,3] (local classical rows)

    @property
    def n_owned(self) -> int:
        return self.halo.n_owned

    @property
    def n_ghost(self) -> int:
        return self.halo.n_ghost

    @property
    def n_local(self) -> int:
        return self.halo.n_local

    @property
    def S(self) -> int:
        return int(self.num_cell_dofs.sum(dtype=np.int64))

    @property
    def S2(self) -> int:
        n = self.num_cell_dofs.astype(np.int64)
        return int((n * n).sum())

    @property
    def has_nonlocal(self) -> bool:
        return self.num_cell_proj is not None and int(self.num_cell_proj.sum()) > 0


# --------------------------------------------------------------------------
# halo pattern derivation (restates what MPIPatternP2P computes; brute force
# with global knowledge — the reference's own test does the same,
# test/utils/src/TestMPIPatternP2P*.cpp)
# --------------------------------------------------------------------------
def owner_of(global_ids: np.ndarray, all_ranges: np.ndarray) -> np.ndarray:
    """all_ranges: u64 [nranks, K, 2].  Returns owning rank of each id."""
    gid = np.asarray(global_ids, dtype=np.int64)
    nranks, K, _ = all_ranges.shape
    owner = np.full(gid.shape, -1, dtype=np.int64)
    for k in range(K):
        starts = all_ranges[:, k, 0].astype(np.int64)
        ends = all_ranges[:, k, 1].astype(np.int64)
        lo, hi = starts.min(), ends.max()
        inside = (gid >= lo) & (gid < hi)
        if not inside.any():
            continue
        # ranges of one rangeId are contiguous and ordered by rank
        r = np.searchsorted(ends, gid[inside], side="right")
        owner[inside] = r
    return owner


def global_to_local_owned(gid: np.ndarray, my_ranges: np.ndarray) -> np.ndarray:
    gid = np.asarray(gid, dtype=np.int64)
    out = np.full(gid.shape, -1, dtype=np.int64)
    cum = 0
    for k in range(my_ranges.shape[0]):
        a, b = int(my_ranges[k, 0]), int(my_ranges[k, 1])
        m = (gid >= a) & (gid < b)
        out[m] = gid[m] - a + cum
        cum += b - a
    return out


def derive_halo_patterns(all_ranges: np.ndarray,
                         ghosts_per_rank: Sequence[np.ndarray]) -> List[HaloPattern]:
    nranks = all_ranges.shape[0]
    ghost_side = []
    requests = [[None] * nranks for _ in range(nranks)]  # requests[owner][requester]
    for r in range(nranks):
        g = np.asarray(ghosts_per_rank[r], dtype=U64)
        assert np.all(g[1:] > g[:-1]), "ghost set must be strictly increasing"
        own = owner_of(g, all_ranges)
        assert not np.any(own == r) and not np.any(own < 0)
        procs = np.unique(own)
        flat, ranges = [], []
        pos = 0
        for q in procs:
            idx = np.nonzero(own == q)[0]
            flat.append(idx)
            ranges += [pos, pos + len(idx)]
            pos += len(idx)
            requests[int(q)][r] = g[idx]
        ghost_side.append((g, procs.astype(U32), np.asarray(ranges, U32),
                           (np.concatenate(flat) if flat else np.zeros(0)).astype(U32)))
    out = []
    for r in range(nranks):
        tprocs, counts, lids = [], [], []
        for q in range(nranks):
            req = requests[r][q]
            if req is None or len(req) == 0:
                continue
            tprocs.append(q)
            counts.append(len(req))
            l = global_to_local_owned(req, all_ranges[r])
            assert np.all(l >= 0)
            lids.append(l)
        g, procs, ranges, flat = ghost_side[r]
        n_owned = int((all_ranges[r, :, 1] - all_ranges[r, :, 0]).sum())
        out.append(HaloPattern(
            n_owned=n_owned, n_ghost=len(g), owned_ranges=all_ranges[r].astype(U64),
            ghost_global_ids=g, ghost_proc_ids=procs, ghost_ranges=ranges, ghost_local_ids=flat,
            target_proc_ids=np.asarray(tprocs, U32), num_owned_for_target=np.asarray(counts, U32),
            owned_local_ids_for_targets=(np.concatenate(lids) if lids else np.zeros(0)).astype(U32)))
    return out


# --------------------------------------------------------------------------
# mesh
# --------------------------------------------------------------------------
def _rank_grid(nranks: int, ncell) -> Tuple[int, int, int]:
    best, best_cost = (1, 1, nranks), None
    for rx in range(1, nranks + 1):
        if nranks % rx:
            continue
        for ry in range(1, nranks // rx + 1):
            if (nranks // rx) % ry:
                continue
            rz = nranks // rx // ry
            if rx > ncell[0] or ry > ncell[1] or rz > ncell[2]:
                continue
            # interface area of the brick decomposition
            cost = (rx - 1) * ncell[1] * ncell[2] + (ry - 1) * ncell[0] * ncell[2] + (rz - 1) * ncell[0] * ncell[1]
            key = (cost, rx, ry)
            if best_cost is None or key < best_cost:
                best, best_cost = (rx, ry, rz), key
    return best


def _key1d(o, s, a, p, n):
    """1-D node key (exact integer identity of a GLL node along one axis).
    o: origin in fine units, s: size in fine units (2 coarse / 1 fine), a: node index."""
    G = 2 * n + 1
    o = np.asarray(o, dtype=np.int64)
    s = np.asarray(s, dtype=np.int64)
    a = np.asarray(a, dtype=np.int64)
    key = np.where(s == 2, G + (o // 2) * (p + 1) + a, G + n * (p + 1) + o * (p + 1) + a)
    key = np.where(a == 0, o, key)
    key = np.where(a == p, o + s, key)
    if p % 2 == 0:
        key = np.where((s == 2) & (a == p // 2), o + 1, key)
    return key


def build_problem(spec: MeshSpec, only_rank: Optional[int] = None) -> List[RankProblem]:
    """Returns one RankProblem per rank (or, with only_rank=r, a list holding just rank r's problem —
    every rank of a multi-GPU job builds the same global mesh and keeps its own part)."""
    nx, ny, nz = spec.ncell
    p = spec.p
    npc = (p + 1) ** 3
    rng = np.random.default_rng(spec.seed)
    x01 = 0.5 * (gll_nodes_weights(p)[0] + 1.0)

    refine = np.zeros((nx, ny, nz), bool) if spec.refine_mask is None else np.asarray(spec.refine_mask, bool)
    assert refine.shape == (nx, ny, nz)
    if spec.boundary == "periodic":
        assert not refine.any(), "periodic + refinement not supported by the generator"

    # ---- cell list (coarse lexicographic, x fastest; children consecutive) ----
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    order = np.lexsort((I.ravel(), J.ravel(), K.ravel()))
    cI, cJ, cK = I.ravel()[order], J.ravel()[order], K.ravel()[order]
    cref = refine[cI, cJ, cK]
    nchild = np.where(cref, 8, 1)
    ncells = int(nchild.sum())
    parent = np.repeat(np.arange(len(cI)), nchild)
    first = np.concatenate(([0], np.cumsum(nchild)[:-1]))
    child = np.arange(ncells) - first[parent]
    size = np.where(cref[parent], 1, 2).astype(np.int64)
    ox = 2 * cI[parent] + np.where(size == 1, child & 1, 0)
    oy = 2 * cJ[parent] + np.where(size == 1, (child >> 1) & 1, 0)
    oz = 2 * cK[parent] + np.where(size == 1, (child >> 2) & 1, 0)

    # ---- node keys & connectivity ----
    a = np.arange(p + 1)
    kx = _key1d(ox[:, None], size[:, None], a[None, :], p, nx)
    ky = _key1d(oy[:, None], size[:, None], a[None, :], p, ny)
    kz = _key1d(oz[:, None], size[:, None], a[None, :], p, nz)
    Kx = 2 * nx + 1 + 3 * nx * (p + 1)
    Ky = 2 * ny + 1 + 3 * ny * (p + 1)
    key3 = (kx[:, None, None, :] + Kx * (ky[:, None, :, None] + Ky * kz[:, :, None, None])).reshape(ncells, npc)
    ukeys, conn = np.unique(key3.ravel(), return_inverse=True)
    conn = conn.reshape(ncells, npc).astype(np.int64)
    nnode = len(ukeys)
    # per-node 1-D keys and coordinates (bohr)
    nkx = ukeys % Kx
    nky = (ukeys // Kx) % Ky
    nkz = ukeys // (Kx * Ky)
    coords = np.zeros((nnode, 3))
    hf = spec.h / 2.0
    px = (ox[:, None] + size[:, None] * x01[None, :]) * hf
    py = (oy[:, None] + size[:, None] * x01[None, :]) * hf
    pz = (oz[:, None] + size[:, None] * x01[None, :]) * hf
    cx3 = np.broadcast_to(px[:, None, None, :], (ncells, p + 1, p + 1, p + 1)).reshape(ncells, npc)
    cy3 = np.broadcast_to(py[:, None, :, None], (ncells, p + 1, p + 1, p + 1)).reshape(ncells, npc)
    cz3 = np.broadcast_to(pz[:, :, None, None], (ncells, p + 1, p + 1, p + 1)).reshape(ncells, npc)
    coords[conn.ravel(), 0] = cx3.ravel()
    coords[conn.ravel(), 1] = cy3.ravel()
    coords[conn.ravel(), 2] = cz3.ravel()

    # ---- constraints (global, on temp node ids) ----
    con_cols = {}   # node -> (cols array, weights array)
    if refine.any():
        # interpolation rows: coarse 1-D basis evaluated at the nodes of each half
        Ih = []
        for half in (0, 1):
            m = lagrange_eval(x01, (half + x01) / 2.0)
            m[np.abs(m) < 1e-13] = 0.0
            m[np.abs(m - 1.0) < 1e-13] = 1.0
            Ih.append(m)
        e0 = np.zeros(p + 1); e0[0] = 1.0
        ep = np.zeros(p + 1); ep[p] = 1.0
        coarse_cell_of_parent = np.full(len(cI), -1, np.int64)
        coarse_cell_of_parent[~cref] = first[~cref]
        pidx = -np.ones((nx, ny, nz), np.int64)
        pidx[cI, cJ, cK] = np.arange(len(cI))
        aidx = np.arange(p + 1)
        for par in np.nonzero(cref)[0]:
            pi, pj, pk = int(cI[par]), int(cJ[par]), int(cK[par])
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        if dx == dy == dz == 0:
                            continue
                        qi, qj, qk = pi + dx, pj + dy, pk + dz
                        if not (0 <= qi < nx and 0 <= qj < ny and 0 <= qk < nz):
                            continue
                        if refine[qi, qj, qk]:
                            continue
                        ccell = coarse_cell_of_parent[pidx[qi, qj, qk]]
                        cconn = conn[ccell].reshape(p + 1, p + 1, p + 1)  # [z,y,x]
                        for ch in range(8):
                            off = (ch & 1, (ch >> 1) & 1, (ch >> 2) & 1)
                            ok = True
                            sel, wrow = [], []
                            for d, o_ in zip((dx, dy, dz), off):
                                if d == 0:
                                    sel.append(aidx); wrow.append(Ih[o_])
                                elif d == 1:
                                    if o_ != 1: ok = False; break
                                    sel.append(np.array([p])); wrow.append(e0[None, :])
                                else:
                                    if o_ != 0: ok = False; break
                                    sel.append(np.array([0])); wrow.append(ep[None, :])
                            if not ok:
                                continue
                            fcell = first[par] + ch
                            fconn = conn[fcell].reshape(p + 1, p + 1, p + 1)
                            for az in range(len(sel[2])):
                                for ay in range(len(sel[1])):
                                    for ax in range(len(sel[0])):
                                        node = fconn[sel[2][az], sel[1][ay], sel[0][ax]]
                                        if node in con_cols:
                                            continue
                                        w3 = (wrow[2][az][:, None, None] * wrow[1][ay][None, :, None]
                                              * wrow[0][ax][None, None, :])
                                        nzm = w3 != 0.0
                                        cols = cconn[nzm]
                                        ws = w3[nzm]
                                        if len(cols) == 1 and cols[0] == node:
                                            continue
                                        assert node not in cols
                                        con_cols[int(node)] = (cols.copy(), ws.copy())
    dirichlet = np.zeros(nnode, bool)
    if spec.boundary == "dirichlet":
        dirichlet = (nkx == 0) | (nkx == 2 * nx) | (nky == 0) | (nky == 2 * ny) | (nkz == 0) | (nkz == 2 * nz)
    periodic_master = None
    if spec.boundary == "periodic":
        mx = np.where(nkx == 2 * nx, 0, nkx)
        my = np.where(nky == 2 * ny, 0, nky)
        mz = np.where(nkz == 2 * nz, 0, nkz)
        mkey = mx + Kx * (my + Ky * mz)
        slave = mkey != ukeys
        master = np.searchsorted(ukeys, mkey[slave])
        assert np.all(ukeys[master] == mkey[slave])
        periodic_master = (np.nonzero(slave)[0], master)

    constrained = dirichlet.copy()
    for n_ in con_cols:
        constrained[n_] = True
    if periodic_master is not None:
        constrained[periodic_master[0]] = True
    # global constraint table: row -> (cols, vals) with constrained columns
    # resolved (deal.II AffineConstraints::close() semantics)
    con_table = {}
    for n_ in np.nonzero(dirichlet)[0]:
        con_table[int(n_)] = (np.zeros(0, np.int64), np.zeros(0))
    for n_, (cols, ws) in con_cols.items():
        if dirichlet[n_]:
            continue
        keep = ~dirichlet[cols]
        assert not np.any(constrained[cols] & ~dirichlet[cols]), "constraint chain"
        con_table[n_] = (cols[keep], ws[keep])
    if periodic_master is not None:
        for s_, m_ in zip(*periodic_master):
            con_table[int(s_)] = (np.array([m_], np.int64), np.array([1.0]))

    # ---- enrichment / projector incidence (global) ----
    ccen = np.stack([(ox + size / 2.0) * hf, (oy + size / 2.0) * hf, (oz + size / 2.0) * hf], axis=1)
    atoms = np.zeros((0, 3)) if spec.atoms is None else np.asarray(spec.atoms, float).reshape(-1, 3)
    nA = len(atoms)

    def incidence(cutoff, per_atom):
        inc = [[] for _ in range(ncells)]
        if per_atom == 0 or nA == 0:
            return inc
        for ia in range(nA):
            d = np.linalg.norm(ccen - atoms[ia][None, :], axis=1)
            for c in np.nonzero(d <= cutoff)[0]:
                inc[c].append(ia)
        return inc

    enr_inc = incidence(spec.enr_cutoff, spec.n_enr_per_atom)
    proj_inc = incidence(spec.proj_cutoff, spec.n_proj_per_atom)

    # ---- partition ----
    rx, ry, rz = _rank_grid(spec.nranks, spec.ncell)
    prank = (cI * rx // nx) + rx * ((cJ * ry // ny) + ry * (cK * rz // nz))
    crank = prank[parent]
    node_owner = np.full(nnode, spec.nranks, np.int64)
    np.minimum.at(node_owner, conn.ravel(), np.repeat(crank, npc))
    # atom owner: rank of the coarse cell containing the atom
    atom_owner = np.zeros(nA, np.int64)
    for ia in range(nA):
        ai = np.clip((atoms[ia] / spec.h).astype(int), 0, [nx - 1, ny - 1, nz - 1])
        atom_owner[ia] = (ai[0] * rx // nx) + rx * ((ai[1] * ry // ny) + ry * (ai[2] * rz // nz))

    # global numbering: rank-major contiguous owned ranges
    ncl_per_rank = np.bincount(node_owner, minlength=spec.nranks)[:spec.nranks]
    cl_off = np.concatenate(([0], np.cumsum(ncl_per_rank)))
    node_gid = np.zeros(nnode, np.int64)
    for r in range(spec.nranks):
        m = np.nonzero(node_owner == r)[0]
        node_gid[m] = cl_off[r] + np.arange(len(m))
    Ncl = nnode
    ne = spec.n_enr_per_atom
    atom_order = np.argsort(atom_owner, kind="stable")
    atom_slot = np.zeros(nA, np.int64); atom_slot[atom_order] = np.arange(nA)
    enr_gid0 = Ncl + atom_slot * ne                       # first enrichment gid of each atom
    nenr_per_rank = np.bincount(atom_owner, minlength=spec.nranks) * ne if nA else np.zeros(spec.nranks, np.int64)
    en_off = Ncl + np.concatenate(([0], np.cumsum(nenr_per_rank)))
    npj = spec.n_proj_per_atom
    proj_gid0 = atom_slot * npj
    nproj_per_rank = np.bincount(atom_owner, minlength=spec.nranks) * npj if nA else np.zeros(spec.nranks, np.int64)
    pj_off = np.concatenate(([0], np.cumsum(nproj_per_rank)))

    all_ranges = np.zeros((spec.nranks, 2, 2), U64)
    all_ranges[:, 0, 0] = cl_off[:-1]; all_ranges[:, 0, 1] = cl_off[1:]
    all_ranges[:, 1, 0] = en_off[:-1]; all_ranges[:, 1, 1] = en_off[1:]
    proj_ranges = np.zeros((spec.nranks, 1, 2), U64)
    proj_ranges[:, 0, 0] = pj_off[:-1]; proj_ranges[:, 0, 1] = pj_off[1:]

    gid_to_node = np.zeros(nnode, np.int64)
    gid_to_node[node_gid] = np.arange(nnode)

    # ---- reference matrices ----
    _, w1, K1, M1 = ref_matrices_1d(p)

    def cell_mats(hc):
        k1 = K1 * (2.0 / hc); m1 = M1 * (hc / 2.0)
        Kc = (np.einsum("ad,be,cf->abcdef", m1, m1, k1) + np.einsum("ad,be,cf->abcdef", m1, k1, m1)
              + np.einsum("ad,be,cf->abcdef", k1, m1, m1)).reshape(npc, npc)
        Mc = np.einsum("ad,be,cf->abcdef", m1, m1, m1).reshape(npc, npc)
        wl = (w1 * hc / 2.0)
        lump = (wl[:, None, None] * wl[None, :, None] * wl[None, None, :]).ravel()
        return 0.5 * (Kc + Kc.T), 0.5 * (Mc + Mc.T), lump
    mats = {2: cell_mats(spec.h), 1: cell_mats(spec.h / 2.0)}

    # smooth-ish local potential per cell: soft Coulomb wells at the atoms + seeded jitter
    vcell = -0.2 + 0.05 * rng.standard_normal(ncells)
    for ia in range(nA):
        d = np.linalg.norm(ccen - atoms[ia][None, :], axis=1)
        vcell += -2.0 / np.sqrt(d * d + 0.5)

    # lumped mass (and the diagonal of the assembled stiffness matrix) assembled globally (every rank later picks
    # its local rows)
    lump_g = np.zeros(nnode)
    kdiag_g = np.zeros(nnode)
    for s_ in (1, 2):
        m = size == s_
        if m.any():
            np.add.at(lump_g, conn[m].ravel(), np.tile(mats[s_][2], int(m.sum())))
            np.add.at(kdiag_g, conn[m].ravel(), np.tile(np.diag(mats[s_][0]), int(m.sum())))

    # The reference condenses the lumped diagonal through the constraints (distributeChildToParent: every parent
    # collects w * child), then zeroes the constrained entries of M's diagonal
    # (basis/OrthoEFEOverlapOperatorContext.t.cpp:1525-1535) and of the reciprocal
    # (basis/CFEOverlapInverseOpContextGLL.t.cpp:447-466), so that M^-1 M = I on the unconstrained rows of
    # non-conforming meshes too.
    lump_c = lump_g.copy()
    for n_, (cols_, ws_) in con_table.items():
        if len(cols_):
            np.add.at(lump_c, cols_, ws_ * lump_g[n_])
    lump_c[constrained] = 0.0

    # per-atom enrichment overlap block (SPD) and projector strengths
    enr_blocks = []
    for ia in range(nA):
        if ne:
            A_ = rng.uniform(-1, 1, (ne, ne)) * 0.1
            enr_blocks.append(np.eye(ne) + A_ @ A_.T)
    proj_strength = rng.uniform(0.5, 2.0, (nA, max(npj, 1))) * np.where(rng.uniform(size=(nA, max(npj, 1))) < 0.3, -1, 1)

    # ---- per-rank assembly ----
    ghosts_per_rank, pghosts_per_rank, per_rank = [], [], []
    for r in range(spec.nranks):
        cells = np.nonzero(crank == r)[0]
        lnodes = np.unique(conn[cells].ravel())
        # constraint closure: parents of local constrained nodes become local
        extra = [con_table[int(n_)][0] for n_ in lnodes[constrained[lnodes]] if int(n_) in con_table]
        if extra:
            lnodes = np.unique(np.concatenate([lnodes] + extra))
        gids = node_gid[lnodes]
        own_m = node_owner[lnodes] == r
        # all owned nodes are local even if untouched by local cells (cannot happen: owner = min rank of touching cells)
        ghost_cl = np.sort(gids[~own_m])
        enr_atoms = sorted({ia for c in cells for ia in enr_inc[c]})
        enr_g = np.array([enr_gid0[ia] + k for ia in enr_atoms for k in range(ne)], np.int64)
        enr_ghost = np.sort(enr_g[(enr_g < en_off[r]) | (enr_g >= en_off[r + 1])]) if len(enr_g) else np.zeros(0, np.int64)
        ghosts_per_rank.append(np.concatenate([ghost_cl, enr_ghost]).astype(U64))
        pj_atoms = sorted({ia for c in cells for ia in proj_inc[c]})
        pj_g = np.array([proj_gid0[ia] + k for ia in pj_atoms for k in range(npj)], np.int64)
        pj_ghost = np.sort(pj_g[(pj_g < pj_off[r]) | (pj_g >= pj_off[r + 1])]) if len(pj_g) else np.zeros(0, np.int64)
        pghosts_per_rank.append(pj_ghost.astype(U64))
        per_rank.append(cells)

    halos = derive_halo_patterns(all_ranges, ghosts_per_rank)
    phalos = derive_halo_patterns(proj_ranges, pghosts_per_rank) if npj and nA else [None] * spec.nranks

    problems = []
    for r in range(spec.nranks):
        if only_rank is not None and r != only_rank:
            continue
        cells = per_rank[r]
        halo = halos[r]
        n_owned, n_ghost = halo.n_owned, halo.n_ghost
        n_own_cl = int(ncl_per_rank[r])
        l2g = np.concatenate([np.arange(cl_off[r], cl_off[r + 1]), np.arange(en_off[r], en_off[r + 1]),
                              halo.ghost_global_ids.astype(np.int64)]).astype(np.int64)

        def g2l(g):
            g = np.asarray(g, np.int64)
            out = global_to_local_owned(g, all_ranges[r])
            miss = out < 0
            if miss.any():
                pos = np.searchsorted(halo.ghost_global_ids.astype(np.int64), g[miss])
                assert np.all(halo.ghost_global_ids.astype(np.int64)[pos] == g[miss])
                out[miss] = n_owned + pos
            return out

        # cell maps (vectorised for cells without enrichment / projectors)
        C = len(cells)
        cl_all = g2l(node_gid[conn[cells]].ravel()).reshape(C, npc)
        nenr_c = np.array([len(enr_inc[c]) * ne for c in cells], np.int64) if (ne and nA) else np.zeros(C, np.int64)
        ncd64 = npc + nenr_c
        off1 = np.concatenate(([0], np.cumsum(ncd64)))
        off2 = np.concatenate(([0], np.cumsum(ncd64 * ncd64)))
        ids = np.zeros(int(off1[-1]), np.int64)
        h_cell = np.empty(int(off2[-1])) if spec.with_h_cell else None
        k_cell = np.empty(int(off2[-1])) if (spec.with_k_cell and spec.with_h_cell) else None
        plain = nenr_c == 0
        for s_ in (1, 2):
            m = plain & (size[cells] == s_)
            if not m.any():
                continue
            Kc, Mc, _ = mats[s_]
            kr, mr = 0.5 * Kc.ravel(), Mc.ravel()
            # plain cells of one size occupy contiguous runs of the flat arrays: fill them as 2-D views, a few hundred
            # cells at a time (a 7 M-DoF order-6 mesh has 31 GB of cell matrices)
            mi = np.nonzero(m)[0]
            brk = np.nonzero(np.diff(mi) != 1)[0] + 1
            for run in (np.split(mi, brk) if h_cell is not None else []):
                for c0 in range(0, len(run), 256):
                    rr = run[c0:c0 + 256]
                    a_, b_ = int(off2[rr[0]]), int(off2[rr[-1] + 1])
                    view = h_cell[a_:b_].reshape(len(rr), npc * npc)
                    np.multiply(vcell[cells[rr]][:, None], mr[None, :], out=view)
                    view += kr[None, :]
                    if k_cell is not None:
                        k_cell[a_:b_].reshape(len(rr), npc * npc)[:] = Kc.ravel()[None, :]
            ids[(off1[:-1][m][:, None] + np.arange(npc)[None, :]).ravel()] = cl_all[m].ravel()
        for ic in np.nonzero(~plain)[0]:
            c = cells[ic]
            eg = np.array([enr_gid0[ia] + k for ia in enr_inc[c] for k in range(ne)], np.int64)
            el = g2l(eg)
            n_c = int(ncd64[ic])
            ids[off1[ic]:off1[ic + 1]] = np.concatenate([cl_all[ic], el])
            Kc, Mc, lump_c_ = mats[int(size[c])]
            Hc = np.zeros((n_c, n_c))
            Hc[:npc, :npc] = 0.5 * Kc + vcell[c] * Mc
            crng = np.random.default_rng(spec.seed + 7919 * (int(c) + 1))
            vol = (spec.h * size[c] / 2.0) ** 3
            # classical-enrichment coupling weighted by the nodal mass (as an integral of N_i against a smooth
            # enrichment function would be), so that the pencil (H, M) keeps a physical spectrum
            Bc = crng.uniform(-1, 1, (npc, len(el))) * lump_c_[:, None] * 0.1
            Ec = crng.uniform(-1, 1, (len(el), len(el))) * 1e-2 * vol
            Hc[:npc, npc:] = Bc
            Hc[npc:, :npc] = Bc.T
            Hc[npc:, npc:] = 0.5 * (Ec + Ec.T) + np.eye(len(el)) * 0.5 * vol
            h_cell[off2[ic]:off2[ic + 1]] = Hc.ravel()
            if k_cell is not None:
                Kx = np.zeros((n_c, n_c))
                Kx[:npc, :npc] = Kc
                Kx[npc:, npc:] = np.eye(len(el)) * vol
                k_cell[off2[ic]:off2[ic + 1]] = Kx.ravel()
        ncd = ncd64.astype(U32)
        ids = ids.astype(U32)
        ncp, pids, cblocks = [], [], []
        if npj and nA:
            ph = phalos[r]
            for ic, c in enumerate(cells):
                pg = np.array([proj_gid0[ia] + k for ia in proj_inc[c] for k in range(npj)], np.int64)
                ncp.append(len(pg))
                if len(pg):
                    pl = global_to_local_owned(pg, proj_ranges[r])
                    miss = pl < 0
                    if miss.any():
                        pos = np.searchsorted(ph.ghost_global_ids.astype(np.int64), pg[miss])
                        pl[miss] = ph.n_owned + pos
                    pids.append(pl)
                    crng = np.random.default_rng(spec.seed + 104729 * (int(c) + 1))
                    vol = (spec.h * size[c] / 2.0) ** 3
                    # projector values integrated against the shape functions: weighted by the nodal mass (see Bc)
                    wdof = np.full(int(ncd64[ic]), np.sqrt(vol) * 0.05)
                    wdof[:npc] = mats[int(size[c])][2] * (0.3 / np.sqrt(vol))
                    Cc = crng.uniform(-1, 1, (len(pg), int(ncd64[ic]))) * wdof[None, :]   # [proj, dof]
                    cblocks.append(Cc.T.ravel())  # column-major nProj_c x n_c  == C[p + j*nP]

        # constraints restricted to local rows; owned rows first then ghosts, ascending gid
        ln = np.concatenate([gid_to_node[l2g[:n_own_cl]],
                             gid_to_node[l2g[n_owned:][l2g[n_owned:] < Ncl]]])
        lrow_local = np.concatenate([np.arange(n_own_cl),
                                     n_owned + np.nonzero(l2g[n_owned:] < Ncl)[0]])
        cm = constrained[ln]
        rows_local = lrow_local[cm]
        rows_node = ln[cm]
        # local-relevant filter: a ghost row whose parents are not local is dropped
        row_ids, row_sizes, row_offsets, col_ids, col_vals = [], [], [], [], []
        off = 0
        ghost_g = halo.ghost_global_ids.astype(np.int64)
        for rl, rn in zip(rows_local, rows_node):
            cols, ws = con_table[int(rn)]
            cg = node_gid[cols]
            cl = global_to_local_owned(cg, all_ranges[r])
            miss = cl < 0
            if miss.any():
                pos = np.searchsorted(ghost_g, cg[miss])
                pos = np.minimum(pos, max(len(ghost_g) - 1, 0))
                okk = (ghost_g[pos] == cg[miss]) if len(ghost_g) else np.zeros(miss.sum(), bool)
                if not np.all(okk):
                    continue  # not locally resolvable (ghost row outside trimmed set)
                cl[miss] = n_owned + pos
            row_ids.append(rl); row_sizes.append(len(cols)); row_offsets.append(off)
            col_ids.append(cl); col_vals.append(ws)
            off += len(cols)
        prob = RankProblem(
            rank=r, nranks=spec.nranks, p=p, halo=halo, n_owned_classical=n_own_cl,
            n_cells=len(cells), num_cell_dofs=ncd, cell_local_ids=ids,
            cell_global_index=cells.astype(np.int64),
            row_ids=np.asarray(row_ids, U32), row_sizes=np.asarray(row_sizes, U32),
            row_offsets=np.asarray(row_offsets, U32),
            col_ids=(np.concatenate(col_ids) if col_ids else np.zeros(0)).astype(U32),
            col_vals=(np.concatenate(col_vals) if col_vals else np.zeros(0)).astype(np.float64),
            inhom=np.zeros(len(row_ids)),
            h_cell=h_cell, diag=np.ones(halo.n_local), diag_inv=np.ones(halo.n_local),
            enr_block=np.zeros(0), enr_block_inv=np.zeros(0), local_to_global=l2g.astype(U64))
        # lumped mass on classical local rows
        clm = l2g < Ncl
        prob.diag[clm] = lump_c[gid_to_node[l2g[clm]]]
        prob.diag_inv = np.where(prob.diag != 0.0, 1.0 / np.where(prob.diag != 0.0, prob.diag, 1.0), 0.0)
        prob.k_cell = k_cell
        prob.cell_edge = spec.h * size[cells].astype(np.float64) / 2.0  # edge length of each local cell
        prob.k_diag = np.ones(halo.n_local)
        prob.k_diag[clm] = kdiag_g[gid_to_node[l2g[clm]]]
        if (~clm).any() and k_cell is not None:  # enrichment rows: sum of the identity*vol blocks of the touching cells
            kd = np.zeros(halo.n_local)
            np.add.at(kd, ids.astype(np.int64), np.concatenate(
                [np.diag(k_cell[off2[i]:off2[i + 1]].reshape(int(ncd64[i]), int(ncd64[i]))) for i in range(C)]))
            prob.k_diag[~clm] = np.maximum(kd[~clm], 1e-12)
        # inhomogeneous Dirichlet data for the Poisson "X" constraints: boundary rows (no parents) carry g(x)
        xyz = coords[gid_to_node[l2g[prob.row_ids.astype(np.int64)]]] if len(row_ids) else np.zeros((0, 3))
        prob.inhom_dirichlet = np.where(prob.row_sizes == 0, 0.3 + 0.1 * xyz[:, 0] - 0.05 * xyz[:, 1] * xyz[:, 2], 0.0) \
            if len(row_ids) else np.zeros(0)
        nat = np.zeros(halo.n_local, np.int64)
        nat[clm] = gid_to_node[l2g[clm]]
        if (~clm).any():
            eg_ = l2g[~clm] - Ncl
            inv_slot_e = np.argsort(atom_slot)
            nat[~clm] = Ncl + inv_slot_e[eg_ // ne] * ne + eg_ % ne
        prob.natural_ids = nat.astype(U64)
        nco = np.full((halo.n_local, 3), np.nan)
        nco[clm] = coords[gid_to_node[l2g[clm]]]
        prob.node_coords = nco
        # atom-block enrichment overlap for locally owned enrichment ids (block diagonal by atom)
        nE = n_owned - n_own_cl
        if nE:
            blk = np.zeros((nE, nE))
            my_atoms = [ia for ia in atom_order if atom_owner[ia] == r]
            for j, ia in enumerate(my_atoms):
                blk[j * ne:(j + 1) * ne, j * ne:(j + 1) * ne] = enr_blocks[ia]
            prob.enr_block = blk.ravel().copy()
            prob.enr_block_inv = np.linalg.inv(blk).ravel().copy()
        if npj and nA:
            ph = phalos[r]
            prob.proj_halo = ph
            prob.num_cell_proj = np.asarray(ncp, U32)
            prob.cell_proj_local_ids = (np.concatenate(pids) if pids else np.zeros(0)).astype(U32)
            prob.cell_c = np.concatenate(cblocks) if cblocks else np.zeros(0)
            pl2g = np.concatenate([np.arange(pj_off[r], pj_off[r + 1]), ph.ghost_global_ids.astype(np.int64)])
            inv_slot = np.argsort(atom_slot)
            prob.proj_v = np.array([proj_strength[inv_slot[g // npj], g % npj] for g in pl2g], float)
        problems.append(prob)
    return problems


# --------------------------------------------------------------------------
# fast path for big uniform benchmark meshes (no refinement, one rank slab per call)
# --------------------------------------------------------------------------
def counter_uniform(seed: int, idx: np.ndarray) -> np.ndarray:
    """Partition-independent U(-0.5,0.5): a counter-based hash (splitmix64) of
    (seed, idx) so X[global dof, vec] does not depend on how the mesh is cut."""
    z = np.asarray(idx, dtype=np.uint64) + np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5


def make_block(prob: RankProblem, B: int, seed: int = 42) -> np.ndarray:
    """X[i*B+v] for local rows (owned + ghost), value keyed by (partition-independent dof id, vec)."""
    g = prob.natural_ids.astype(np.uint64)
    idx = g[:, None] * np.uint64(4096) + np.arange(B, dtype=np.uint64)[None, :]
    return counter_uniform(seed, idx).reshape(prob.n_local, B)


def refine_ball(ncell, h, centers, radius) -> np.ndarray:
    nx, ny, nz = ncell
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cen = np.stack([(I + 0.5) * h, (J + 0.5) * h, (K + 0.5) * h], axis=-1)
    mask = np.zeros((nx, ny, nz), bool)
    for c in np.asarray(centers, float).reshape(-1, 3):
        mask |= np.linalg.norm(cen - c[None, None, None, :], axis=-1) <= radius
    return mask


def fe_basis_data(prob: RankProblem, nq1d: Optional[int] = None, seed: int = 77):
    """Synthetic FEBasisDataStorage arrays for FEBasisOperations::computeFEMatrices
    (reference src/basis/FEBasisOperations.t.cpp:111-160): Gauss-Legendre tensor rule with nq1d points per direction
    (default p + 2), basis values per cell as getBasisDataInCellRange lays them out (nq x n_c, DoF index fastest),
    JxW per quadrature point.  Classical columns: the tensor-product Lagrange (GLL) shape functions in the cell-local
    DoF order of the generator; enrichment columns: seeded smooth-ish values (counter-based on the global cell index, so
    partition independent).  Returns dict(num_cell_quad, basis, jxw, same_basis, quad_weights_ref)."""
    from numpy.polynomial import legendre as L
    p = prob.p
    nq1d = nq1d or (p + 2)
    x, _ = gll_nodes_weights(p)
    q1, w1 = L.leggauss(nq1d)
    Lq = lagrange_eval(x, q1)                                   # [nq1d, p+1]
    Ncl = np.einsum("ia,jb,kc->ijkabc", Lq, Lq, Lq).reshape(nq1d ** 3, (p + 1) ** 3)
    wq = np.einsum("i,j,k->ijk", w1, w1, w1).ravel()
    C = prob.n_cells
    nq = nq1d ** 3
    npc = (p + 1) ** 3
    ncd = prob.num_cell_dofs.astype(np.int64)
    same = bool(np.all(ncd == npc))
    jxw = (wq[None, :] * ((prob.cell_edge / 2.0) ** 3)[:, None]).ravel()
    if same:
        basis = np.ascontiguousarray(Ncl).ravel()
    else:
        parts = []
        for c in range(C):
            ne = int(ncd[c] - npc)
            if ne == 0:
                parts.append(Ncl.ravel())
            else:
                crng = np.random.default_rng(seed + 15485863 * (int(prob.cell_global_index[c]) + 1))
                E = 0.5 * crng.uniform(-1.0, 1.0, (nq, ne))
                parts.append(np.concatenate([Ncl, E], axis=1).ravel())
        basis = np.concatenate(parts)
    return {"num_cell_quad": np.full(C, nq, np.uint32), "basis": basis, "jxw": jxw, "same_basis": same,
            "quad_weights_ref": wq, "classical_basis": Ncl}


def potential_at_quad_points(prob: RankProblem, nq: int, seed: int = 5) -> np.ndarray:
    """A smooth-ish synthetic local potential at the quadrature points (counter-based on the global cell index)."""
    out = np.empty((prob.n_cells, nq))
    for c in range(prob.n_cells):
        crng = np.random.default_rng(seed + 32452843 * (int(prob.cell_global_index[c]) + 1))
        out[c] = -0.8 + 0.3 * crng.uniform(-1.0, 1.0) + 0.05 * crng.uniform(-1.0, 1.0, nq)
    return out.ravel()
