"""CPU oracle: orchestration of the C restatement in hx_oracle.c.

TEST INFRASTRUCTURE ONLY (see hx_oracle.c header): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by anything under dft_efe_b200/.

An `OracleWorld` holds every rank's arrays of one partitioned problem in this
process and replays, rank by rank, what each reference MPI rank does; the MPI
messages of MPICommunicatorP2P become buffer copies between the rank objects.
All block vectors are numpy float64 [n_local, B] (vector index fastest — the
reference MultiVector layout, src/linearAlgebra/MultiVector.h:134-160).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_u32p = C.POINTER(C.c_uint32)
c_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/liborc.so."""
    src = os.path.join(_HERE, "hx_oracle.c")
    out = os.path.join(_HERE, "liborc.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11",
                               "-o", out, src, "-lm"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def use_scipy_dgemm(enable: bool = True) -> bool:
    """Route every dgemm of the oracle through SciPy's bundled OpenBLAS
    (Fortran-ABI pointer from scipy.linalg.cython_blas) — the stand-in for the
    optimised BLAS (MKL/BLIS) a reference build links."""
    L = lib()
    L.orc_set_dgemm.argtypes = [C.c_void_p]
    if not enable:
        L.orc_set_dgemm(None)
        return False
    try:
        import scipy.linalg.cython_blas as cb
        cap = cb.__pyx_capi__["dgemm"]
        C.pythonapi.PyCapsule_GetName.restype = C.c_char_p
        C.pythonapi.PyCapsule_GetName.argtypes = [C.py_object]
        C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
        C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
        ptr = C.pythonapi.PyCapsule_GetPointer(cap, C.pythonapi.PyCapsule_GetName(cap))
        L.orc_set_dgemm(ptr)
        return True
    except Exception:
        L.orc_set_dgemm(None)
        return False


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(c_u32p)


def _f64(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_f64p)


class OracleRank:
    """One reference MPI rank: FEBasisManager arrays + operator data."""

    def __init__(self, prob):
        self.p = prob
        self.n_owned, self.n_ghost, self.n_local = prob.n_owned, prob.n_ghost, prob.n_local
        self.C = prob.n_cells
        self.ncd, self.ncd_p = _u32(prob.num_cell_dofs)
        self.ids, self.ids_p = _u32(prob.cell_local_ids)
        self.row_ids, self.row_ids_p = _u32(prob.row_ids)
        self.row_sizes, self.row_sizes_p = _u32(prob.row_sizes)
        self.row_offsets, self.row_offsets_p = _u32(prob.row_offsets)
        self.col_ids, self.col_ids_p = _u32(prob.col_ids)
        self.col_vals = np.ascontiguousarray(prob.col_vals, dtype=np.float64)
        self.inhom = np.ascontiguousarray(prob.inhom, dtype=np.float64)
        self.h_cell = np.ascontiguousarray(prob.h_cell, dtype=np.float64)
        self.nonlocal_ = prob.has_nonlocal or (prob.num_cell_proj is not None)
        if self.nonlocal_:
            self.ncp, self.ncp_p = _u32(prob.num_cell_proj)
            self.pids, self.pids_p = _u32(prob.cell_proj_local_ids)
            self.cell_c = np.ascontiguousarray(prob.cell_c, dtype=np.float64)
            self.proj_v = np.ascontiguousarray(prob.proj_v, dtype=np.float64)
            self.n_proj_local = prob.proj_halo.n_local
        self._xcell = None

    # -- constraints (a3 / a7) --
    def p2c(self, X, inhom=None):
        """inhom: inhomogeneities of another ConstraintsLocal on the same rows (the Poisson X basis manager)"""
        B = X.shape[1]
        ih = self.inhom if inhom is None else np.ascontiguousarray(inhom, dtype=np.float64)
        lib().orc_p2c(_f64(X), C.c_uint32(B), C.c_uint32(len(self.row_ids)), self.row_ids_p, self.row_sizes_p,
                      self.row_offsets_p, self.col_ids_p, _f64(self.col_vals), _f64(ih))

    def c2p(self, Y):
        B = Y.shape[1]
        lib().orc_c2p(_f64(Y), C.c_uint32(B), C.c_uint32(len(self.row_ids)), self.row_ids_p, self.row_sizes_p,
                      self.row_offsets_p, self.col_ids_p, _f64(self.col_vals))

    def xcell(self, B):
        S = int(self.ncd.sum(dtype=np.int64))
        if self._xcell is None or self._xcell.shape != (S, B):
            self._xcell = np.zeros((S, B))
        return self._xcell

    def loop_a(self, X, CX, use_nonlocal=True):
        B = X.shape[1]
        xc = self.xcell(B)
        if self.nonlocal_ and use_nonlocal:
            lib().orc_hx_loop_a(_f64(X), C.c_uint32(B), C.c_uint32(self.C), self.ncd_p, self.ids_p, _f64(xc),
                                self.ncp_p, self.pids_p, _f64(self.cell_c), _f64(CX))
        else:
            lib().orc_hx_loop_a(_f64(X), C.c_uint32(B), C.c_uint32(self.C), self.ncd_p, self.ids_p, _f64(xc),
                                None, None, None, None)

    def loop_b(self, Y, CX, h_cell=None, long_double=False, use_nonlocal=True):
        B = Y.shape[1]
        xc = self.xcell(B)
        h = self.h_cell if h_cell is None else h_cell
        if self.nonlocal_ and use_nonlocal:
            lib().orc_hx_loop_b(_f64(xc), _f64(Y), C.c_uint32(B), C.c_uint32(self.C), self.ncd_p, self.ids_p,
                                _f64(h), self.ncp_p, self.pids_p, _f64(self.cell_c), _f64(CX), C.c_int(long_double))
        else:
            lib().orc_hx_loop_b(_f64(xc), _f64(Y), C.c_uint32(B), C.c_uint32(self.C), self.ncd_p, self.ids_p,
                                _f64(h), None, None, None, None, C.c_int(long_double))


def _exchange_update(ranks_halo, Xs, n_owned_of):
    """MPICommunicatorP2P::updateGhostValues (utils/MPICommunicatorP2P.t.cpp:77-273)."""
    L = lib()
    nr = len(Xs)
    B = Xs[0].shape[1]
    send = []
    for r in range(nr):
        h = ranks_halo[r]
        ids, ids_p = _u32(h.owned_local_ids_for_targets)
        buf = np.zeros((len(ids), B))
        L.orc_pack(_f64(Xs[r]), C.c_uint32(B), ids_p, C.c_uint32(len(ids)), _f64(buf))
        send.append(buf)
    for r in range(nr):
        h = ranks_halo[r]
        recv = np.zeros((h.n_ghost, B))
        pos = 0
        for i, q in enumerate(h.ghost_proc_ids):
            q = int(q)
            cnt = int(h.ghost_ranges[2 * i + 1] - h.ghost_ranges[2 * i])
            hq = ranks_halo[q]
            t = list(hq.target_proc_ids).index(r)
            off = int(np.sum(hq.num_owned_for_target[:t], dtype=np.int64))
            assert int(hq.num_owned_for_target[t]) == cnt
            recv[pos:pos + cnt] = send[q][off:off + cnt]
            pos += cnt
        ids, ids_p = _u32(h.ghost_local_ids)
        ghost = Xs[r][n_owned_of[r]:]
        L.orc_unpack(_f64(recv), C.c_uint32(B), ids_p, C.c_uint32(len(ids)),
                     C.cast(C.c_void_p(Xs[r].ctypes.data + n_owned_of[r] * B * 8), c_f64p))


def _exchange_accumulate(ranks_halo, Ys, n_owned_of):
    """MPICommunicatorP2P::accumulateAddLocallyOwned (utils/MPICommunicatorP2P.t.cpp:278-470).
    Ghost rows are NOT cleared afterwards."""
    L = lib()
    nr = len(Ys)
    B = Ys[0].shape[1]
    send = []
    for r in range(nr):
        h = ranks_halo[r]
        ids, ids_p = _u32(h.ghost_local_ids)
        buf = np.zeros((len(ids), B))
        L.orc_pack(C.cast(C.c_void_p(Ys[r].ctypes.data + n_owned_of[r] * B * 8), c_f64p), C.c_uint32(B), ids_p,
                   C.c_uint32(len(ids)), _f64(buf))
        send.append(buf)
    for r in range(nr):
        h = ranks_halo[r]
        tot = int(np.sum(h.num_owned_for_target, dtype=np.int64))
        recv = np.zeros((tot, B))
        pos = 0
        for i, t in enumerate(h.target_proc_ids):
            t = int(t)
            cnt = int(h.num_owned_for_target[i])
            ht = ranks_halo[t]
            g = list(ht.ghost_proc_ids).index(r)
            a, b = int(ht.ghost_ranges[2 * g]), int(ht.ghost_ranges[2 * g + 1])
            assert b - a == cnt
            recv[pos:pos + cnt] = send[t][a:b]
            pos += cnt
        ids, ids_p = _u32(h.owned_local_ids_for_targets)
        L.orc_add_from_buf(_f64(recv), C.c_uint32(B), ids_p, C.c_uint32(len(ids)), _f64(Ys[r]))


class OracleWorld:
    def __init__(self, problems: Sequence):
        self.problems = list(problems)
        self.ranks = [OracleRank(p) for p in problems]
        self.nr = len(problems)
        self.halos = [p.halo for p in problems]
        self.n_owned = [p.n_owned for p in problems]
        self.has_nonlocal = any(r.nonlocal_ for r in self.ranks)
        if self.has_nonlocal:
            self.phalos = [p.proj_halo for p in problems]
            self.np_owned = [p.proj_halo.n_owned for p in problems]
        self.pool = None  # optional concurrent.futures executor: rank-local sections run one partition per thread

    def _each_rank(self, fn):
        """Run fn(i) for every rank i - concurrently when a pool is attached (the stand-in for `mpirun -n nr`:
        the C routines release the GIL), else in rank order.  Rank-local work only."""
        if self.pool is not None and self.nr > 1:
            list(self.pool.map(fn, range(self.nr)))
        else:
            for i in range(self.nr):
                fn(i)

    # a2 / a8
    def update_ghost_values(self, Xs):
        if self.nr > 1:
            _exchange_update(self.halos, Xs, self.n_owned)

    def accumulate_add_locally_owned(self, Ys):
        if self.nr > 1:
            _exchange_accumulate(self.halos, Ys, self.n_owned)

    # a10: KohnShamOperatorContextFE::apply (ksdft/KohnShamOperatorContextFE.t.cpp:1313-1443)
    def hx_apply(self, Xs, Ys, update_ghost_x=False, update_ghost_y=False, h_cells=None,
                 long_double=False, use_nonlocal=True, x_inhoms=None):
        B = Xs[0].shape[1]
        if update_ghost_x:
            self.update_ghost_values(Xs)
        nl = self.has_nonlocal and use_nonlocal
        CXs = [np.zeros((r.n_proj_local, B)) if nl else None for r in self.ranks]

        def sec_a(i):
            self.ranks[i].p2c(Xs[i], None if x_inhoms is None else x_inhoms[i])
            Ys[i][...] = 0.0
            self.ranks[i].loop_a(Xs[i], CXs[i], use_nonlocal=nl)
        self._each_rank(sec_a)
        if nl:
            # applyAllReduceOnCconjtransX + applyVOnCconjtransX
            # (basis/AtomCenterNonLocalOpContextFE.t.cpp:944-986)
            if self.nr > 1:
                _exchange_accumulate(self.phalos, CXs, self.np_owned)
                _exchange_update(self.phalos, CXs, self.np_owned)
            for r, CX in zip(self.ranks, CXs):
                lib().orc_row_scale(_f64(r.proj_v), _f64(CX), _f64(CX), C.c_uint32(B), C.c_size_t(r.n_proj_local))

        def sec_b(i):
            self.ranks[i].loop_b(Ys[i], CXs[i], h_cell=None if h_cells is None else h_cells[i],
                                 long_double=long_double, use_nonlocal=nl)
            self.ranks[i].c2p(Ys[i])
        self._each_rank(sec_b)
        self.accumulate_add_locally_owned(Ys)
        if update_ghost_y:
            self.update_ghost_values(Ys)

    # 8f rank 1: electrostatics::LaplaceOperatorContextFE::apply (electrostatics/LaplaceOperatorContextFE.t.cpp:
    # 395-470): the H.X path with the grad N_i . grad N_j cell matrices and no nonlocal part; X is filled through the
    # constraints of feBasisManagerX (inhomogeneous Dirichlet values), Y condensed through those of feBasisManagerY
    def laplace_apply(self, Xs, Ys, update_ghost_x=False, update_ghost_y=False, inhomogeneous=True):
        self.hx_apply(Xs, Ys, update_ghost_x, update_ghost_y, h_cells=[p.k_cell for p in self.problems],
                      use_nonlocal=False,
                      x_inhoms=[p.inhom_dirichlet for p in self.problems] if inhomogeneous else None)

    # linearAlgebra::PreconditionerJacobi::apply (linearAlgebra/PreconditionerJacobi.t.cpp:52-82)
    def jacobi_apply(self, Xs, Ys, update_ghost_x=False, update_ghost_y=False):
        B = Xs[0].shape[1]
        if update_ghost_x:
            self.update_ghost_values(Xs)
        for r, p, X, Y in zip(self.ranks, self.problems, Xs, Ys):
            d = np.ascontiguousarray(1.0 / p.k_diag)
            lib().orc_row_scale(_f64(d), _f64(X), _f64(Y), C.c_uint32(B), C.c_size_t(r.n_local))
        if update_ghost_y:
            self.update_ghost_values(Ys)

    def col_dots(self, Us, Vs):
        """MultiVector dot over the owned rows, summed over ranks (linearAlgebra/MultiVector.t.cpp:805-870)"""
        B = Us[0].shape[1]
        out = np.zeros(B)
        for n, U, V in zip(self.n_owned, Us, Vs):
            out += np.einsum("ij,ij->j", U[:n], V[:n])
        return out

    # linearAlgebra::CGLinearSolver::solve (linearAlgebra/CGLinearSolver.t.cpp:68-300)
    def cg_solve(self, apply_A, apply_PC, bs, xs, max_iter, abs_tol, rel_tol, div_tol):
        """bs, xs: per-rank [n_local, B]; xs holds the initial guess and receives xConverged.
        Returns (iterations, error code [0 success, 1 failed to converge, 2 divergence, 4 other], residual norms)."""
        B = bs[0].shape[1]
        bnorm = np.sqrt(self.col_dots(bs, bs))
        xconv = [x.copy() for x in xs]
        r = [np.zeros_like(b) for b in bs]
        w = [np.zeros_like(b) for b in bs]
        z = [np.zeros_like(b) for b in bs]
        pd = [np.zeros_like(b) for b in bs]
        converged = np.zeros(B, bool)
        err, diverged = 4, False
        rnorm = np.zeros(B)
        it = 0
        while it <= max_iter:
            if it == 0:
                apply_A(xs, w, True, True)
                for i, n in enumerate(self.n_owned):
                    r[i][:n] = 1.0 * bs[i][:n] + (-1.0) * w[i][:n]
                apply_PC(r, z, False, False)
                for i in range(self.nr):
                    pd[i][...] = z[i]
            else:
                apply_A(pd, w, True, True)
                zdotr = self.col_dots(z, r)
                pdotw = self.col_dots(pd, w)
                alpha = zdotr / pdotw
                for i, n in enumerate(self.n_owned):
                    xs[i][:n] = 1.0 * xs[i][:n] + alpha[None, :] * pd[i][:n]
                    r[i][:n] = 1.0 * r[i][:n] + (-alpha)[None, :] * w[i][:n]
                apply_PC(r, z, False, False)
                beta = self.col_dots(z, r) / zdotr
                for i, n in enumerate(self.n_owned):
                    pd[i][:n] = 1.0 * z[i][:n] + beta[None, :] * pd[i][:n]
            rnorm = np.sqrt(self.col_dots(r, r))
            for j in range(B):
                if rnorm[j] < max(abs_tol, bnorm[j] * rel_tol) and not converged[j]:
                    err = 0
                    converged[j] = True
                    for i, n in enumerate(self.n_owned):
                        xconv[i][:n, j] = xs[i][:n, j]
                if rnorm[j] > div_tol and not diverged:
                    err, diverged = 2, True
            if diverged or converged.all():
                break
            it += 1
        for i in range(self.nr):
            xs[i][...] = xconv[i]
        if it > max_iter:
            err = 1
        return it, err, rnorm

    # a11 / a12 (mass-lumped): diag row-scale + atom-block enrichment GEMM
    def diag_apply(self, Xs, Ys, diags, enr_blocks, update_ghost_x=False, update_ghost_y=False,
                   variant="oefe_atomblock"):
        """variant 'cfe'  : basis/CFEOverlapInverseOpContextGLL.t.cpp:529-558
           variant 'oefe_atomblock': basis/OEFEAtomBlockOverlapInvOpContextGLL.t.cpp:953-1108 and the
             mass-lumped atom-block branch of basis/OrthoEFEOverlapOperatorContext.t.cpp:2093-2235
             (the latter forces both ghost flags to false — do that at the call site)."""
        B = Xs[0].shape[1]
        if update_ghost_x:
            self.update_ghost_values(Xs)

        def sec(i):
            r, X, Y = self.ranks[i], Xs[i], Ys[i]
            r.p2c(X)
            Y[...] = 0.0
            d = np.ascontiguousarray(diags[i], dtype=np.float64)
            lib().orc_row_scale(_f64(d), _f64(X), _f64(Y), C.c_uint32(B), C.c_size_t(r.n_local))
            if variant == "oefe_atomblock":
                ncl = r.p.n_owned_classical
                nE = r.n_owned - ncl
                if nE:
                    xe = np.ascontiguousarray(X[ncl:ncl + nE])
                    ye = np.zeros((nE, B))
                    blk = np.ascontiguousarray(enr_blocks[i], dtype=np.float64)
                    lib().orc_enr_block_apply(_f64(xe), _f64(ye), C.c_uint32(B), C.c_uint32(nE), _f64(blk))
                    Y[ncl:ncl + nE] = ye
        self._each_rank(sec)
        if variant == "oefe_atomblock":
            self.update_ghost_values(Ys)
        self._each_rank(lambda i: self.ranks[i].c2p(Ys[i]))
        if update_ghost_y:
            self.update_ghost_values(Ys)

    def minv_apply_global_enrichment(self, Xs, Ys, block_global, update_ghost_x=False, update_ghost_y=False):
        """OrthoEFEOverlapInverseOpContextGLL::apply (basis/OrthoEFEOverlapInverseOpContextGLL.t.cpp:1182-1282): diag_inv on
        every local row, then one dense block over ALL enrichment functions of the system - the owned enrichment rows of every
        rank placed at their global offset, summed over the ranks (MPI_Allreduce, :1228-1234), times block^T (gemm 'N','T'),
        own rows kept; ghost update, child->parent."""
        B = Xs[0].shape[1]
        nEs = [r.n_owned - r.p.n_owned_classical for r in self.ranks]
        offs = np.concatenate(([0], np.cumsum(nEs)))
        nEg = int(offs[-1])
        blk = np.asarray(block_global, dtype=np.float64).reshape(nEg, nEg, order="F")   # column-major like the reference
        if update_ghost_x:
            self.update_ghost_values(Xs)
        xg = np.zeros((nEg, B))
        for i, r in enumerate(self.ranks):
            r.p2c(Xs[i])
            Ys[i][...] = 0.0
            d = np.ascontiguousarray(r.p.diag_inv, dtype=np.float64)
            lib().orc_row_scale(_f64(d), _f64(Xs[i]), _f64(Ys[i]), C.c_uint32(B), C.c_size_t(r.n_local))
            ncl = r.p.n_owned_classical
            xg[offs[i]:offs[i + 1]] += Xs[i][ncl:ncl + nEs[i]]          # the all-reduce
        for i, r in enumerate(self.ranks):
            ncl = r.p.n_owned_classical
            for j in range(nEs[i]):                                    # Y[j,v] = sum_k block[j,k] X[k,v], k ascending
                acc = np.zeros(B)
                for k in range(nEg):
                    acc += xg[k] * blk[offs[i] + j, k]
                Ys[i][ncl + j] = acc
        self.update_ghost_values(Ys)
        self._each_rank(lambda i: self.ranks[i].c2p(Ys[i]))
        if update_ghost_y:
            self.update_ghost_values(Ys)

    def minv_apply(self, Xs, Ys, update_ghost_x=False, update_ghost_y=False, variant="oefe_atomblock"):
        self.diag_apply(Xs, Ys, [p.diag_inv for p in self.problems], [p.enr_block_inv for p in self.problems],
                        update_ghost_x, update_ghost_y, variant)

    def m_apply(self, Xs, Ys, update_ghost_x=False, update_ghost_y=False, variant="oefe_atomblock"):
        # OrthoEFEOverlapOperatorContext.t.cpp:2095-2096: mass-lumped path ignores the flags
        self.diag_apply(Xs, Ys, [p.diag for p in self.problems], [p.enr_block for p in self.problems],
                        False, False, variant)

    # a13: linearAlgebra/ChebyshevFilter.t.cpp:39-134
    def chebyshev_filter(self, Xs, degree, a0, a, b, minv_variant="oefe_atomblock"):
        """Returns the filtered block per rank (and overwrites Xs with it, as the reference's final
        `eigenSubspaceGuess = filteredSubspace` does)."""
        L = lib()
        e = 0.5 * (b - a)
        c = 0.5 * (b + a)
        sigma = e / (a0 - c)
        sigma1 = sigma
        gamma = 2.0 / sigma1
        Xs = list(Xs)
        Fs = [np.zeros_like(X) for X in Xs]
        s1 = [np.zeros_like(X) for X in Xs]
        s2 = [np.zeros_like(X) for X in Xs]
        B = Xs[0].shape[1]
        nown = [n * B for n in self.n_owned]
        self.hx_apply(Xs, s1, True, False)
        self.minv_apply(s1, s2, False, False, minv_variant)

        def first(i):
            L.orc_axpby(C.c_size_t(nown[i]), C.c_double(sigma1 / e), _f64(s2[i]), C.c_double(-sigma1 / e * c),
                        _f64(Xs[i]), _f64(Fs[i]))
        self._each_rank(first)
        for _deg in range(2, degree + 1):
            sigma2 = 1.0 / (gamma - sigma)
            self.hx_apply(Fs, s1, True, False)
            self.minv_apply(s1, s2, False, False, minv_variant)

            def rec(i, Xs=Xs, Fs=Fs, sigma=sigma, sigma2=sigma2):
                L.orc_axpby(C.c_size_t(nown[i]), C.c_double(2.0 * sigma2 / e), _f64(s2[i]),
                            C.c_double(-2.0 * sigma2 / e * c), _f64(Fs[i]), _f64(s1[i]))
                L.orc_axpby(C.c_size_t(nown[i]), C.c_double(1.0), _f64(s1[i]), C.c_double(-sigma * sigma2),
                            _f64(Xs[i]), _f64(Xs[i]))
            self._each_rank(rec)
            Xs, Fs = Fs, Xs
            sigma = sigma2
        return Fs

    # a13: ResidualChebyshevFilterGEP, linearAlgebra/ChebyshevFilter.t.cpp:242-445
    def residual_chebyshev_filter(self, Xs, eigenvalues, degree, a0, a, b, minv_variant="oefe_atomblock"):
        L = lib()
        B = Xs[0].shape[1]
        e = 0.5 * (b - a)
        c = 0.5 * (b + a)
        sigma = e / (a0 - c)
        sigma1 = sigma
        gamma = 2.0 / sigma1
        nr = self.nr
        Ys = [np.zeros_like(X) for X in Xs]
        s1 = [np.zeros_like(X) for X in Xs]
        s2 = [np.zeros_like(X) for X in Xs]
        s3 = [np.zeros_like(X) for X in Xs]
        Res = [np.zeros_like(X) for X in Xs]
        ones = np.ones(B)
        ev = np.ascontiguousarray(eigenvalues, dtype=np.float64).copy()
        ev1 = np.ones(B)
        ev2 = ev.copy()
        alpha1, alpha2 = sigma1 / e, -c
        self.m_apply(Xs, Ys, True, False, minv_variant)
        self.hx_apply(Xs, s3, True, False)
        for i in range(nr):
            L.orc_axpby_blocked(C.c_size_t(self.n_owned[i]), C.c_uint32(B), C.c_double(1.0), _f64(ones), _f64(s3[i]),
                                C.c_double(-1.0), _f64(ev), _f64(Ys[i]), _f64(Ys[i]))
        ResNew = [Y.copy() for Y in Ys]
        ev2[:] = alpha1 * alpha2
        L.orc_axpby_blocked(C.c_size_t(1), C.c_uint32(B), C.c_double(1.0), _f64(ones), _f64(ev2),
                            C.c_double(alpha1), _f64(ev), _f64(ev1), _f64(ev2))
        for i in range(nr):
            L.orc_ascale(C.c_size_t(self.n_owned[i] * B), C.c_double(alpha1), _f64(ResNew[i]), _f64(ResNew[i]))
        for _deg in range(2, degree + 1):
            sigma2 = 1.0 / (gamma - sigma)
            alpha1, alpha2 = 2.0 * sigma2 / e, -(sigma * sigma2)
            self.minv_apply(ResNew, s1, True, False, minv_variant)
            self.hx_apply(s1, s2, False, False)
            for i in range(nr):
                n = C.c_size_t(self.n_owned[i] * B)
                L.orc_axpby(n, C.c_double(alpha1), _f64(s2[i]), C.c_double(-c * alpha1), _f64(ResNew[i]), _f64(s1[i]))
                L.orc_axpby(n, C.c_double(1.0), _f64(s1[i]), C.c_double(alpha2), _f64(Res[i]), _f64(Res[i]))
                L.orc_axpby_blocked(C.c_size_t(self.n_owned[i]), C.c_uint32(B), C.c_double(1.0), _f64(ones),
                                    _f64(Res[i]), C.c_double(alpha1), _f64(ev2), _f64(Ys[i]), _f64(Res[i]))
            L.orc_axpby(C.c_size_t(B), C.c_double(-c * alpha1), _f64(ev2), C.c_double(alpha2), _f64(ev1), _f64(ev1))
            L.orc_axpby_blocked(C.c_size_t(1), C.c_uint32(B), C.c_double(1.0), _f64(ones), _f64(ev1),
                                C.c_double(alpha1), _f64(ev), _f64(ev2), _f64(ev1))
            ResNew, Res = Res, ResNew
            ev1, ev2 = ev2, ev1
            sigma = sigma2
        self.minv_apply(ResNew, Res, True, True, minv_variant)
        for i in range(nr):
            L.orc_axpby_blocked(C.c_size_t(self.n_owned[i]), C.c_uint32(B), C.c_double(1.0), _f64(ones), _f64(Res[i]),
                                C.c_double(1.0), _f64(ev2), _f64(Xs[i]), _f64(Ys[i]))
        return Ys

    # a15: computeXTransOpX, linearAlgebra/RayleighRitzEigenSolver.t.cpp:685-844
    def xtopx(self, Xs, op, batch):
        """Returns the B x B matrix S (numpy, S[row, col]) with only the lower triangle filled,
        exactly the entries the reference writes into the ScaLAPACK matrix."""
        L = lib()
        B = Xs[0].shape[1]
        S = np.zeros((B, B))
        for j0 in range(0, B, batch):
            b = min(batch, B - j0)
            Xin = [np.ascontiguousarray(X[:, j0:j0 + b]) for X in Xs]
            Xout = [np.zeros_like(x) for x in Xin]
            op(Xin, Xout, True, False)
            blk = np.zeros(((B - j0) * b,))
            for i in range(self.nr):
                part = np.zeros(((B - j0) * b,))
                L.orc_gram_block(_f64(Xs[i]), C.c_uint32(B), C.c_uint32(j0), _f64(Xout[i]), C.c_uint32(b),
                                 C.c_size_t(self.n_owned[i]), _f64(part))
                blk += part  # MPI_Allreduce(SUM)
            blk = blk.reshape(b, B - j0)  # blk[i, j] = SBlock[j + i*(B-j0)]
            for i in range(b):
                for j in range(j0 + i, B):
                    S[j, i + j0] = blk[i, j - j0]
            for X, xin in zip(Xs, Xin):
                X[:, j0:j0 + b] = xin  # the reference copies the (possibly modified) batch back
        return S

    # a16
    def subspace_rotation(self, Xs, Q, transpose, lower_tri, dof_block=20000, vec_block=2000):
        B = Xs[0].shape[1]
        Qc = np.asfortranarray(Q, dtype=np.float64)
        for i, X in enumerate(Xs):
            xo = np.ascontiguousarray(X[:self.n_owned[i]])
            lib().orc_subspace_rotation(_f64(xo), C.c_size_t(self.n_owned[i]), C.c_uint32(B),
                                        Qc.ctypes.data_as(c_f64p), C.c_uint32(dof_block), C.c_uint32(vec_block),
                                        C.c_int(int(transpose)), C.c_int(int(lower_tri)))
            X[:self.n_owned[i]] = xo

    # a18 (local part + allreduce): MultiVector::l2Norms
    def l2_norms(self, Xs):
        B = Xs[0].shape[1]
        tot = np.zeros(B)
        for i, X in enumerate(Xs):
            part = np.zeros(B)
            lib().orc_col_sumsq(_f64(X), C.c_uint32(B), C.c_size_t(self.n_owned[i]), _f64(part))
            tot += part
        return np.sqrt(tot)


# 8f rank 2: FEBasisOperations::computeFEMatrices (basis/FEBasisOperations.t.cpp:41-427, 2210-2243)
def compute_fe_matrices(num_cell_dofs, num_cell_quad, basis, jxw, f, zero_stride):
    """Cell matrices C_c[i, j] = sum_q N_c[q, i] f[q] JxW[q] N_c[q, j], concatenated (S2 doubles).  basis: per cell
    nq_c x n_c with the DoF index fastest (one shared matrix when zero_stride)."""
    ncd, ncd_p = _u32(num_cell_dofs)
    ncq, ncq_p = _u32(num_cell_quad)
    basis, jxw, f = (np.ascontiguousarray(a, dtype=np.float64) for a in (basis, jxw, f))
    out = np.zeros(int(np.sum(ncd.astype(np.int64) ** 2)))
    lib().orc_compute_fe_matrices(C.c_uint32(len(ncd)), ncd_p, ncq_p, _f64(basis), C.c_int(int(zero_stride)), _f64(jxw),
                                  _f64(f), _f64(out))
    return out


# 8f rank 3: DensityCalculator::computeRho (ksdft/DensityCalculator.t.cpp:283-437)
def compute_rho(prob, num_cell_quad, basis, zero_stride, X, occupation, batch):
    """rho[q] = sum_i 2 occ_i |psi_i(q)|^2 at the quadrature points of the rank's cells; X [n_local, B] as given."""
    ncd, ncd_p = _u32(prob.num_cell_dofs)
    ncq, ncq_p = _u32(num_cell_quad)
    ids, ids_p = _u32(prob.cell_local_ids)
    basis, X, occ = (np.ascontiguousarray(a, dtype=np.float64) for a in (basis, X, occupation))
    rho = np.zeros(int(np.sum(ncq.astype(np.int64))))
    lib().orc_compute_rho(C.c_uint32(len(ncd)), ncd_p, ncq_p, ids_p, _f64(basis), C.c_int(int(zero_stride)), _f64(X),
                          C.c_uint32(X.shape[1]), C.c_uint32(batch), _f64(occ), _f64(rho))
    return rho


# ---------------------------------------------------------------------------
# integer work the CUDA library derives at plan creation (bit-exact parity targets)
# ---------------------------------------------------------------------------
def cell_colouring(prob, shared_threshold: int = 8):
    """Greedy colouring of the cell-DoF graph: cells in ascending order take the lowest colour not used
    by an earlier cell that shares a local row with them; rows touched by more than `shared_threshold`
    cells (enrichment DoFs) do not constrain the colouring (they are reduced separately).  This is the
    specification of the deterministic scatter that replaces the reference's sequential
    addCellWiseDataToFieldData (basis/FECellWiseDataOperations.t.cpp:87-153)."""
    ids = prob.cell_local_ids.astype(np.int64)
    ncd = prob.num_cell_dofs.astype(np.int64)
    off = np.concatenate(([0], np.cumsum(ncd)))
    inc = np.bincount(ids, minlength=prob.n_local)
    used = [0] * prob.n_local
    colour = np.zeros(prob.n_cells, np.uint32)
    for c in range(prob.n_cells):
        rows = [int(r) for r in ids[off[c]:off[c + 1]] if inc[r] <= shared_threshold]
        mask = 0
        for r in rows:
            mask |= used[r]
        col = 0
        while mask & (1 << col):
            col += 1
        colour[c] = col
        for r in rows:
            used[r] |= (1 << col)
    return int(colour.max()) + 1 if prob.n_cells else 0, colour


def processing_order(prob, delay: int, shared_threshold: int = 8):
    """Specification of the ordered scatter's processing order (delay-D list schedule) and of the per-cell
    predecessor lists.  Cells are swept in the caller's order - the order of the reference's sequential
    scatter loop, basis/FECellWiseDataOperations.t.cpp:87-153 - but a cell is placed only when every
    already-placed neighbour (a cell sharing a non-shared row) sits at least `delay` positions back; if no
    cell qualifies, the one that qualifies soonest (ties: lowest sweep key) is taken.
    On a multi-rank partition the cells that read ghost rows are all keyed at a quarter of the sweep (ahead of
    the interior cell with that index, among themselves in caller order): the halo exchange runs while the
    interior cells before and after them are contracted.
    Returns (order[C], wait_off[C+1], wait_list): position -> cell, and for each position the sorted
    positions of the immediately preceding toucher of each of the cell's non-shared rows."""
    import heapq
    ids = prob.cell_local_ids.astype(np.int64)
    ncd = prob.num_cell_dofs.astype(np.int64)
    off = np.concatenate(([0], np.cumsum(ncd)))
    inc = np.bincount(ids, minlength=prob.n_local)
    C = prob.n_cells
    rows_of = [[int(r) for r in ids[off[c]:off[c + 1]] if inc[r] <= shared_threshold] for c in range(C)]
    multi = getattr(prob, "nranks", 1) > 1
    boundary = [multi and bool((ids[off[c]:off[c + 1]] >= prob.n_owned).any()) for c in range(C)]
    k0 = C // 4
    key_of = [((k0, 0, c) if boundary[c] else (c, 1, c)) for c in range(C)]
    cells_of = {}
    for c in range(C):
        for r in rows_of[c]:
            cells_of.setdefault(r, []).append(c)
    ready = [0] * C
    placed = [False] * C
    eligible = [key_of[c] for c in range(C)]
    heapq.heapify(eligible)
    waiting = []
    order = []
    for t in range(C):
        while waiting and waiting[0][0] <= t:
            key, kx = heapq.heappop(waiting)
            x = kx[2]
            if not placed[x]:
                if ready[x] <= t:
                    heapq.heappush(eligible, kx)
                elif ready[x] != key:
                    heapq.heappush(waiting, (ready[x], kx))
        c = None
        while eligible:
            kx = heapq.heappop(eligible)
            x = kx[2]
            if placed[x]:
                continue
            if ready[x] > t:
                heapq.heappush(waiting, (ready[x], kx))
                continue
            c = x
            break
        while c is None:
            key, kx = heapq.heappop(waiting)
            x = kx[2]
            if placed[x]:
                continue
            if ready[x] != key:
                heapq.heappush(waiting, (ready[x], kx))
                continue
            c = x
        placed[c] = True
        order.append(c)
        for r in rows_of[c]:
            for nb in cells_of[r]:
                if not placed[nb]:
                    ready[nb] = t + delay
    last = {}
    wait_off, wait_list = [0], []
    for w, c in enumerate(order):
        pr = sorted({last[r] for r in rows_of[c] if r in last})
        wait_list.extend(pr)
        wait_off.append(len(wait_list))
        for r in rows_of[c]:
            last[r] = w
    return np.array(order, np.uint32), np.array(wait_off, np.uint32), np.array(wait_list, np.uint32)


def c2p_transpose(prob):
    """Parent-side view of the constraint CSR: parents ascending, entries in the reference's (row, entry)
    order — the order in which basis/ConstraintsInternal.cpp:110-170 adds into each parent."""
    par = {}
    for i in range(len(prob.row_ids)):
        o = int(prob.row_offsets[i])
        for j in range(int(prob.row_sizes[i])):
            par.setdefault(int(prob.col_ids[o + j]), []).append((int(prob.row_ids[i]), float(prob.col_vals[o + j])))
    ids = np.array(sorted(par), np.uint32)
    off = [0]
    ch, w = [], []
    for p_ in ids:
        for r, v in par[int(p_)]:
            ch.append(r)
            w.append(v)
        off.append(len(ch))
    return ids, np.array(off, np.uint32), np.array(ch, np.uint32), np.array(w, np.float64)
