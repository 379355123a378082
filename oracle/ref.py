"""ctypes access to oracle/_ref/libdftefe_ref.so — the REFERENCE's own sources compiled by
oracle/Makefile (`make ref`).  TEST INFRASTRUCTURE ONLY.  `available()` is False when the
library has not been built (it needs /root/reference at build time; the built .so travels)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libdftefe_ref.so")
_LIB = None

c_u32p = C.POINTER(C.c_uint32)
c_f64p = C.POINTER(C.c_double)
APPLY_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, c_f64p, c_f64p, C.c_uint32, C.c_uint32, C.c_int, C.c_int)


def build() -> bool:
    """Build _ref if the reference tree is present; return availability."""
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])
    return available()


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_PATH)
    return _LIB


def f64(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_f64p)


def u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(c_u32p)


def p2c(X, prob):
    r, rp = u32(prob.row_ids); s, sp = u32(prob.row_sizes); o, op = u32(prob.row_offsets); c, cp = u32(prob.col_ids)
    v = np.ascontiguousarray(prob.col_vals); ih = np.ascontiguousarray(prob.inhom)
    lib().ref_p2c(f64(X), C.c_uint32(X.shape[0]), C.c_uint32(X.shape[1]), C.c_uint32(len(r)), rp, sp, op, cp,
                  C.c_uint32(len(c)), f64(v), f64(ih))


def c2p(Y, prob):
    r, rp = u32(prob.row_ids); s, sp = u32(prob.row_sizes); o, op = u32(prob.row_offsets); c, cp = u32(prob.col_ids)
    v = np.ascontiguousarray(prob.col_vals)
    lib().ref_c2p(f64(Y), C.c_uint32(Y.shape[0]), C.c_uint32(Y.shape[1]), C.c_uint32(len(r)), rp, sp, op, cp,
                  C.c_uint32(len(c)), f64(v))


def cell_gemm_batched(xcell, h_cell, ncd, B):
    """The gemmStridedVarBatched call of KohnShamOperatorContextFE.t.cpp:1155-1175 with the sizes of
    storeSizes (:713-760), all cells in one batch.  Returns yCell [S, B]."""
    n = np.ascontiguousarray(ncd, dtype=np.uint32)
    nm = len(n)
    m = np.full(nm, B, np.uint32)
    sa = (m * n).astype(np.uint32); sb = (n * n).astype(np.uint32); sc = sa.copy()
    ta = (C.c_char * nm)(*([b"N"] * nm)); tb = (C.c_char * nm)(*([b"N"] * nm))
    y = np.zeros_like(xcell)
    p = lambda a: a.ctypes.data_as(c_u32p)
    lib().ref_gemm_strided_var_batched(C.c_uint32(nm), ta, tb, p(sa), p(sb), p(sc), p(m), p(n), p(n),
                                       C.c_double(1.0), f64(xcell), p(m), f64(h_cell), p(n), C.c_double(0.0),
                                       f64(y), p(m))
    return y


def compute_fe_matrices(num_cell_dofs, num_cell_quad, basis, jxw, f, zero_stride, cell_block):
    """FEBasisOperations::computeFEMatrices(IDENTITY, MULT, MULT, IDENTITY, f) assembled from the reference's own
    hadamardProduct / scaleStridedVarBatched / gemmStridedVarBatched (ref_compute_fe_matrices in ref_shim.cpp)."""
    ncd = np.ascontiguousarray(num_cell_dofs, dtype=np.uint32)
    ncq = np.ascontiguousarray(num_cell_quad, dtype=np.uint32)
    basis, jxw, f = (np.ascontiguousarray(a, dtype=np.float64) for a in (basis, jxw, f))
    out = np.zeros(int(np.sum(ncd.astype(np.int64) ** 2)))
    lib().ref_compute_fe_matrices(C.c_uint32(len(ncd)), ncd.ctypes.data_as(c_u32p), ncq.ctypes.data_as(c_u32p), f64(basis),
                                  C.c_int(int(zero_stride)), f64(jxw), f64(f), C.c_uint32(cell_block), f64(out))
    return out


def interpolate(prob, num_cell_quad, basis, zero_stride, X):
    """FEBasisOperations::interpolate assembled from the reference's own gather (copyFieldToCellWiseData) and
    gemmStridedVarBatched with the operand conventions of basis/FEBasisOperations.t.cpp:1166-1262.  Returns psiQuad
    (concatenated per cell, each nq_c x B with the vector index fastest)."""
    B = X.shape[1]
    n = np.ascontiguousarray(prob.num_cell_dofs, dtype=np.uint32)
    nq = np.ascontiguousarray(num_cell_quad, dtype=np.uint32)
    nm = len(n)
    xcell = gather(np.ascontiguousarray(X), prob)
    m = np.full(nm, B, np.uint32)
    sa = (m * n).astype(np.uint32)
    sb = np.zeros(nm, np.uint32) if zero_stride else (n * nq).astype(np.uint32)
    sc = (m * nq).astype(np.uint32)
    ta = (C.c_char * nm)(*([b"N"] * nm)); tb = (C.c_char * nm)(*([b"N"] * nm))
    out = np.zeros(int(np.sum(sc.astype(np.int64))))
    basis = np.ascontiguousarray(basis, dtype=np.float64)
    p = lambda a: a.ctypes.data_as(c_u32p)
    lib().ref_gemm_strided_var_batched(C.c_uint32(nm), ta, tb, p(sa), p(sb), p(sc), p(m), p(nq), p(n), C.c_double(1.0),
                                       f64(np.ascontiguousarray(xcell)), p(m), f64(basis), p(n), C.c_double(0.0), f64(out),
                                       p(m))
    return out


def gather(X, prob):
    """FECellWiseDataOperations::copyFieldToCellWiseData (basis/FECellWiseDataOperations.t.cpp:58-86), reference body."""
    B = X.shape[1]
    ids, ip = u32(prob.cell_local_ids); ncd, np_ = u32(prob.num_cell_dofs)
    out = np.zeros((int(ncd.sum(dtype=np.int64)), B))
    lib().ref_gather(f64(X), C.c_uint32(B), ip, np_, C.c_uint32(len(ncd)), f64(out))
    return out


def scatter_add(ycell, prob, Y):
    """FECellWiseDataOperations::addCellWiseDataToFieldData (:87-153), reference body; adds into Y in place."""
    B = Y.shape[1]
    ids, ip = u32(prob.cell_local_ids); ncd, np_ = u32(prob.num_cell_dofs)
    lib().ref_scatter_add(f64(ycell), C.c_uint32(B), ip, np_, C.c_uint32(len(ncd)), f64(Y))


def hx_apply_serial(prob, X, cell_block=3, use_nonlocal=True, out=None):
    """KohnShamOperatorContextFE::apply on one rank assembled from reference-compiled routines only
    (ref_shim_cellwise.cpp: ref_hx_apply_serial).  X is modified in place like the reference; returns Y (`out` when
    given: the routine overwrites it, as the reference's apply does with its Y)."""
    assert prob.nranks == 1
    B = X.shape[1]
    Y = np.zeros_like(X) if out is None else out
    assert Y.shape == X.shape and Y.dtype == np.float64 and Y.flags["C_CONTIGUOUS"]
    ids, ip = u32(prob.cell_local_ids); ncd, ncdp = u32(prob.num_cell_dofs)
    r, rp = u32(prob.row_ids); s, sp = u32(prob.row_sizes); o, op = u32(prob.row_offsets); c, cp = u32(prob.col_ids)
    v = np.ascontiguousarray(prob.col_vals); ih = np.ascontiguousarray(prob.inhom)
    h = np.ascontiguousarray(prob.h_cell)
    nl = use_nonlocal and prob.num_cell_proj is not None
    if nl:
        ncp, ncpp = u32(prob.num_cell_proj); pids, pp = u32(prob.cell_proj_local_ids)
        cc = np.ascontiguousarray(prob.cell_c); V = np.ascontiguousarray(prob.proj_v)
        nproj = prob.proj_halo.n_local
        lib().ref_hx_apply_serial(f64(X), f64(Y), C.c_uint32(X.shape[0]), C.c_uint32(B), C.c_uint32(len(ncd)), ncdp, ip,
                                  f64(h), C.c_uint32(len(r)), rp, sp, op, cp, C.c_uint32(len(c)), f64(v), f64(ih),
                                  ncpp, pp, f64(cc), f64(V), C.c_uint32(nproj), C.c_uint32(cell_block))
    else:
        lib().ref_hx_apply_serial(f64(X), f64(Y), C.c_uint32(X.shape[0]), C.c_uint32(B), C.c_uint32(len(ncd)), ncdp, ip,
                                  f64(h), C.c_uint32(len(r)), rp, sp, op, cp, C.c_uint32(len(c)), f64(v), f64(ih),
                                  None, None, None, None, C.c_uint32(0), C.c_uint32(cell_block))
    return Y


class PartitionedApply:
    """KohnShamOperatorContextFE::apply over a PARTITIONED mesh with every rank-local routine the reference's own
    (ref_hx_phase_a / ref_hx_phase_b of ref_shim_cellwise.cpp) and the exchanges of an oracle.OracleWorld (the in-process
    stand-in for MPI): ghost update, phase A on every rank, projector all-reduce + V scaling, phase B, accumulate.  Used by
    `bench.py --impl reference` (one partition per host core, rank-local phases on a thread pool: the C routines release the
    GIL) and pinned against the oracle's hx_apply in tests/test_oracle.py."""

    def __init__(self, world, cell_block=1):
        self.W = world
        self.cell_block = cell_block
        self.args = []
        for q in world.problems:
            ids, ip = u32(q.cell_local_ids); ncd, ncdp = u32(q.num_cell_dofs)
            r, rp = u32(q.row_ids); s_, sp = u32(q.row_sizes); o, op = u32(q.row_offsets); c, cp = u32(q.col_ids)
            v = np.ascontiguousarray(q.col_vals); ih = np.ascontiguousarray(q.inhom)
            h = np.ascontiguousarray(q.h_cell)
            nl = q.num_cell_proj is not None
            d = dict(ids=(ids, ip), ncd=(ncd, ncdp), rows=(r, rp), sizes=(s_, sp), offs=(o, op), cols=(c, cp), v=v, ih=ih, h=h, nl=nl,
                     n=q.n_local, C=len(ncd), S=int(np.sum(q.num_cell_dofs.astype(np.int64))))
            if nl:
                d["ncp"] = u32(q.num_cell_proj); d["pids"] = u32(q.cell_proj_local_ids)
                d["cc"] = np.ascontiguousarray(q.cell_c); d["V"] = np.ascontiguousarray(q.proj_v); d["nproj"] = q.proj_halo.n_local
            self.args.append(d)
        self.xcell = None

    def __call__(self, Xs, Ys, update_ghost_x=False, update_ghost_y=False):
        from . import oracle as orc
        W, B = self.W, Xs[0].shape[1]
        if self.xcell is None or self.xcell[0].shape[0] != self.args[0]["S"] * B:
            self.xcell = [np.empty(d["S"] * B) for d in self.args]
            self.cx = [np.zeros((d.get("nproj", 0), B)) for d in self.args]
        if update_ghost_x:
            W.update_ghost_values(Xs)
        L = lib()

        def a(i):
            d = self.args[i]
            L.ref_hx_phase_a(f64(Xs[i]), f64(Ys[i]), C.c_uint32(d["n"]), C.c_uint32(B), C.c_uint32(d["C"]), d["ncd"][1], d["ids"][1],
                             C.c_uint32(len(d["rows"][0])), d["rows"][1], d["sizes"][1], d["offs"][1], d["cols"][1],
                             C.c_uint32(len(d["cols"][0])), f64(d["v"]), f64(d["ih"]),
                             d["ncp"][1] if d["nl"] else None, d["pids"][1] if d["nl"] else None, f64(d["cc"]) if d["nl"] else None,
                             C.c_uint32(d.get("nproj", 0)), C.c_uint32(self.cell_block), f64(self.xcell[i]),
                             f64(self.cx[i]) if d["nl"] else None)
        W._each_rank(a)
        if W.has_nonlocal:
            if W.nr > 1:
                orc._exchange_accumulate(W.phalos, self.cx, W.np_owned)
                orc._exchange_update(W.phalos, self.cx, W.np_owned)
            for d, cx in zip(self.args, self.cx):
                L.ref_scale_rows_strided(f64(d["V"]), f64(cx), C.c_uint32(B), C.c_uint32(d["nproj"]))

        def b(i):
            d = self.args[i]
            L.ref_hx_phase_b(f64(Ys[i]), C.c_uint32(d["n"]), C.c_uint32(B), C.c_uint32(d["C"]), d["ncd"][1], d["ids"][1], f64(d["h"]),
                             C.c_uint32(len(d["rows"][0])), d["rows"][1], d["sizes"][1], d["offs"][1], d["cols"][1],
                             C.c_uint32(len(d["cols"][0])), f64(d["v"]),
                             d["ncp"][1] if d["nl"] else None, d["pids"][1] if d["nl"] else None, f64(d["cc"]) if d["nl"] else None,
                             C.c_uint32(self.cell_block), f64(self.xcell[i]), f64(self.cx[i]) if d["nl"] else None)
        W._each_rank(b)
        W.accumulate_add_locally_owned(Ys)
        if update_ghost_y:
            W.update_ghost_values(Ys)


def cg_solve(apply_A, apply_PC, b, x0, max_iter, abs_tol, rel_tol, div_tol):
    """The reference's own CGLinearSolver::solve (linearAlgebra/CGLinearSolver.t.cpp) on one rank over Python
    operators: apply_X(X, Y, update_ghost_x, update_ghost_y) with [n, B] arrays.  Returns (x, isSuccess)."""
    n, B = b.shape

    def cb(_user, op_id, xp, yp, n_, B_, ugx, ugy):
        X = np.ctypeslib.as_array(xp, shape=(n_, B_))
        Y = np.ctypeslib.as_array(yp, shape=(n_, B_))
        (apply_A if op_id == 0 else apply_PC)(X, Y, bool(ugx), bool(ugy))

    c = APPLY_CB(cb)
    x = np.ascontiguousarray(x0, dtype=np.float64).copy()
    lib().ref_cg_solve.restype = C.c_int
    ok = lib().ref_cg_solve(c, None, f64(np.ascontiguousarray(b)), f64(x), C.c_uint32(n), C.c_uint32(B),
                            C.c_uint32(max_iter), C.c_double(abs_tol), C.c_double(rel_tol), C.c_double(div_tol))
    return x, bool(ok)
