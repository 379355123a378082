"""ctypes binding of the C ABI in include/hxb200.h (harness for tests and bench; the product is the .so).

Fails loudly when the library is missing or no CUDA device is usable: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HXB200_LIB") or os.path.join(_HERE, "lib", "libhxb200.so")  # HXB200_LIB: A/B build (tools/build_variants.py)

u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)

EXPORTS = [
    "hx_last_error", "hx_version", "hx_device_count", "hx_set_device", "hx_device_alloc", "hx_device_free",
    "hx_host_alloc_pinned", "hx_host_free_pinned", "hx_memcpy_h2d", "hx_memcpy_d2h", "hx_memset_zero",
    "hx_plan_create", "hx_plan_destroy", "hx_plan_synchronize", "hx_comm_unique_id", "hx_plan_attach_comm", "hx_plan_halo_transport",
    "hx_plan_set_scatter_mode", "hx_plan_get_wait_lists", "hx_plan_get_processing_order", "hx_plan_num_colours", "hx_plan_get_cell_colours", "hx_plan_get_c2p_transpose", "hx_plan_get_fusable_rows", "hx_update_ghost_values",
    "hx_accumulate_add_locally_owned", "hx_distribute_parent_to_child", "hx_distribute_child_to_parent",
    "hx_set_constrained_nodes_to_zero", "hx_plan_add_constraints", "hx_distribute_parent_to_child_set",
    "hx_distribute_child_to_parent_set", "hx_cellop_set_constraint_sets", "hx_cg_solve", "hx_cellop_create", "hx_cellop_set_matrices", "hx_cellop_set_nonlocal",
    "hx_diagop_create", "hx_diagop_create_global_enrichment", "hx_op_destroy", "hx_op_apply", "hx_op_apply_host", "hx_chebyshev_filter", "hx_chebyshev_filter_host", "hx_chebyshev_filter_host_batches", "hx_multipass_cgs", "hx_chfsi_solve_ortho", "hx_plan_global_size",
    "hx_residual_chebyshev_filter", "hx_xtopx", "hx_subspace_rotation", "hx_l2_norms", "hx_axpby",
    "hx_axpby_blocked", "hx_plan_launch_count", "hx_plan_cell_kernel_time_ms", "hx_plan_cell_kernel_sm_clock_mhz", "hx_plan_enable_kernel_timing", "hx_plan_trace", "hx_plan_trace_report",
    "hx_microbench", "hx_programmatic_launch_enabled",
    "hx_xtopx_device", "hx_subspace_rotation_device", "hx_dense_cholesky_inverse", "hx_dense_sym_eig",
    "hx_cholesky_gram_schmidt", "hx_rayleigh_ritz", "hx_chfsi_solve", "hx_eigen_residual_norms", "hx_lanczos_extreme",
    "hx_chebyshev_polynomial_degree", "hx_fe_basis_create", "hx_fe_basis_destroy", "hx_compute_fe_matrices", "hx_compute_rho",
    "hx_cellop_set_matrix_sharing", "hx_cellop_num_unique_matrices", "hx_cellop_assemble_matrices",
]


class HaloDesc(C.Structure):
    _fields_ = [("n_owned", C.c_uint32), ("n_ghost", C.c_uint32), ("n_ghost_procs", C.c_uint32),
                ("ghost_proc_ids", u32p), ("ghost_ranges", u32p), ("ghost_local_ids", u32p),
                ("n_target_procs", C.c_uint32), ("target_proc_ids", u32p), ("num_owned_for_target", u32p),
                ("owned_local_ids_for_targets", u32p)]


class MeshDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("rank", C.c_int32), ("nranks", C.c_int32), ("halo", HaloDesc),
                ("n_owned_classical", C.c_uint32), ("n_cells", C.c_uint32), ("num_cell_dofs", u32p),
                ("cell_local_ids", u32p), ("n_constraint_rows", C.c_uint32), ("row_ids", u32p), ("row_sizes", u32p),
                ("row_offsets", u32p), ("col_ids", u32p), ("col_vals", f64p), ("inhom", f64p),
                ("max_block", C.c_uint32)]


class NonlocalDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("proj_halo", HaloDesc), ("num_cell_proj", u32p),
                ("cell_proj_local_ids", u32p), ("cell_c", f64p), ("v", f64p)]


class FeBasisDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("same_basis_in_all_cells", C.c_int32), ("num_cell_quad", u32p),
                ("basis_data", f64p), ("jxw", f64p)]


class HxError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HxError(f"{LIB_PATH} not built (run python dft_efe_b200/build.py or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        _lib.hx_last_error.restype = C.c_char_p
    return _lib


def check(rc: int):
    if rc != 0:
        raise HxError(f"hxb200 error {rc}: {lib().hx_last_error().decode()}")


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(u32p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(f64p)


def _halo_desc(h, keep):
    d = HaloDesc()
    d.n_owned, d.n_ghost = h.n_owned, h.n_ghost
    d.n_ghost_procs = len(h.ghost_proc_ids)
    d.n_target_procs = len(h.target_proc_ids)
    for name in ("ghost_proc_ids", "ghost_ranges", "ghost_local_ids", "target_proc_ids", "num_owned_for_target",
                 "owned_local_ids_for_targets"):
        a, p = _u32(getattr(h, name))
        keep.append(a)
        setattr(d, name, p)
    return d


def device_count() -> int:
    n = C.c_int(0)
    check(lib().hx_device_count(C.byref(n)))
    return n.value


class DeviceBlock:
    """A block vector [n_rows, B] of doubles in device memory."""

    def __init__(self, n_rows: int, B: int, host: Optional[np.ndarray] = None):
        self.n_rows, self.B = n_rows, B
        self.ptr = C.c_void_p()
        check(lib().hx_device_alloc(C.byref(self.ptr), C.c_size_t(max(n_rows * B, 1) * 8)))
        if host is not None:
            self.upload(host)
        else:
            check(lib().hx_memset_zero(self.ptr, C.c_size_t(n_rows * B * 8)))

    @property
    def p(self):
        return C.cast(self.ptr, f64p)

    def upload(self, host):
        a, p = _f64(host)
        assert a.size == self.n_rows * self.B
        check(lib().hx_memcpy_h2d(self.ptr, p, C.c_size_t(a.size * 8)))

    def download(self) -> np.ndarray:
        out = np.empty((self.n_rows, self.B))
        check(lib().hx_memcpy_d2h(out.ctypes.data_as(f64p), self.ptr, C.c_size_t(out.size * 8)))
        return out

    def free(self):
        if self.ptr:
            lib().hx_device_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Plan:
    def __init__(self, prob, max_block: int, stream=None):
        self.prob = prob
        keep = []
        m = MeshDesc()
        m.struct_size = C.sizeof(MeshDesc)
        m.rank, m.nranks = prob.rank, prob.nranks
        m.halo = _halo_desc(prob.halo, keep)
        m.n_owned_classical = prob.n_owned_classical
        m.n_cells = prob.n_cells
        for name, src in (("num_cell_dofs", prob.num_cell_dofs), ("cell_local_ids", prob.cell_local_ids),
                          ("row_ids", prob.row_ids), ("row_sizes", prob.row_sizes), ("row_offsets", prob.row_offsets),
                          ("col_ids", prob.col_ids)):
            a, p = _u32(src)
            keep.append(a)
            setattr(m, name, p)
        m.n_constraint_rows = len(prob.row_ids)
        for name, src in (("col_vals", prob.col_vals), ("inhom", prob.inhom)):
            a, p = _f64(src)
            keep.append(a)
            setattr(m, name, p)
        m.max_block = max_block
        self.h = C.c_void_p()
        check(lib().hx_plan_create(C.byref(self.h), C.byref(m), C.c_void_p(stream) if stream else None))
        self.max_block = max_block
        self.n_local, self.n_owned = prob.n_local, prob.n_owned

    def destroy(self):
        if self.h:
            lib().hx_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def synchronize(self):
        check(lib().hx_plan_synchronize(self.h))

    def attach_comm(self, uid: bytes):
        assert len(uid) == 128
        check(lib().hx_plan_attach_comm(self.h, C.c_char_p(uid)))

    def halo_transport(self) -> str:
        t = C.c_int()
        check(lib().hx_plan_halo_transport(self.h, C.byref(t)))
        return {0: "none", 1: "nccl", 2: "nvlink-peer"}[t.value]

    def set_scatter_mode(self, mode: int):
        check(lib().hx_plan_set_scatter_mode(self.h, C.c_int(mode)))

    def processing_order(self):
        order = np.zeros(self.prob.n_cells, np.uint32)
        check(lib().hx_plan_get_processing_order(self.h, order.ctypes.data_as(u32p)))
        return order

    def wait_lists(self):
        n = C.c_uint32()
        check(lib().hx_plan_get_wait_lists(self.h, C.byref(n), None, None))
        off = np.zeros(self.prob.n_cells + 1, np.uint32)
        preds = np.zeros(max(n.value, 1), np.uint32)
        check(lib().hx_plan_get_wait_lists(self.h, C.byref(n), off.ctypes.data_as(u32p), preds.ctypes.data_as(u32p)))
        return off, preds[:n.value]

    def colours(self):
        n = C.c_uint32()
        check(lib().hx_plan_num_colours(self.h, C.byref(n)))
        col = np.zeros(self.prob.n_cells, np.uint32)
        check(lib().hx_plan_get_cell_colours(self.h, col.ctypes.data_as(u32p)))
        return n.value, col

    def c2p_transpose(self):
        n = C.c_uint32()
        check(lib().hx_plan_get_c2p_transpose(self.h, C.byref(n), None, None, None, None))
        ids = np.zeros(n.value, np.uint32)
        off = np.zeros(n.value + 1, np.uint32)
        nnz = int(self.prob.row_sizes.sum())
        ch = np.zeros(nnz, np.uint32)
        w = np.zeros(nnz)
        check(lib().hx_plan_get_c2p_transpose(self.h, C.byref(n), ids.ctypes.data_as(u32p), off.ctypes.data_as(u32p),
                                              ch.ctypes.data_as(u32p), w.ctypes.data_as(f64p)))
        return ids, off, ch, w

    def fusable_rows(self):
        a, b = C.c_uint32(), C.c_uint32()
        check(lib().hx_plan_get_fusable_rows(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def block(self, B, host=None) -> DeviceBlock:
        return DeviceBlock(self.n_local, B, host)

    def update_ghost_values(self, X: DeviceBlock):
        check(lib().hx_update_ghost_values(self.h, X.p, C.c_uint32(X.B)))

    def accumulate_add_locally_owned(self, Y: DeviceBlock):
        check(lib().hx_accumulate_add_locally_owned(self.h, Y.p, C.c_uint32(Y.B)))

    def p2c(self, X: DeviceBlock):
        check(lib().hx_distribute_parent_to_child(self.h, X.p, C.c_uint32(X.B)))

    def c2p(self, Y: DeviceBlock):
        check(lib().hx_distribute_child_to_parent(self.h, Y.p, C.c_uint32(Y.B)))

    def add_constraints(self, row_ids, row_sizes, row_offsets, col_ids, col_vals, inhom) -> int:
        """another ConstraintsLocal on the same DoF numbering; returns its set id (0 is the mesh's own)"""
        keep = []
        ptrs = []
        for a in (row_ids, row_sizes, row_offsets, col_ids):
            arr, ptr = _u32(a); keep.append(arr); ptrs.append(ptr)
        cv, cvp = _f64(col_vals); ih, ihp = _f64(inhom)
        sid = C.c_uint32()
        check(lib().hx_plan_add_constraints(self.h, C.c_uint32(len(keep[0])), ptrs[0], ptrs[1], ptrs[2], ptrs[3], cvp, ihp,
                                            C.byref(sid)))
        return sid.value

    def p2c_set(self, set_id: int, X: DeviceBlock):
        check(lib().hx_distribute_parent_to_child_set(self.h, C.c_uint32(set_id), X.p, C.c_uint32(X.B)))

    def c2p_set(self, set_id: int, Y: DeviceBlock):
        check(lib().hx_distribute_child_to_parent_set(self.h, C.c_uint32(set_id), Y.p, C.c_uint32(Y.B)))

    def l2_norms(self, X: DeviceBlock) -> np.ndarray:
        out = np.zeros(X.B)
        check(lib().hx_l2_norms(self.h, X.p, C.c_uint32(X.B), out.ctypes.data_as(f64p)))
        return out

    def subspace_rotation(self, X: DeviceBlock, Q: np.ndarray, transpose: bool, lower_tri: bool):
        Qc = np.asfortranarray(Q, dtype=np.float64)
        check(lib().hx_subspace_rotation(self.h, X.p, C.c_uint32(X.B), Qc.ctypes.data_as(f64p), C.c_int(int(transpose)),
                                         C.c_int(int(lower_tri))))

    def trace(self, on=True):
        check(lib().hx_plan_trace(self.h, C.c_int(int(on))))

    def trace_report(self) -> dict:
        import json
        buf = C.create_string_buffer(8192)
        check(lib().hx_plan_trace_report(self.h, buf, C.c_size_t(8192)))
        return json.loads(buf.value.decode())

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(lib().hx_plan_launch_count(self.h, C.byref(n)))
        return n.value

    def enable_kernel_timing(self, on=True):
        check(lib().hx_plan_enable_kernel_timing(self.h, C.c_int(int(on))))

    def cell_kernel_time_ms(self):
        ms = C.c_double()
        n = C.c_uint64()
        check(lib().hx_plan_cell_kernel_time_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def global_size(self) -> int:
        n = C.c_uint64()
        check(lib().hx_plan_global_size(self.h, C.byref(n)))
        return n.value

    def cell_kernel_sm_clock_mhz(self) -> float:
        mhz = C.c_double()
        check(lib().hx_plan_cell_kernel_sm_clock_mhz(self.h, C.byref(mhz)))
        return mhz.value


class Op:
    def __init__(self, plan: Plan):
        self.plan = plan
        self.h = C.c_void_p()

    def apply(self, X: DeviceBlock, Y: DeviceBlock, update_ghost_x=False, update_ghost_y=False):
        assert X.B == Y.B
        check(lib().hx_op_apply(self.h, X.p, Y.p, C.c_uint32(X.B), C.c_int(int(update_ghost_x)),
                                C.c_int(int(update_ghost_y))))

    def apply_host(self, Xh: np.ndarray, Yh: np.ndarray, update_ghost_x=False, update_ghost_y=False):
        assert Xh.flags["C_CONTIGUOUS"] and Yh.flags["C_CONTIGUOUS"] and Xh.dtype == np.float64
        check(lib().hx_op_apply_host(self.h, Xh.ctypes.data_as(f64p), Yh.ctypes.data_as(f64p), C.c_uint32(Xh.shape[1]),
                                     C.c_int(int(update_ghost_x)), C.c_int(int(update_ghost_y))))

    def apply_host_ptr(self, xptr, yptr, B, update_ghost_x=False, update_ghost_y=False):
        check(lib().hx_op_apply_host(self.h, C.cast(xptr, f64p), C.cast(yptr, f64p), C.c_uint32(B),
                                     C.c_int(int(update_ghost_x)), C.c_int(int(update_ghost_y))))

    def xtopx(self, X: DeviceBlock, batch: int) -> np.ndarray:
        S = np.zeros((X.B, X.B), order="F")
        check(lib().hx_xtopx(self.h, X.p, C.c_uint32(X.B), C.c_uint32(batch), S.ctypes.data_as(f64p)))
        return S

    def destroy(self):
        if self.h:
            lib().hx_op_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class CellOp(Op):
    """KohnShamOperatorContextFE-shaped operator (cell matrices + optional nonlocal projectors)."""

    def __init__(self, plan: Plan, h_cell=None, with_nonlocal=True, share_identical=False, matrices=True):
        """matrices=False: no cell matrices yet (they are assembled on the device: FeBasis.assemble_into)"""
        super().__init__(plan)
        check(lib().hx_cellop_create(plan.h, C.byref(self.h)))
        if share_identical:
            check(lib().hx_cellop_set_matrix_sharing(self.h, C.c_int(1)))
        prob = plan.prob
        if with_nonlocal and prob.num_cell_proj is not None:
            keep = []
            d = NonlocalDesc()
            d.struct_size = C.sizeof(NonlocalDesc)
            d.proj_halo = _halo_desc(prob.proj_halo, keep)
            a, d.num_cell_proj = _u32(prob.num_cell_proj); keep.append(a)
            a, d.cell_proj_local_ids = _u32(prob.cell_proj_local_ids); keep.append(a)
            a, d.cell_c = _f64(prob.cell_c); keep.append(a)
            a, d.v = _f64(prob.proj_v); keep.append(a)
            check(lib().hx_cellop_set_nonlocal(self.h, C.byref(d)))
        if matrices:
            self.set_matrices(prob.h_cell if h_cell is None else h_cell)

    def set_matrices(self, h_cell: np.ndarray):
        a, p = _f64(h_cell)
        assert a.size == self.plan.prob.S2
        check(lib().hx_cellop_set_matrices(self.h, p, C.c_int(0)))

    def num_unique_matrices(self) -> int:
        n = C.c_uint32()
        check(lib().hx_cellop_num_unique_matrices(self.h, C.byref(n)))
        return n.value

    def set_constraint_sets(self, x_set: int, y_set: int):
        check(lib().hx_cellop_set_constraint_sets(self.h, C.c_uint32(x_set), C.c_uint32(y_set)))

    def set_matrices_device(self, dev_ptr):
        check(lib().hx_cellop_set_matrices(self.h, C.cast(dev_ptr, f64p), C.c_int(1)))


DIAG_CFE, DIAG_OEFE_ATOMBLOCK, DIAG_OEFE_MASS, DIAG_JACOBI = 0, 1, 2, 3
CG_SUCCESS, CG_FAILED_TO_CONVERGE, CG_RESIDUAL_DIVERGENCE, CG_DIVISION_BY_ZERO, CG_OTHER_ERROR = 0, 1, 2, 3, 4


class DiagOp(Op):
    def __init__(self, plan: Plan, diag: np.ndarray, enr_block: Optional[np.ndarray], variant: int):
        super().__init__(plan)
        a, p = _f64(diag)
        assert a.size == plan.n_local
        if enr_block is not None and enr_block.size:
            e, ep = _f64(enr_block)
        else:
            e, ep = None, None
        check(lib().hx_diagop_create(plan.h, p, ep, C.c_int(variant), C.byref(self.h)))


class DiagOpGlobalEnrichment(Op):
    """OrthoEFEOverlapInverseOpContextGLL: diag_inv + one dense block (column-major nEg x nEg) over all enrichment functions"""

    def __init__(self, plan: Plan, diag_inv: np.ndarray, block_global: np.ndarray, n_enr_global: int, owned_offset: int):
        super().__init__(plan)
        a, p = _f64(diag_inv)
        assert a.size == plan.n_local
        e, ep = _f64(np.asarray(block_global).ravel()) if n_enr_global else (None, None)
        check(lib().hx_diagop_create_global_enrichment(plan.h, p, ep, C.c_uint32(n_enr_global), C.c_uint32(owned_offset),
                                                       C.byref(self.h)))


def chebyshev_filter(A: Op, BInv: Op, X: DeviceBlock, Y: DeviceBlock, degree, a0, a, b):
    check(lib().hx_chebyshev_filter(A.h, BInv.h, X.p, Y.p, C.c_uint32(X.B), C.c_uint32(degree), C.c_double(a0),
                                    C.c_double(a), C.c_double(b)))


def chebyshev_filter_host_ptr(A: Op, BInv: Op, xptr, yptr, B, degree, a0, a, b, write_back_x=False):
    check(lib().hx_chebyshev_filter_host(A.h, BInv.h, C.cast(xptr, f64p), C.cast(yptr, f64p), C.c_uint32(B),
                                         C.c_uint32(degree), C.c_double(a0), C.c_double(a), C.c_double(b),
                                         C.c_int(int(write_back_x))))


def chebyshev_filter_host_batches(A, BInv, xptrs, yptrs, B, degree, a0, a, b):
    """pinned HOST buffers, one pointer per column batch (n_local x B each): copies overlap the filter of the batch between"""
    n = len(xptrs)
    XP = (f64p * n)(*[C.cast(p_, f64p) for p_ in xptrs])
    YP = (f64p * n)(*[C.cast(p_, f64p) for p_ in yptrs])
    check(lib().hx_chebyshev_filter_host_batches(A.h, BInv.h, XP, YP, C.c_uint32(n), C.c_uint32(B), C.c_uint32(degree),
                                                 C.c_double(a0), C.c_double(a), C.c_double(b)))


def chebyshev_filter_host(A: Op, BInv: Op, Xh: np.ndarray, Yh: np.ndarray, degree, a0, a, b, write_back_x=True):
    assert Xh.flags["C_CONTIGUOUS"] and Yh.flags["C_CONTIGUOUS"] and Xh.dtype == np.float64 and Yh.dtype == np.float64
    chebyshev_filter_host_ptr(A, BInv, Xh.ctypes.data, Yh.ctypes.data, Xh.shape[1], degree, a0, a, b, write_back_x)


def residual_chebyshev_filter(A: Op, Bop: Op, BInv: Op, eig: np.ndarray, X: DeviceBlock, Y: DeviceBlock, degree, a0, a, b):
    e, ep = _f64(eig)
    check(lib().hx_residual_chebyshev_filter(A.h, Bop.h, BInv.h, ep, X.p, Y.p, C.c_uint32(X.B), C.c_uint32(degree),
                                             C.c_double(a0), C.c_double(a), C.c_double(b)))


def cg_solve(A: Op, PC: Op, b: DeviceBlock, x: DeviceBlock, max_iter, abs_tol, rel_tol, div_tol):
    """CGLinearSolver::solve; x holds the initial guess and receives xConverged.  Returns (iterations, status, norms)."""
    it, st = C.c_uint32(), C.c_int()
    rn = np.zeros(b.B)
    check(lib().hx_cg_solve(A.h, PC.h, b.p, x.p, C.c_uint32(b.B), C.c_uint32(max_iter), C.c_double(abs_tol),
                            C.c_double(rel_tol), C.c_double(div_tol), C.byref(it), C.byref(st), rn.ctypes.data_as(f64p)))
    return it.value, st.value, rn


# ---- the eigensolve around the path (device-resident) ----
EIG_SUCCESS, EIG_LAPACK_ERROR, EIG_LANCZOS_BETA_ZERO, EIG_LANCZOS_SUBSPACE_INSUFFICIENT = 0, 1, 2, 3
EIG_CHFSI_ORTHONORMALIZATION_ERROR, EIG_CHFSI_RAYLEIGH_RITZ_ERROR = 4, 5


class DenseMatrix(DeviceBlock):
    """B x B column-major matrix in device memory."""

    def __init__(self, B: int, host: Optional[np.ndarray] = None):
        super().__init__(B, B, None if host is None else np.asfortranarray(host).T)

    def download(self) -> np.ndarray:  # as a [row, col] numpy array
        return super().download().T.copy()


def xtopx_device(op: Op, X: DeviceBlock, batch: int) -> DenseMatrix:
    S = DenseMatrix(X.B)
    check(lib().hx_xtopx_device(op.h, X.p, C.c_uint32(X.B), C.c_uint32(batch), S.p))
    return S


def subspace_rotation_device(plan: Plan, X: DeviceBlock, Q: DenseMatrix, transpose: bool, lower_tri: bool):
    check(lib().hx_subspace_rotation_device(plan.h, X.p, C.c_uint32(X.B), Q.p, C.c_int(int(transpose)),
                                            C.c_int(int(lower_tri))))


def dense_cholesky_inverse(plan: Plan, S: DenseMatrix) -> int:
    info = C.c_int()
    check(lib().hx_dense_cholesky_inverse(plan.h, S.p, C.c_uint32(S.B), C.byref(info)))
    return info.value


def dense_sym_eig(plan: Plan, S: DenseMatrix):
    info = C.c_int()
    w = np.zeros(S.B)
    check(lib().hx_dense_sym_eig(plan.h, S.p, C.c_uint32(S.B), w.ctypes.data_as(f64p), C.byref(info)))
    return w, info.value


def cholesky_gram_schmidt(Bop: Op, X: DeviceBlock, ortho: DeviceBlock, batch: int) -> int:
    st = C.c_int()
    check(lib().hx_cholesky_gram_schmidt(Bop.h, X.p, ortho.p, C.c_uint32(X.B), C.c_uint32(batch), C.byref(st)))
    return st.value


def multipass_cgs(Bop: Op, X: DeviceBlock, ortho: DeviceBlock, batch: int, max_pass=50, shift_tol=1e-12, identity_tol=1e-12):
    """OrthonormalizationFunctions::MultipassCGS: returns (OrthonormalizationErrorCode, Cholesky passes performed)."""
    st, np_ = C.c_int(), C.c_uint32()
    check(lib().hx_multipass_cgs(Bop.h, X.p, ortho.p, C.c_uint32(X.B), C.c_uint32(batch), C.c_uint32(max_pass),
                                 C.c_double(shift_tol), C.c_double(identity_tol), C.byref(st), C.byref(np_)))
    return st.value, np_.value


def rayleigh_ritz(A: Op, X: DeviceBlock, vecs: DeviceBlock, batch: int, compute_vectors=True):
    st = C.c_int()
    w = np.zeros(X.B)
    check(lib().hx_rayleigh_ritz(A.h, X.p, vecs.p, C.c_uint32(X.B), C.c_uint32(batch), w.ctypes.data_as(f64p),
                                 C.c_int(int(compute_vectors)), C.byref(st)))
    return w, st.value


def chfsi_solve(A: Op, Bop: Op, BInv: Op, guess: DeviceBlock, vecs: DeviceBlock, batch, degree, a0, a, b,
                eigenvalues=None, residual_filter=False, compute_vectors=True, multipass_cgs=False):
    """ChebyshevFilteredEigenSolver::solve: returns (Ritz values, EigenSolverErrorCode)."""
    st = C.c_int()
    w = np.zeros(guess.B) if eigenvalues is None else np.ascontiguousarray(eigenvalues, dtype=np.float64).copy()
    check(lib().hx_chfsi_solve_ortho(A.h, Bop.h, BInv.h, guess.p, vecs.p, C.c_uint32(guess.B), C.c_uint32(batch),
                                     C.c_uint32(degree), C.c_double(a0), C.c_double(a), C.c_double(b),
                                     C.c_int(int(residual_filter)), w.ctypes.data_as(f64p), C.c_int(int(compute_vectors)),
                                     C.c_int(1 if multipass_cgs else 0), C.byref(st)))
    return w, st.value


def eigen_residual_norms(A: Op, Mop: Op, X: DeviceBlock, eigenvalues, batch) -> np.ndarray:
    e, ep = _f64(eigenvalues)
    out = np.zeros(X.B)
    check(lib().hx_eigen_residual_norms(A.h, Mop.h, X.p, C.c_uint32(X.B), C.c_uint32(batch), ep,
                                        out.ctypes.data_as(f64p)))
    return out


def lanczos_extreme(A: Op, Bop: Op, BInv: Op, guess: DeviceBlock, max_krylov, n_lower=1, n_upper=1, tol=None,
                    beta_tol=1e-14, adaptive=False):
    """LanczosExtremeEigenSolver::solve (eigenvalues only): returns (eigenvalues, diagonal, subdiagonal, status)."""
    assert guess.B == 1
    ev = np.zeros(n_lower + n_upper)
    diag, sub = np.zeros(max_krylov), np.zeros(max_krylov)
    k, st = C.c_uint32(), C.c_int()
    t, tp = _f64(np.full(n_lower + n_upper, 1e-6) if tol is None else tol)
    check(lib().hx_lanczos_extreme(A.h, Bop.h, BInv.h, guess.p, C.c_uint32(max_krylov), C.c_uint32(n_lower),
                                   C.c_uint32(n_upper), tp, C.c_double(beta_tol), C.c_int(int(adaptive)),
                                   ev.ctypes.data_as(f64p), diag.ctypes.data_as(f64p), sub.ctypes.data_as(f64p),
                                   C.byref(k), C.byref(st)))
    return ev, diag[:k.value], sub[:k.value], st.value


class FeBasis:
    """FEBasisDataStorage arrays on the device + FEBasisOperations::computeFEMatrices."""

    def __init__(self, plan: Plan, num_cell_quad, basis_data, jxw, same_basis: bool):
        self.plan = plan
        d = FeBasisDesc()
        d.struct_size = C.sizeof(FeBasisDesc)
        d.same_basis_in_all_cells = int(same_basis)
        a, d.num_cell_quad = _u32(num_cell_quad)
        b, d.basis_data = _f64(basis_data)
        c, d.jxw = _f64(jxw)
        self.n_quad = int(np.sum(a.astype(np.int64)))
        self.h = C.c_void_p()
        check(lib().hx_fe_basis_create(plan.h, C.byref(d), C.byref(self.h)))

    def compute_fe_matrices(self, f, out: DeviceBlock, add_to: Optional[DeviceBlock] = None, f_device=None):
        """out: DeviceBlock of S2 doubles.  f: host array (one value per quadrature point) or f_device: DeviceBlock."""
        if f_device is not None:
            check(lib().hx_compute_fe_matrices(self.h, f_device.p, C.c_int(1), add_to.p if add_to is not None else None, out.p))
        else:
            a, p = _f64(f)
            assert a.size == self.n_quad
            check(lib().hx_compute_fe_matrices(self.h, p, C.c_int(0), add_to.p if add_to is not None else None, out.p))

    def assemble_into(self, op: "CellOp", f=None, add_to: Optional[DeviceBlock] = None, f_device=None):
        """computeFEMatrices + component sum + reinit in one kernel: straight into the operator's packed stream"""
        if f_device is not None:
            check(lib().hx_cellop_assemble_matrices(op.h, self.h, f_device.p, C.c_int(1), add_to.p if add_to is not None else None))
        else:
            a, p = _f64(f)
            assert a.size == self.n_quad
            check(lib().hx_cellop_assemble_matrices(op.h, self.h, p, C.c_int(0), add_to.p if add_to is not None else None))

    def compute_rho(self, X: DeviceBlock, occupation) -> np.ndarray:
        """DensityCalculator::computeRho: rho at the quadrature points (host array)."""
        o, op = _f64(occupation)
        assert o.size == X.B
        rho = np.zeros(self.n_quad)
        check(lib().hx_compute_rho(self.h, X.p, C.c_uint32(X.B), op, rho.ctypes.data_as(f64p), C.c_int(0)))
        return rho

    def compute_rho_device(self, X: DeviceBlock, occupation, rho: DeviceBlock):
        """same, rho left on the device (one value per quadrature point)"""
        o, op = _f64(occupation)
        check(lib().hx_compute_rho(self.h, X.p, C.c_uint32(X.B), op, rho.p, C.c_int(1)))

    def destroy(self):
        if self.h:
            lib().hx_fe_basis_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def chebyshev_polynomial_degree(unwanted_upper: float) -> int:
    d = C.c_uint32()
    check(lib().hx_chebyshev_polynomial_degree(C.c_double(unwanted_upper), C.byref(d)))
    return d.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().hx_comm_unique_id(buf))
    return buf.raw


def pdl_enabled() -> bool:
    return bool(lib().hx_programmatic_launch_enabled())


def microbench():
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    check(lib().hx_microbench(C.byref(a), C.byref(b), C.byref(c)))
    return {"dmma_tflops": a.value, "dfma_tflops": b.value, "copy_gbs": c.value}
