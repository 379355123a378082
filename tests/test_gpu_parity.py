"""GPU parity tests: every call goes through the C ABI (dft_efe_b200.capi -> libhxb200.so) and is compared
with the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): integer/index work bit-exact; FP64 H.X within 1e-12 relative in the
L2 norm per vector.
"""
import numpy as np
import pytest

from dft_efe_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

RTOL_HX = 1e-12


def rel_l2_per_vector(a, b):
    num = np.linalg.norm(a - b, axis=0)
    den = np.linalg.norm(b, axis=0)
    den[den == 0] = 1.0
    return (num / den).max()


@pytest.fixture(scope="module")
def capi():
    from dft_efe_b200 import capi as c
    assert c.device_count() >= 1, "no CUDA device"
    return c


def spec_full(nranks=1, p=3, nc=(4, 4, 4), refine=True, enr=3, proj=2, boundary="dirichlet"):
    L = float(nc[0])
    atoms = np.array([[L / 2, L / 2, L / 2], [0.3 * L, 0.72 * L, 0.28 * L]])
    return synth.MeshSpec(ncell=nc, p=p, refine_mask=synth.refine_ball(nc, 1.0, [atoms[0]], 0.9) if refine else None,
                          atoms=atoms if (enr or proj) else None, n_enr_per_atom=enr, enr_cutoff=1.2,
                          n_proj_per_atom=proj, proj_cutoff=1.0, nranks=nranks, boundary=boundary)


@pytest.fixture(scope="module")
def prob_full():
    return synth.build_problem(spec_full())[0]


@pytest.fixture(scope="module")
def prob_plain():
    return synth.build_problem(synth.MeshSpec(ncell=(5, 4, 3), p=4, boundary="none"))[0]


# ------------------------------------------------------------------ integer work: bit exact ----
def test_colouring_and_constraint_transpose_bit_exact(capi, prob_full):
    plan = capi.Plan(prob_full, max_block=8)
    n, col = plan.colours()
    n_o, col_o = orc.cell_colouring(prob_full)
    assert n == n_o and np.array_equal(col, col_o)
    # a colouring is valid: cells of one colour share no non-shared row
    ids = prob_full.cell_local_ids.astype(np.int64)
    off = np.concatenate(([0], np.cumsum(prob_full.num_cell_dofs.astype(np.int64))))
    inc = np.bincount(ids, minlength=prob_full.n_local)
    for k in range(n):
        seen = set()
        for c in np.nonzero(col == k)[0]:
            rows = [r for r in ids[off[c]:off[c + 1]] if inc[r] <= 8]
            assert not (seen & set(rows))
            seen |= set(rows)
    a = plan.c2p_transpose()
    b = orc.c2p_transpose(prob_full)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# --------------------------------------------------------------------------------- leaf ops ----
@pytest.mark.parametrize("B", [1, 3, 8, 32])
def test_constraints_match_oracle(capi, prob_full, B):
    plan = capi.Plan(prob_full, max_block=32)
    R = orc.OracleRank(prob_full)
    X = synth.make_block(prob_full, B)
    d = plan.block(B, X)
    plan.p2c(d)
    xo = X.copy(); R.p2c(xo)
    assert rel_l2_per_vector(d.download(), xo) < 1e-14
    d = plan.block(B, X)
    plan.c2p(d)
    xo = X.copy(); R.c2p(xo)
    assert rel_l2_per_vector(d.download(), xo) < 1e-14


# ---------------------------------------------------------------------------------- H.X ----
@pytest.mark.parametrize("B", [1, 2, 5, 8, 16, 32, 40, 64, 96])
def test_hx_plain_mesh(capi, prob_plain, B):
    """uniform mesh, no constraints, uniform n_c: the minimum slice of SURVEY 7 step 2."""
    p = prob_plain
    plan = capi.Plan(p, max_block=96)
    op = capi.CellOp(plan)
    W = orc.OracleWorld([p])
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    op.apply(dX, dY, True, False)
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.hx_apply([Xo], [Yo], True, False, long_double=True)
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX


@pytest.mark.parametrize("B", [1, 4, 7, 32, 48])
def test_hx_full_features(capi, prob_full, B):
    """hanging nodes + Dirichlet rows + enrichment (variable n_c, shared rows) + nonlocal projectors."""
    p = prob_full
    plan = capi.Plan(p, max_block=48)
    op = capi.CellOp(plan)
    W = orc.OracleWorld([p])
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    op.apply(dX, dY, True, False)
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.hx_apply([Xo], [Yo], True, False)
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX
    # X is modified in place exactly like the reference (constrained rows filled)
    assert rel_l2_per_vector(dX.download(), Xo) < 1e-14
    # constrained rows of Y are zero
    assert np.all(dY.download()[p.row_ids.astype(np.int64)] == 0.0)


def test_hx_without_nonlocal_and_reinit(capi, prob_full):
    p = prob_full
    B = 8
    plan = capi.Plan(p, max_block=B)
    op = capi.CellOp(plan, with_nonlocal=False)
    W = orc.OracleWorld([p])
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    op.apply(dX, dY)
    Yo = np.zeros_like(X)
    W.hx_apply([X.copy()], [Yo], use_nonlocal=False)
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX
    # reinit with new cell matrices (one SCF iteration later): index maps unchanged
    h2 = p.h_cell * 1.5 + 0.01
    op.set_matrices(h2)
    dX = plan.block(B, X)
    op.apply(dX, dY)
    Yo = np.zeros_like(X)
    W.hx_apply([X.copy()], [Yo], use_nonlocal=False, h_cells=[h2])
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX


def test_hx_bitwise_deterministic(capi, prob_full):
    p = prob_full
    B = 16
    plan = capi.Plan(p, max_block=B)
    op = capi.CellOp(plan)
    X = synth.make_block(p, B)
    outs = []
    for _ in range(3):
        dX, dY = plan.block(B, X), plan.block(B)
        op.apply(dX, dY)
        outs.append(dY.download())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("delay", [0, 5, 23, 592])
def test_processing_order_and_wait_lists_bit_exact(capi, prob_full, delay, monkeypatch):
    """Ordered scatter: delay-D list schedule + per-cell predecessor lists match the oracle's specification."""
    p = prob_full
    monkeypatch.setenv("HXB200_ORDER_DELAY", str(delay))
    plan = capi.Plan(p, max_block=8)
    off, preds = plan.wait_lists()
    order = plan.processing_order()
    o_order, o_off, o_preds = orc.processing_order(p, delay)
    assert sorted(order.tolist()) == list(range(p.n_cells))
    assert np.array_equal(order, o_order)
    assert np.array_equal(off, o_off)
    assert np.array_equal(preds, o_preds)
    assert all((preds[off[w]:off[w + 1]] < w).all() for w in range(p.n_cells))
    # H.X is independent of the processing order up to rounding
    B = 8
    op = capi.CellOp(plan)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    op.apply(dX, dY, True, False)
    Yo = np.zeros_like(X)
    orc.OracleWorld([p]).hx_apply([X.copy()], [Yo], True, False)
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX


@pytest.mark.parametrize("delay", [7, 592])
def test_processing_order_of_a_partition_keys_its_boundary_cells_at_a_quarter(capi, delay, monkeypatch):
    """Multi-rank plans (halo overlap): the cells that read ghost rows are keyed at C/4 of the sweep; the order and the
    predecessor lists of every partition match the oracle's specification bit for bit.  (Plan creation needs no
    communicator, so a partition of a 2-rank mesh can be planned on one GPU.)"""
    nc = (4, 4, 8)
    atoms = np.array([[2.0, 2.0, 4.0]])
    spec = synth.MeshSpec(ncell=nc, p=3, refine_mask=synth.refine_ball(nc, 1.0, atoms, 0.9), atoms=atoms,
                          n_enr_per_atom=2, enr_cutoff=1.2, n_proj_per_atom=2, proj_cutoff=1.0, nranks=2)
    monkeypatch.setenv("HXB200_ORDER_DELAY", str(delay))
    for q in synth.build_problem(spec):
        plan = capi.Plan(q, max_block=8)
        off, preds = plan.wait_lists()
        order = plan.processing_order()
        o_order, o_off, o_preds = orc.processing_order(q, delay)
        assert np.array_equal(order, o_order)
        assert np.array_equal(off, o_off)
        assert np.array_equal(preds, o_preds)
        ids = q.cell_local_ids.astype(np.int64)
        cell_off = np.concatenate(([0], np.cumsum(q.num_cell_dofs.astype(np.int64))))
        bnd = np.array([(ids[cell_off[c]:cell_off[c + 1]] >= q.n_owned).any() for c in range(q.n_cells)])
        assert bnd.any() and not bnd.all()


@pytest.mark.parametrize("B", [3, 8, 32, 40])
def test_ordered_and_coloured_scatter_agree(capi, prob_full, B):
    p = prob_full
    plan = capi.Plan(p, max_block=B)
    op = capi.CellOp(plan)
    X = synth.make_block(p, B)
    Yo = np.zeros_like(X)
    orc.OracleWorld([p]).hx_apply([X.copy()], [Yo], True, False)
    res = []
    for mode in (0, 1):
        plan.set_scatter_mode(mode)
        dX = plan.block(B, X)
        dY = plan.block(B, np.full_like(X, 1e300))  # Y is fully overwritten: garbage must not leak
        op.apply(dX, dY, True, False)
        res.append(dY.download())
        assert rel_l2_per_vector(res[-1], Yo) < RTOL_HX
    assert rel_l2_per_vector(res[0], res[1]) < 1e-14


def test_ordered_scatter_many_applies(capi, prob_plain):
    """Epoch stamps / work counters survive repeated launches with changing block widths."""
    p = prob_plain
    plan = capi.Plan(p, max_block=64)
    op = capi.CellOp(plan)
    W = orc.OracleWorld([p])
    for B in (64, 8, 1, 32, 64, 2, 16):
        X = synth.make_block(p, B)
        dX, dY = plan.block(B, X), plan.block(B)
        for _ in range(3):
            op.apply(dX, dY)
        Yo = np.zeros_like(X)
        W.hx_apply([X.copy()], [Yo])
        assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX


def test_hx_host_entry_point(capi, prob_full):
    p = prob_full
    B = 8
    plan = capi.Plan(p, max_block=B)
    op = capi.CellOp(plan)
    X = synth.make_block(p, B)
    Xh, Yh = X.copy(), np.zeros_like(X)
    op.apply_host(Xh, Yh, True, False)
    Xo, Yo = X.copy(), np.zeros_like(X)
    orc.OracleWorld([p]).hx_apply([Xo], [Yo], True, False)
    assert rel_l2_per_vector(Yh, Yo) < RTOL_HX
    assert rel_l2_per_vector(Xh, Xo) < 1e-14


@pytest.mark.parametrize("p_order,B", [(6, 128), (4, 256), (5, 256), (7, 32), (8, 16)])
def test_hx_and_filter_at_the_baseline_block_widths(capi, p_order, B):
    """The shapes of BASELINE configs[2] (order 6, B = 128) and of a configs[3] column batch (order 5, B = 256), wide blocks at
    order 4 and the large cells of the sweep at narrow B: H.X within 1e-12 and the fused Chebyshev filter within 1e-11 of the
    oracle, on a mesh with hanging nodes, enrichment and projectors (several column tiles, several m-chunks per cell)."""
    nc = (3, 3, 3)
    atoms = np.array([[1.5, 1.5, 1.5]])
    spec = synth.MeshSpec(ncell=nc, p=p_order, refine_mask=synth.refine_ball(nc, 1.0, atoms, 0.9), atoms=atoms,
                          n_enr_per_atom=2, enr_cutoff=1.2, n_proj_per_atom=2, proj_cutoff=1.0)
    p = synth.build_problem(spec)[0]
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    H.apply(dX, dY, True, False)
    W = orc.OracleWorld([p])
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.hx_apply([Xo], [Yo], True, False)
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX
    dX, dF = plan.block(B, X), plan.block(B)
    capi.chebyshev_filter(H, minv, dX, dF, 4, -3.0, 1.0, 60.0)
    F = W.chebyshev_filter([X.copy()], 4, -3.0, 1.0, 60.0)[0]
    assert rel_l2_per_vector(dF.download()[:p.n_owned], F[:p.n_owned]) < 1e-11


@pytest.mark.parametrize("p_order", [1, 2, 5, 6])
def test_hx_other_orders(capi, p_order):
    nc = (3, 3, 3) if p_order >= 5 else (4, 4, 4)
    p = synth.build_problem(synth.MeshSpec(ncell=nc, p=p_order, refine_mask=synth.refine_ball(nc, 1.0, [[1.5, 1.5, 1.5]], 0.9)))[0]
    B = 12
    plan = capi.Plan(p, max_block=B)
    op = capi.CellOp(plan)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    op.apply(dX, dY)
    Yo = np.zeros_like(X)
    orc.OracleWorld([p]).hx_apply([X.copy()], [Yo])
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX


def test_empty_and_error_paths(capi, prob_full):
    plan = capi.Plan(prob_full, max_block=4)
    op = capi.CellOp(plan)
    dX, dY = plan.block(8), plan.block(8)
    with pytest.raises(capi.HxError):       # B > max_block
        op.apply(dX, dY)
    d4 = plan.block(4)
    with pytest.raises(capi.HxError):       # aliasing
        op.apply(d4, d4)
    # zero cells / zero constraints
    import copy
    e = copy.copy(prob_full)
    e.n_cells = 0
    e.num_cell_dofs = np.zeros(0, np.uint32); e.cell_local_ids = np.zeros(0, np.uint32); e.h_cell = np.zeros(0)
    e.num_cell_proj = None
    e.row_ids = e.row_sizes = e.row_offsets = e.col_ids = np.zeros(0, np.uint32)
    e.col_vals = e.inhom = np.zeros(0)
    pl = capi.Plan(e, max_block=4)
    o = capi.CellOp(pl)
    dX, dY = pl.block(4, synth.make_block(e, 4)), pl.block(4)
    o.apply(dX, dY)
    assert np.all(dY.download() == 0.0)


# ---------------------------------------------------------------------------- M, M^-1 ----
@pytest.mark.parametrize("variant", ["cfe", "oefe_atomblock"])
def test_diag_ops(capi, prob_full, variant):
    p = prob_full
    B = 8
    plan = capi.Plan(p, max_block=B)
    W = orc.OracleWorld([p])
    vid = capi.DIAG_CFE if variant == "cfe" else capi.DIAG_OEFE_ATOMBLOCK
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, vid)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    minv.apply(dX, dY, True, True)
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.minv_apply([Xo], [Yo], True, True, variant)
    assert rel_l2_per_vector(dY.download(), Yo) < 1e-14
    m = capi.DiagOp(plan, p.diag, p.enr_block, capi.DIAG_OEFE_MASS)
    dX = plan.block(B, X)
    m.apply(dX, dY, True, True)
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.m_apply([Xo], [Yo], True, True)
    assert rel_l2_per_vector(dY.download(), Yo) < 1e-14


@pytest.mark.parametrize("B", [1, 8, 32])
def test_global_enrichment_overlap_inverse(capi, prob_full, B):
    """HX_DIAG_OEFE_GLOBAL = OrthoEFEOverlapInverseOpContextGLL::apply (OrthoEFEOverlapInverseOpContextGLL.t.cpp:1182-1282):
    diagonal + one dense block over all enrichment functions; also as the M^-1 of the (then unfused) Chebyshev filter."""
    p = prob_full
    nE = p.n_owned - p.n_owned_classical
    assert nE > 0
    rng = np.random.default_rng(17)
    Rm = rng.standard_normal((nE, nE))
    blk = np.asfortranarray(Rm @ Rm.T / nE + np.eye(nE) + 0.05 * rng.standard_normal((nE, nE)))  # not symmetric: order matters
    plan = capi.Plan(p, max_block=B)
    MI = capi.DiagOpGlobalEnrichment(plan, p.diag_inv, blk.ravel(order="F"), nE, 0)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    MI.apply(dX, dY, True, True)
    W = orc.OracleWorld([p])
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.minv_apply_global_enrichment([Xo], [Yo], blk.ravel(order="F"), True, True)
    assert rel_l2_per_vector(dY.download(), Yo) < 1e-13
    assert rel_l2_per_vector(dX.download(), Xo) < 1e-14   # X is filled through the constraints, like the reference's apply
    if B % 2 == 0:
        H = capi.CellOp(plan)
        dX, dF = plan.block(B, X), plan.block(B)
        capi.chebyshev_filter(H, MI, dX, dF, 4, -3.0, 1.0, 60.0)
        assert np.isfinite(dF.download()).all()


def test_hx_against_reference_golden_fixture(capi):
    """tests/golden/ref_hx_small.npz = KohnShamOperatorContextFE::apply assembled from the reference's own compiled
    routines (tests/golden/make_golden.py): the CUDA path must match it within the H.X tolerance."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_hx_small.npz"))
    p = synth.build_problem(spec_full(p=int(g["p"]), nc=tuple(int(v) for v in g["nc"])))[0]
    B = int(g["B"])
    plan = capi.Plan(p, max_block=B)
    op = capi.CellOp(plan)
    dX, dY = plan.block(B, np.ascontiguousarray(g["X"])), plan.block(B)
    op.apply(dX, dY, True, False)
    assert rel_l2_per_vector(dY.download(), g["Y"]) < RTOL_HX
    assert rel_l2_per_vector(dX.download(), g["X_after"]) < 1e-14


def test_hx_enrichment_rows_shared_by_every_cell(capi):
    """an enrichment function whose cutoff covers the whole mesh: its row is touched by all 216 cells and goes
    through the two-stage fixed-order reduction of the staging slots."""
    nc = (6, 6, 6)
    spec = synth.MeshSpec(ncell=nc, p=3, atoms=np.array([[2.9, 3.1, 3.0]]), n_enr_per_atom=3, enr_cutoff=1e3,
                          n_proj_per_atom=2, proj_cutoff=1.1)
    p = synth.build_problem(spec)[0]
    assert int(np.bincount(p.cell_local_ids.astype(np.int64)).max()) == 216
    B = 16
    plan = capi.Plan(p, max_block=B)
    op = capi.CellOp(plan)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    op.apply(dX, dY, True, False)
    Y1 = dY.download()
    Xo, Yo = X.copy(), np.zeros_like(X)
    orc.OracleWorld([p]).hx_apply([Xo], [Yo], True, False)
    assert rel_l2_per_vector(Y1, Yo) < RTOL_HX
    op.apply(plan.block(B, X), dY, True, False)
    assert np.array_equal(Y1, dY.download())  # fixed reduction order: bitwise reproducible


# ------------------------------------------------------------------------- filters ----
@pytest.mark.parametrize("variant", ["cfe", "oefe_atomblock"])
def test_chebyshev_filter(capi, prob_full, variant):
    p = prob_full
    B, deg = 8, 9
    a0, a, b = -3.0, 1.0, 60.0
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    vid = capi.DIAG_CFE if variant == "cfe" else capi.DIAG_OEFE_ATOMBLOCK
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, vid)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    capi.chebyshev_filter(H, minv, dX, dY, deg, a0, a, b)
    F = orc.OracleWorld([p]).chebyshev_filter([X.copy()], deg, a0, a, b, minv_variant=variant)[0]
    own = p.n_owned
    assert rel_l2_per_vector(dY.download()[:own], F[:own]) < 1e-11   # degree-9 recurrence of 1e-12-accurate applies
    assert np.array_equal(dX.download()[:own], dY.download()[:own])  # both hold the result


@pytest.mark.parametrize("B", [2, 7, 8, 32, 40])
@pytest.mark.parametrize("mesh", ["full", "plain"])
def test_chebyshev_epilogue_fusion_is_bitwise_the_unfused_recurrence(capi, prob_full, prob_plain, mesh, B):
    """The recurrence applied by the last toucher inside the cell kernel's scatter (most rows) + the row-list pass
    (constrained / parent / enrichment / halo rows) must equal the separate full-vector pass bit for bit, and the
    oracle within the filter tolerance."""
    import os
    p = prob_full if mesh == "full" else prob_plain
    deg = 5
    a0, a, b = -3.0, 1.0, 60.0
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(p, B)
    res = []
    for off in (False, True):
        if off:
            os.environ["HXB200_NO_EPILOGUE_FUSION"] = "1"
        try:
            dX, dY = plan.block(B, X), plan.block(B)
            capi.chebyshev_filter(H, minv, dX, dY, deg, a0, a, b)
            res.append(dY.download()[:p.n_owned])
        finally:
            os.environ.pop("HXB200_NO_EPILOGUE_FUSION", None)
    assert np.array_equal(res[0], res[1])
    F = orc.OracleWorld([p]).chebyshev_filter([X.copy()], deg, a0, a, b)[0]
    assert rel_l2_per_vector(res[0], F[:p.n_owned]) < 1e-11


@pytest.mark.parametrize("B", [1, 8, 40])
@pytest.mark.parametrize("mesh", ["full", "plain"])
def test_programmatic_dependent_launch_is_bitwise_the_serialised_launches(capi, prob_full, prob_plain, mesh, B):
    """Programmatic dependent launch (HXB200_PDL) only overlaps the launch of kernel k+1 with the tail of kernel k:
    every kernel waits for its predecessor's completion before touching memory, so an apply and a filter (many short
    kernels back to back, repeated to give an ordering bug a chance to show) must give the same bits either way."""
    import os
    p = prob_full if mesh == "full" else prob_plain
    deg = 12
    a0, a, b = -3.0, 1.0, 60.0
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(p, B)
    old = os.environ.get("HXB200_PDL")
    res = {}
    try:
        for mode in ("0", "1"):
            os.environ["HXB200_PDL"] = mode
            out = []
            for _ in range(6):
                dX, dY = plan.block(B, X), plan.block(B)
                H.apply(dX, dY, True, False)
                out.append(dY.download())
                dX, dY = plan.block(B, X), plan.block(B)
                capi.chebyshev_filter(H, minv, dX, dY, deg, a0, a, b)
                out.append(dY.download()[:p.n_owned])
            res[mode] = out
    finally:
        if old is None:
            os.environ.pop("HXB200_PDL", None)
        else:
            os.environ["HXB200_PDL"] = old
    for k, (u, v) in enumerate(zip(res["0"], res["1"])):
        assert np.array_equal(u, v), f"result {k} differs between serialised and programmatic dependent launch"
        assert np.array_equal(u, res["0"][k % 2]), f"result {k} is not reproducible"


@pytest.mark.parametrize("B", [2, 8, 32])
def test_split_row_list_is_bitwise_the_single_launch(capi, prob_full, B):
    """The row-list pass of the fused filter runs as two launches (rows without a child list, parent rows; first run on a
    B200 in round 2); HXB200_SPLIT_ROWLIST=0 keeps the single launch.  Same kernels, same work per row."""
    import os
    p = prob_full
    deg = 7
    a0, a, b = -3.0, 1.0, 60.0
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(p, B)
    res = {}
    try:
        for mode in ("0", "1"):
            os.environ["HXB200_SPLIT_ROWLIST"] = mode
            dX, dY = plan.block(B, X), plan.block(B)
            capi.chebyshev_filter(H, minv, dX, dY, deg, a0, a, b)
            res[mode] = dY.download()[:p.n_owned]
    finally:
        os.environ.pop("HXB200_SPLIT_ROWLIST", None)
    assert np.array_equal(res["0"], res["1"])


def test_chebyshev_filter_host_entry(capi, prob_full):
    """hx_chebyshev_filter_host (HOST buffers in/out) == the device entry point, bit for bit."""
    p = prob_full
    B, deg = 8, 6
    a0, a, b = -3.0, 1.0, 60.0
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    capi.chebyshev_filter(H, minv, dX, dY, deg, a0, a, b)
    Xh, Yh = X.copy(), np.zeros_like(X)
    capi.chebyshev_filter_host(H, minv, Xh, Yh, deg, a0, a, b, write_back_x=True)
    own = p.n_owned
    assert np.array_equal(Yh[:own], dY.download()[:own])
    assert np.array_equal(Xh[:own], Yh[:own])
    Xh2, Yh2 = X.copy(), np.zeros_like(X)
    capi.chebyshev_filter_host(H, minv, Xh2, Yh2, deg, a0, a, b, write_back_x=False)
    assert np.array_equal(Xh2, X) and np.array_equal(Yh2[:own], Yh[:own])


def test_chebyshev_filter_host_batches_pipeline(capi, prob_full):
    """hx_chebyshev_filter_host_batches (column batches from HOST memory, copies overlapped with the filter of the batch in
    between) == the device entry point on every batch, bit for bit - 5 batches through 2 device buffers per direction."""
    import torch
    p = prob_full
    B, deg = 8, 6
    a0, a, b = -3.0, 1.0, 60.0
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    rng = np.random.default_rng(5)
    Xs = [synth.make_block(p, B) * (1.0 + 0.1 * k) + 1e-3 * rng.standard_normal((p.n_local, B)) for k in range(5)]
    xb = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in Xs]
    yb = [torch.zeros_like(x).pin_memory() for x in xb]
    for rep in range(2):   # second call: buffers and events are reused
        capi.chebyshev_filter_host_batches(H, minv, [x.data_ptr() for x in xb], [y.data_ptr() for y in yb], B, deg, a0, a, b)
        for k in range(5):
            dX, dY = plan.block(B, Xs[k]), plan.block(B)
            capi.chebyshev_filter(H, minv, dX, dY, deg, a0, a, b)
            assert np.array_equal(yb[k].numpy(), dY.download()), f"batch {k} differs (call {rep})"


def test_residual_chebyshev_filter(capi, prob_full):
    p = prob_full
    B, deg = 6, 7
    a0, a, b = -3.0, 1.0, 60.0
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    m = capi.DiagOp(plan, p.diag, p.enr_block, capi.DIAG_OEFE_MASS)
    ev = np.linspace(-2.5, 0.5, B)
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    capi.residual_chebyshev_filter(H, m, minv, ev, dX, dY, deg, a0, a, b)
    Yo = orc.OracleWorld([p]).residual_chebyshev_filter([X.copy()], ev, deg, a0, a, b)[0]
    own = p.n_owned
    assert rel_l2_per_vector(dY.download()[:own], Yo[:own]) < 1e-11


# ------------------------------------- electrostatics: Laplace operator + CG ----
def _laplace_ops(capi, p, B):
    plan = capi.Plan(p, max_block=B)
    xset = plan.add_constraints(p.row_ids, p.row_sizes, p.row_offsets, p.col_ids, p.col_vals, p.inhom_dirichlet)
    A = capi.CellOp(plan, h_cell=p.k_cell, with_nonlocal=False)
    pc = capi.DiagOp(plan, 1.0 / p.k_diag, None, capi.DIAG_JACOBI)
    return plan, xset, A, pc


@pytest.mark.parametrize("B", [1, 2, 5])
def test_laplace_operator_two_constraint_sets(capi, prob_full, B):
    """LaplaceOperatorContextFE::apply: X filled through the inhomogeneous-Dirichlet constraints of feBasisManagerX,
    Y condensed through the homogeneous ones of feBasisManagerY"""
    p = prob_full
    plan, xset, A, pc = _laplace_ops(capi, p, B)
    A.set_constraint_sets(xset, 0)
    W = orc.OracleWorld([p])
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    A.apply(dX, dY, True, True)
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.laplace_apply([Xo], [Yo], True, True, inhomogeneous=True)
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX
    assert rel_l2_per_vector(dX.download(), Xo) < 1e-14
    assert np.abs(Xo[p.row_ids.astype(np.int64)][p.row_sizes == 0] - p.inhom_dirichlet[p.row_sizes == 0][:, None]).max() == 0.0
    # Jacobi preconditioner
    pc.apply(dX, dY, False, False)
    Zo = np.zeros_like(X)
    W.jacobi_apply([Xo], [Zo], False, False)
    assert rel_l2_per_vector(dY.download(), Zo) < 1e-15


@pytest.mark.parametrize("B", [1, 4])
def test_cg_poisson_solve(capi, prob_full, B):
    """CGLinearSolver::solve on the device against the oracle's restatement (itself pinned against the reference's
    compiled CGLinearSolver, tests/test_oracle.py): same iteration count, same solution."""
    p = prob_full
    plan, xset, A, pc = _laplace_ops(capi, p, B)
    rng = np.random.default_rng(5)
    b = rng.standard_normal((p.n_local, B)); b[p.row_ids.astype(np.int64)] = 0.0
    x0 = 0.1 * rng.standard_normal((p.n_local, B))
    db, dx = plan.block(B, b), plan.block(B, x0)
    it, st, rn = capi.cg_solve(A, pc, db, dx, 400, 1e-12, 1e-10, 1e10)
    W = orc.OracleWorld([p])
    xs = [x0.copy()]
    ito, erro, rno = W.cg_solve(lambda X, Y, a, c: W.laplace_apply(X, Y, a, c, inhomogeneous=False),
                                lambda X, Y, a, c: W.jacobi_apply(X, Y, a, c), [b], xs, 400, 1e-12, 1e-10, 1e10)
    assert st == capi.CG_SUCCESS and erro == 0
    assert abs(it - ito) <= 1
    own = p.n_owned
    xg = dx.download()
    assert np.abs(xg[:own] - xs[0][:own]).max() < 1e-8 * np.abs(xs[0][:own]).max()
    # residual of the device solution through the device operator
    dy = plan.block(B)
    A.apply(plan.block(B, xg), dy, True, True)
    r = dy.download()[:own] - b[:own]
    assert np.linalg.norm(r) < 1e-8 * np.linalg.norm(b[:own])
    # failure mode: an iteration cap that cannot be met reports FAILED_TO_CONVERGE
    dx2 = plan.block(B, x0)
    it2, st2, _ = capi.cg_solve(A, pc, db, dx2, 3, 1e-14, 1e-14, 1e10)
    assert st2 == capi.CG_FAILED_TO_CONVERGE and it2 == 4


@pytest.mark.parametrize("B", [1, 8, 32])
def test_identical_cell_matrices_share_one_copy(capi, prob_full, B):
    """hx_cellop_set_matrix_sharing: the Laplace matrices of equally sized cells are bitwise identical and stream from one
    re-tiled copy; the result does not change by a bit.  A Hamiltonian that differs cell by cell keeps every copy."""
    p = prob_full
    plan = capi.Plan(p, max_block=B)
    A = capi.CellOp(plan, h_cell=p.k_cell, with_nonlocal=False)
    As = capi.CellOp(plan, h_cell=p.k_cell, with_nonlocal=False, share_identical=True)
    assert A.num_unique_matrices() == p.n_cells
    nu = As.num_unique_matrices()
    assert 2 <= nu < p.n_cells // 2, (nu, p.n_cells)  # two cell sizes + the enriched cells
    X = synth.make_block(p, B)
    y0, y1 = plan.block(B), plan.block(B)
    A.apply(plan.block(B, X), y0, True, True)
    As.apply(plan.block(B, X), y1, True, True)
    assert np.array_equal(y0.download(), y1.download())
    # re-setting different matrices undoes the sharing where they differ
    As.set_matrices(p.h_cell)
    assert As.num_unique_matrices() == p.n_cells
    Hn = capi.CellOp(plan, with_nonlocal=False)
    Hn.apply(plan.block(B, X), y0, True, False)
    As.apply(plan.block(B, X), y1, True, False)
    assert np.array_equal(y0.download(), y1.download())


@pytest.mark.parametrize("B", [1, 3])
def test_cg_fused_blas1_agrees_with_the_separate_passes(capi, prob_full, B, monkeypatch):
    """the two fused BLAS-1 passes of an iteration (Jacobi preconditioner) against the launch-per-operation path"""
    p = prob_full
    plan, xset, A, pc = _laplace_ops(capi, p, B)
    rng = np.random.default_rng(9)
    b = rng.standard_normal((p.n_local, B)); b[p.row_ids.astype(np.int64)] = 0.0
    x0 = 0.1 * rng.standard_normal((p.n_local, B))
    res = []
    for unfused in (False, True):
        if unfused:
            monkeypatch.setenv("HXB200_CG_UNFUSED", "1")
        dx = plan.block(B, x0)
        l0 = plan.launch_count()
        it, st, rn = capi.cg_solve(A, pc, plan.block(B, b), dx, 400, 1e-12, 1e-10, 1e10)
        res.append((it, st, dx.download()[:p.n_owned], plan.launch_count() - l0))
    monkeypatch.delenv("HXB200_CG_UNFUSED")
    assert res[0][1] == res[1][1] == capi.CG_SUCCESS and abs(res[0][0] - res[1][0]) <= 1
    assert np.abs(res[0][2] - res[1][2]).max() < 1e-9 * np.abs(res[1][2]).max()
    assert res[0][3] < res[1][3]  # fewer launches


# --------------------------------------------------------------- subspace projections ----
@pytest.mark.parametrize("B,batch", [(6, 4), (32, 32), (40, 16), (96, 64)])
def test_xtopx(capi, prob_full, B, batch):
    p = prob_full
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    W = orc.OracleWorld([p])
    X = synth.make_block(p, B)
    dX = plan.block(B, X)
    S = H.xtopx(dX, batch)
    So = W.xtopx([X.copy()], lambda a, b, c, d: W.hx_apply(a, b, c, d), batch)
    assert np.abs(S - So).max() < 1e-12 * np.abs(So).max()
    assert np.all(np.triu(S, 1) == 0.0)


@pytest.mark.parametrize("B", [6, 32, 72])
@pytest.mark.parametrize("transpose,lower", [(True, False), (False, True), (True, True)])
def test_subspace_rotation(capi, prob_full, B, transpose, lower):
    p = prob_full
    plan = capi.Plan(p, max_block=B)
    W = orc.OracleWorld([p])
    X = synth.make_block(p, B)
    Q = np.random.default_rng(5).standard_normal((B, B))
    if lower:
        Q = np.tril(Q)
    dX = plan.block(B, X)
    plan.subspace_rotation(dX, Q, transpose, lower)
    Xo = [X.copy()]
    W.subspace_rotation(Xo, Q, transpose, lower)
    own = p.n_owned
    assert rel_l2_per_vector(dX.download()[:own], Xo[0][:own]) < 1e-13


def test_l2_norms(capi, prob_full):
    p = prob_full
    B = 24
    plan = capi.Plan(p, max_block=B)
    X = synth.make_block(p, B)
    n = plan.l2_norms(plan.block(B, X))
    no = orc.OracleWorld([p]).l2_norms([X])
    assert np.abs(n - no).max() < 1e-13 * no.max()


# ----------------------------------------------------- size-independent properties at scale ----
def test_properties_at_bench_scale(capi):
    """C2-shaped problem (order 4, ~1M DoFs would take the oracle minutes): a 12^3 mesh of the same cell
    shape checked through properties that do not need the oracle: symmetry <HX,Z> = <X,HZ>, linearity,
    and agreement of two column tilings (B=32 as one tile vs 4 applies of 8 columns)."""
    p = synth.build_problem(synth.MeshSpec(ncell=(12, 12, 12), p=4))[0]
    B = 32
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    X = synth.make_block(p, B, seed=1); Z = synth.make_block(p, B, seed=2)
    X[p.row_ids.astype(np.int64)] = 0.0; Z[p.row_ids.astype(np.int64)] = 0.0
    dX, dZ, dHX, dHZ = plan.block(B, X), plan.block(B, Z), plan.block(B), plan.block(B)
    H.apply(dX, dHX); H.apply(dZ, dHZ)
    HX, HZ = dHX.download(), dHZ.download()
    a = np.sum(HX * Z, axis=0); b = np.sum(X * HZ, axis=0)
    assert np.abs(a - b).max() < 1e-11 * np.abs(a).max()
    dS = plan.block(B, 2.0 * X - 0.5 * Z); dHS = plan.block(B)
    H.apply(dS, dHS)
    assert rel_l2_per_vector(dHS.download(), 2.0 * HX - 0.5 * HZ) < 1e-12
    for j0 in range(0, B, 8):
        d8, dh8 = plan.block(8, np.ascontiguousarray(X[:, j0:j0 + 8])), plan.block(8)
        H.apply(d8, dh8)
        assert rel_l2_per_vector(dh8.download(), HX[:, j0:j0 + 8]) < 1e-13


def owned_rows_by_natural_id(whole, part):
    """Rows of the single-partition problem `whole` that hold the owned DoFs of partition `part` (RankProblem.natural_ids
    is the partition-independent DoF id the generator keys its values on)."""
    nat = whole.natural_ids[:whole.n_owned].astype(np.int64)
    order = np.argsort(nat, kind="stable")
    want = part.natural_ids[:part.n_owned].astype(np.int64)
    pos = np.searchsorted(nat[order], want)
    assert np.array_equal(nat[order][pos], want)
    return order[pos]


def test_c2_size_hx_and_filter_against_the_partitioned_oracle(capi):
    """BASELINE configs[1] at FULL size (25^3 cells, order 4, 1 030 321 DoFs, 32 vectors, 5 atoms x 4 enrichment functions
    + 4 projectors - the problem bench.py times) against the oracle itself, not only through properties: the oracle runs the
    same mesh cut into one partition per host core (its result is partition independent, tests/test_oracle.py), rows are
    matched through the partition-independent DoF ids.  One H.X apply and one fused degree-6 Chebyshev filter."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    nt = max(2, min(16, os.cpu_count() or 2))
    B, h, nc = 32, 0.8, (25, 25, 25)
    atoms = (0.25 + 0.5 * np.random.default_rng(7).uniform(size=(5, 3))) * (np.array(nc) * h)[None, :]

    def spec(nranks):
        return synth.MeshSpec(ncell=nc, p=4, h=h, atoms=atoms, n_enr_per_atom=4, enr_cutoff=1.6 * h, n_proj_per_atom=4,
                              proj_cutoff=1.3 * h, nranks=nranks, boundary="dirichlet", with_k_cell=False)

    whole = synth.build_problem(spec(1))[0]
    parts = synth.build_problem(spec(nt))
    assert whole.n_owned == 1030321 == sum(q.n_owned for q in parts)
    deg, a0, a, b = 6, -3.0, 1.0, 400.0
    # ---- GPU, through the C ABI ----
    plan = capi.Plan(whole, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, whole.diag_inv, whole.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    X = synth.make_block(whole, B)
    dX, dY = plan.block(B, X), plan.block(B)
    H.apply(dX, dY, True, False)
    Y = dY.download()
    dX = plan.block(B, X)
    capi.chebyshev_filter(H, minv, dX, dY, deg, a0, a, b)
    F = dY.download()
    del plan, H, minv, dX, dY
    # ---- oracle, one partition per thread ----
    orc.use_scipy_dgemm(True)
    W = orc.OracleWorld(parts)
    W.pool = ThreadPoolExecutor(max_workers=nt)
    Xs = [synth.make_block(q, B) for q in parts]
    Ys = [np.zeros_like(x) for x in Xs]
    W.hx_apply([x.copy() for x in Xs], Ys, True, False)
    Fs = W.chebyshev_filter([x.copy() for x in Xs], deg, a0, a, b)
    Yo, Fo = np.zeros((whole.n_owned, B)), np.zeros((whole.n_owned, B))
    for q, y, f in zip(parts, Ys, Fs):
        rows = owned_rows_by_natural_id(whole, q)
        Yo[rows] = y[:q.n_owned]
        Fo[rows] = f[:q.n_owned]
    assert rel_l2_per_vector(Y[:whole.n_owned], Yo) < RTOL_HX
    assert rel_l2_per_vector(F[:whole.n_owned], Fo) < 1e-11


def test_chebyshev_filter_against_reference_golden_fixture(capi):
    """tests/golden/ref_filter_small.npz = the reference's own compiled ChebyshevFilter template driving
    KohnShamOperatorContextFE::apply assembled from its compiled routines (tests/golden/make_golden.py): the CUDA filter
    (epilogue-fused recurrence, dependent launches and all) must match it within the filter tolerance."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_filter_small.npz"))
    p = synth.build_problem(spec_full(p=int(g["p"]), nc=tuple(int(v) for v in g["nc"])))[0]
    B = int(g["B"])
    a0, a, b = (float(v) for v in g["bounds"])
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    dX, dY = plan.block(B, np.ascontiguousarray(g["X"])), plan.block(B)
    capi.chebyshev_filter(H, minv, dX, dY, int(g["degree"]), a0, a, b)
    assert rel_l2_per_vector(dY.download()[:p.n_owned], g["F"]) < 1e-11


@pytest.mark.parametrize("B", [1, 4, 32])
def test_hx_and_filter_periodic_wrap(capi, B):
    """BASELINE configs[3] is periodic: the wrap as one-entry constraint rows (slave -> master, weight 1; corner masters
    with 7 slaves), cf. tests/test_oracle.py::test_periodic_wrap_constraints."""
    p = synth.build_problem(spec_full(p=3, nc=(4, 4, 4), refine=False, enr=2, proj=2, boundary="periodic"))[0]
    plan = capi.Plan(p, max_block=B)
    H = capi.CellOp(plan)
    W = orc.OracleWorld([p])
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    H.apply(dX, dY, True, False)
    Xo, Yo = X.copy(), np.zeros_like(X)
    W.hx_apply([Xo], [Yo], True, False)
    assert rel_l2_per_vector(dY.download(), Yo) < RTOL_HX
    assert rel_l2_per_vector(dX.download(), Xo) < 1e-14
    assert np.all(dY.download()[p.row_ids.astype(np.int64)] == 0.0)
    minv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    dX, dY = plan.block(B, X), plan.block(B)
    capi.chebyshev_filter(H, minv, dX, dY, 7, -3.0, 1.0, 60.0)
    F = W.chebyshev_filter([X.copy()], 7, -3.0, 1.0, 60.0)[0]
    assert rel_l2_per_vector(dY.download()[:p.n_owned], F[:p.n_owned]) < 1e-11

