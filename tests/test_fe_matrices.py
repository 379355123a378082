"""FEBasisOperations::computeFEMatrices (cell matrices of a local potential) - the oracle restatement
against the reference's own compiled routines (oracle/_ref: hadamardProduct, scaleStridedVarBatched,
gemmStridedVarBatched in the reference's call sequence) and a known answer (f = 1 reproduces the consistent mass
matrix); the CUDA path against the oracle through the C ABI, and end to end into the H.X operator."""
import numpy as np
import pytest

from dft_efe_b200 import synth
from oracle import oracle as orc


def fe_spec(enr=2, refine=True, p=3, nc=(3, 3, 3)):
    L = np.array(nc, dtype=float)
    atoms = np.array([[L[0] / 2, L[1] / 2, L[2] / 2]])
    return synth.MeshSpec(ncell=nc, p=p, refine_mask=synth.refine_ball(nc, 1.0, [atoms[0]], 0.8) if refine else None,
                          atoms=atoms if enr else None, n_enr_per_atom=enr, enr_cutoff=1.2, n_proj_per_atom=0,
                          boundary="dirichlet")


def test_oracle_matches_reference_assembled_compute_fe_matrices(ref_lib):
    """variable DoFs per cell (EFE: the reference's non-zero-stride branch), cell blocks of 1, 4 and all cells"""
    p = synth.build_problem(fe_spec())[0]
    fe = synth.fe_basis_data(p)
    assert not fe["same_basis"]
    f = synth.potential_at_quad_points(p, int(fe["num_cell_quad"][0]))
    mine = orc.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], f, False)
    for blk in (1, 4, p.n_cells):
        ref = ref_lib.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], f, False, blk)
        assert np.abs(mine - ref).max() <= 1e-14 * np.abs(ref).max()


def test_shared_basis_branch_against_reference(ref_lib):
    """classical FE (one basis matrix for all cells, the reference's zeroStrideBasisVal branch): the reference passes
    stride 0 for the SCALED operand there (FEBasisOperations.t.cpp:345-346), so with more than one cell per block every
    cell of a block receives the block's first matrix; with one cell per block it computes the intended integral, which
    is what the oracle (and the CUDA path) compute for every cell."""
    p = synth.build_problem(fe_spec(enr=0))[0]
    fe = synth.fe_basis_data(p)
    assert fe["same_basis"]
    f = synth.potential_at_quad_points(p, int(fe["num_cell_quad"][0]))
    mine = orc.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], f, True)
    ref1 = ref_lib.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], f, True, 1)
    assert np.abs(mine - ref1).max() <= 1e-14 * np.abs(ref1).max()
    ref4 = ref_lib.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], f, True, 4)
    n2 = int(p.num_cell_dofs[0]) ** 2
    assert np.array_equal(ref4[n2:2 * n2], ref4[:n2])  # the reference's behaviour with cell blocks > 1, documented
    assert np.abs(ref4[:n2] - mine[:n2]).max() <= 1e-14 * np.abs(mine).max()


def test_unit_potential_reproduces_the_consistent_mass_matrix():
    """known answer: f = 1 with a Gauss rule exact for degree 2p gives M_c = int N_i N_j, which the generator builds
    independently from 1-D tensor products (cell-matrix path of CFEOverlapOperatorContext)."""
    spec = fe_spec(enr=0, refine=True)
    p = synth.build_problem(spec)[0]
    fe = synth.fe_basis_data(p)
    out = orc.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], np.ones(fe["jxw"].size), True)
    _, _, _, M1 = synth.ref_matrices_1d(p.p)
    n = (p.p + 1) ** 3
    for c in (0, p.n_cells // 2, p.n_cells - 1):
        m1 = M1 * (p.cell_edge[c] / 2.0)
        Mc = np.einsum("ad,be,cf->abcdef", m1, m1, m1).reshape(n, n)
        got = out[c * n * n:(c + 1) * n * n].reshape(n, n)
        assert np.abs(got - Mc).max() < 1e-13 * np.abs(Mc).max()
        assert np.abs(got - got.T).max() == 0.0 or np.abs(got - got.T).max() < 1e-17


@pytest.fixture(scope="module")
def capi():
    from dft_efe_b200 import capi as c
    assert c.device_count() >= 1, "no CUDA device"
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("enr,p_order,nq1d", [(2, 3, None), (0, 3, None), (0, 4, 6), (3, 2, 5), (0, 5, None)])
def test_gpu_compute_fe_matrices_matches_oracle(capi, enr, p_order, nq1d):
    p = synth.build_problem(fe_spec(enr=enr, p=p_order))[0]
    fe = synth.fe_basis_data(p, nq1d)
    f = synth.potential_at_quad_points(p, int(fe["num_cell_quad"][0]))
    want = orc.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], f, fe["same_basis"])
    plan = capi.Plan(p, max_block=8)
    feb = capi.FeBasis(plan, fe["num_cell_quad"], fe["basis"], fe["jxw"], fe["same_basis"])
    out = capi.DeviceBlock(p.S2, 1)
    feb.compute_fe_matrices(f, out)
    got = out.download().ravel()
    assert np.abs(got - want).max() < 1e-13 * np.abs(want).max()
    # exactly symmetric cell matrices (the mirror image of every off-diagonal tile is a copy)
    off = 0
    for n in p.num_cell_dofs[:5].astype(np.int64):
        m = got[off:off + n * n].reshape(n, n)
        assert np.array_equal(m[:64, 64:], m[64:, :64].T) if n > 64 else True
        off += n * n
    # f on the device + the component sum of reinit fused into the epilogue: H = K/2 + V
    base = capi.DeviceBlock(p.S2, 1, 0.5 * p.k_cell)
    fdev = capi.DeviceBlock(f.size, 1, f)
    out2 = capi.DeviceBlock(p.S2, 1)
    feb.compute_fe_matrices(None, out2, add_to=base, f_device=fdev)
    assert np.abs(out2.download().ravel() - (want + 0.5 * p.k_cell)).max() < 1e-13 * np.abs(want + 0.5 * p.k_cell).max()


@pytest.mark.gpu
def test_assembled_matrices_feed_the_hx_operator(capi):
    """SCF call sequence: potential at quadrature points -> cell matrices on the device -> reinit -> H.X; against the
    oracle doing the same on the host."""
    p = synth.build_problem(fe_spec(enr=2, p=3))[0]
    fe = synth.fe_basis_data(p)
    f = synth.potential_at_quad_points(p, int(fe["num_cell_quad"][0]))
    plan = capi.Plan(p, max_block=8)
    feb = capi.FeBasis(plan, fe["num_cell_quad"], fe["basis"], fe["jxw"], fe["same_basis"])
    base = capi.DeviceBlock(p.S2, 1, 0.5 * p.k_cell)
    hdev = capi.DeviceBlock(p.S2, 1)
    feb.compute_fe_matrices(f, hdev, add_to=base)
    H = capi.CellOp(plan, with_nonlocal=False)
    H.set_matrices_device(hdev.ptr)
    B = 8
    X = synth.make_block(p, B)
    dX, dY = plan.block(B, X), plan.block(B)
    H.apply(dX, dY, True, False)
    h_host = 0.5 * p.k_cell + orc.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], f,
                                                      fe["same_basis"])
    W = orc.OracleWorld([p])
    Yo = np.zeros_like(X)
    W.hx_apply([X.copy()], [Yo], True, False, h_cells=[h_host], use_nonlocal=False)
    err = np.linalg.norm(dY.download() - Yo, axis=0) / np.linalg.norm(Yo, axis=0)
    assert err.max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("enr,p_order,proj", [(2, 3, 0), (0, 4, 0), (3, 2, 2), (2, 5, 2)])
def test_assembly_straight_into_the_packed_stream(capi, enr, p_order, proj):
    """hx_cellop_assemble_matrices (assembly + component sum + reinit in one kernel) gives bit for bit the operator the
    two-step path (flat array, then hx_cellop_set_matrices) gives - with and without a nonlocal part, whose projector
    columns stay in place."""
    spec = fe_spec(enr=enr, p=p_order)
    if proj:
        spec.n_proj_per_atom, spec.proj_cutoff = proj, 1.0
        if spec.atoms is None:
            spec.atoms = np.array([[1.5, 1.5, 1.5]])
    p = synth.build_problem(spec)[0]
    fe = synth.fe_basis_data(p)
    f = synth.potential_at_quad_points(p, int(fe["num_cell_quad"][0]))
    B = 8
    plan = capi.Plan(p, max_block=B)
    feb = capi.FeBasis(plan, fe["num_cell_quad"], fe["basis"], fe["jxw"], fe["same_basis"])
    base = capi.DeviceBlock(p.S2, 1, 0.5 * p.k_cell)
    flat = capi.DeviceBlock(p.S2, 1)
    feb.compute_fe_matrices(f, flat, add_to=base)
    H2 = capi.CellOp(plan, with_nonlocal=bool(proj))
    H2.set_matrices_device(flat.ptr)
    H1 = capi.CellOp(plan, with_nonlocal=bool(proj))   # starts with the generator's matrices: fully overwritten
    feb.assemble_into(H1, f, add_to=base)
    X = synth.make_block(p, B)
    y1, y2 = plan.block(B), plan.block(B)
    H1.apply(plan.block(B, X), y1, True, False)
    H2.apply(plan.block(B, X), y2, True, False)
    assert np.array_equal(y1.download(), y2.download())
    # a fresh operator (nothing packed yet) and a second assembly with another potential
    H3 = capi.CellOp.__new__(capi.CellOp)
    capi.Op.__init__(H3, plan)
    capi.check(capi.lib().hx_cellop_create(plan.h, __import__("ctypes").byref(H3.h)))
    feb.assemble_into(H3, 2.0 * f)
    feb.compute_fe_matrices(2.0 * f, flat)
    H2.set_matrices_device(flat.ptr)
    if not proj:
        H3.apply(plan.block(B, X), y1, True, False)
        H2.apply(plan.block(B, X), y2, True, False)
        assert np.array_equal(y1.download(), y2.download())


# ------------------------------------------------------------------ 8f rank 3: density ----
def test_oracle_interpolation_matches_reference_routines(ref_lib):
    """psi at the quadrature points: the oracle's interpolate against the reference's own gather + batched GEMM with
    the reference's operand conventions; rho then follows computeRhoInBatch's three-line loop."""
    p = synth.build_problem(fe_spec())[0]
    fe = synth.fe_basis_data(p)
    B = 5
    X = synth.make_block(p, B)
    psi = ref_lib.interpolate(p, fe["num_cell_quad"], fe["basis"], fe["same_basis"], X)
    occ = np.array([1.0, 1.0, 0.7, 0.2, 0.0])
    rho_ref = np.zeros(fe["jxw"].size)
    off = 0
    for c, nq in enumerate(fe["num_cell_quad"].astype(np.int64)):
        blk = psi[off * B:(off + nq) * B].reshape(nq, B)
        rho_ref[off:off + nq] = (2.0 * blk * blk * occ[None, :]).sum(axis=1)
        off += nq
    for batch in (B, 2):
        rho = orc.compute_rho(p, fe["num_cell_quad"], fe["basis"], fe["same_basis"], X, occ, batch)
        assert np.abs(rho - rho_ref).max() < 1e-13 * np.abs(rho_ref).max()


def test_density_integrates_to_the_electron_count():
    """known answer: for M-orthonormal orbitals int rho = sum_i 2 f_i (consistent mass matrix, exact quadrature)."""
    p = synth.build_problem(fe_spec(enr=0, refine=False))[0]
    fe = synth.fe_basis_data(p)
    B = 4
    X = synth.make_block(p, B)
    X[p.row_ids.astype(np.int64)] = 0.0
    # orthonormalise against the consistent mass matrix assembled from the cell matrices
    Mc = orc.compute_fe_matrices(p.num_cell_dofs, fe["num_cell_quad"], fe["basis"], fe["jxw"], np.ones(fe["jxw"].size), True)
    W = orc.OracleWorld([p])
    MX = np.zeros_like(X)
    W.hx_apply([X.copy()], [MX], False, False, h_cells=[Mc], use_nonlocal=False)
    S = X[:p.n_owned].T @ MX[:p.n_owned]
    X = X @ np.linalg.inv(np.linalg.cholesky(S)).T
    occ = np.array([1.0, 1.0, 0.5, 0.25])
    rho = orc.compute_rho(p, fe["num_cell_quad"], fe["basis"], True, X, occ, B)
    assert abs(np.dot(rho, fe["jxw"]) - 2.0 * occ.sum()) < 1e-11
    assert rho.min() >= 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("enr,p_order,B", [(2, 3, 5), (0, 3, 8), (0, 4, 32), (3, 2, 70), (0, 5, 130)])
def test_gpu_compute_rho_matches_oracle(capi, enr, p_order, B):
    p = synth.build_problem(fe_spec(enr=enr, p=p_order))[0]
    fe = synth.fe_basis_data(p)
    X = synth.make_block(p, B)
    occ = np.clip(np.linspace(1.2, -0.2, B), 0.0, 1.0)
    want = orc.compute_rho(p, fe["num_cell_quad"], fe["basis"], fe["same_basis"], X, occ, 16)
    plan = capi.Plan(p, max_block=B)
    feb = capi.FeBasis(plan, fe["num_cell_quad"], fe["basis"], fe["jxw"], fe["same_basis"])
    got = feb.compute_rho(plan.block(B, X), occ)
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
    # bitwise reproducible (fixed summation order, no atomics)
    assert np.array_equal(got, feb.compute_rho(plan.block(B, X), occ))
