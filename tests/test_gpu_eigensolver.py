"""GPU parity tests of the device-resident eigensolve around the H.X path (8f rank 4): every
call goes through the C ABI and is compared with the CPU oracle (oracle/eigensolver.py) on identical inputs, and the
converged Kohn-Sham eigenvalues with a dense generalized eigensolve of the assembled pencil.

Tolerances: dense B x B factors 1e-12 relative; Ritz values 1e-10 relative after one pass; converged Kohn-Sham
eigenvalues within 1e-8 Ha (BASELINE.json north_star)."""
import numpy as np
import pytest
import scipy.linalg as sla

from dft_efe_b200 import synth
from oracle import eigensolver as es
from oracle import oracle as orc
from tests.test_eigensolver_oracle import dense_pencil, eig_spec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from dft_efe_b200 import capi as c
    assert c.device_count() >= 1, "no CUDA device"
    return c


@pytest.fixture(scope="module")
def setup(capi):
    ps = synth.build_problem(eig_spec(1))
    p = ps[0]
    W = orc.OracleWorld(ps)
    plan = capi.Plan(p, max_block=16)
    H = capi.CellOp(plan)
    M = capi.DiagOp(plan, p.diag, p.enr_block, capi.DIAG_OEFE_MASS)
    MInv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    return p, W, plan, H, M, MInv


@pytest.mark.parametrize("B", [1, 5, 32, 200])
def test_dense_cholesky_inverse_and_sym_eig(capi, setup, B):
    _, _, plan, *_ = setup
    rng = np.random.default_rng(B)
    A = rng.standard_normal((B, B))
    S = A @ A.T + B * np.eye(B)
    d = capi.DenseMatrix(B, np.tril(S))  # only the lower triangle is read
    assert capi.dense_cholesky_inverse(plan, d) == 0
    Linv = d.download()
    ref = np.linalg.inv(np.linalg.cholesky(S))
    assert np.all(np.triu(Linv, 1) == 0.0)
    assert np.abs(Linv - ref).max() < 1e-12 * np.abs(ref).max() * B
    T = A + A.T
    d = capi.DenseMatrix(B, np.tril(T))
    w, info = capi.dense_sym_eig(plan, d)
    assert info == 0
    Q = d.download()
    wref = np.linalg.eigvalsh(T)
    assert np.abs(w - wref).max() < 1e-12 * max(1.0, np.abs(wref).max())
    assert np.abs(Q.T @ Q - np.eye(B)).max() < 1e-12
    assert np.abs(Q.T @ T @ Q - np.diag(w)).max() < 1e-11 * max(1.0, np.abs(wref).max())
    # a matrix that is not positive definite is reported, not factorised
    bad = capi.DenseMatrix(B, -np.eye(B))
    assert capi.dense_cholesky_inverse(plan, bad) > 0


@pytest.mark.parametrize("B,batch", [(7, 3), (16, 16), (12, 5)])
def test_xtopx_and_rotation_device_variants(capi, setup, B, batch):
    p, W, plan, H, M, MInv = setup
    X = synth.make_block(p, B)
    S_host = H.xtopx(plan.block(B, X), batch)
    S_dev = capi.xtopx_device(H, plan.block(B, X), batch).download()
    assert np.array_equal(np.tril(S_host), np.tril(S_dev)) and np.all(np.triu(S_dev, 1) == 0.0)
    Q = np.random.default_rng(1).standard_normal((B, B))
    for transpose, lower in ((True, False), (False, False), (False, True), (True, True)):
        Qm = np.tril(Q) if lower else Q
        a, b = plan.block(B, X), plan.block(B, X)
        plan.subspace_rotation(a, Qm, transpose, lower)
        capi.subspace_rotation_device(plan, b, capi.DenseMatrix(B, Qm), transpose, lower)
        assert np.array_equal(a.download(), b.download())


@pytest.mark.parametrize("B,batch", [(7, 3), (16, 16)])
def test_cholesky_gram_schmidt_and_rayleigh_ritz(capi, setup, B, batch):
    p, W, plan, H, M, MInv = setup
    X = synth.make_block(p, B)
    dX, dO = plan.block(B, X), plan.block(B)
    assert capi.cholesky_gram_schmidt(M, dX, dO, batch) == 0
    Xo = [X.copy()]
    Mop = lambda a, b, gx, gy: W.m_apply(a, b, gx, gy)  # noqa: E731
    Aop = lambda a, b, gx, gy: W.hx_apply(a, b, gx, gy)  # noqa: E731
    st, _ = es.cholesky_gram_schmidt(W, Xo, Mop, batch)
    assert st == 0
    n = p.n_owned
    got = dO.download()
    assert np.abs(got[:n] - Xo[0][:n]).max() < 1e-10 * np.abs(Xo[0][:n]).max()
    assert np.array_equal(dX.download(), got)  # X rotated in place, orthogonalizedX = X
    dV = plan.block(B)
    w, st = capi.rayleigh_ritz(H, dO, dV, batch, True)
    wo, _ = es.rayleigh_ritz(W, Xo, Aop, batch)
    assert st == 0 and np.abs(w - wo).max() < 1e-10 * np.abs(wo).max()
    V, Vo = dV.download()[:n], Xo[0][:n]
    sign = np.sign(np.einsum("ij,ij->j", V, Vo))  # eigenvectors are defined up to a sign
    assert np.abs(V * sign[None, :] - Vo).max() < 1e-8 * np.abs(Vo).max()
    # a rank-deficient block is refused like the reference does (Cholesky fails / diagonal of L below 1e-14)
    Xd = X.copy(); Xd[:, 1] = 0.0
    assert capi.cholesky_gram_schmidt(M, plan.block(B, Xd), plan.block(B), batch) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("B,batch", [(8, 3), (16, 16)])
def test_multipass_cgs_matches_oracle(capi, setup, B, batch):
    """OrthonormalizationFunctions::MultipassCGS (OrthonormalizationFunctions.t.cpp:440-785): a well conditioned block
    needs one pass; a block with nearly dependent columns is shifted and re-orthonormalised in further passes.  Same pass
    count and same block as the oracle; the result is M-orthonormal to 1e-10; a ChFSI pass with MULTIPASS_CGS gives the
    Ritz values of the CholGS pass."""
    p, W, plan, H, M, MInv = setup
    Mop = lambda a, b, gx, gy: W.m_apply(a, b, gx, gy)  # noqa: E731
    n = p.n_owned
    X = synth.make_block(p, B)
    Xbad = X.copy()
    Xbad[:, 1] = Xbad[:, 0] * (1.0 + 1e-9) + 1e-7 * X[:, 1]   # kappa(X^T M X) ~ 1e14: CholGS alone loses orthogonality
    for blk, min_passes in ((X, 1), (Xbad, 2)):
        dX, dO = plan.block(B, blk), plan.block(B)
        st, passes = capi.multipass_cgs(M, dX, dO, batch)
        Xo = [blk.copy()]
        sto, passes_o = es.multipass_cgs(W, Xo, Mop, batch)
        assert st == 0 and sto == 0 and passes == passes_o and passes >= min_passes, (st, sto, passes, passes_o)
        got = dO.download()
        assert np.array_equal(dX.download(), got)
        if min_passes == 1:
            assert np.abs(got[:n] - Xo[0][:n]).max() < 1e-9 * np.abs(Xo[0][:n]).max()
        G = [got.copy()]
        S = W.xtopx(G, Mop, batch)
        S = S + S.T - np.diag(np.diag(S))
        assert np.abs(S - np.eye(B)).max() < 1e-10
    dG, dV = plan.block(B, X), plan.block(B)
    w1, st1 = capi.chfsi_solve(H, M, MInv, dG, dV, batch, 12, -1.0, 6.0, 2500.0, None, False, True, multipass_cgs=True)
    dG, dV = plan.block(B, X), plan.block(B)
    w0, st0 = capi.chfsi_solve(H, M, MInv, dG, dV, batch, 12, -1.0, 6.0, 2500.0, None, False, True)
    assert st0 == 0 and st1 == 0 and np.abs(w1 - w0).max() < 1e-9 * np.abs(w0).max()
    assert plan.global_size() == p.n_owned


@pytest.mark.parametrize("residual_filter", [False, True])
def test_chfsi_pass_matches_oracle(capi, setup, residual_filter):
    p, W, plan, H, M, MInv = setup
    B, batch, deg = 8, 3, 20
    X = synth.make_block(p, B)
    ev0 = np.linspace(-0.5, 4.0, B)
    dG, dV = plan.block(B, X), plan.block(B)
    w, st = capi.chfsi_solve(H, M, MInv, dG, dV, batch, deg, -1.0, 6.0, 2500.0, ev0, residual_filter, True)
    g = [X.copy()]
    wo, sto, vo = es.chfsi_solve(W, g, ev0, batch, deg, -1.0, 6.0, 2500.0, residual_filter=residual_filter)
    assert st == 0 and sto == 0
    assert np.abs(w - wo).max() < 1e-9 * np.abs(wo).max()
    n = p.n_owned
    V, Vo = dV.download()[:n], vo[0][:n]
    sign = np.sign(np.einsum("ij,ij->j", V, Vo))
    # Ritz vectors of well separated Ritz values agree up to sign
    sep = np.minimum(np.diff(wo, prepend=-np.inf), np.diff(wo, append=np.inf)) > 1e-3 * np.abs(wo).max()
    assert sep.sum() >= 2
    assert np.abs(V * sign[None, :] - Vo)[:, sep].max() < 1e-6 * np.abs(Vo).max()
    assert np.array_equal(dG.download()[:n], dV.download()[:n])  # the guess holds the Ritz vectors for the next pass
    r = capi.eigen_residual_norms(H, M, dV, w, batch)
    ro = es.eigen_residual_norms(W, vo, wo, batch)
    assert np.abs(r - ro).max() < 1e-8 * max(ro.max(), 1e-30) + 1e-12


def test_lanczos_matches_oracle(capi, setup):
    p, W, plan, H, M, MInv = setup
    g = np.random.default_rng(5).uniform(-0.5, 0.5, (p.n_local, 1))
    ev, diag, sub, st = capi.lanczos_extreme(H, M, MInv, plan.block(1, g), 20)
    A = lambda X, Y, gx, gy: W.hx_apply(X, Y, gx, gy)  # noqa: E731
    Mo = lambda X, Y, gx, gy: W.m_apply(X, Y, gx, gy)  # noqa: E731
    MI = lambda X, Y, gx, gy: W.minv_apply(X, Y, gx, gy)  # noqa: E731
    evo, do, so, sto = es.lanczos_extreme(W, A, Mo, MI, [g.copy()], 20)
    assert st == 0 and sto == 0 and len(diag) == 20 and len(sub) == 20
    # the recurrence amplifies rounding differences step by step: tight at the start, looser at the end
    assert np.abs(diag[:5] - do[:5]).max() < 1e-10 * np.abs(do).max()
    assert np.abs(diag - do).max() < 1e-6 * np.abs(do).max() and np.abs(sub - so).max() < 1e-6 * np.abs(so).max()
    assert np.abs(ev - evo).max() < 1e-7 * np.abs(evo).max()
    # adaptive mode on a conforming mesh reaches the extreme eigenvalue of the pencil
    ps = synth.build_problem(eig_spec(1, refine=False, enr=0))
    q = ps[0]
    Wc = orc.OracleWorld(ps)
    Hd, Md, _ = dense_pencil(Wc, ps)
    lam = sla.eigh(Hd, Md, eigvals_only=True)
    plan2 = capi.Plan(q, max_block=4)
    H2 = capi.CellOp(plan2)
    M2 = capi.DiagOp(plan2, q.diag, q.enr_block, capi.DIAG_OEFE_MASS)
    MI2 = capi.DiagOp(plan2, q.diag_inv, q.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
    g2 = np.random.default_rng(6).uniform(-0.5, 0.5, (q.n_local, 1))
    ev2, d2, s2, st2 = capi.lanczos_extreme(H2, M2, MI2, plan2.block(1, g2), 200, tol=[1e-9, 1e-9], adaptive=True)
    assert st2 == 0 and abs(ev2[1] - lam[-1]) < 1e-6 * abs(lam[-1]) and ev2[0] >= lam[0] - 1e-9


def ks_eigen_solve_gpu(capi, plan, H, M, MInv, p, X0, lanczos_guess, n_el, T, fermi_tol, occ_tol, res_tol, max_pass, batch,
                       residual_filter=False, bounds=None):
    """KohnShamEigenSolver::solve (src/ksdft/KohnShamEigenSolver.t.cpp:214-560) driven through the C ABI: the host
    scalar logic (bounds, degree, Fermi level) is the reference's, every vector operation is a library call."""
    B = X0.shape[1]
    ev_l, diag, sub, st = capi.lanczos_extreme(H, M, MInv, plan.block(1, lanczos_guess), es.LANCZOS_MAX_KRYLOV_SUBSPACE)
    assert st in (0, 3)
    residual = sub[-1] / 10.0
    unwanted = ev_l[1] + residual
    lower = ev_l[0]
    upper = (unwanted - ev_l[0]) * (B * 200.0 / p.n_owned) + ev_l[0]
    if upper >= unwanted:
        upper = (unwanted + ev_l[0]) * 0.5
    if bounds is not None:
        lower, upper = bounds
    deg = es.cheby_polynomial_degree(unwanted)
    dG, dV = plan.block(B, X0), plan.block(B)
    evals = np.zeros(B)
    for ipass in range(max_pass):
        evals, st = capi.chfsi_solve(H, M, MInv, dG, dV, batch, deg, lower, upper, unwanted, evals, residual_filter, True)
        assert st == 0
        plan.update_ghost_values(dV)
        mu, ok = es.fermi_energy(list(evals), n_el, T, fermi_tol)
        occ = np.array([es.fermi_dirac(e, mu, es.BOLTZMANN_CONST_HARTREE, T) for e in evals])
        res = capi.eigen_residual_norms(H, M, dV, evals, batch)
        dG.upload(dV.download())
        if int(np.sum(occ > occ_tol)) == int(np.sum((occ > occ_tol) & (res <= res_tol))) or not ok:
            return evals, occ, res, ipass + 1, deg, dV
        lower, upper = evals[0], evals[B - 1]
    raise AssertionError(f"not converged in {max_pass} passes: {res}")


@pytest.mark.parametrize("residual_filter,refined", [(False, True), (True, True), (False, False)])
def test_converged_kohn_sham_eigenvalues(capi, setup, residual_filter, refined):
    """north_star: converged Kohn-Sham eigenvalues within 1e-8 Ha of the reference solution — here of (i) the oracle
    running the same eigensolver on the same inputs and (ii) the dense generalized eigensolve of the assembled pencil."""
    if refined:
        p, W, plan, H, M, MInv = setup
        bounds = (-1.0, 6.0)
    else:
        ps = synth.build_problem(eig_spec(1, refine=False, enr=0))
        p, W = ps[0], orc.OracleWorld(ps)
        plan = capi.Plan(p, max_block=16)
        H = capi.CellOp(plan)
        M = capi.DiagOp(plan, p.diag, p.enr_block, capi.DIAG_OEFE_MASS)
        MInv = capi.DiagOp(plan, p.diag_inv, p.enr_block_inv, capi.DIAG_OEFE_ATOMBLOCK)
        bounds = None
    Hd, Md, _ = dense_pencil(W, [p])
    lam = sla.eigh(Hd, Md, eigvals_only=True)
    B, n_el = 8, 8
    X0 = synth.make_block(p, B)
    lg = np.random.default_rng(11).uniform(-0.5, 0.5, (p.n_local, 1))
    ev, occ, res, passes, deg, dV = ks_eigen_solve_gpu(capi, plan, H, M, MInv, p, X0, lg, n_el, 500.0, 1e-10, 1e-8, 1e-9,
                                                       60, 3, residual_filter, bounds)
    out = es.ks_eigen_solve(W, [X0.copy()], [lg.copy()], n_el, 500.0, 1e-10, 1e-8, 1e-9, 60, batch=3,
                            residual_filter=residual_filter, bounds=bounds)
    n_occ = int(np.sum(occ > 1e-8))
    assert out["status"] == 0 and n_occ >= 4 and deg == out["degree"]
    assert np.abs(ev[:n_occ] - out["eigenvalues"][:n_occ]).max() < 1e-8
    assert np.abs(ev[:n_occ] - lam[:n_occ]).max() < 1e-8
    assert np.all(res[:n_occ] <= 1e-9)
    # band energy 2 sum f_i eps_i (the eigenvalue part of the total energy): 1e-7 Ha
    band = 2.0 * np.sum(occ * ev)
    band_o = 2.0 * np.sum(out["occupancy"] * out["eigenvalues"])
    assert abs(band - band_o) < 1e-7
    # the converged vectors are M-orthonormal
    V = dV.download()[:p.n_owned]
    MV = [np.zeros((p.n_local, B))]
    Vl = np.zeros((p.n_local, B)); Vl[:p.n_owned] = V
    W.m_apply([Vl], MV, True, False)
    assert np.abs(V.T @ MV[0][:p.n_owned] - np.eye(B)).max() < 1e-10


def test_error_paths_of_the_eigensolve_entry_points(capi, setup):
    """the error convention of the boundary: bad arguments come back as HxError (HX_ERR_INVALID), never a crash"""
    p, W, plan, H, M, MInv = setup
    B = plan.max_block
    X = synth.make_block(p, B)
    dG, dV = plan.block(B, X), plan.block(B)
    with pytest.raises(capi.HxError):  # guess and eigenvectors must be different blocks
        capi.chfsi_solve(H, M, MInv, dG, dG, 4, 5, -1.0, 6.0, 2500.0)
    big = capi.DeviceBlock(p.n_local, B + 1)
    with pytest.raises(capi.HxError):  # wider than max_block
        capi.chfsi_solve(H, M, MInv, big, capi.DeviceBlock(p.n_local, B + 1), 4, 5, -1.0, 6.0, 2500.0)
    with pytest.raises(capi.HxError):  # degree 0
        capi.chfsi_solve(H, M, MInv, dG, dV, 4, 0, -1.0, 6.0, 2500.0)
    with pytest.raises(capi.HxError):  # Krylov space smaller than the number of wanted eigenvalues
        capi.lanczos_extreme(H, M, MInv, plan.block(1, X[:, :1].copy()), 1, 1, 1)
    other = capi.Plan(p, max_block=4)
    M2 = capi.DiagOp(other, p.diag, p.enr_block, capi.DIAG_OEFE_MASS)
    with pytest.raises(capi.HxError):  # operators of different plans
        capi.eigen_residual_norms(H, M2, dV, np.zeros(B), 4)
    # a block that is fine afterwards: the failed calls left the plan usable
    w, st = capi.chfsi_solve(H, M, MInv, dG, dV, 4, 5, -1.0, 6.0, 2500.0)
    assert st == 0 and np.all(np.diff(w) >= 0)
