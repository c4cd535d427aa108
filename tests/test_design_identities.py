"""The algebraic identities the CUDA path relies on, checked in numpy on the ORACLE's
landmark blocks (which are computed the reference's way: dense Householder / Givens on the full block, Givens damping
with un-doing).  CPU only: this does not test the kernels (tests/test_gpu_parity.py does), it documents and pins the
mathematics they are built on.

Block layout (qr/landmark_block_dynamic.hpp:49-69): rows 0..2 = Q1^T [Jp | Jl | r] (damped), rows 3..2n-1 = Q2^T [...],
rows 2n..2n+2 = the three damping rows; columns [Jp_0 (9) ... Jp_{n-1} (9) | pad | Jl (3) | r (1)]."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import oracle_py as orc


def _weighted_scaled_jacobians(prob, lm, scaling):
    """[Jp * D | Jl * s | r] of one landmark before marginalisation (2n x (9n + 3 + 1)), from linearize_point"""
    k0, k1 = int(prob.lm_off[lm]), int(prob.lm_off[lm + 1])
    n = k1 - k0
    Jp, Jl, r = np.zeros((2 * n, 9 * n)), np.zeros((2 * n, 3)), np.zeros(2 * n)
    for i, k in enumerate(range(k0, k1)):
        c = int(prob.obs_cam[k])
        res, jp, ji, jl, _ = orc.linearize_point(prob.obs_xy[k], prob.lms[lm], prob.cams[c])
        Jp[2 * i:2 * i + 2, 9 * i:9 * i + 9] = np.hstack([jp, ji]) * scaling[9 * c:9 * c + 9]
        Jl[2 * i:2 * i + 2] = jl
        r[2 * i:2 * i + 2] = res
    eps = np.sqrt(1e-10)
    return Jp, Jl * (1.0 / (eps + np.linalg.norm(Jl, axis=0))), r


@pytest.mark.parametrize("use_householder", [1, 0])
@pytest.mark.parametrize("lam", [0.0, 0.1])
def test_orthogonality_identities(small_problem, use_householder, lam):
    prob = small_problem
    o = orc.Oracle(prob, np.float64, orc.default_options(use_householder=use_householder, num_threads=1))
    assert o.linearize()
    o.solve(lam)  # stage 2: Jp scaled, landmark damping lam applied
    D = o.get_scaling()
    n_all = prob.track_lengths()
    for lm in [0, 1, int(np.argmax(n_all)), int(np.argmin(n_all)), prob.nl - 1]:
        blk, lm_idx, res_idx, _ = o.get_block(lm)
        n = int(n_all[lm])
        Jp, Jl, r = _weighted_scaled_jacobians(prob, lm, D)
        Q1d, P = blk[:3, :9 * n], blk[3:, :9 * n]          # damped Q1^T Jp (3 x 9n); Q2^T Jp incl. damping rows (2n x 9n)
        q1r, q2r = blk[:3, res_idx], blk[3:, res_idx]
        Rd = np.triu(blk[:3, lm_idx:lm_idx + 3])
        # (1) [Q1d; P] is an orthogonal transform of [Jp; 0]:  P^T P = Jp^T Jp - Q1d^T Q1d   (section 2.4, 9)
        assert rel_err(P.T @ P, Jp.T @ Jp - Q1d.T @ Q1d) < 1e-11
        # (2) gradient of the reduced system without touching the panel: P^T (Q2^T r) = Jp^T r - Q1d^T (Q1^T r)_d
        assert rel_err(P.T @ q2r, Jp.T @ r - Q1d.T @ q1r) < 1e-10
        # (3) the damped R factor: R_d^T R_d = Jl^T Jl + lam I  (what back-substitution inverts, ipp:233-239)
        assert rel_err(Rd.T @ Rd, Jl.T @ Jl + lam * np.eye(3)) < 1e-11
        # (4) the reduced system is the Schur complement of the damped landmark block
        M = np.linalg.inv(Jl.T @ Jl + lam * np.eye(3))
        assert rel_err(P.T @ P, Jp.T @ Jp - Jp.T @ Jl @ M @ Jl.T @ Jp) < 1e-9
        # (5) implicit operator (section 9): y = Jp^T (Jp x) - Q1d^T (Q1d x) equals the dense product P^T (P x)
        x = np.random.default_rng(lm).uniform(-1, 1, 9 * n)
        assert rel_err(Jp.T @ (Jp @ x) - Q1d.T @ (Q1d @ x), P.T @ (P @ x)) < 1e-10
        # (6) model cost change in the un-rotated basis (section 2.4): |Q^T v| = |v|
        assert abs(np.linalg.norm(blk[:, res_idx]) - np.linalg.norm(r)) <= 1e-12 * np.linalg.norm(r)


def test_compact_wy_equals_sequential_householder():
    """section 2.2: applying H2 H1 H0 (Eigen makeHouseholder convention: v[0] = 1, H = I - tau v v^T) one after the other
    equals  a - V (T^T (V^T a))  with the 3x3 triangular T accumulated as in k_linearize_qr (tau, g10, g20, g21)."""
    rng = np.random.default_rng(0)
    m = 14
    A = rng.normal(size=(m, 3))
    V, tau = np.zeros((m, 3)), np.zeros(3)
    W = A.copy()
    for k in range(3):
        x = W[k:, k]
        beta = -np.sign(x[0]) * np.linalg.norm(x)
        tau[k] = (beta - x[0]) / beta
        v = np.concatenate([[1.0], x[1:] / (x[0] - beta)])
        V[k:, k] = v
        W[k:, :] -= tau[k] * np.outer(v, v @ W[k:, :])
    assert np.allclose(np.tril(W[:, :3], -1), 0, atol=1e-13)  # A is upper triangular after the three reflectors
    a = rng.normal(size=m)
    seq = a.copy()
    for k in range(3):
        seq -= tau[k] * V[:, k] * (V[:, k] @ seq)
    z = V.T @ a
    g10, g20, g21 = V[:, 1] @ V[:, 0], V[:, 2] @ V[:, 0], V[:, 2] @ V[:, 1]
    w0 = tau[0] * z[0]
    w1 = tau[1] * (z[1] - g10 * w0)
    w2 = tau[2] * (z[2] - g20 * w0 - g21 * w1)
    assert np.allclose(a - (w0 * V[:, 0] + w1 * V[:, 1] + w2 * V[:, 2]), seq, rtol=1e-13, atol=1e-13)
