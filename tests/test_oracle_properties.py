"""The reference's own property tests, re-run on the CPU oracle (this is what pins the oracle;
the reference has no golden vectors and cannot be built here -- see oracle/rootba_oracle.hpp).

  * Jacobians vs central differences: bal/bal_bundle_adjustment_helper.test.cpp:54-148
  * projection formula:               bal/snavely_projection.test.cpp:155-188
  * QR == SC equivalence:             qr/linearization_qr.test.cpp:120-222
  * matvec vs explicit sparse:        qr/linearization_qr.test.cpp:63-110
"""
import numpy as np
import pytest

from conftest import rel_err
from oracle import oracle_py as orc
from rootba_b200 import synthetic as syn

# testing/float_utils.hpp:62-69: default_test_precision = 1e-5 (f32) / 1e-12 (f64), the reference's own bar
PREC = {np.float32: 1e-5, np.float64: 1e-12}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_projection_formula(dtype):
    rng = np.random.default_rng(0)
    for _ in range(50):
        cam = np.zeros(10)
        cam[3] = 1.0
        cam[7:] = [rng.uniform(300, 2000), rng.normal(0, 1e-2), rng.normal(0, 1e-3)]
        p = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(0.5, 8)])
        res, *_ = orc.linearize_point([0, 0], p, cam, dtype)
        m = p[:2] / p[2]
        r2 = m @ m
        want = cam[7] * (1 + cam[8] * r2 + cam[9] * r2 * r2) * m
        assert rel_err(res, want) < (1e-6 if dtype == np.float32 else 1e-14)


def test_jacobians_vs_central_differences():
    # double only for the tight check (reference: eps 1e-8, tol 1e-3 for double)
    rng = np.random.default_rng(1)
    prob = syn.synth_bal(6, 40, 3.5, seed=3)
    for k in range(20):
        o = rng.integers(0, prob.nobs)
        lm = np.searchsorted(prob.lm_off, o, side="right") - 1
        cam = prob.cams[prob.obs_cam[o]].copy()
        p_w = prob.lms[lm].copy()
        obs = prob.obs_xy[o]
        res, Jp, Ji, Jl, _ = orc.linearize_point(obs, p_w, cam)
        h = 1e-6
        # pose: left-multiplicative decoupled update (Camera::inc_pose, bal_problem.hpp:99-101)
        for j in range(6):
            inc = np.zeros(9)
            inc[j] = h
            rp, *_ = orc.linearize_point(obs, p_w, orc.camera_apply_inc(cam, inc))
            rm, *_ = orc.linearize_point(obs, p_w, orc.camera_apply_inc(cam, -inc))
            num = (rp - rm) / (2 * h)
            assert np.allclose(num, Jp[:, j], rtol=1e-5, atol=1e-5 * (1 + np.abs(Jp).max())), (j, num, Jp[:, j])
        for j in range(3):
            inc = np.zeros(9)
            hj = h * max(1.0, abs(cam[7 + j]))
            inc[6 + j] = hj
            rp, *_ = orc.linearize_point(obs, p_w, orc.camera_apply_inc(cam, inc))
            rm, *_ = orc.linearize_point(obs, p_w, orc.camera_apply_inc(cam, -inc))
            num = (rp - rm) / (2 * hj)
            assert np.allclose(num, Ji[:, j], rtol=1e-5, atol=1e-5 * (1 + np.abs(Ji[:, j]).max()))
        for j in range(3):
            d = np.zeros(3)
            d[j] = h
            rp, *_ = orc.linearize_point(obs, p_w + d, cam)
            rm, *_ = orc.linearize_point(obs, p_w - d, cam)
            num = (rp - rm) / (2 * h)
            assert np.allclose(num, Jl[:, j], rtol=1e-5, atol=1e-5 * (1 + np.abs(Jl).max()))


def _explicit_Q2TJp(o, nc):
    """dense (sum 2n) x 9nc matrix from the blocks (get_Q2TJp, qr/linearization_qr.hpp:181-220)"""
    rows = []
    for l in range(o.nl):
        blk, lm_idx, res_idx, _ = o.get_block(l)
        n = (blk.shape[0] - 3) // 2
        cams = o.obs_cam[o.lm_off[l]:o.lm_off[l + 1]]
        M = np.zeros((blk.shape[0] - 3, 9 * nc))
        for i in range(n):
            M[:, 9 * cams[i]:9 * cams[i] + 9] = blk[3:, 9 * i:9 * i + 9]
        rows.append(M)
    return np.vstack(rows)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_basic_linear_algebra(tiny_problem, dtype):
    o = orc.Oracle(tiny_problem, dtype)
    d2, ok = o.stage1()
    assert ok
    M = _explicit_Q2TJp(o, o.nc)
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, 9 * o.nc)
    y = o.right_multiply(x)
    want = M.T @ (M @ x)
    assert rel_err(y, want) < (1e-5 if dtype == np.float32 else 1e-12)


@pytest.mark.parametrize("use_householder", [1, 0])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_qr_sc_equivalence(small_problem, dtype, use_householder):
    prec = PREC[dtype]
    lam = 0.1
    opts = orc.default_options(use_householder=use_householder)
    qr = orc.Oracle(small_problem, dtype, opts)
    sc = orc.Oracle(small_problem, dtype, opts)
    rng = np.random.default_rng(7)
    eps = np.sqrt(1e-10) if dtype == np.float64 else np.sqrt(np.float32(1e-5))
    for pose_damping in (0.0, lam):
        d2, ok = qr.stage1()
        assert ok
        scaling = (1.0 / (eps + np.sqrt(d2.astype(np.float64)))).astype(dtype)
        d2_sc = sc.sc_linearize()
        assert rel_err(d2, d2_sc) < 2 * prec  # solver/bal_bundle_adjustment.test.cpp:60,83-86
        qr.set_pose_damping(pose_damping)
        b_qr, blocks_qr = qr.stage2(lam, scaling, schur_blocks=True)
        sc.sc_scale_Jp(scaling)
        x = rng.uniform(-1, 1, 9 * qr.nc).astype(dtype)
        b_sc, blocks_sc, y_sc = sc.sc_get_Hb(lam, pose_damping, x)
        assert rel_err(b_qr, b_sc) < prec
        for c in range(qr.nc):
            assert rel_err(blocks_qr[c], blocks_sc[c]) < prec * 5, c
        y_qr = qr.right_multiply(x)
        assert rel_err(y_qr, y_sc) < prec
        pose_inc = (rng.uniform(-1, 1, 9 * qr.nc) * 0.01).astype(dtype)
        l_qr, ok = qr.back_substitute(pose_inc)
        assert ok
        l_sc, lms_sc = sc.sc_back_substitute(lam, pose_inc)
        assert abs(l_qr - l_sc) / (abs(l_qr) + abs(l_sc)) < prec * 10
        _, lms_qr = qr.get_state()
        assert rel_err(lms_qr, lms_sc) < prec
        sc.set_state(*qr.get_state())  # keep both at the same state for the second pass


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_staged_equals_unstaged(tiny_problem, dtype):
    a = orc.Oracle(tiny_problem, dtype, orc.default_options(staged_execution=1))
    b = orc.Oracle(tiny_problem, dtype, orc.default_options(staged_execution=0))
    assert a.linearize() and b.linearize()
    inc_a, it_a = a.solve(1e-4)
    inc_b, it_b = b.solve(1e-4)
    assert rel_err(inc_a, inc_b) < (1e-4 if dtype == np.float32 else 1e-10)


def test_block_inverse():
    rng = np.random.default_rng(3)
    A = rng.normal(size=(9, 9))
    B = A @ A.T + 0.1 * np.eye(9)
    inv = orc.invert_block9(np.triu(B))  # only the upper triangle is read (selfadjointView<Upper>)
    assert np.allclose(inv, np.linalg.inv(B), rtol=1e-9, atol=1e-11)
    d = rng.uniform(0.1, 1, 9)
    inv2 = orc.invert_block9(B, d)
    assert np.allclose(inv2, np.linalg.inv(B + np.diag(d)), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("precond", [1, 0])
def test_lm_converges(small_problem, dtype, precond):
    o = orc.Oracle(small_problem, dtype, orc.default_options(preconditioner_type=precond, max_num_iterations=10))
    rows, term = o.optimize()
    assert term in (0, 1)
    costs = [r["cost"] for r in rows if r["step_is_successful"]]
    assert costs[-1] < 0.2 * costs[0]
    # successful steps are monotone
    assert all(c1 <= c0 * (1 + 1e-9) for c0, c1 in zip(costs, costs[1:]))


def test_threads_match_single(small_problem):
    a = orc.Oracle(small_problem, np.float64, orc.default_options(num_threads=1))
    b = orc.Oracle(small_problem, np.float64, orc.default_options(num_threads=4))
    assert a.linearize() and b.linearize()
    inc_a, _ = a.solve(1e-4)
    inc_b, _ = b.solve(1e-4)
    assert rel_err(inc_a, inc_b) < 1e-9
    assert abs(a.apply(inc_a) - b.apply(inc_b)) < 1e-6 * abs(a.apply(inc_a) * 0 + 1)
