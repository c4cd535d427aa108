"""BalProblem::perturb (bal/bal_problem.cpp:507-554) in the loader path (rba_bal_perturb): checked against an independent
Python restatement of the reference's code AND of the libstdc++ pieces it draws from -- std::default_random_engine
(= minstd_rand0), std::generate_canonical<double, 53> and the polar std::normal_distribution, a FRESH distribution object
per 3-vector as `perturbation<T, N>` creates it (:105-114).  Host only, no GPU."""
import math

import numpy as np
import pytest

from conftest import rel_err


class MinstdRand0:
    M, A = 2147483647, 16807

    def __init__(self, seed):
        self.x = seed % self.M or 1

    def __call__(self):
        self.x = (self.A * self.x) % self.M
        return self.x


def canonical(eng):  # std::generate_canonical<double, 53>: two draws of ~31 bits
    r = 2147483646.0
    s = float(eng() - 1)
    s += float(eng() - 1) * r
    ret = s / (r * r)
    return math.nextafter(1.0, 0.0) if ret >= 1.0 else ret


def perturbation3(eng, sigma):
    """three draws from one fresh std::normal_distribution<double>: the polar method yields pairs (y first, x saved)"""
    out, saved = [], None
    for _ in range(3):
        if saved is not None:
            v, saved = saved, None
        else:
            while True:
                x = 2.0 * canonical(eng) - 1.0
                y = 2.0 * canonical(eng) - 1.0
                r2 = x * x + y * y
                if not (r2 > 1.0 or r2 == 0.0):
                    break
            mult = math.sqrt(-2.0 * math.log(r2) / r2)
            saved, v = x * mult, y * mult
        out.append(0.0 + v * sigma)
    return np.array(out)


def test_perturb_matches_the_reference_random_stream(tmp_path):
    import rootba_b200 as rb
    from rootba_b200.synthetic import quat_to_rot, so3_exp, synth_bal, write_bal
    prob = synth_bal(9, 60, 3.5, seed=5, normalize_scale=None, perturb_lm=0, perturb_rot=0, perturb_trans=0)
    path = str(tmp_path / "p.txt")
    write_bal(prob, path)
    base = rb.BalProblem.load_bal(path)
    rs, ts, ps, seed = 0.01, 0.02, 0.05, 38401
    got = rb.BalProblem.load_bal(path, rotation_sigma=rs, translation_sigma=ts, point_sigma=ps, random_seed=seed)
    eng = MinstdRand0(seed)
    cams, lms = base.cams.copy(), base.lms.copy()
    for c in cams:
        R = quat_to_rot(c[None, :4])[0]
        ctr = -R.T @ c[4:7]
        ctr = ctr + perturbation3(eng, ts)                 # translation first, then rotation (bal_problem.cpp:533-544)
        c[4:7] = -R @ ctr
        q = so3_exp(perturbation3(eng, rs)[None])[0]       # exp(w) * so3
        a, b = q, c[:4].copy()
        c[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]
        c[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1]
        c[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2]
        c[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0]
        c[:4] /= np.linalg.norm(c[:4])
    for l in lms:
        l += perturbation3(eng, ps)
    assert rel_err(got.lms, lms) < 1e-14 and rel_err(got.cams[:, 4:7], cams[:, 4:7]) < 1e-12
    assert rel_err(got.cams[:, :4], cams[:, :4]) < 1e-12 and np.array_equal(got.cams[:, 7:], base.cams[:, 7:])
    # the perturbation has the requested size, is deterministic in the seed and absent for zero sigmas
    assert 0.5 * ps < np.std(got.lms - base.lms) < 1.5 * ps
    again = rb.BalProblem.load_bal(path, rotation_sigma=rs, translation_sigma=ts, point_sigma=ps, random_seed=seed)
    other = rb.BalProblem.load_bal(path, rotation_sigma=rs, translation_sigma=ts, point_sigma=ps, random_seed=7)
    assert np.array_equal(again.lms, got.lms) and np.array_equal(again.cams, got.cams) and not np.array_equal(other.lms, got.lms)
    only_points = rb.BalProblem.load_bal(path, point_sigma=ps)
    assert np.array_equal(only_points.cams, base.cams) and not np.array_equal(only_points.lms, base.lms)
    with pytest.raises(rb.RbaError):
        rb.BalProblem.load_bal(path, point_sigma=-1.0, rotation_sigma=0.1)
