"""Property-based tests (hypothesis) of the two integer / byte paths of the product that run on the host:
the data-model construction behind rba_create (layout.hpp, replaces the reference's block factory
qr/landmark_block.cpp:51-80 and allocate_landmark ipp:67-86) and the BAL text loader (bal_problem.cpp:189-282).
Both must be bit-exact, so they are driven with random topologies / random token layouts instead of a few fixed shapes."""
import ctypes as C
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from rootba_b200 import _lib


def _random_topology(rng, nc, nl, max_n, heavy_tail):
    """landmark -> ascending camera list with 2 <= n <= min(nc, max_n)"""
    hi = max(2, min(nc, max_n))
    if heavy_tail:
        n = np.minimum(hi, 2 + rng.geometric(0.35, nl) - 1 + (rng.random(nl) < 0.03) * rng.integers(0, hi, nl))
    else:
        n = rng.integers(2, hi + 1, nl)
    off = np.zeros(nl + 1, np.int64)
    off[1:] = np.cumsum(n)
    cam = np.empty(int(off[-1]), np.int32)
    for l in range(nl):
        cam[off[l]:off[l + 1]] = np.sort(rng.choice(nc, int(n[l]), replace=False))
    return off, cam


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 2**31 - 1), nc=st.integers(2, 260), nl=st.integers(1, 600), max_n=st.sampled_from([2, 3, 8, 33, 70, 130, 260]),
       heavy_tail=st.booleans(), nranks=st.sampled_from([1, 2, 3, 5, 8]))
def test_layout_invariants_hold_for_random_topologies(seed, nc, nl, max_n, heavy_tail, nranks):
    """rba_layout_selftest rebuilds the layout rba_create uses and checks: every observation has exactly one slot with
    its camera and landmark, tiles are homogeneous in track length, the matvec row chunks tile every panel exactly once,
    both camera-major CSRs list every slot exactly once under the right camera, shards partition the landmarks."""
    rng = np.random.default_rng(seed)
    off, cam = _random_topology(rng, nc, nl, max_n, heavy_tail)
    xy = np.zeros(2 * len(cam))
    pv = _lib.ProblemView(nc, nl, len(cam), off.ctypes.data, cam.ctypes.data, xy.ctypes.data)
    L = _lib.lib()
    bounds = np.zeros(nranks + 1, np.int32)
    assert L.rba_partition_landmarks(C.c_int32(nl), off.ctypes.data_as(C.c_void_p), C.c_int32(nranks), bounds.ctypes.data_as(C.c_void_p)) == 0
    assert bounds[0] == 0 and bounds[-1] == nl and np.all(np.diff(bounds) >= 0)
    for rank in range(nranks):
        for ssz in (4, 8):
            rc = L.rba_layout_selftest(C.byref(pv), rank, nranks, ssz)
            assert rc == 0, (rank, nranks, ssz, L.rba_last_error())


# ---- BAL loader vs a pure-Python statement of the format ---------------------------------------------------------------

def _py_load_bal(text):
    """the file format in the most literal form: fscanf-style whitespace tokens (bal_problem.cpp:205-282)"""
    tok = text.split()
    nc, nl, nobs = int(tok[0]), int(tok[1]), int(tok[2])
    p = 3
    per_lm = [dict() for _ in range(nl)]
    for _ in range(nobs):
        c, l, x, y = int(tok[p]), int(tok[p + 1]), float(tok[p + 2]), float(tok[p + 3])
        p += 4
        assert c not in per_lm[l]
        per_lm[l][c] = (x, -y)  # std::map keyed by camera; image y flipped (:243)
    cams = np.array([float(t) for t in tok[p:p + 9 * nc]]).reshape(nc, 9)
    p += 9 * nc
    lms = np.array([float(t) for t in tok[p:p + 3 * nl]]).reshape(nl, 3)
    off = np.zeros(nl + 1, np.int64)
    oc, xy = [], []
    for l in range(nl):
        for c in sorted(per_lm[l]):
            oc.append(c)
            xy.append(per_lm[l][c])
        off[l + 1] = len(oc)
    return cams, lms, off, np.array(oc, np.int32), np.array(xy).reshape(-1, 2)


_number = st.one_of(
    st.floats(-1e4, 1e4, allow_nan=False, allow_infinity=False).map(lambda v: f"{v:.6e}"),
    st.floats(-1e3, 1e3, allow_nan=False, allow_infinity=False).map(repr),
    st.floats(-10, 10, allow_nan=False, allow_infinity=False).map(lambda v: f"{v:+.17g}"),
    st.integers(-50, 50).map(str),
    st.sampled_from(["0", "-0.0", "1e-320", "4.9e-324", "1.7976931348623157e308", "0.1", "1E5", ".5", "5."]),
)
_angle = st.floats(-3.0, 3.0, allow_nan=False, allow_infinity=False).map(lambda v: f"{v:.10e}")
_sep = st.sampled_from([" ", "\n", "\t", "  ", " \n", "\r\n", "\n\n"])


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(data=st.data(), seed=st.integers(0, 2**31 - 1), nc=st.integers(1, 9), nl=st.integers(1, 25), threads=st.sampled_from([1, 2, 7]))
def test_bal_loader_matches_a_literal_python_reading(tmp_path_factory, data, seed, nc, nl, threads):
    """rba_bal_load (multi-threaded, in place) against `text.split()` + float(): numbers in every style the format allows
    (%e, repr, explicit '+', integers, subnormals, DBL_MAX, '.5', '5.'), arbitrary whitespace between tokens, observations
    in random order.  Observation values, landmarks and the topology must be bit-identical; the camera rotation goes
    through Rodrigues -> quaternion (Sophus SO3::exp) and is checked against scipy to 1e-15."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    pairs = [(c, l) for l in range(nl) for c in rng.choice(nc, int(rng.integers(1, nc + 1)), replace=False)]
    rng.shuffle(pairs)
    toks = [str(nc), str(nl), str(len(pairs))]
    for c, l in pairs:
        toks += [str(int(c)), str(int(l)), data.draw(_number), data.draw(_number)]
    for _ in range(nc):  # Rodrigues vector of moderate size (a rotation), the other six values in any style
        toks += [data.draw(_angle) for _ in range(3)] + [data.draw(_number) for _ in range(6)]
    toks += [data.draw(_number) for _ in range(3 * nl)]
    text = "".join(t + data.draw(_sep) for t in toks)
    path = os.path.join(str(tmp_path_factory.mktemp("bal")), "p.txt")
    with open(path, "w", newline="") as f:
        f.write(text)
    import rootba_b200 as rb
    bp = rb.BalProblem.load_bal(path, np.float64, normalize=False, num_threads=threads)
    cams, lms, off, oc, xy = _py_load_bal(text)
    assert np.array_equal(bp.lm_off, off) and np.array_equal(bp.obs_cam, oc)
    assert bp.obs_xy.tobytes() == xy.tobytes() and bp.lms.tobytes() == lms.tobytes()  # bit-identical, incl. -0.0
    assert np.array_equal(bp.cams[:, 7:10], cams[:, 6:9])
    assert np.array_equal(bp.cams[:, 4:7], cams[:, 3:6] * np.array([1.0, -1.0, -1.0]))   # camera y/z axes flipped (:247-258)
    # T_c_w.so3 = diag(1,-1,-1) * exp(r): compare as rotation matrices (q and -q are the same rotation)
    want = np.diag([1.0, -1.0, -1.0]) @ Rotation.from_rotvec(cams[:, 0:3]).as_matrix()
    got = Rotation.from_quat(bp.cams[:, 0:4]).as_matrix()
    assert np.abs(got - want).max() < 1e-12
