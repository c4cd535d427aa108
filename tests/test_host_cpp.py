"""C++ host mirror (rootba_b200/host): BAL loader + normalisation vs the oracle's loader and the generator;
bal_qr end-to-end on the GPU vs the Python host mirror."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_err

BAL_QR = os.path.join(ROOT, "rootba_b200", "host", "bal_qr")


def _build():
    from rootba_b200 import _lib
    _lib.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "rootba_b200", "host"), "-s"])


def _read_dump(path):
    raw = open(path, "rb").read()
    nc, nl, nobs = np.frombuffer(raw, np.int64, 3)
    o = 24
    cams = np.frombuffer(raw, np.float64, 10 * nc, o).reshape(nc, 10); o += 80 * nc
    lms = np.frombuffer(raw, np.float64, 3 * nl, o).reshape(nl, 3); o += 24 * nl
    off = np.frombuffer(raw, np.int64, nl + 1, o); o += 8 * (nl + 1)
    oc = np.frombuffer(raw, np.int32, nobs, o); o += 4 * nobs
    xy = np.frombuffer(raw, np.float64, 2 * nobs, o).reshape(nobs, 2)
    return cams, lms, off, oc, xy


def test_loader_matches_oracle_and_generator(tmp_path):
    from oracle import oracle_py as orc
    from rootba_b200.synthetic import synth_bal, write_bal
    _build()
    prob = synth_bal(7, 60, 3.2, seed=5, normalize_scale=None, perturb_lm=0, perturb_rot=0, perturb_trans=0)
    path = str(tmp_path / "p.txt")
    write_bal(prob, path)
    # raw load (no normalisation): round trip of the generator's arrays through the BAL text format
    subprocess.check_call([BAL_QR, "--input", path, "--no-normalize", "--dump-problem", str(tmp_path / "raw.bin")])
    cams, lms, off, oc, xy = _read_dump(str(tmp_path / "raw.bin"))
    assert np.array_equal(off, prob.lm_off) and np.array_equal(oc, prob.obs_cam)  # indexing bit-exact
    assert np.allclose(xy, prob.obs_xy, rtol=0, atol=1e-12) and np.allclose(lms, prob.lms, rtol=0, atol=1e-12)
    q_sign = np.sign(np.sum(cams[:, :4] * prob.cams[:, :4], axis=1, keepdims=True))  # q and -q are the same rotation
    assert np.allclose(cams[:, :4] * q_sign, prob.cams[:, :4], atol=1e-12)
    assert np.allclose(cams[:, 4:], prob.cams[:, 4:], atol=1e-10)
    # normalised load: two independent implementations (C++ host, oracle) of bal_problem.cpp:189-282,428-469
    subprocess.check_call([BAL_QR, "--input", path, "--dump-problem", str(tmp_path / "norm.bin")])
    cams, lms, off, oc, xy = _read_dump(str(tmp_path / "norm.bin"))
    ref = orc.load_bal(path, normalize=True)
    assert np.array_equal(off, ref["lm_off"]) and np.array_equal(oc, ref["obs_cam"])
    assert rel_err(cams, ref["cams"]) < 1e-14 and rel_err(lms, ref["lms"]) < 1e-14 and np.array_equal(xy, ref["obs_xy"])


@pytest.mark.parametrize("loader", ["parallel", "map"])
def test_rejects_duplicate_observation(tmp_path, loader):
    _build()
    p = tmp_path / "dup.txt"
    p.write_text("1 1 2\n0 0 1.0 2.0\n0 0 1.5 2.5\n" + "0\n" * 9 + "0\n0\n1\n")
    r = subprocess.run([BAL_QR, "--input", str(p), "--loader", loader, "--dump-problem", str(tmp_path / "x.bin")], capture_output=True)
    assert r.returncode != 0  # the reference CHECK-fails on a duplicate (cam, lm) pair (bal_problem.cpp:229-230)


def _shuffled_bal_text(prob, seed, style):
    """BAL text of `prob` with the observation lines in random order and mixed token separators / number styles:
    the loaders tokenise on whitespace like fscanf, and must sort the observations of a landmark by camera."""
    from scipy.spatial.transform import Rotation
    from rootba_b200.synthetic import quat_to_rot
    rng = np.random.default_rng(seed)
    flip = np.diag([1.0, -1.0, -1.0])
    rv = Rotation.from_matrix(np.einsum("ij,mjk->mik", flip, quat_to_rot(prob.cams[:, :4]))).as_rotvec()
    tb = prob.cams[:, 4:7] @ flip.T
    lm_of_obs = np.repeat(np.arange(prob.nl), np.diff(prob.lm_off))
    fmt = {"g17": lambda v: f"{v:.17g}", "e6": lambda v: f"{v:.6e}", "plus": lambda v: f"{v:+.12e}"}[style]
    sep = {"g17": " ", "e6": "\t", "plus": "  "}[style]
    lines = [f"{prob.nc}{sep}{prob.nl}{sep}{prob.nobs}"]
    for k in rng.permutation(prob.nobs):
        x, y = prob.obs_xy[k]
        lines.append(sep.join([str(prob.obs_cam[k]), str(lm_of_obs[k]), fmt(x), fmt(-y)]))
    for i in range(prob.nc):
        lines += [fmt(v) for v in (*rv[i], *tb[i], *prob.cams[i, 7:10])]
    # landmark coordinates three per line: the format is token-based, not line-based
    lines += [sep.join(fmt(v) for v in prob.lms[i]) for i in range(prob.nl)]
    eol = "\r\n" if style == "e6" else "\n"
    return eol.join(lines) + ("" if style == "plus" else eol)


@pytest.mark.parametrize("style", ["g17", "e6", "plus"])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_parallel_loader_is_bit_identical_to_map_loader(tmp_path, style, threads):
    """the mmap + multi-threaded loader (bal_io_fast.hpp) and the reference-style fscanf + std::map loader
    (bal_problem.hpp, after bal_problem.cpp:189-282) must produce the same bytes: topology, observations, state."""
    from rootba_b200.synthetic import synth_bal
    _build()
    prob = synth_bal(23, 3000, 3.7, seed=17, normalize_scale=None)
    path = tmp_path / "shuffled.txt"
    path.write_text(_shuffled_bal_text(prob, seed=3, style=style))
    dumps = {}
    for loader in ("parallel", "map"):
        out = str(tmp_path / f"{loader}.bin")
        subprocess.check_call([BAL_QR, "--input", str(path), "--loader", loader, "--num-threads", str(threads), "--dump-problem", out],
                              stdout=subprocess.DEVNULL)
        dumps[loader] = open(out, "rb").read()
    assert dumps["parallel"] == dumps["map"]
    cams, lms, off, oc, xy = _read_dump(str(tmp_path / "parallel.bin"))
    assert np.array_equal(off, prob.lm_off) and np.array_equal(oc, prob.obs_cam)  # sorted back: indexing bit-exact


@pytest.mark.parametrize("text", [
    "2 1 2\n0 0 1.0 2.0\n1 0 1.5 2.5\n" + "0\n" * 18 + "0\n0\n",          # file ends early
    "2 1 2\n0 0 1.0 2.0\n2 0 1.5 2.5\n" + "0\n" * 18 + "0\n0\n1\n",       # camera index out of range
    "2 1 2\n0 0 1.0 2.0\n1 0 1.5 abc\n" + "0\n" * 18 + "0\n0\n1\n",       # not a number
    "0 1 2\n",                                                                 # bad header
])
def test_parallel_loader_rejects_malformed_files(tmp_path, text):
    _build()
    p = tmp_path / "bad.txt"
    p.write_text(text)
    r = subprocess.run([BAL_QR, "--input", str(p), "--loader", "parallel", "--dump-problem", str(tmp_path / "x.bin")], capture_output=True)
    assert r.returncode != 0 and b"FATAL" in r.stderr


def test_abi_loader_matches_map_loader_and_oracle(tmp_path):
    """rba_bal_load (C ABI, used by BalProblem.load_bal) == bal_qr --loader map (bytes) == oracle loader (to rounding)"""
    import rootba_b200 as rb
    from oracle import oracle_py as orc
    from rootba_b200.synthetic import synth_bal
    _build()
    prob = synth_bal(31, 2500, 4.2, seed=23, normalize_scale=None)
    path = tmp_path / "p.txt"
    path.write_text(_shuffled_bal_text(prob, seed=8, style="e6"))
    for normalize in (False, True):
        bp = rb.BalProblem.load_bal(str(path), np.float64, normalize=normalize, num_threads=4)
        out = str(tmp_path / "map.bin")
        subprocess.check_call([BAL_QR, "--input", str(path), "--loader", "map", "--dump-problem", out] + ([] if normalize else ["--no-normalize"]),
                              stdout=subprocess.DEVNULL)
        cams, lms, off, oc, xy = _read_dump(out)
        assert np.array_equal(bp.lm_off, off) and np.array_equal(bp.obs_cam, oc)
        assert np.array_equal(bp.obs_xy, xy) and np.array_equal(bp.cams, cams) and np.array_equal(bp.lms, lms)
        ref = orc.load_bal(str(path), normalize=normalize)
        assert np.array_equal(bp.lm_off, ref["lm_off"]) and np.array_equal(bp.obs_cam, ref["obs_cam"]) and np.array_equal(bp.obs_xy, ref["obs_xy"])
        assert rel_err(bp.cams, ref["cams"]) < 1e-14 and rel_err(bp.lms, ref["lms"]) < 1e-14
    assert set(bp.load_timings) == {"read", "count", "parse", "csr", "normalize"}
    f32 = rb.BalProblem.load_bal(str(path), np.float32)
    assert f32.cams.dtype == np.float32 and np.array_equal(f32.cams, bp.cams.astype(np.float32))  # cast after normalising (:813-832)
    with pytest.raises(rb.RbaError):
        rb.BalProblem.load_bal(str(tmp_path / "missing.txt"))


@pytest.mark.parametrize("threshold", [0.0, 60.0, 95.0, 1e6])
def test_filter_obs_matches_oracle_and_numpy(tmp_path, threshold):
    """BalProblem::filter_obs (bal_problem.cpp:471-505, BalDatasetOptions::init_depth_threshold) in the four places it exists --
    both C++ loaders (byte-compared), the C ABI (rba_bal_filter_obs via BalProblem.load_bal) and the oracle -- against a numpy
    statement of the rule: keep an observation iff z(T_c_w p_w) >= threshold, keep a landmark iff >= 2 observations are left."""
    import rootba_b200 as rb
    from oracle import oracle_py as orc
    from rootba_b200.synthetic import quat_to_rot, synth_bal, write_bal
    _build()
    prob = synth_bal(15, 400, 4.0, seed=4, normalize_scale=None)
    path = str(tmp_path / "p.txt")
    write_bal(prob, path)
    base = orc.load_bal(path, normalize=True)  # normalised, unfiltered: the input of filter_obs in the reference pipeline
    R = quat_to_rot(base["cams"][:, :4])
    lm_of_obs = np.repeat(np.arange(len(base["lm_off"]) - 1), np.diff(base["lm_off"]))
    z = np.einsum("kj,kj->k", R[base["obs_cam"], 2, :], base["lms"][lm_of_obs]) + base["cams"][base["obs_cam"], 6]
    keep_obs = z >= threshold if threshold > 0 else np.ones(len(z), bool)
    n_left = np.bincount(lm_of_obs[keep_obs], minlength=len(base["lm_off"]) - 1)
    keep_lm = n_left >= 2
    keep = keep_obs & keep_lm[lm_of_obs]
    want_off = np.concatenate([[0], np.cumsum(n_left[keep_lm])])
    if threshold in (60.0, 95.0):
        assert 0 < keep.sum() < len(keep) and keep_lm.sum() < len(keep_lm)  # the case really filters something
    got = orc.load_bal(path, normalize=True, init_depth_threshold=threshold)
    assert np.array_equal(got["lm_off"], want_off) and np.array_equal(got["obs_cam"], base["obs_cam"][keep])
    assert np.array_equal(got["obs_xy"], base["obs_xy"][keep]) and np.array_equal(got["lms"], base["lms"][keep_lm])
    dumps = {}
    for loader in ("parallel", "map"):
        out = str(tmp_path / f"{loader}.bin")
        subprocess.check_call([BAL_QR, "--input", path, "--loader", loader, "--init-depth-threshold", str(threshold), "--dump-problem", out],
                              stdout=subprocess.DEVNULL)
        dumps[loader] = open(out, "rb").read()
    assert dumps["parallel"] == dumps["map"]
    cams, lms, off, oc, xy = _read_dump(str(tmp_path / "parallel.bin"))
    assert np.array_equal(off, want_off) and np.array_equal(oc, got["obs_cam"]) and np.array_equal(xy, got["obs_xy"])
    bp = rb.BalProblem.load_bal(path, np.float64, init_depth_threshold=threshold)
    assert np.array_equal(bp.lm_off, off) and np.array_equal(bp.obs_cam, oc) and np.array_equal(bp.lms, lms) and np.array_equal(bp.cams, cams)


def _write_bundler(prob, path, rng, dead_cams):
    """`prob` as a Bundler "bundle.out" v0.3 file (inverse of load_bundler, bal_problem.cpp:284-404): cameras listed in `dead_cams`
    are written with focal length 0 (uninitialised) and still referenced by some view lists; views in random order."""
    from rootba_b200.synthetic import quat_to_rot
    flip = np.diag([1.0, -1.0, -1.0])
    R = np.einsum("ij,mjk->mik", flip, quat_to_rot(prob.cams[:, :4]))
    t = prob.cams[:, 4:7] @ flip.T
    file_idx, lines = [], ["# Bundle file v0.3", None]
    n_file = 0
    for c in range(prob.nc):
        while n_file in dead_cams:  # an uninitialised camera before camera c
            lines += ["0 0 0", "0 0 0", "0 0 0", "0 0 0", "0 0 0"]
            n_file += 1
        file_idx.append(n_file)
        lines.append(" ".join(f"{v:.17g}" for v in prob.cams[c, 7:10]))
        lines += [" ".join(f"{v:.17g}" for v in R[c, r]) for r in range(3)]
        lines.append(" ".join(f"{v:.17g}" for v in t[c]))
        n_file += 1
    lines[1] = f"{n_file} {prob.nl}"
    for l in range(prob.nl):
        lines.append(" ".join(f"{v:.17g}" for v in prob.lms[l]))
        lines.append("255 0 0")
        views = [(file_idx[int(prob.obs_cam[k])], k) for k in range(int(prob.lm_off[l]), int(prob.lm_off[l + 1]))]
        order = rng.permutation(len(views))
        items = [f"{views[i][0]} {7 * views[i][1]} {prob.obs_xy[views[i][1], 0]:.17g} {-prob.obs_xy[views[i][1], 1]:.17g}" for i in order]
        if dead_cams and l % 3 == 0:
            items.insert(0, f"{sorted(dead_cams)[0]} 1 0.5 0.5")  # a view of an uninitialised camera: must be ignored
        lines.append(f"{len(items)} " + " ".join(items))
    open(path, "w").write("\n".join(lines) + "\n")


@pytest.mark.parametrize("dead_cams", [(), (0, 3)])
def test_bundler_loader(tmp_path, dead_cams):
    """Bundler "bundle.out" files (bal_problem.cpp:284-404; chosen by file name like autodetect_input_type :124-135): the three
    loaders (SoA / C ABI, reference-style AoS, oracle) against the generator's arrays; uninitialised cameras are dropped and the
    remaining ones renumbered; rotation via Eigen's matrix -> quaternion conversion."""
    import rootba_b200 as rb
    from oracle import oracle_py as orc
    from rootba_b200.synthetic import synth_bal
    _build()
    prob = synth_bal(9, 120, 3.5, seed=12, normalize_scale=None)
    path = str(tmp_path / "bundle.out")
    _write_bundler(prob, path, np.random.default_rng(0), set(dead_cams))
    dumps = {}
    for loader in ("parallel", "map"):
        out = str(tmp_path / f"{loader}.bin")
        subprocess.check_call([BAL_QR, "--input", path, "--loader", loader, "--no-normalize", "--dump-problem", out], stdout=subprocess.DEVNULL)
        dumps[loader] = open(out, "rb").read()
    assert dumps["parallel"] == dumps["map"]
    cams, lms, off, oc, xy = _read_dump(str(tmp_path / "parallel.bin"))
    assert np.array_equal(off, prob.lm_off) and np.array_equal(oc, prob.obs_cam)  # dead cameras gone, indices compacted
    assert np.array_equal(xy, prob.obs_xy) and np.array_equal(lms, prob.lms)      # %.17g round trip: exact
    q_sign = np.sign(np.sum(cams[:, :4] * prob.cams[:, :4], axis=1, keepdims=True))
    assert np.allclose(cams[:, :4] * q_sign, prob.cams[:, :4], atol=1e-14) and np.allclose(cams[:, 4:], prob.cams[:, 4:], rtol=1e-15, atol=1e-15)
    ref = orc.load_bal(path, normalize=False)
    assert np.array_equal(ref["lm_off"], off) and np.array_equal(ref["obs_cam"], oc) and np.array_equal(ref["obs_xy"], xy)
    assert rel_err(ref["cams"], cams) < 1e-15 and np.array_equal(ref["lms"], lms)
    bp = rb.BalProblem.load_bal(path, np.float64, normalize=True)   # C ABI, with normalisation
    refn = orc.load_bal(path, normalize=True)
    assert np.array_equal(bp.lm_off, off) and rel_err(bp.cams, refn["cams"]) < 1e-14 and rel_err(bp.lms, refn["lms"]) < 1e-14
    bad = tmp_path / "bundle_bad.out"
    bad.write_text("no comment line\n1 1\n")
    for loader in ("parallel", "map"):
        assert subprocess.run([BAL_QR, "--input", str(bad), "--loader", loader, "--dump-problem", str(tmp_path / "x.bin")], capture_output=True).returncode != 0


def _load_ba_log(path):
    """ba_log.json as the reference's python/rootba/log.py reads it: top-level columns -> numpy arrays, `_static` nested"""
    d = json.load(open(path))
    assert d["_type"] == "rootba"
    cols = {k: np.array(v) for k, v in d.items() if not k.startswith("_")}
    return cols, d["_static"]


def test_ba_log_has_the_reference_layout(tmp_path):
    """`bal_qr` writes ba_log.json in the reference's format (bal/ba_log.hpp:139-252, ba_log.cpp:62-150): the field lists
    come from tests/golden/ba_log_fields.json (extracted from the reference header by tests/golden/make_ba_log_fields.py);
    the values follow log_summary (bal/ba_log_utils.cpp:97-166) and finish_solve (bal_bundle_adjustment.cpp:117-140)."""
    _build()
    out = str(tmp_path / "ba_log.json")
    subprocess.check_call([BAL_QR, "--selftest-log", out])
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ba_log_fields.json")))["structs"]
    raw = json.load(open(out))
    cols, static = _load_ba_log(out)
    names = lambda s: [f["name"] for f in golden[s]]
    assert [k for k in raw if not k.startswith("_")] == names("BaIteration")       # same columns, same order
    assert list(static) == names("Static")
    assert list(static["problem_info"]) == names("ProblemInfo") and list(static["timing"]) == names("PipelineTiming")
    assert list(static["solver"]) == names("BaSolver")
    assert list(static["problem_info"]["per_lm_obs"]) == names("Stats") == list(static["problem_info"]["per_host_lms"])
    n = len(cols["iteration"])
    assert n == 4 and all(len(v) == n for v in cols.values())
    for f in golden["BaIteration"]:  # JSON types follow the C++ member types
        v = raw[f["name"]][0]
        want = {"bool": bool, "int": int, "uint64_t": int, "double": (int, float), "std::string": str}[f["type"]]
        assert isinstance(v, want) and (f["type"] == "bool" or not isinstance(v, bool)), f
    # the fabricated run: accepted, rejected, accepted.  A rejected iteration repeats the previous cost columns
    # ("for monotonic plots", ba_log_utils.cpp:119-137); cost_change is "previous logged cost - this cost"
    assert cols["step_is_successful"].tolist() == [True, True, False, True]
    assert cols["cost"].tolist() == [100.0, 40.0, 40.0, 39.99999]
    assert np.allclose(cols["cost_change"], [0, 60, 0, 55 - 39.99999])
    assert cols["relative_decrease"][2] == 0 and cols["linear_solver_type"].tolist() == ["", "bal_qr", "bal_qr", "bal_qr"]
    assert np.allclose(cols["step_solver_time"], cols["stage2_time"] + cols["solve_reduced_system_time"] + cols["back_substitution_time"])
    sol = static["solver"]
    assert (sol["solver_type"], sol["termination_type"], sol["num_successful_steps"], sol["num_unsuccessful_steps"]) == ("bal_qr", 0, 2, 1)
    assert sol["linear_solver_time_in_seconds"] == pytest.approx(cols["step_solver_time"].sum())
    pi = static["problem_info"]
    assert (pi["type"], pi["num_cameras"], pi["num_landmarks"], pi["num_observations"]) == ("bal", 3, 2, 5)
    assert pi["per_lm_obs"] == {"mean": 2.5, "min": 2, "max": 3, "stddev": 0.5} and pi["rcs_sparsity"] == 0
    assert pi["input_path"] == 'selftest "quoted" path'
    assert static["timing"]["total"] == pytest.approx(0.58)


def test_python_ba_log_writer_agrees_with_the_cpp_one(tmp_path):
    """rootba_b200.ba_log (Python host) and rootba_b200/host/ba_log.hpp (C++ host) on the same fabricated run"""
    import types
    import rootba_b200 as rb
    from rootba_b200.ba_log import BA_ITERATION_FIELDS
    _build()
    out = str(tmp_path / "cpp.json")
    subprocess.check_call([BAL_QR, "--selftest-log", out])
    cpp = json.load(open(out))
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ba_log_fields.json")))["structs"]
    assert list(BA_ITERATION_FIELDS) == [f["name"] for f in golden["BaIteration"]]
    costs, ok = [100.0, 40.0, 55.0, 39.99999], [True, True, False, True]
    its = []
    for i in range(4):  # the summary `bal_qr --selftest-log` fabricates (rootba_b200/host/bal_qr.cpp)
        its.append({"iteration": i, "cost": {"all": {"num_obs": 10, "error": costs[i], "residual_sum": 10 * np.sqrt(costs[i])},
                                             "valid": {"num_obs": 9, "error": 0.9 * costs[i], "residual_sum": 9 * np.sqrt(costs[i])}},
                    "step_is_valid": True, "step_is_successful": ok[i], "trust_region_radius": 1e4 * (i + 1),
                    "relative_decrease": 0.5 if i else 0.0, "linear_solver_iterations": 3 * i, "stage1_time": 0.001 if i else 0.0,
                    "stage2_time": 0.002 * i, "solve_reduced_system_time": 0.01 * i, "back_substitution_time": 0.0005 * i,
                    "iteration_time": 0.02, "cumulative_time": 0.02 * (i + 1)})
    summary = {"iterations": its, "termination_type": "CONVERGENCE", "message": "Function tolerance reached.", "num_linear_solves": 3,
               "num_residual_evaluations": 7, "num_jacobian_evaluations": 2, "total_time": 0.08, "minimizer_time": 0.07, "preprocessor_time": 0.01}
    prob = types.SimpleNamespace(lm_off=np.array([0, 2, 5]), obs_cam=np.array([0, 1, 0, 1, 2]), num_cameras=lambda: 3,
                                 num_landmarks=lambda: 2, num_observations=lambda: 5)
    py = rb.make_ba_log(summary, rb.summarize_problem(prob, 'selftest "quoted" path'), {"load": 0.5, "optimize": 0.08})
    py = json.loads(json.dumps(py))  # through JSON, like the file
    assert list(py) == list(cpp)
    for k in cpp:
        if k.startswith("_"):
            continue
        assert len(py[k]) == len(cpp[k]) == 4
        for a, b in zip(py[k], cpp[k]):
            assert (a == pytest.approx(b, rel=1e-15, abs=1e-300)) if isinstance(b, float) else (a == b and type(a) is type(b)), (k, a, b)
    assert py["_type"] == cpp["_type"]
    for sect in ("problem_info", "timing", "solver"):
        assert list(py["_static"][sect]) == list(cpp["_static"][sect])
        for k, b in cpp["_static"][sect].items():
            a = py["_static"][sect][k]
            assert a == (pytest.approx(b, rel=1e-12) if isinstance(b, float) else b), (sect, k, a, b)
    rb.save_ba_log(str(tmp_path / "py.json"), summary, prob, "x")
    assert json.load(open(tmp_path / "py.json"))["_static"]["problem_info"]["input_path"] == "x"


@pytest.mark.gpu
@pytest.mark.parametrize("use_double", [True, False])
def test_bal_qr_matches_python_host(tmp_path, use_double):
    import rootba_b200 as rb
    from oracle import oracle_py as orc
    from rootba_b200.synthetic import BalArrays, synth_bal, write_bal
    _build()
    prob = synth_bal(20, 400, 4.0, seed=9, normalize_scale=None)
    path = str(tmp_path / "p.txt")
    write_bal(prob, path)
    log = str(tmp_path / "ba_log.json")
    args = [BAL_QR, "--input", path, "--max-num-iterations", "4", "--log-path", log] + ([] if use_double else ["--no-use-double"])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    cols, static = _load_ba_log(log)
    d = orc.load_bal(path, normalize=True)
    arrays = BalArrays(d["cams"], d["lms"], d["lm_off"], d["obs_cam"], d["obs_xy"])
    dtype = np.float64 if use_double else np.float32
    bp = rb.BalProblem.from_arrays(arrays, dtype)
    summ = rb.bundle_adjust_manual(bp, rb.SolverOptions(max_num_iterations=4))
    assert len(cols["iteration"]) == len(summ["iterations"])
    assert static["problem_info"]["num_observations"] == arrays.nobs and static["solver"]["solver_type"] == "bal_qr"
    prev = None
    for k, b in enumerate(summ["iterations"]):
        # a rejected iteration repeats the previous cost in the log (ba_log_utils.cpp:119-137)
        cb = b["cost"]["all"]["error"] if (b["step_is_successful"] or prev is None) else prev
        prev = cb
        # the two hosts load the file with independent loaders (inputs differ in the last ulp), so the f64 trajectories
        # agree to ~1e-7 after a few LM iterations (threshold-based PCG stopping amplifies the ulp), not to 1e-9
        assert abs(cols["cost"][k] - cb) <= (1e-6 if use_double else 5e-3) * cb + 1e-12 * cols["cost"][0]
