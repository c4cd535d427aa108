#!/usr/bin/env python
"""Time the two BAL loaders of the C++ host (rootba_b200/host) on a synthetic file of a BASELINE shape.

  python scripts/bench_loader.py --workload venice-1778 [--threads 0] [--keep]

Writes the BAL text file (same '%.6e'-style numbers as the real BAL files) under $TMPDIR, runs
`bal_qr --dump-problem` with --loader map (reference-style fscanf + std::map, bal_problem.cpp:189-282) and
--loader parallel (bal_io_fast.hpp), checks that both dumps are byte-identical and prints one JSON line.
Host-only: no GPU is touched (--dump-problem returns before the solver is created)."""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BAL_QR = os.path.join(ROOT, "rootba_b200", "host", "bal_qr")


def write_bal_fast(prob, path):
    from scipy.spatial.transform import Rotation
    from rootba_b200.synthetic import quat_to_rot
    flip = np.diag([1.0, -1.0, -1.0])
    rv = Rotation.from_matrix(np.einsum("ij,mjk->mik", flip, quat_to_rot(prob.cams[:, :4]))).as_rotvec()
    tb = prob.cams[:, 4:7] @ flip.T
    lm_of_obs = np.repeat(np.arange(prob.nl), np.diff(prob.lm_off))
    with open(path, "w") as f:
        f.write(f"{prob.nc} {prob.nl} {prob.nobs}\n")
    obs = pd.DataFrame({"c": prob.obs_cam, "l": lm_of_obs, "x": prob.obs_xy[:, 0], "y": -prob.obs_xy[:, 1]})
    obs.to_csv(path, sep=" ", header=False, index=False, float_format="%.6e", mode="a")
    tail = np.concatenate([np.concatenate([rv, tb, prob.cams[:, 7:10]], axis=1).ravel(), prob.lms.ravel()])
    pd.DataFrame({"v": tail}).to_csv(path, header=False, index=False, float_format="%.16e", mode="a")


def run(loader, path, out, threads):
    t0 = time.time()
    r = subprocess.run([BAL_QR, "--input", path, "--loader", loader, "--num-threads", str(threads), "--dump-problem", out],
                       capture_output=True, text=True, check=True)
    wall = time.time() - t0
    m = re.search(r"load time ([0-9.]+)s", r.stdout)
    return float(m.group(1)), wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="venice-1778")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    from rootba_b200.synthetic import synth_config
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "rootba_b200", "host"), "-s"])
    prob = synth_config(args.workload, scale=args.scale)
    tmp = tempfile.mkdtemp(prefix="rba_loader_")
    path = os.path.join(tmp, "problem.txt")
    write_bal_fast(prob, path)
    res = {"workload": args.workload, "num_cameras": int(prob.nc), "num_landmarks": int(prob.nl), "num_observations": int(prob.nobs),
           "file_mb": round(os.path.getsize(path) / 1e6, 1), "host_threads": os.cpu_count() if args.threads <= 0 else args.threads}
    digests = {}
    for loader in ("parallel", "map"):
        out = os.path.join(tmp, loader + ".bin")
        run(loader, path, out, args.threads)  # warm the page cache / first-touch
        load_s, wall = run(loader, path, out, args.threads)
        res[f"{loader}_load_s"] = load_s
        h = hashlib.sha256()
        with open(out, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        digests[loader] = h.hexdigest()
        os.remove(out)
    res["identical"] = digests["parallel"] == digests["map"]
    res["speedup"] = round(res["map_load_s"] / res["parallel_load_s"], 2)
    if not args.keep:
        os.remove(path)
        os.rmdir(tmp)
    print(json.dumps(res))
    return 0 if res["identical"] else 1


if __name__ == "__main__":
    sys.exit(main())
