"""N > 1 host logic on CPU (gloo, world size 2): landmark sharding by rba_partition_landmarks and the reductions the
sharded solver performs (column norms, gradient, preconditioner blocks, operator output, residual info, l_diff) give the
single-rank result.  The per-shard arithmetic is done by the CPU oracle (test infrastructure); the GPU path performs the
same all-reduces with NCCL at the same points (rootba_b200/csrc/solver.cu: Solver::allreduce)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rel_err


def _shard(arrays, b0, b1):
    from rootba_b200.synthetic import BalArrays
    o0, o1 = arrays.lm_off[b0], arrays.lm_off[b1]
    return BalArrays(arrays.cams.copy(), arrays.lms[b0:b1].copy(), (arrays.lm_off[b0:b1 + 1] - o0).copy(),
                     arrays.obs_cam[o0:o1].copy(), arrays.obs_xy[o0:o1].copy())


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rootba_b200 as rb
    from oracle import oracle_py as orc
    from rootba_b200.synthetic import synth_bal
    arrays = synth_bal(30, 900, 4.2, seed=21)
    bounds = rb.partition_landmarks(arrays.lm_off, world)
    sh = _shard(arrays, bounds[rank], bounds[rank + 1])
    o = orc.Oracle(sh, np.float64)

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    res = {}
    e = o.compute_error()
    res["error"] = allreduce(np.array([e["all"]["num_obs"], e["all"]["error"], e["valid"]["num_obs"]], dtype=np.float64))
    d2, ok = o.stage1()
    d2 = allreduce(d2)
    scaling = 1.0 / (1e-5 + np.sqrt(d2))
    b, blocks = o.stage2(0.1, scaling, schur_blocks=True)
    res["diag2"], res["b"], res["blocks"] = d2, allreduce(b), allreduce(blocks.ravel())
    x = np.random.default_rng(0).uniform(-1, 1, 9 * arrays.nc)
    res["Hx"] = allreduce(o.right_multiply(x))
    inc = np.random.default_rng(1).uniform(-1, 1, 9 * arrays.nc) * 0.01
    l, ok = o.back_substitute(inc)
    res["l_diff"] = allreduce(np.array([l]))
    _, lms = o.get_state()
    full = np.zeros((arrays.nl, 3))
    full[bounds[rank]:bounds[rank + 1]] = lms
    res["lms"] = allreduce(full)
    if rank == 0:
        np.savez(out_path, bounds=bounds, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    from oracle import oracle_py as orc
    from rootba_b200.synthetic import synth_bal
    orc.build()
    out = str(tmp_path / "r.npz")
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    arrays = synth_bal(30, 900, 4.2, seed=21)
    o = orc.Oracle(arrays, np.float64)
    e = o.compute_error()
    assert got["error"][0] == e["all"]["num_obs"] and got["error"][2] == e["valid"]["num_obs"]
    assert abs(got["error"][1] - e["all"]["error"]) <= 1e-12 * e["all"]["error"]
    d2, ok = o.stage1()
    assert rel_err(got["diag2"], d2) < 1e-13
    b, blocks = o.stage2(0.1, 1.0 / (1e-5 + np.sqrt(d2)), schur_blocks=True)
    assert rel_err(got["b"], b) < 1e-12 and rel_err(got["blocks"], blocks) < 1e-12
    x = np.random.default_rng(0).uniform(-1, 1, 9 * arrays.nc)
    assert rel_err(got["Hx"], o.right_multiply(x)) < 1e-12
    inc = np.random.default_rng(1).uniform(-1, 1, 9 * arrays.nc) * 0.01
    l, ok = o.back_substitute(inc)
    assert abs(got["l_diff"][0] - l) <= 1e-11 * abs(l)
    _, lms = o.get_state()
    assert rel_err(got["lms"], lms) < 1e-14
    # shards are contiguous, cover everything, and balance sum(n^2 + 4n)
    bnd = got["bounds"]
    n = np.diff(arrays.lm_off).astype(np.float64)
    w = n * n + 4 * n
    assert bnd[0] == 0 and bnd[-1] == arrays.nl
    loads = [w[bnd[r]:bnd[r + 1]].sum() for r in range(2)]
    assert max(loads) <= 1.05 * np.mean(loads) + w.max()
