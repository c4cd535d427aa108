"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun), landmarks sharded over the ranks, compared on
rank 0 with the single-rank CPU oracle.  Usage: torchrun --nproc-per-node N multirank_gpu_worker.py <out.json> <f32|f64>"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(a) + np.linalg.norm(b) + 1e-300))


def main():
    out_path, sfx = sys.argv[1], sys.argv[2]
    dtype = np.float32 if sfx == "f32" else np.float64
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import rootba_b200 as rb
    from rootba_b200.synthetic import synth_bal
    # sequence-like visibility: every shard misses most cameras (the case the in-place all-reduce must get right)
    arrays = synth_bal(300, 6000, 4.5, seed=17, locality=2.0, max_track=40)
    # landmarks in order of their first camera: a shard (contiguous landmark range) then sees only part of the cameras
    from rootba_b200.synthetic import BalArrays
    n = np.diff(arrays.lm_off)
    order = np.argsort(arrays.obs_cam[arrays.lm_off[:-1]], kind="stable")
    off = np.concatenate([[0], np.cumsum(n[order])]).astype(np.int64)
    idx = np.concatenate([np.arange(arrays.lm_off[l], arrays.lm_off[l + 1]) for l in order])
    arrays = BalArrays(arrays.cams, arrays.lms[order].copy(), off, arrays.obs_cam[idx].copy(), arrays.obs_xy[idx].copy())
    bp = rb.BalProblem.from_arrays(arrays, dtype)
    so = rb.SolverOptions(device=local, rank=rank, nranks=world)
    lin = rb.LinearizorQR.create(bp, so)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(rb.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    lin.comm_init(bytes(uid.cpu().numpy().tobytes()))
    mine = torch.frombuffer(bytearray(lin.ipc_export()), dtype=torch.uint8).cuda()
    allh = [torch.zeros(len(mine), dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine)
    lin.ipc_import(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))  # no-op with RBA_PEER_AR=0
    st = lin.stats()
    # a camera without observations in this shard must exist for the test to mean anything
    o0, o1 = arrays.lm_off[st["landmark_begin"]], arrays.lm_off[st["landmark_end"]]
    cams_here = np.unique(arrays.obs_cam[o0:o1])
    res = {"rank": rank, "world": world, "peer": os.environ.get("RBA_PEER_AR", "1"), "cams_without_obs_here": int(arrays.nc - cams_here.size)}
    e = lin.compute_error()
    lin.linearize()
    lam = 1e-3
    inc = lin.solve(lam)
    b = lin.get_rhs()
    inv, _ = lin.get_preconditioner()
    x = np.random.default_rng(3).uniform(-1, 1, 9 * lin.nc).astype(dtype)
    Hx = lin.right_multiply(x)
    Hx2 = lin.right_multiply(x)  # a second application must not see stale sums
    cg_it, cg_term = lin.last_cg.num_iterations, lin.last_cg.termination_type
    l_diff = lin.apply(inc)
    lin.download_state()  # cameras (replicated) + this shard's landmarks
    lms = torch.from_numpy(np.where(np.arange(arrays.nl)[:, None] >= st["landmark_begin"], bp.lms, 0) *
                           (np.arange(arrays.nl)[:, None] < st["landmark_end"])).double().cuda()
    dist.all_reduce(lms)
    e2 = lin.compute_error()
    # every rank must hold the same replicated vectors
    chk = torch.from_numpy(np.concatenate([inc, b, Hx]).astype(np.float64)).cuda()
    mx, mn = chk.clone(), chk.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    res["replicas_identical"] = bool(torch.equal(mx, mn))
    if rank == 0:
        from oracle import oracle_py as orc
        o = orc.Oracle(arrays, dtype, orc.default_options(num_threads=0))
        ec = o.compute_error()
        o.linearize()
        inc_c, dbg = o.solve(lam, want_debug=True)
        res.update(
            error=[e["all"]["num_obs"], ec["all"]["num_obs"], abs(e["all"]["error"] - ec["all"]["error"]) / ec["all"]["error"]],
            b=rel(b, dbg["b"]), inv=max(rel(inv[c], dbg["inv_blocks"][c]) for c in range(lin.nc)),
            Hx=rel(Hx, o.right_multiply(x)), Hx_repeat_identical=bool(np.array_equal(Hx, Hx2)),
            inc=rel(inc, inc_c), cg=[cg_it, dbg["cg_iterations"]], cg_term=[cg_term, dbg["cg_termination"]])
        l_c = o.apply(inc_c)
        cams_c, lms_c = o.get_state()
        ec2 = o.compute_error()
        res.update(l_diff=abs(l_diff - l_c) / abs(l_c), lms=rel(lms.cpu().numpy(), lms_c), cams=rel(bp.cams, cams_c),
                   error_after=abs(e2["all"]["error"] - ec2["all"]["error"]) / ec2["all"]["error"])
        with open(out_path, "w") as f:
            json.dump(res, f)
    lin.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
