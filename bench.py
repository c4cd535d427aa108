#!/usr/bin/env python
"""bench.py -- ms per LM iteration (linearize + QR + PCG + back-substitution) of the square-root BA inner loop.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A "step" is ONE Levenberg-Marquardt iteration of the reference's loop (solver/bal_bundle_adjustment.cpp:291-521):
[compute_error + linearize at a new linearization point] + solve(lambda) + apply + compute_error + accept/reject.
Both arms run the same LM trajectory from the same seeded synthetic problem, W warm-up iterations then K timed.
The workload is BASELINE.json configs[1] (BAL ladybug problem-1723-156502, float32) as a synthetic stand-in of
the same shape (no BAL file exists on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ms/LM-iter (linearize+QR+PCG) on BAL at 1/2/4/8 B200 vs CPU ref"


# ----------------------------------------------------------------------------------------------
# LM loop as a step machine (one call = one `it` of optimize_lm_ours)
# ----------------------------------------------------------------------------------------------
class LMStepper:
    def __init__(self, backend, dtype, initial_trust_region_radius=1e4, min_trust_region_radius=1e-32,
                 max_trust_region_radius=1e16, min_relative_decrease=0.0, initial_vee=2.0, vee_factor=2.0):
        self.b = backend
        self.S = np.float32 if np.dtype(dtype) == np.float32 else np.float64
        S = self.S
        self.min_lambda, self.max_lambda = S(1.0 / max_trust_region_radius), S(1.0 / min_trust_region_radius)
        self.vee_factor, self.initial_vee = S(vee_factor), S(initial_vee)
        self.lam, self.vee = S(1.0 / initial_trust_region_radius), S(initial_vee)
        self.min_relative_decrease = min_relative_decrease
        self.new_outer = True
        self.ri = None
        self.log = []

    def step(self):
        b, S = self.b, self.S
        rec = {"lambda": float(self.lam)}
        with np.errstate(all="ignore"):
            if self.new_outer:
                self.ri = b.compute_error()
                b.linearize()
                self.new_outer = False
            inc_ok = b.solve(float(self.lam))
            rec["cg_iterations"] = b.cg_iterations()
            if not inc_ok:
                self.lam = S(self.vee * self.lam); self.vee = S(self.vee * self.vee_factor)
                rec["accepted"] = False
                self.log.append(rec)
                return rec
            b.backup()
            l_diff = S(b.apply())
            ri2 = b.compute_error()
            ok = bool(np.isfinite(l_diff)) and ri2["is_numerically_valid"]
            success = False
            if ok:
                f_diff = S(self.ri["all"]["error"] - ri2["all"]["error"])
                q = S(f_diff / l_diff)
                success = bool(l_diff > 0 and q > self.min_relative_decrease)
            rec.update(cost=ri2["all"]["error"], accepted=success)
            if success:
                self.lam = max(self.min_lambda, S(self.lam * S(max(1.0 / 3, 1 - (2 * float(q) - 1) ** 3))))
                self.vee = self.initial_vee
                self.new_outer = True
            else:
                self.lam = S(self.vee * self.lam); self.vee = S(self.vee * self.vee_factor)
                b.restore()
        self.log.append(rec)
        return rec


class GpuBackend:
    """device-resident state and increment (what `value` measures)"""

    def __init__(self, lin):
        self.lin = lin

    def compute_error(self): return self.lin.compute_error()
    def linearize(self): self.lin.linearize()
    def solve(self, lam):
        self.lin.solve(lam, to_host=False)
        return self.lin.last_cg.termination_type != 2
    def cg_iterations(self): return self.lin.last_cg.num_iterations
    def backup(self): self.lin._backup()
    def apply(self): return self.lin.apply(None)
    def restore(self): self.lin._restore()


class GpuE2EBackend(GpuBackend):
    """every step moves its inputs host->device from pinned memory and its results device->host through the
    reference-facing API: state up, increment down, increment up (apply takes a host vector like
    LinearizorQR::apply(VecX&&)), state + l_diff + ResidualInfo down."""

    def __init__(self, lin):
        super().__init__(lin)
        import torch
        bp = lin.bal_problem
        # re-home the host arrays in pinned memory
        self._pin = [torch.from_numpy(bp.cams.copy()).pin_memory(), torch.from_numpy(bp.lms.copy()).pin_memory()]
        bp.cams, bp.lms = self._pin[0].numpy(), self._pin[1].numpy()
        self._bk = (bp.cams.copy(), bp.lms.copy())
        self._inc_pin = torch.zeros(9 * lin.nc, dtype=torch.float32 if lin.dtype == np.float32 else torch.float64).pin_memory()
        self.h2d = self.d2h = 0
        self.item = lin.dtype.itemsize

    def solve(self, lam):
        lin = self.lin
        lin.upload_state()
        self.h2d += (10 * lin.nc + 3 * lin.nl) * self.item
        inc = lin.solve(lam, to_host=True)
        self._inc_pin.numpy()[:] = inc
        self.d2h += 9 * lin.nc * self.item
        return bool(np.all(np.isfinite(inc)))

    def backup(self):
        bp = self.lin.bal_problem
        self._bk[0][:] = bp.cams; self._bk[1][:] = bp.lms
        self.lin._backup()

    def apply(self):
        lin = self.lin
        l = lin.apply(self._inc_pin.numpy())
        lin.download_state()
        self.h2d += 9 * lin.nc * self.item
        self.d2h += (10 * lin.nc + 3 * lin.nl) * self.item + self.item
        return l

    def compute_error(self):
        self.d2h += 56
        return self.lin.compute_error()

    def restore(self):
        bp = self.lin.bal_problem
        bp.cams[:] = self._bk[0]; bp.lms[:] = self._bk[1]
        self.lin._restore()


class OracleBackend:
    def __init__(self, o):
        self.o = o
        self._inc = None
        self._it = 0

    def compute_error(self): return self.o.compute_error()
    def linearize(self):
        if not self.o.linearize():
            raise RuntimeError("numerical failure during linearization")
    def solve(self, lam):
        self._inc, self._it = self.o.solve(lam)
        return bool(np.all(np.isfinite(self._inc)))
    def cg_iterations(self): return self._it
    def backup(self): self.o.backup()
    def apply(self): return self.o.apply(self._inc)
    def restore(self): self.o.restore()


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 8:
                    self.samples.append(f)
        except Exception:
            pass

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def host_threads() -> int:
    """threads for the CPU arm: one per physical core (the landmark loop is memory-bound; SMT siblings only add contention)"""
    try:
        import psutil
        n = psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        n = os.cpu_count()
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return max(1, int(n))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_problem(args):
    """the synthetic stand-in; with RBA_BENCH_CACHE_DIR set the generated arrays are kept there (npz) so that several
    bench invocations on one box (dense / implicit, the ranks of one torchrun) generate a large problem only once"""
    from rootba_b200.synthetic import BalArrays, synth_config
    cache = os.environ.get("RBA_BENCH_CACHE_DIR")
    path = os.path.join(cache, f"{args.workload}_s{args.seed}_x{args.scale}.npz") if cache else None
    if path and os.path.exists(path):
        z = np.load(path)
        return BalArrays(z["cams"], z["lms"], z["lm_off"], z["obs_cam"], z["obs_xy"])
    arrays = synth_config(args.workload, seed=args.seed, scale=args.scale)
    if path:
        os.makedirs(cache, exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp.npz"
        np.savez(tmp, cams=arrays.cams, lms=arrays.lms, lm_off=arrays.lm_off, obs_cam=arrays.obs_cam, obs_xy=arrays.obs_xy)
        os.replace(tmp, path)
    return arrays


def tune_oracle_threads(o, n_vec, dtype, max_threads):
    """The CPU restatement scatters under per-camera locks like the reference (reduction_alg=1), which stops scaling well
    before 64 threads; so that the CPU arm is not handicapped, time the PCG operator at a few thread counts (after the
    warm-up steps, when the oracle is linearised) and keep the fastest.  Returns the thread count left configured."""
    from oracle import oracle_py as orc
    best, best_t = max_threads, None
    try:
        x = np.ones(n_vec, dtype=dtype)
        cands = sorted({max(1, max_threads >> k) for k in range(4)} | {min(8, max_threads)})
        for T in cands:
            o.set_options(orc.default_options(num_threads=T))
            o.right_multiply(x)
            t0 = time.perf_counter()
            for _ in range(3):
                o.right_multiply(x)
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best, best_t = T, dt
    except Exception:  # the sweep is a courtesy to the CPU arm; never let it take the measurement down
        best = max_threads
    o.set_options(orc.default_options(num_threads=best))
    return best


def run_lm(backend, dtype, warmup, steps, timer=None, barrier=None, after_warmup=None):
    """returns (seconds for the K timed steps, stepper)"""
    st = LMStepper(backend, dtype)
    for _ in range(warmup):
        st.step()
    if after_warmup:
        after_warmup()
    if barrier:
        barrier()
    t0 = time.perf_counter()
    if timer:
        timer[0]()
    for _ in range(steps):
        st.step()
    dev = timer[1]() if timer else None
    wall = time.perf_counter() - t0
    if barrier:
        barrier()
    return (dev if dev is not None else wall), wall, st


def bench_reference(args):
    """the reference's algorithm (CPU restatement, oracle/) on the host cores; rank 0 only"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as orc
    dtype = np.float32 if args.dtype == "f32" else np.float64
    arrays = make_problem(args)
    cores = host_threads()
    o = orc.Oracle(arrays, dtype, orc.default_options(num_threads=cores))
    used = [cores]
    secs, wall, st = run_lm(OracleBackend(o), dtype, args.warmup, args.steps,
                            after_warmup=lambda: used.__setitem__(0, tune_oracle_threads(o, 9 * arrays.nc, dtype, cores)))
    avail, cores = cores, used[0]
    ms = 1e3 * secs / args.steps
    st_ = arrays.stats()
    out = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms/LM-iter", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"synthetic {args.workload} (BAL ladybug problem-1723-156502 shape)" if args.workload == "ladybug-1723" else f"synthetic {args.workload}",
                   "scale": args.scale, "seed": args.seed, **st_, "solver": "SQUARE_ROOT/SCHUR_JACOBI/Householder, reference defaults"},
        "cpu_baseline": {"value": ms, "unit": "ms/LM-iter", "cores": cores, "kind": "port",
                         "sample": f"LM iterations {args.warmup + 1}..{args.warmup + args.steps} of the same trajectory, OpenMP over landmarks "
                                   f"with per-camera locks (reduction_alg=1); CPU restatement of the reference (reference itself not buildable here); "
                                   f"{cores} of {avail} physical cores = the fastest of a short sweep of the PCG operator"},
        "e2e": {"value": ms, "unit": "ms/LM-iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cg_iterations": [r.get("cg_iterations") for r in st.log[args.warmup:]],
        "accepted": [bool(r.get("accepted")) for r in st.log[args.warmup:]],
        "final_cost": st.log[-1].get("cost"),
    }
    # "run with T = all cores and T = 1; report both".  One LM iteration (the first timed one) on a single thread,
    # reached with the same warm-up on all cores; bounded to the cheapest timed iteration so the arm stays within minutes.
    if not args.no_single_thread:
        o1 = orc.Oracle(arrays, dtype, orc.default_options(num_threads=cores))
        st1 = LMStepper(OracleBackend(o1), dtype)
        for _ in range(args.warmup):
            st1.step()
        o1.set_options(orc.default_options(num_threads=1))
        t0 = time.perf_counter()
        st1.step()
        out["cpu_baseline_1thread"] = {"value": 1e3 * (time.perf_counter() - t0), "unit": "ms/LM-iter", "cores": 1, "kind": "port",
                                       "sample": f"LM iteration {args.warmup + 1} only", "cg_iterations": st1.log[-1].get("cg_iterations")}
    print(json.dumps(out), flush=True)


def bench_ours(args):
    import torch
    import rootba_b200 as rb
    dtype = np.float32 if args.dtype == "f32" else np.float64
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    arrays = make_problem(args)

    def make_linearizor():
        bp = rb.BalProblem.from_arrays(arrays, dtype)
        so = rb.SolverOptions(use_double=(dtype == np.float64), device=local_rank, rank=rank, nranks=world,
                              operator_form=args.operator.upper())
        lin = rb.LinearizorQR.create(bp, so)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(rb.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            lin.comm_init(bytes(uid.cpu().numpy().tobytes()))
            # peer-memory all-reduce fused into the PCG vector kernel: exchange the CUDA IPC handles
            mine = torch.frombuffer(bytearray(lin.ipc_export()), dtype=torch.uint8).cuda()
            allh = [torch.zeros(128, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            lin.ipc_import(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))
        return lin

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident run (value) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lin = make_linearizor()
    l0 = lin.timings()["kernel_launches"]
    # warm-up outside, then K timed steps bracketed by barrier + synchronize, CUDA events on the solver stream
    st = LMStepper(GpuBackend(lin), dtype)
    for _ in range(args.warmup):
        st.step()
    barrier()
    l0 = lin.timings()["kernel_launches"]
    lin.timer_start()
    t0 = time.perf_counter()
    phase = {k: 0.0 for k in ("stage1_time", "stage2_time", "compute_preconditioner_time", "solve_reduced_system_time",
                              "back_substitution_time", "update_cameras_time", "residual_evaluation_time")}
    for _ in range(args.steps):
        was_outer = st.new_outer
        st.step()
        t = lin.timings()
        for k in phase:
            if k == "stage1_time" and not was_outer:
                continue
            phase[k] += t[k]
    dev_s = lin.timer_stop()
    wall_s = time.perf_counter() - t0
    barrier()
    launches = lin.timings()["kernel_launches"] - l0
    secs = torch.tensor([dev_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
    ms = 1e3 * float(secs.item()) / args.steps
    # ---- roofline of the dominant kernel (rcs_matvec), timed alone with CUDA events ----
    mv_s = lin.time_matvec(200)
    stats = lin.stats()
    clocks = sampler.stop() if rank == 0 else {}
    mv_t = torch.tensor([mv_s], dtype=torch.float64, device="cuda")
    sz = 4 if dtype == np.float32 else 8
    dense_bytes = float(stats["matvec_algorithmic_bytes"])  # 18 M2 s + 18 Nobs s + 4 Nobs (dense Q2 panels)
    # implicit form: jp (18) + q1d (27) records + x gather (9) + per-observation y written and read back (9 + 9), cam index
    implicit_bytes = float(stats["num_observations_local"]) * (72 * sz + 4)
    mv_bytes = torch.tensor([implicit_bytes if args.operator == "implicit" else dense_bytes], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mv_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(mv_bytes, op=dist.ReduceOp.SUM)
    peak, peak_src = hbm_peak()
    traffic = None
    try:  # DRAM bytes of the dominant kernel from the committed ncu --set full capture (same workload), per launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_matvec_traffic.json")))
        if args.workload == "ladybug-1723" and args.scale == 1.0 and args.dtype == "f32" and world == 1 and args.operator == "dense":
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
    achieved = float(mv_bytes.item()) / float(mv_t.item()) / 1e9 / world  # per GPU
    cg_log = [r.get("cg_iterations") for r in st.log[args.warmup:]]
    final_cost = st.log[-1].get("cost")
    lin.close()
    # ---- end-to-end run through the host-buffer API ----
    lin2 = make_linearizor()
    be = GpuE2EBackend(lin2)
    st2 = LMStepper(be, dtype)
    for _ in range(args.warmup):
        st2.step()
    barrier()
    be.h2d = be.d2h = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st2.step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = 1e3 * float(e2e_t.item()) / args.steps
    h2d, d2h = be.h2d // args.steps, be.d2h // args.steps
    lin2.close()
    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_py as orc
        o = orc.Oracle(arrays, dtype, orc.default_options(num_threads=host_threads()))
        n_cpu = min(args.steps, 2)
        used = [host_threads()]
        cs, _, stc = run_lm(OracleBackend(o), dtype, args.warmup, n_cpu,
                            after_warmup=lambda: used.__setitem__(0, tune_oracle_threads(o, 9 * arrays.nc, dtype, host_threads())))
        cpu = {"value": 1e3 * cs / n_cpu, "unit": "ms/LM-iter", "cores": used[0], "cores_available": host_threads(), "kind": "port",
               "sample": f"LM iterations {args.warmup + 1}..{args.warmup + n_cpu} of the same trajectory on the host cores "
                         f"(CPU restatement of the reference, OpenMP over landmarks; warm-up iterations untimed)",
               "cg_iterations": [r.get("cg_iterations") for r in stc.log[args.warmup:]]}
    if rank == 0:
        full = arrays.stats()
        out = {
            "metric": METRIC, "value": ms, "unit": "ms/LM-iter", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": f"synthetic {args.workload}" + (" (BAL ladybug problem-1723-156502 shape, BASELINE configs[1])" if args.workload == "ladybug-1723" else ""),
                       "scale": args.scale, "seed": args.seed, **full, "parallelism": f"landmark-shard x{world}",
                       "solver": "SQUARE_ROOT/SCHUR_JACOBI/Householder, reference defaults" + (", operator_form=IMPLICIT (opt-in)" if args.operator == "implicit" else ""),
                       "l2": "inputs larger than L2 (Q2 panels %.0f MB vs 126 MB L2)" % (stats["panel_scalars"] * (4 if dtype == np.float32 else 8) / 1e6)},
            "e2e": {"value": e2e_ms, "unit": "ms/LM-iter", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": clocks.get("reasons", [])},
            "roofline": {"bound": "hbm", "kernel": "rcs_matvec (k_matvec_implicit + k_cam_reduce), bytes of the implicit form" if args.operator == "implicit"
                         else "rcs_matvec (k_matvec_small + k_cam_reduce)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": float(mv_bytes.item()) / world, "us_per_launch": 1e6 * float(mv_t.item())},
            "cpu_baseline": cpu,
            "phases_ms_per_step": {k: 1e3 * v / args.steps for k, v in phase.items()},
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "cg_iterations": cg_log, "final_cost": final_cost,
        }
        # also report microseconds per PCG iteration (solve_reduced_system_time over the PCG iterations)
        n_cg = sum(int(c) for c in cg_log if c is not None)
        out["pcg"] = {"iterations": n_cg,
                      "us_per_iteration": (1e6 * phase["solve_reduced_system_time"] / n_cg) if n_cg > 0 else None}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ladybug-1723")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=38401)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-thread", action="store_true", help="reference arm: skip the extra single-thread LM iteration")
    ap.add_argument("--operator", default="dense", choices=["dense", "implicit"],
                    help="PCG operator form: dense = the reference's Q2-panel product (default, contract kernel); implicit = opt-in")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
