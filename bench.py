#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native batched IDA engine.

Metric (BASELINE.json): ensemble DAE solves/sec -- Roberts (n = 3), rate constants perturbed +-10 % per
instance, each instance integrated from t = 0 to t = 4e10 through the reference's 12-decade output ladder
(tout = 0.4 * 10^i, examples/roberts.rs) -- instances finished per second, whole job over all GPUs.

A "step" is one pass of the hot path over one batch: re-initialise the ensemble state (Ida::new) and
integrate every instance to t = 4e10 (one solve-ladder kernel launch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

    value      instances/s with params / yy0 / yp0 already resident in HBM
    e2e        the same through the C-ABI host-buffer path: H2D of the inputs from pinned memory, the
               solve, D2H of the final states and step counts, every step
    roofline   FP64-issue roofline of the solve kernel (n = 3 ensembles are FP64-pipe-bound, not HBM-bound;
               SURVEY.md 8d): algorithmic FP64 ops / kernel time vs the measured un-fused FP64 issue rate
    cpu_baseline  the CPU oracle (a C++ restatement of rust-ida, NOT rust-ida itself: there is no Rust
               toolchain here) on a bounded sample, one instance per thread on all host cores

N > 1 (torchrun): one process per GPU, instances sharded by contiguous id range (weak scaling: --batch
instances per GPU), no collective inside the step; one NCCL all_gather of the per-instance step counts at
the end of every step.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RTOL = 1e-4
ATOL = np.array([1e-8, 1e-6, 1e-6])
TOUTS = 0.4 * 10.0 ** np.arange(12)
SEED = 1234
# FP64 arithmetic per step attempt of the reference algorithm at n = 3 (each add, mul, div, sqrt, pow counted
# as ONE op): 98,244 ops / 377 attempts for the default Roberts instance (SURVEY.md 6 / BASELINE.md 2).
OPS_PER_ATTEMPT = 98244.0 / 377.0
# algorithmic HBM bytes per instance and launch: the SoA state (10n + 45 + n^2 doubles + 16 + n int32, n = 3) is read
# once and written once
ALG_BYTES_PER_INSTANCE = 2 * ((10 * 3 + 45 + 9) * 8 + (16 + 3) * 4)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1 << 20, help="instances per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---- clocks sampler (nvidia-smi during the timed region) -----------------------------------------------------
class ClockSampler:
    """one `nvidia-smi -lms 50` process for the duration of the timed region (B200_PROFILING.md clocks line)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)  # let the first samples land before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            try:
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(self.NAMES, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU arm: the oracle port on the host cores --------------------------------------------------------------
def cpu_run(sample, first_id=0):
    """integrate `sample` perturbed Roberts instances to t=4e10 with the CPU oracle on all cores -> (inst/s, cores)"""
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params
    cores = os.cpu_count() or 1
    p = roberts_ensemble_params(sample, seed=SEED, first_id=first_id)
    yy0, yp0 = Roberts.consistent_ic(p)
    t0 = time.perf_counter()
    out = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, RTOL, ATOL, TOUTS, nthreads=cores)
    dt = time.perf_counter() - t0
    assert (out["status"] == 0).all()
    return sample / dt, cores, dt


def auto_cpu_sample(target_s=12.0):
    rate, cores, _ = cpu_run(256 * (os.cpu_count() or 1))
    n = int(rate * target_s)
    return max(1024, min(n, 1 << 20))


def workload_config(B):
    """the `config` object both arms print (the reference arm must run on our arm's config)"""
    return {"workload": "Roberts ensemble, n=3, %d instances per GPU, rate constants +-10%% (splitmix64 seed 1234 "
                        "keyed by instance id), rtol 1e-4, atol [1e-8,1e-6,1e-6], tout = 0.4*10^i i=0..11 "
                        "(BASELINE.json configs[1])" % B,
            "batch_per_gpu": B, "sharding": "contiguous instance-id ranges, no collective in the step loop",
            "l2": "per-step state (%.0f MB/GPU) is larger than the 126 MB L2" % (B * 0.78e-3)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for the same metric/config.  rust-ida cannot be built here
    (no Rust toolchain, tree mid-refactor), so this is the oracle port, all host threads, bounded sample."""
    if rank != 0:
        return
    sample = args.cpu_sample or auto_cpu_sample(6.0)
    for _ in range(args.warmup):
        cpu_run(max(256, sample // 8))
    t0 = time.perf_counter()
    tot = 0
    for k in range(args.steps):
        cpu_run(sample, first_id=k * sample)
        tot += sample
    dt = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    val = tot / dt
    desc = "%d perturbed Roberts instances per step (of the %d-per-GPU workload), to t=4e10" % (sample, args.batch)
    line = {
        "impl": "reference", "metric": "ensemble DAE solves/sec (Roberts, t=4e10)", "value": val, "unit": "instances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.batch),
        "cpu_baseline": {"value": val, "unit": "instances/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- GPU arm -----------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import rust_ida_b200 as r
    from rust_ida_b200 import _lib
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    first_id = rank * B  # contiguous id range per rank; parameters are keyed by instance id
    p_h = roberts_ensemble_params(B, seed=SEED, first_id=first_id)
    yy0_h, yp0_h = Roberts.consistent_ic(p_h)

    # pinned host staging for the end-to-end path
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t

    p_pin, yy0_pin, yp0_pin = pinned(p_h), pinned(yy0_h), pinned(yp0_h)
    yy_out_pin = torch.empty((3, B), dtype=torch.float64, pin_memory=True)
    nst_out_pin = torch.empty((B,), dtype=torch.int64, pin_memory=True)

    p_d, yy0_d, yp0_d = (torch.from_numpy(v).to(dev) for v in (p_h, yy0_h, yp0_h))
    tol = r.TolControlSV(RTOL, ATOL)
    ida = r.Ida(Roberts(p_d), yy0_d, yp0_d, tol, device=local_rank)
    status_d = torch.empty((len(TOUTS), B), dtype=torch.int32, device=dev)
    nst_d = torch.empty((B,), dtype=torch.int64, device=dev)
    gathered = torch.empty((world * B,), dtype=torch.int64, device=dev) if world > 1 else None
    L = _lib.lib()
    h = ida._h
    touts = np.ascontiguousarray(TOUTS)
    null = None

    ev_k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def step_device(k=None):
        _lib.check(L.b200ida_reinit(h, null, yy0_d.data_ptr(), yp0_d.data_ptr(), _lib.MEM_DEVICE))
        if k is not None:
            ev_k0[k].record()
        _lib.check(L.b200ida_solve_ladder(h, touts.ctypes.data, len(touts), 0, null, status_d.data_ptr(), _lib.MEM_DEVICE))
        if k is not None:
            ev_k1[k].record()
        _lib.check(L.b200ida_get_counters(h, nst_d.data_ptr(), null, null, null, null, null, null, null, _lib.MEM_DEVICE))
        if world > 1:  # the only collective: final gather of per-instance statistics
            dist.all_gather_into_tensor(gathered, nst_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for k in range(args.steps):
            step_device(k)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)]))
    if world > 1:
        t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kernel_ms = float(t[0]), float(t[1])
    ok = bool((status_d == 0).all().item())
    attempts = ida.get_num_step_attempts()
    tot_attempts = float(attempts.sum().item())
    nst = nst_d
    nst_minmax = (int(nst.min().item()), int(nst.max().item()), float(nst.double().mean().item()))

    # ---- end to end through the host-buffer C-ABI path: H2D inputs (pinned), solve, D2H results --------------
    def step_e2e():
        _lib.check(L.b200ida_reinit(h, p_pin.data_ptr(), yy0_pin.data_ptr(), yp0_pin.data_ptr(), _lib.MEM_HOST))
        _lib.check(L.b200ida_solve_ladder(h, touts.ctypes.data, len(touts), 0, null, null, _lib.MEM_DEVICE))
        _lib.check(L.b200ida_get_state(h, yy_out_pin.data_ptr(), null, null, _lib.MEM_HOST))
        _lib.check(L.b200ida_get_counters(h, nst_out_pin.data_ptr(), null, null, null, null, null, null, null, _lib.MEM_HOST))

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_ok = bool(np.array_equal(nst_out_pin.numpy(), nst_d.cpu().numpy()))
    h2d = int(p_pin.numel() + yy0_pin.numel() + yp0_pin.numel()) * 8
    d2h = int(yy_out_pin.numel() + nst_out_pin.numel()) * 8

    # ---- roofline denominator: measured un-fused FP64 issue rate of this GPU ------------------------------------
    import ctypes as C
    rate = {}
    for kind, name in ((0, "dfma"), (1, "dadd"), (2, "dmul")):
        v = C.c_double()
        _lib.check(L.b200ida_fp64_issue_rate(kind, local_rank, C.byref(v)))
        rate[name] = v.value
    peak_gops = min(rate["dadd"], rate["dmul"])  # un-fused: one instruction per algorithmic op
    achieved_gops = tot_attempts * OPS_PER_ATTEMPT / (kernel_ms * 1e-3) * 1e-9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world >= 1:
        sample = args.cpu_sample or auto_cpu_sample(12.0)
        val, cores, dt = cpu_run(sample)
        cpu = {"value": val, "unit": "instances/s", "cores": cores, "kind": "port",
               "sample": "%d perturbed Roberts instances of the same workload to t=4e10 (%.1f s), C++ restatement of "
                         "rust-ida (oracle/), one instance per thread" % (sample, dt)}

    total = world * B * args.steps
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0  # fallback stated in /opt/skills/guides/B200_PROFILING.md
    line = {
        "metric": "ensemble DAE solves/sec (Roberts, t=4e10)",
        "value": total / (ms * 1e-3), "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(B), parity="all status==0: %s; nst min/max/mean %d/%d/%.1f" % ((ok,) + nst_minmax)),
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "instances/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "results_match_device_run": e2e_ok},
        "gpu_launches": args.steps * 3,  # per step: ida_init_kernel, ida_solve_thread_kernel, widen_i32_kernel
        "roofline": {"bound": "fp64", "achieved": achieved_gops, "peak": peak_gops, "unit": "GFLOP/s (un-fused FP64 ops)",
                     "frac": achieved_gops / peak_gops,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at batch 2^20 (ncu --set full,
                     # profiles/r1g_ida_solve_thread_ncu_summary.csv), scaled to this batch
                     "traffic": int(2185871000 * (B / float(1 << 20))),
                     "hbm_view": {"algorithmic_bytes_per_launch": B * ALG_BYTES_PER_INSTANCE,
                                  "achieved_gbs": B * ALG_BYTES_PER_INSTANCE / (kernel_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak,
                                  "frac": B * ALG_BYTES_PER_INSTANCE / (kernel_ms * 1e-3) * 1e-9 / hbm_peak,
                                  "note": "state read once + written once per launch; three orders of magnitude below the "
                                          "FP64 bound, which is why the roofline above is the FP64 issue rate"},
                     # hardware-side view of the same kernel from ncu (profiles/r1g_ida_solve_thread_ncu_summary.csv):
                     # the FP64 pipe is busy 29 % of the cycles, the schedulers issue on 43 % -- division, sqrt and pow
                     # expand to 10-60 FP64-pipe instructions per algorithmic op, which is the gap to `frac`
                     "ncu": {"sm__pipe_fp64_cycles_active_pct": 28.95, "smsp__issue_active_pct": 43.47,
                             "active_lanes_per_instruction": 20.63, "source": "profiles/r1g_ida_solve_thread_ncu_summary.csv"},
                     "kernel": "ida_solve_thread_kernel<Roberts>", "kernel_ms": kernel_ms,
                     "ops_per_launch": tot_attempts * OPS_PER_ATTEMPT, "attempts_per_launch": tot_attempts,
                     "peak_source": "measured on this GPU by b200ida_fp64_issue_rate (MEASURED_PEAKS.json has no FP64 entry): "
                                    "dfma %.0f, dadd %.0f, dmul %.0f Ginstr/s" % (rate["dfma"], rate["dadd"], rate["dmul"])},
        "cpu_baseline": cpu,
        "clocks": clk.summary(),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = 1


def emit(line):
    """The one JSON line of the contract, written to the process's ORIGINAL stdout (see main)."""
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


def main():
    # Libraries may write to fd 1 behind Python's back (NCCL prints "NCCL version ..." to stdout when the box sets
    # NCCL_DEBUG=VERSION): keep the real stdout for the JSON line only and send everything else to stderr.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
