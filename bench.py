#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native batched IDA engine.

Metric (BASELINE.json): ensemble DAE solves/sec -- Roberts (n = 3), rate constants perturbed +-10 % per
instance, each instance integrated from t = 0 to t = 4e10 through the reference's 12-decade output ladder
(tout = 0.4 * 10^i, examples/roberts.rs) -- instances finished per second, whole job over all GPUs.

A "step" is one pass of the hot path over one batch: re-initialise the ensemble state (Ida::new) and
integrate every instance to t = 4e10 (one solve-ladder kernel launch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--scaling weak|strong] [--impl ours|reference]

    value      instances/s with params / yy0 / yp0 already resident in HBM
    e2e        the same through the C-ABI host-buffer path: H2D of the inputs from pinned memory, the
               solve, D2H of the final states and step counts, every step
    roofline   FP64-issue roofline of the solve kernel (n = 3 ensembles are FP64-pipe-bound, not HBM-bound;
               SURVEY.md 8d): algorithmic FP64 ops / kernel time vs the measured un-fused FP64 issue rate
    cpu_baseline  the CPU oracle (a C++ restatement of rust-ida, NOT rust-ida itself: there is no Rust
               toolchain here) on a bounded sample, one instance per thread on all host cores

    parity     the CPU leg integrates EVERY instance of the (rank-0) batch with the libm-pow oracle and the device
               results (final states and step counts) are compared with it bit for bit
    extra      the rest of BASELINE.json's metric on the same GPU: the config-4 batched LU table, Lorenz-63 (config 3)
               and heat 2-D (config 5), each with its own roofline fraction (N = 1 only; --no-extra skips it)

N > 1 (torchrun): one process per GPU, instances sharded by contiguous id range, no collective inside the step; one
NCCL all_gather of the per-instance statistics block {nst, nre, nje, nni, netf, ncfn, status, kused} at the end of
every step.  --scaling weak (default): --batch instances per GPU; --scaling strong: --batch instances in total.
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
# FP64 arithmetic per step attempt of the reference algorithm at n = 3 (each add, mul, div, sqrt, pow counted as ONE
# op).  Measured at run time by the operation counter of the oracle (oracle/ida_oracle.cpp built with
# -DORACLE_COUNT_OPS) on a sample of this workload; this constant is only the fallback when the CPU leg is switched
# off: 274.7 = the counter's figure for the first 20,000 instances of the workload (default Roberts instance: 103,604
# ops / 377 attempts = 274.8, with the 2,030 sqrt and 528 pow of SURVEY.md section 6).
OPS_PER_ATTEMPT_FALLBACK = 274.7
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch instances per GPU; strong: --batch instances in total, split over the GPUs")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 3 / 4 / 5 measurements of `extra`")
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
def cpu_run(sample, first_id=0, keep=False):
    """integrate `sample` perturbed Roberts instances to t=4e10 with the CPU oracle (libm pow, the reference's
    f64::powf) on all cores -> (inst/s, cores, seconds[, oracle outputs])"""
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params
    cores = os.cpu_count() or 1
    p = roberts_ensemble_params(sample, seed=SEED, first_id=first_id)
    yy0, yp0 = Roberts.consistent_ic(p)
    ob.set_pow(None)
    t0 = time.perf_counter()
    out = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, RTOL, ATOL, TOUTS, nthreads=cores)
    dt = time.perf_counter() - t0
    assert (out["status"] == 0).all()
    if keep:
        return sample / dt, cores, dt, out
    return sample / dt, cores, dt


def ops_per_attempt(sample=8192, first_id=0):
    """algorithmic FP64 ops per step attempt of this workload, counted by the instrumented oracle on a sample"""
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params
    p = roberts_ensemble_params(sample, seed=SEED, first_id=first_id)
    yy0, yp0 = Roberts.consistent_ic(p)
    ops, natt = ob.ensemble_ops(ob.PK_ROBERTS, 3, p, yy0, yp0, RTOL, ATOL, TOUTS, nthreads=os.cpu_count() or 1)
    tot = ops.sum(axis=0)
    return float(tot.sum()) / float(natt.sum()), {k: int(v) for k, v in zip(("add", "mul", "div", "sqrt", "pow"), tot)}, sample


def workload_config(B, world=1, scaling="weak"):
    """the `config` object both arms print (the reference arm runs on our arm's config)"""
    return {"workload": "Roberts ensemble, n=3, %d instances per GPU, rate constants +-10%% (splitmix64 seed 1234 "
                        "keyed by instance id), rtol 1e-4, atol [1e-8,1e-6,1e-6], tout = 0.4*10^i i=0..11 "
                        "(BASELINE.json configs[1])" % B,
            "batch_per_gpu": B, "sharding": "contiguous instance-id ranges, no collective in the step loop",
            "l2": "per-step state (%.0f MB/GPU) is larger than the 126 MB L2" % (B * 0.78e-3)}


def batch_per_gpu(args, world):
    return args.batch if args.scaling == "weak" else max(1, args.batch // world)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for the same metric/config.  rust-ida cannot be built here
    (no Rust toolchain, tree mid-refactor), so this is the oracle port, all host threads, bounded sample."""
    if rank != 0:
        return
    B = batch_per_gpu(args, world)
    # one step = the whole per-GPU batch of our arm (the same instances: ids 0..B-1), all host threads; --cpu-sample
    # bounds the step for slow hosts (then `sample` says so)
    sample = args.cpu_sample or B
    for _ in range(min(args.warmup, 1)):
        cpu_run(max(256, sample // 16))
    t0 = time.perf_counter()
    tot = 0
    for k in range(args.steps):
        cpu_run(sample, first_id=0)
        tot += sample
    dt = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    val = tot / dt
    desc = "%d perturbed Roberts instances per step (%s the %d-per-GPU workload), to t=4e10" % (
        sample, "all of" if sample == B else "a sample of", B)
    line = {
        "impl": "reference", "metric": "ensemble DAE solves/sec (Roberts, t=4e10)", "value": val, "unit": "instances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(B, world, args.scaling),
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

    B = batch_per_gpu(args, world)
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
    # the per-instance statistics block of the final gather: nst, nre, nje, nni, netf, ncfn, status, kused
    stats_d = torch.empty((8, B), dtype=torch.int64, device=dev)
    kused_d = torch.empty((B,), dtype=torch.int32, device=dev)
    gathered = torch.empty((world, 8, B), dtype=torch.int64, device=dev) if world > 1 else None
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
        row = lambda i: stats_d[i].data_ptr()
        _lib.check(L.b200ida_get_counters(h, row(0), row(1), row(2), row(3), row(4), row(5), null, null, _lib.MEM_DEVICE))
        _lib.check(L.b200ida_get_step_info(h, kused_d.data_ptr(), null, null, null, null, null, _lib.MEM_DEVICE))
        stats_d[6].copy_(status_d[-1])
        stats_d[7].copy_(kused_d)
        if world > 1:  # the only collective: final gather of the per-instance statistics
            dist.all_gather_into_tensor(gathered, stats_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.b200ida_kernel_launches()
    with ClockSampler(local_rank) as clk:
        e0.record()
        for k in range(args.steps):
            step_device(k)
        e1.record()
        barrier()
    launches = int(L.b200ida_kernel_launches() - launches0)
    ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)]))
    if world > 1:
        t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kernel_ms = float(t[0]), float(t[1])
    ok = bool((status_d == 0).all().item())
    attempts = ida.get_num_step_attempts()
    tot_attempts = float(attempts.sum().item())
    nst = stats_d[0]
    nst_minmax = (int(nst.min().item()), int(nst.max().item()), float(nst.double().mean().item()))
    yy_dev = ida.get_yy().cpu().numpy()
    nst_dev = nst.cpu().numpy()

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
    e2e_ok = bool(np.array_equal(nst_out_pin.numpy(), nst_dev) and
                  np.array_equal(yy_out_pin.numpy().view(np.uint64), yy_dev.view(np.uint64)))
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU leg (rank 0, N = 1): every instance of the batch through the libm-pow oracle; it is the cpu_baseline AND
    # the parity check of the device results; the instrumented oracle gives the roofline's work unit ---------------------
    cpu = None
    parity = None
    opa, opa_src = OPS_PER_ATTEMPT_FALLBACK, "fallback constant (CPU leg off)"
    if not args.no_cpu_baseline and world == 1:
        sample = args.cpu_sample or B
        val, cores, dt, ref = cpu_run(sample, first_id=first_id, keep=True)
        cpu = {"value": val, "unit": "instances/s", "cores": cores, "kind": "port",
               "sample": "%s %d perturbed Roberts instances of the workload to t=4e10 (%.1f s), C++ restatement of "
                         "rust-ida (oracle/) with libm pow, one instance per thread" % ("all" if sample == B else "the first", sample, dt)}
        bad_nst = int(np.count_nonzero(ref["counters"]["nst"] != nst_dev[:sample]))
        bad_yy = int(np.count_nonzero((ref["yy"].view(np.uint64) != yy_dev[:, :sample].view(np.uint64)).any(axis=0)))
        parity = {"checked": int(sample), "mismatch": int(max(bad_nst, bad_yy)), "mismatch_nst": bad_nst, "mismatch_yy": bad_yy,
                  "pow": "libm", "compared": "final yy (3 doubles, bit-exact) and nst of every checked instance, device run vs CPU oracle",
                  "all_status_zero": ok, "nst_min_max_mean": list(nst_minmax)}
        opa_v, opa_mix, opa_n = ops_per_attempt(min(8192, B), first_id=first_id)
        opa, opa_src = opa_v, "oracle op counter (-DORACLE_COUNT_OPS) on the first %d instances: %s" % (opa_n, json.dumps(opa_mix))
    achieved_gops = tot_attempts * opa / (kernel_ms * 1e-3) * 1e-9

    # ---- the rest of BASELINE.json's metric on this GPU (N = 1 only) -------------------------------------------------------
    extra = None
    if world == 1 and not args.no_extra:
        extra = run_extra(peak_gops)

    quoted = quoted_profile()
    total = world * B * args.steps
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak = 6650.0  # fallback stated in /opt/skills/guides/B200_PROFILING.md
        hbm_src = "fallback of B200_PROFILING.md"
    alg_bytes = B * ALG_BYTES_PER_INSTANCE
    line = {
        "metric": "ensemble DAE solves/sec (Roberts, t=4e10)",
        "value": total / (ms * 1e-3), "unit": "instances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(B, world, args.scaling),
        "parity": parity,
        "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "instances/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "results_match_device_run": e2e_ok},
        "gpu_launches": launches,  # counted by the library (b200ida_kernel_launches) over the timed region
        "roofline": {"bound": "fp64", "achieved": achieved_gops, "peak": peak_gops, "unit": "GFLOP/s (un-fused FP64 ops)",
                     "frac": achieved_gops / peak_gops,
                     "traffic": quoted.get("traffic_bytes_per_launch"),
                     "hbm_view": {"algorithmic_bytes_per_launch": alg_bytes,
                                  "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) * 1e-9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                                  "frac": alg_bytes / (kernel_ms * 1e-3) * 1e-9 / hbm_peak,
                                  "note": "state read once + written once per launch; three orders of magnitude below the "
                                          "FP64 bound, which is why the roofline above is the FP64 issue rate"},
                     "quoted_from_profile": quoted,
                     "kernel": "ida_solve_thread_kernel<Roberts, FAST>: one solve = its segment launches (+ the slot re-grouping "
                               "between them, the empty exact pass and ida_interp_kernel<3>), all inside the same CUDA-event "
                               "bracket; `achieved` divides by that bracket",
                     "kernel_ms": kernel_ms, "ops_per_attempt": opa, "ops_per_attempt_source": opa_src,
                     "ops_per_launch": tot_attempts * opa, "attempts_per_launch": tot_attempts,
                     "peak_source": "measured on this GPU by b200ida_fp64_issue_rate (MEASURED_PEAKS.json has no FP64 entry): "
                                    "dfma %.0f, dadd %.0f, dmul %.0f Ginstr/s" % (rate["dfma"], rate["dadd"], rate["dmul"])},
        "cpu_baseline": cpu,
        "extra": extra,
        "clocks": clk.summary(),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def quoted_profile():
    """ncu figures of the solve kernel, READ from the newest committed summary under profiles/ (never typed in here):
    they describe the build that summary was taken from, which the entry names.  A summary may hold several launches
    (the segments of one solve): byte counts, durations and instruction counts are summed, percentages and ratios are
    averaged with the launch durations as weights."""
    import csv
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ida_solve_thread*_ncu_summary.csv")))
    if not cands:
        return {}
    path = cands[-1]
    launches = []  # one dict per launch, in file order
    note = ""
    want = ("dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "gpu__time_duration.sum", "smsp__inst_executed.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}
    try:
        for row in csv.reader(open(path)):
            if row and row[0].startswith("#"):
                note = ",".join(row).lstrip("# ").strip()
            # kernel names hold commas ("<Roberts, 0, 0, 0, 1>"): the metric, unit and value are the LAST three fields
            if len(row) >= 4 and row[-3] in want:
                if not launches or row[-3] in launches[-1]:
                    launches.append({})
                launches[-1][row[-3]] = float(row[-1]) * scale.get(row[-2], 1.0)
    except Exception:
        return {}
    launches = [l for l in launches if "gpu__time_duration.sum" in l]
    if not launches:
        return {}
    tsum = sum(l["gpu__time_duration.sum"] for l in launches)
    wavg = lambda k: sum(l.get(k, 0.0) * l["gpu__time_duration.sum"] for l in launches) / tsum
    out = {"source": os.path.relpath(path, ROOT), "source_note": note, "launches_in_summary": len(launches),
           "applies_to": "the build and batch (2^20) that file was captured from, not necessarily this run",
           "traffic_bytes_per_launch": int(sum(l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0) for l in launches)),
           "sm__pipe_fp64_cycles_active_pct": wavg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
           "smsp__issue_active_pct": wavg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "active_lanes_per_instruction": wavg("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "kernel_ms_under_ncu": tsum, "warp_instructions": sum(l.get("smsp__inst_executed.sum", 0.0) for l in launches)}
    return out


def run_extra(fp64_peak_gops):
    """the other halves of BASELINE.json's metric, measured in the same process on the same GPU: batched LU (config 4),
    Lorenz-63 (config 3, 4 * 2^20 instances) and heat 2-D (config 5, 65,536 instances).  Failures are reported, not fatal."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    out = {}
    try:
        import bench_lu
        rows = bench_lu.run_lu(reps=2, fp64_peak_gops=fp64_peak_gops)
        # the tolerance mode of the same table: FP64 tensor-core (DMMA) trailing update where it applies (even n >= 115)
        rows += [r_ for r_ in bench_lu.run_lu(sizes=(128, 256), reps=2, fp64_peak_gops=fp64_peak_gops, flags=0x100)
                 if r_["op"].startswith("getrf")]
        out["lu_config4"] = [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in row.items()
                              if k in ("op", "n", "batch", "layout", "ms", "matrices_per_s", "gflops", "algorithmic_gbs",
                                       "hbm_frac_of_measured", "fp64_frac", "pivots_bit_exact_vs_oracle")} for row in rows]
    except Exception as exc:  # noqa: BLE001
        out["lu_config4"] = {"error": repr(exc)}
    try:
        import bench_configs
        from oracle import binding as ob
        from rust_ida_b200.problems import Lorenz63, lorenz_ensemble_params
        lo = bench_configs.run_lorenz(4 << 20, reps=2)
        p = lorenz_ensemble_params(4096, seed=1234)
        yy0, yp0 = Lorenz63.consistent_ic(p)
        ops, natt = ob.ensemble_ops(ob.PK_LORENZ63, 3, p, yy0, yp0, 1e-6, np.full(3, 1e-8), 0.5 * np.arange(1, 11),
                                    nthreads=os.cpu_count() or 1)
        opa = float(ops.sum()) / float(natt.sum())
        g = lo["attempts"] * opa / (lo["ms"] * 1e-3) * 1e-9
        lo["roofline"] = {"bound": "fp64", "ops_per_attempt": opa, "achieved_gflops": g, "peak_gflops": fp64_peak_gops,
                          "frac": g / fp64_peak_gops}
        out["lorenz_config3"] = lo
    except Exception as exc:  # noqa: BLE001
        out["lorenz_config3"] = {"error": repr(exc)}
    try:
        import bench_configs
        he = bench_configs.run_heat(65536, reps=1)
        he["roofline"] = {"bound": "fp64", "achieved_gflops": he["lu_plus_solve_gflops"], "peak_gflops": fp64_peak_gops,
                          "frac": he["lu_plus_solve_gflops"] / fp64_peak_gops,
                          "note": "LU + triangular-solve flops only (the iteration matrix stays in shared memory)"}
        out["heat_config5"] = he
    except Exception as exc:  # noqa: BLE001
        out["heat_config5"] = {"error": repr(exc)}
    try:
        # the warp-per-system kernel (the AUTO path for 3 < n <= 32) on ensembles of its own size class: heat 2-D on a
        # 4 x 4 and a 5 x 5 grid (n = 16 / 25), 65,536 instances each
        import bench_configs
        rows = []
        for M in (4, 5):
            hs = bench_configs.run_heat(65536, reps=1, M=M)
            hs["kernel"] = "warp"
            rows.append({k: hs[k] for k in ("config", "batch", "n", "kernel", "ms", "instances_per_s", "all_success")})
        out["heat_small_warp_team"] = rows
    except Exception as exc:  # noqa: BLE001
        out["heat_small_warp_team"] = {"error": repr(exc)}
    return out


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
