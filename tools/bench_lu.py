#!/usr/bin/env python3
"""Batched getrf / getrs microbenchmark (BASELINE.json configs[3]).

n in {3,4,8,16,32,64,100,128,256}; batch(n) = ceil(9*2^20 / n^2); the timed passes rotate over enough independent
buffer copies that each pass touches >= 512 MB (the 126 MB L2 would otherwise hide HBM).  Matrices are the
reference's own (commented-out) generator: U(0, 1/n) + anti-identity (benches/dense.rs:236-255), seed 1234.
Reports, per n: matrices/s, GFLOP/s (getrf: sum_k[(n-k-1) + 2(n-k-1)^2], getrs: 2n^2 - n + n divisions) and the
algorithmic HBM traffic rate (getrf 16n^2 + 8n B, getrs 8n^2 + 24n B per system) against MEASURED_PEAKS.json.
One JSON line per (n, op).  Layout: SoA for n <= 8 (thread-per-system kernels), AoS above.
"""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flops_getrf(n):
    return sum((n - k - 1) + 2 * (n - k - 1) ** 2 for k in range(n))


def pivots_match_oracle(a, piv_dev, n, layout, sample=24):
    """pivot sequences of `sample` matrices of the batch against the CPU oracle's dense_get_rf (bit-exact integers)"""
    from oracle import binding as ob
    batch = a.shape[0]
    ids = np.unique(np.linspace(0, batch - 1, min(sample, batch)).astype(np.int64))
    piv = piv_dev.cpu().numpy()
    piv = piv.reshape(n, batch).T if layout == "soa" else piv.reshape(batch, n)
    for b in ids:
        _, pref, info = ob.getrf(a[b].numpy())
        if info != 0 or not np.array_equal(pref, piv[b]):
            return False
    return True


def run_lu(sizes=(3, 4, 8, 16, 32, 64, 100, 128, 256), reps=3, min_mb=512.0, batch_scale=1, fp64_peak_gops=None, emit=None,
           check_pivots=True, flags=0):
    """the config-4 table as a list of dicts (one per (n, op)); `emit` (if given) is called with every row"""
    import torch
    from rust_ida_b200 import _lib

    class A:
        pass
    args = A()
    args.sizes, args.reps, args.min_mb, args.batch_scale = ",".join(str(x) for x in sizes), reps, min_mb, batch_scale
    rows = []
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    for n in [int(x) for x in args.sizes.split(",")]:
        batch = math.ceil(9 * (1 << 20) / (n * n)) * args.batch_scale
        layout = "soa" if n <= 8 else "aos"
        lay = _lib.LAYOUT_SOA if layout == "soa" else _lib.LAYOUT_AOS
        bytes_mat = 8 * n * n * batch
        copies = max(4, math.ceil(args.min_mb * 1e6 / (2 * bytes_mat)))
        g = torch.Generator(device="cpu").manual_seed(1234 + n)
        a = torch.rand((batch, n, n), generator=g, dtype=torch.float64) / n
        a += torch.flip(torch.eye(n, dtype=torch.float64), dims=[0])[None]
        rhs = torch.rand((batch, n), generator=g, dtype=torch.float64)
        if layout == "soa":
            master = a.permute(1, 2, 0).contiguous().to(dev)   # [i, j, b]
            bmaster = rhs.t().contiguous().to(dev)
        else:
            master = a.transpose(1, 2).contiguous().to(dev)   # column-major per matrix
            bmaster = rhs.contiguous().to(dev)
        bufs = [torch.empty_like(master) for _ in range(copies)]
        xs = [torch.empty_like(bmaster) for _ in range(copies)]
        piv = [torch.empty((n * batch,), dtype=torch.int64, device=dev) for _ in range(copies)]
        info = torch.empty((batch,), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        best = {"getrf": 1e30, "getrs": 1e30}
        for rep in range(args.reps + 1):
            for c in range(copies):
                bufs[c].copy_(master)
                xs[c].copy_(bmaster)
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            for c in range(copies):
                _lib.check(L.b200ida_getrf_batched(bufs[c].data_ptr(), piv[c].data_ptr(), info.data_ptr(), n, batch, lay | flags,
                                                   _lib.MEM_DEVICE, st))
            e1.record()
            for c in range(copies):
                _lib.check(L.b200ida_getrs_batched(bufs[c].data_ptr(), piv[c].data_ptr(), xs[c].data_ptr(), n, batch, lay,
                                                   _lib.MEM_DEVICE, st))
            e2.record()
            torch.cuda.synchronize()
            if rep > 0:
                best["getrf"] = min(best["getrf"], e0.elapsed_time(e1) / copies)
                best["getrs"] = min(best["getrs"], e1.elapsed_time(e2) / copies)
        assert int(info.abs().max().item()) == 0
        # residual check on the last copy: ||A x - b|| small
        if layout == "soa":
            x = xs[-1].t().cpu()
        else:
            x = xs[-1].cpu()
        res = (torch.einsum("bij,bj->bi", a, x) - rhs).abs().max().item()
        piv_ok = pivots_match_oracle(a, piv[-1], n, layout) if check_pivots else None
        for op, ms in best.items():
            fl = flops_getrf(n) if op == "getrf" else (2 * n * n - n + n)
            by = (16 * n * n + 8 * n) if op == "getrf" else (8 * n * n + 24 * n)
            rate = batch / (ms * 1e-3)
            row = {"op": op + ("+dmma" if (flags and op == "getrf") else ""), "n": n, "batch": batch, "layout": layout, "copies": copies, "ms": ms,
                   "matrices_per_s": rate, "gflops": rate * fl * 1e-9, "algorithmic_gbs": rate * by * 1e-9,
                   "hbm_frac_of_measured": rate * by * 1e-9 / hbm, "max_abs_residual": res}
            if fp64_peak_gops:
                row["fp64_frac"] = rate * fl * 1e-9 / fp64_peak_gops
            if op == "getrf" and piv_ok is not None:
                row["pivots_bit_exact_vs_oracle"] = piv_ok
            rows.append(row)
            if emit:
                emit(row)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="3,4,8,16,32,64,100,128,256")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--min-mb", type=float, default=512.0)
    ap.add_argument("--batch-scale", type=int, default=1,
                    help="multiply the config's batch (throughput regime: many systems per resident CTA / warp)")
    ap.add_argument("--dmma", action="store_true", help="B200IDA_LU_DMMA: tensor-core trailing update (tolerance mode)")
    args = ap.parse_args()
    run_lu([int(x) for x in args.sizes.split(",")], args.reps, args.min_mb, args.batch_scale,
           emit=lambda row: print(json.dumps(row), flush=True), flags=0x100 if args.dmma else 0)


if __name__ == "__main__":
    main()
