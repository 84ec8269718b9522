"""Lane-divergence simulator for the thread-per-instance kernel (analysis tool, CPU only; uses the oracle's per-attempt
event log).  Not a test: run `python tests/analysis/sim_lanes.py [log2_B]`.

Model: a warp executes one step attempt of each of its 32 lanes per loop trip; the trip costs, per code segment, the
segment's instruction estimate times the max over the live lanes of the segment's trip count.  Compares the static
Morton order of the engine with periodic re-grouping by a behaviour key.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rust_ida_b200.problems import Roberts, roberts_ensemble_params  # noqa: E402

ROBERTS_RTOL = 1e-4
ROBERTS_ATOL = np.array([1e-8, 1e-6, 1e-6])
TOUTS = np.array([0.4 * 10.0 ** i for i in range(12)])


def morton_keys(p):
    P, B = p.shape
    bits = min(10, 30 // P)
    key = np.zeros(B, dtype=np.uint64)
    for j in range(P):
        lo, hi = p[j].min(), p[j].max()
        u = np.clip((p[j] - lo) / (hi - lo), 0, 1)
        q = (u * ((1 << bits) - 1)).astype(np.uint64)
        for k in range(bits):
            key |= ((q >> np.uint64(k)) & np.uint64(1)) << np.uint64(k * P + j)
    return key


def events(p, yy0, yp0, max_att=448, nthreads=8):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libida_oracle.so"))
    B = p.shape[1]
    ev = np.zeros((B, max_att), dtype=np.uint32)
    evh = np.zeros((B, max_att), dtype=np.float32)
    natt = np.zeros(B, dtype=np.int32)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.oracle_ensemble_events.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, dp, dp, dp, ctypes.c_double, dp,
                                           dp, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p]
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(dp)
    lib.oracle_ensemble_events(0, 3, B, 3, f(p), f(yy0), f(yp0), ROBERTS_RTOL, f(ROBERTS_ATOL), f(TOUTS), len(TOUTS), nthreads,
                               max_att, ev.ctypes.data, evh.ctypes.data, natt.ctypes.data)
    return ev, evh, natt


# instruction estimates per segment execution (warp instructions)
C = dict(base=300, coef_trip=66, alpha0_trip=4, scale_trip=10, pred_trip=12, sys0=25, it0=183, it1=219, it2=309, setup=130,
         te2=55, te3=55, cs_base=30, cs_rec_trip=10, cs_pow=110, cs_kp1=73, fail_base=20, fail_rest_trip=5, fail_scale_trip=26,
         fail_pow=90, out=240)


def seg_counts(ev):
    """per (instance, attempt) trip counts of every segment, as a dict of int arrays"""
    kk = (ev & 7).astype(np.int32)
    rec = ((ev >> 3) & 1).astype(np.int32)
    ns = ((ev >> 4) & 7).astype(np.int32)
    nset = ((ev >> 7) & 3).astype(np.int32)
    nit = ((ev >> 9) & 15).astype(np.int32)
    te = ((ev >> 13) & 3).astype(np.int32)
    conv = ((ev >> 15) & 1).astype(np.int32)
    powf = ((ev >> 16) & 1).astype(np.int32)
    kp1 = ((ev >> 17) & 1).astype(np.int32)
    outf = ((ev >> 22) & 1).astype(np.int32)
    scale = np.maximum(kk - ns + 1, 0)
    live = (kk > 0).astype(np.int32)  # an attempt always starts with kk >= 1
    s = {}
    s["base"] = live
    s["coef_trip"] = rec * kk
    s["alpha0_trip"] = kk
    s["scale_trip"] = scale
    s["pred_trip"] = kk
    s["sys0"] = live
    s["it0"] = (nit >= 1).astype(np.int32)
    s["it1"] = (nit >= 2).astype(np.int32)
    s["it2"] = np.maximum(nit - 2, 0)
    s["setup"] = nset
    s["te2"] = (te >= 2).astype(np.int32)
    s["te3"] = (te >= 3).astype(np.int32)
    s["cs_base"] = conv
    s["cs_rec_trip"] = conv * (kk + 1)
    s["cs_pow"] = conv * powf
    s["cs_kp1"] = kp1
    fail = live * (1 - conv)
    s["fail_base"] = fail
    s["fail_rest_trip"] = fail * kk
    s["fail_scale_trip"] = fail * scale
    s["fail_pow"] = fail * powf
    s["out"] = outf
    return s


def cost_of(order, segs, a0, a1):
    """(lane cost, warp cost) over attempts [a0, a1) with instances grouped 32 at a time in `order`"""
    nw = len(order) // 32
    lane = 0.0
    warp = 0.0
    for name, c in C.items():
        x = segs[name][order[: nw * 32], a0:a1].reshape(nw, 32, -1)
        lane += c * float(x.sum())
        warp += c * float(x.max(axis=1).sum())
    return lane, warp


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    Bfull = 1 << 20
    B = 1 << lg
    p_all = roberts_ensemble_params(Bfull, seed=1234)
    order_all = np.argsort(morton_keys(p_all), kind="stable")
    sel = order_all[:B]  # a contiguous range of the engine's slot order: same density as the full ensemble
    p = np.ascontiguousarray(p_all[:, sel])
    yy0, yp0 = Roberts.consistent_ic(p)
    ev, evh, natt = events(p, yy0, yp0)
    A = int(natt.max())
    print("B", B, "attempts", natt.min(), natt.mean(), natt.max())
    segs = seg_counts(ev)
    ident = np.arange(B)
    lane, warp = cost_of(ident, segs, 0, A)
    print("static Morton order: lane cost/attempt %.0f, warp cost/trip-lane %.0f, efficiency %.3f" %
          (lane / natt.sum(), warp * 32 / natt.sum(), lane / (32 * warp)))
    rnd = np.random.default_rng(0).permutation(B)
    l2, w2 = cost_of(rnd, segs, 0, A)
    print("random order: efficiency %.3f, warp cost x%.3f" % (l2 / (32 * w2), w2 / warp))
    # tail only: how much is lost to lanes finishing early
    trips = natt[: B // 32 * 32].reshape(-1, 32).max(axis=1).sum() * 32
    print("attempt-count efficiency (static) %.3f" % (natt.sum() / trips))

    kk_after = ((ev >> 18) & 7).astype(np.int64)
    ph_after = ((ev >> 21) & 1).astype(np.int64)
    ns_after = ((ev >> 4) & 7).astype(np.int64)
    for K in (16, 32, 64, 128):
        for keyname in ("kk", "kk_ns", "kk_ns_h", "h_kk"):
            tot = 0.0
            for a0 in range(0, A, K):
                if a0 == 0:
                    order = ident
                else:
                    done = natt <= a0
                    kka = kk_after[:, a0 - 1]
                    nsa = ns_after[:, a0 - 1]
                    pha = ph_after[:, a0 - 1]
                    hq = np.round(np.log2(np.maximum(evh[:, a0 - 1], 1e-300)) * 4).astype(np.int64)
                    if keyname == "kk":
                        keys = (done, kka * 2 + pha)
                    elif keyname == "kk_ns":
                        keys = (done, (kka * 2 + pha) * 8 + nsa)
                    elif keyname == "kk_ns_h":
                        keys = (done, (kka * 2 + pha) * 8 + nsa, hq)
                    else:
                        keys = (done, hq, (kka * 2 + pha) * 8 + nsa)
                    # np.lexsort: last key is primary; ties keep Morton (index) order
                    order = np.lexsort((ident,) + tuple(reversed(keys[1:])) + (keys[0],))
                _, w = cost_of(order, segs, a0, min(a0 + K, A))
                tot += w
            print("re-sort every %3d attempts by %-8s: warp cost x%.3f of static (efficiency %.3f)" %
                  (K, keyname, tot / warp, lane / (32 * tot)))


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[2] == "pool"):
    main()


def pooled(lg=15, pools=(128, 256, 384, 512, 1024), Ks=(8, 16, 32, 64), hq_scale=4):
    """re-grouping restricted to a pool of consecutive slots (a CTA that re-sorts its own lanes through shared memory)"""
    Bfull = 1 << 20
    B = 1 << lg
    p_all = roberts_ensemble_params(Bfull, seed=1234)
    order_all = np.argsort(morton_keys(p_all), kind="stable")
    sel = order_all[:B]
    p = np.ascontiguousarray(p_all[:, sel])
    yy0, yp0 = Roberts.consistent_ic(p)
    ev, evh, natt = events(p, yy0, yp0)
    A = int(natt.max())
    segs = seg_counts(ev)
    ident = np.arange(B)
    lane, warp = cost_of(ident, segs, 0, A)
    kk_after = ((ev >> 18) & 7).astype(np.int64)
    ph_after = ((ev >> 21) & 1).astype(np.int64)
    ns_after = ((ev >> 4) & 7).astype(np.int64)
    with np.errstate(all="ignore"):
        hq_all = np.round(np.log2(np.maximum(evh, 1e-30)) * hq_scale).astype(np.int64)
    for pool in pools:
        nb = B // pool
        Bp = nb * pool
        blk = (ident[:Bp] // pool).astype(np.int64)
        for K in Ks:
            tot = 0.0
            for a0 in range(0, A, K):
                if a0 == 0:
                    order = ident[:Bp]
                else:
                    done = (natt[:Bp] <= a0).astype(np.int64)
                    k2 = (kk_after[:Bp, a0 - 1] * 2 + ph_after[:Bp, a0 - 1]) * 8 + ns_after[:Bp, a0 - 1]
                    order = np.lexsort((ident[:Bp], hq_all[:Bp, a0 - 1], k2, done, blk))
                _, w = cost_of(order, segs, a0, min(a0 + K, A))
                tot += w
            l0, w0 = cost_of(ident[:Bp], segs, 0, A)
            print("pool %5d, re-group every %3d trips: warp cost x%.3f of static (efficiency %.3f)" % (pool, K, tot / w0, l0 / (32 * tot)))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "pool":
    pooled(int(sys.argv[1]))


def breakdown(lg=14):
    Bfull = 1 << 20
    B = 1 << lg
    p_all = roberts_ensemble_params(Bfull, seed=1234)
    order_all = np.argsort(morton_keys(p_all), kind="stable")
    sel = order_all[:B]
    p = np.ascontiguousarray(p_all[:, sel])
    yy0, yp0 = Roberts.consistent_ic(p)
    ev, evh, natt = events(p, yy0, yp0)
    segs = seg_counts(ev)
    nw = B // 32
    tl = tw = 0
    rows = []
    for name, c in C.items():
        x = segs[name].reshape(nw, 32, -1)
        l = c * float(x.sum()) / 32
        w = c * float(x.max(axis=1).sum())
        rows.append((name, l, w))
        tl += l
        tw += w
    for name, l, w in rows:
        print("%-16s lane-avg %5.1f%%  warp %5.1f%%  loss %5.1f%% of warp total   seg-eff %.2f" % (name, 100 * l / tl, 100 * w / tw, 100 * (w - l) / tw, l / w if w else 1))
    print("total eff %.3f" % (tl / tw))


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "breakdown":
    breakdown(int(sys.argv[1]))
