#!/usr/bin/env python3
"""Integrate one Roberts instance of an ensemble on the GPU with the per-attempt trace on and write the reference's
`data_trace` JSON (what `cargo run --example roberts --features data_trace` writes to roberts_rs.json), so that
scripts/data_trace.ipynb of the reference can load a B200 run next to the reference's own.

    python tools/trace_to_json.py [--instance I] [--batch B] [--out roberts_b200.json]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200 import trace_json
    from rust_ida_b200.problems import Roberts
    ap = argparse.ArgumentParser()
    ap.add_argument("--instance", type=int, default=0)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--out", default="roberts_b200.json")
    a = ap.parse_args()
    # instance 0 carries the reference's own rate constants (examples/roberts.rs); the others are perturbed
    p = r.problems.roberts_ensemble_params(a.batch, seed=1234)
    p[:, 0] = Roberts.DEFAULT
    yy0, yp0 = Roberts.consistent_ic(p)
    rtol, atol = 1e-4, np.array([1e-8, 1e-6, 1e-6])
    dev = torch.device("cuda", 0)
    ida = r.Ida(Roberts(torch.from_numpy(p).to(dev)), torch.from_numpy(yy0).to(dev), torch.from_numpy(yp0).to(dev),
                r.TolControlSV(rtol, atol))
    tr = ida.trace(a.instance)
    ida.solve_ladder(0.4 * 10.0 ** np.arange(12), want_traj=False)
    d, q = tr.rows()
    n = trace_json.write(a.out, d, q, 3, rtol, atol)
    print("%s: %d frames (step attempts) of instance %d" % (a.out, n, a.instance))


if __name__ == "__main__":
    main()
