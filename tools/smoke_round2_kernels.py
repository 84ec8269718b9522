#!/usr/bin/env python3
"""a tiny workload that touches the round-2 kernels (team-flavour blocked LU, CTA and warp team IDA kernels, register warp
LU, packed n = 16 getrs, TMA getrs): residual checks plus run-to-run bit determinism (every case is repeated and the
factors, pivots and solutions must be identical -- a shared-memory ordering bug shows up as a difference).
compute-sanitizer is not available on the GPU pool this was developed on, so these kernels have NOT been race-checked by a
tool; bit parity with the CPU oracle (tests/) and this repeat check are the evidence."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import rust_ida_b200 as r  # noqa: E402
from rust_ida_b200.problems import Heat2D, heat_ensemble_params  # noqa: E402

rng = np.random.default_rng(5)
for n, batch, kw in ((100, 3, {"team_panel": True}), (64, 3, {"team_panel": True}), (100, 3, {}), (16, 5, {}), (32, 3, {}), (24, 3, {})):
    a = rng.uniform(0.0, 1.0 / n, size=(batch, n, n)) + np.eye(n)[::-1][None]
    b = rng.uniform(0.0, 1.0, size=(batch, n))
    A = torch.from_numpy(np.ascontiguousarray(a.transpose(0, 2, 1))).cuda()
    X = torch.from_numpy(b.copy()).cuda()
    A0, X0 = A.clone(), X.clone()
    outs = []
    for rep in range(4):
        A.copy_(A0)
        X.copy_(X0)
        ls = r.Dense.new(n, batch, layout="aos", **kw)
        ls.setup(A)
        ls.solve(A, X, X.clone(), 0.0)
        torch.cuda.synchronize()
        outs.append((A.cpu().numpy().tobytes(), ls.pivots.cpu().numpy().tobytes(), X.cpu().numpy().tobytes()))
    assert all(o == outs[0] for o in outs), "n=%d %s: results differ between identical runs" % (n, kw)
    res = np.abs(np.einsum("bij,bj->bi", a, X.cpu().numpy()) - b).max()
    print("getrf/getrs n=%d %s residual %.2e, 4 runs bit-identical" % (n, kw, res))
    assert res < 1e-10
for M, B, nt in ((10, 4, 3), (4, 8, 5), (5, 8, 4)):
    p = heat_ensemble_params(B, M=M, seed=1)
    yy0, yp0 = Heat2D.consistent_ic(p, M)
    outs = []
    for rep in range(3):
        ida = r.Ida(Heat2D(torch.from_numpy(p).cuda(), M=M), torch.from_numpy(yy0).cuda(), torch.from_numpy(yp0).cuda(),
                    r.TolControlSV(1e-6, np.full(M * M, 1e-8)))
        st, _ = ida.solve_ladder(0.01 * 2.0 ** np.arange(nt), want_traj=False)
        torch.cuda.synchronize()
        outs.append((ida.get_yy().cpu().numpy().tobytes(), ida.get_num_steps().cpu().numpy().tobytes()))
        ok, nst = bool((st == 0).all().item()), ida.get_num_steps().cpu().numpy().tolist()
        ida.close()
    assert all(o == outs[0] for o in outs), "heat M=%d: results differ between identical runs" % M
    print("heat M=%d status ok %s nst %s, 3 runs bit-identical" % (M, ok, nst))
print("done")
