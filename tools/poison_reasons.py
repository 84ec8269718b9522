import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import rust_ida_b200 as r
from rust_ida_b200 import _lib
from rust_ida_b200.problems import Lorenz63, lorenz_ensemble_params
B = 1 << 18
p = lorenz_ensemble_params(B, seed=1234); yy0, yp0 = Lorenz63.consistent_ic(p)
dev = "cuda"
ida = r.Ida(Lorenz63(torch.from_numpy(p).to(dev)), torch.from_numpy(yy0).to(dev), torch.from_numpy(yp0).to(dev), r.TolControlSV(1e-6, np.full(3, 1e-8)))
for nt in (1, 2, 10):
    ida.reinit(torch.from_numpy(yy0).to(dev), torch.from_numpy(yp0).to(dev))
    st, _ = ida.solve_ladder(0.5 * np.arange(1, nt + 1), want_traj=False)
    print(nt, "fallbacks", ida.fallback_count(), ida.fallback_histogram())
# reason histogram: read the poisoned bytes through a tiny hack: fallback_count counts nonzero; use ctypes to read the device buffer? not exposed -> use statuses instead
