"""Adding a problem without touching engine sources (INTEGRATION.md "Adding a problem"; src/traits.rs:28-89):
`make user PROBLEM_HDR=examples/user_problem_lorenz_root.cuh` compiles one functor into libb200ida_user.so, which exports
the same C ABI and serves it as B200IDA_PROBLEM_USER.  The GPU leg runs it (in a subprocess: the library is chosen at
import time through B200IDA_LIB) against the oracle's Lorenz problem with the same root function."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import LORENZ_ATOL, LORENZ_RTOL, LORENZ_TOUTS, ROOT

CSRC = os.path.join(ROOT, "rust_ida_b200", "csrc")
USER_LIB = os.path.join(CSRC, "libb200ida_user.so")
HDR = os.path.join(ROOT, "examples", "user_problem_lorenz_root.cuh")


def build_user_lib():
    subprocess.check_call(["make", "-C", CSRC, "-s", "user", "PROBLEM_HDR=" + HDR])
    return USER_LIB


def test_user_problem_library_builds_and_exports_the_abi():
    import ctypes as C
    from rust_ida_b200 import _lib
    L = C.CDLL(build_user_lib())
    for name in _lib.SYMBOLS:
        assert hasattr(L, name), name
    out = subprocess.run(["cuobjdump", "-sass", USER_LIB], capture_output=True, text=True).stdout
    assert out.count("UserProblem") >= 2  # production + ROOTS instantiations of the thread kernel for the user's functor


CHILD = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
import rust_ida_b200 as r
from rust_ida_b200.problems import UserProblem, lorenz_ensemble_params, Lorenz63
B = 1500
p = lorenz_ensemble_params(B, seed=21)
yy0, yp0 = Lorenz63.consistent_ic(p)
touts = 0.5 * np.arange(1, 11)
dev = "cuda"
ida = r.Ida(UserProblem(torch.from_numpy(p).to(dev), n=3, nroots=1), torch.from_numpy(yy0).to(dev), torch.from_numpy(yp0).to(dev),
            r.TolControlSV(1e-6, np.full(3, 1e-8)))
ev = ida.enable_roots(64)
status, traj = ida.solve_ladder(touts)
e = ev.numpy()
np.savez(%(out)r, status=status.cpu().numpy(), traj=traj.cpu().numpy(), yy=ida.get_yy().cpu().numpy(),
         nst=ida.get_num_steps().cpu().numpy(), ev_t=e["t"], ev_iroots=e["iroots"], ev_count=e["count"], nge=e["nge"])
# the same library still serves the built-in problems
ida2 = r.Ida(Lorenz63(torch.from_numpy(p).to(dev)), torch.from_numpy(yy0).to(dev), torch.from_numpy(yp0).to(dev),
             r.TolControlSV(1e-6, np.full(3, 1e-8)))
st2, _ = ida2.solve_ladder(touts, want_traj=False)
np.savez(%(out2)r, yy=ida2.get_yy().cpu().numpy(), nst=ida2.get_num_steps().cpu().numpy())
print("child ok")
"""


@pytest.mark.gpu
def test_user_problem_with_roots_matches_the_oracle(tmp_path):
    from oracle import binding as ob
    from rust_ida_b200.problems import Lorenz63, lorenz_ensemble_params
    build_user_lib()
    out, out2 = str(tmp_path / "user.npz"), str(tmp_path / "builtin.npz")
    env = dict(os.environ, B200IDA_LIB=USER_LIB)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "out": out, "out2": out2}], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "child ok" in r.stdout, r.stderr[-2000:]
    g, g2 = np.load(out), np.load(out2)
    B = 1500
    p = lorenz_ensemble_params(B, seed=21)
    yy0, yp0 = Lorenz63.consistent_ic(p)
    ob.set_pow(None)
    nev_total = 0
    for b in list(range(0, B, 7)):
        ref = ob.roots_run(ob.PK_LORENZ63, p[:, b], yy0[:, b], yp0[:, b], LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS, max_ev=64)
        nev = len(ref["t"])
        nev_total += nev
        assert g["ev_count"][b] == nev
        assert np.array_equal(g["ev_t"][:nev, b].view(np.uint64), ref["t"].view(np.uint64)), b      # root times, bit for bit
        assert np.array_equal(g["ev_iroots"][:nev, :, b], ref["iroots"])                             # get_root_info
        assert np.array_equal(g["yy"][:, b].view(np.uint64), ref["yy"].view(np.uint64))
        assert g["nst"][b] == ref["counters"]["nst"] and g["nge"][b] == ref["counters"]["nge"]
    assert nev_total > 50  # the x = 0 crossings really happen
    # the user functor without its root function IS the built-in Lorenz problem: same states from the same library
    plain = ob.ensemble_solve(ob.PK_LORENZ63, 3, p, yy0, yp0, LORENZ_RTOL, LORENZ_ATOL, LORENZ_TOUTS, nthreads=8)
    assert np.array_equal(g2["yy"].view(np.uint64), plain["yy"].view(np.uint64))
    assert np.array_equal(g2["nst"], plain["counters"]["nst"])
