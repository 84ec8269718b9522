"""Host-side logic that needs no GPU: problem descriptors, parameter generators, tolerance control, argument
checking of the Python mirror of the reference interfaces."""
import numpy as np
import pytest

import rust_ida_b200 as r
from rust_ida_b200 import _lib
from rust_ida_b200.problems import (Heat2D, Lorenz63, Roberts, heat_ensemble_params, lorenz_ensemble_params,
                                    roberts_ensemble_params, splitmix64_uniform)


def test_splitmix_uniform_range_and_determinism():
    u = splitmix64_uniform(1234, np.arange(100000), 0)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    assert np.array_equal(u, splitmix64_uniform(1234, np.arange(100000), 0))
    assert not np.array_equal(u, splitmix64_uniform(1234, np.arange(100000), 1))


def test_roberts_descriptor_matches_reference_example():
    p = Roberts.default(2).params
    assert p[:, 0].tolist() == [0.04, 1.0e4, 3.0e7]           # src/sample_problems/roberts.rs:58-59
    yy0, yp0 = Roberts.consistent_ic(p)
    assert yy0[:, 0].tolist() == [1.0, 0.0, 0.0]               # examples/roberts.rs:67
    assert yp0[:, 0].tolist() == [-0.04, 0.04, 0.0]            # examples/roberts.rs:68
    assert Roberts(p).model_size() == 3                        # ModelSpec::model_size
    with pytest.raises(ValueError):
        Roberts(np.zeros((2, 4)))


def test_lorenz_and_heat_descriptors():
    p = lorenz_ensemble_params(1000)
    assert (p[0] >= 8).all() and (p[0] <= 12).all() and (p[1] >= 20).all() and (p[1] <= 35).all()
    yy0, yp0 = Lorenz63.consistent_ic(p)
    np.testing.assert_allclose(yp0[0], p[0] * 0.0)             # p (y1 - y0) = 0 at y = (1,1,1)
    h = heat_ensemble_params(16, M=10)
    assert abs(h / 81.0 - 1).max() <= 0.1
    y, yp = Heat2D.consistent_ic(h, 10)
    assert y.shape == (100, 16) and (y[:10] == 0).all() and (yp[:10] == 0).all() and y.max() <= 1.0
    assert Heat2D(h, M=10).model_size() == 100
    # the residual of the consistent IC is zero on the interior (checked with the oracle's residual)
    from oracle import binding as ob
    res = ob.res(ob.PK_HEAT2D, h[:, 0], y[:, 0].copy(), yp[:, 0].copy())
    assert np.abs(res).max() < 1e-12


def test_tol_control():
    sv = r.TolControlSV.new(1e-4, [1e-8, 1e-6, 1e-6])
    assert sv.ida_rtol == 1e-4 and sv.atol_vector(3).tolist() == [1e-8, 1e-6, 1e-6]
    with pytest.raises(ValueError):
        sv.atol_vector(4)
    ss = r.TolControlSS(1e-4, 1e-4)                            # src/tests/*.rs use TolControlSS::new(1e-4, 1e-4)
    assert ss.atol_vector(3).tolist() == [1e-4] * 3


def test_dense_mirror_metadata_and_shape_checks():
    d = r.Dense.new(3, 5)
    assert d.get_type() == r.LSolverType.Direct and d.num_iters() == 0 and d.res_norm() == 0.0
    assert d.set_scaling_vectors(None, None) is None
    if _lib.device_count() == 0:
        with pytest.raises(r.B200IdaError):
            d.setup(np.zeros((3, 3, 5)))


def test_status_enum_matches_header():
    import os, re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "b200ida.h")).read()
    codes = dict(re.findall(r"#define B200IDA_ST_(\w+)\s+\(?(-?\d+)\)?", hdr))
    assert int(codes["SUCCESS"]) == r.IdaSolveStatus.Success
    assert int(codes["TOO_MUCH_WORK"]) == r.IdaSolveStatus.TooMuchWork
    assert int(codes["LSETUP_FAIL"]) == r.IdaSolveStatus.LinearSetupFail
    assert int(codes["BAD_T"]) == r.IdaSolveStatus.BadTimeValue
    assert int(codes["CLOSE_ROOTS"]) == r.IdaSolveStatus.CloseRoots and int(codes["ROOT"]) == r.IdaSolveStatus.Root
    assert int(codes["BAD_K"]) == r.IdaSolveStatus.BadK
    assert len(codes) == len(r.IdaSolveStatus)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: one JSON line, the same metric / unit / config keys as the CUDA arm,
    the reference-arm extras, and a positive instances/s figure from the oracle port."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "instances/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 0 and j["dtype"] == "f64"
    assert j["value"] > 0 and j["ms_per_step"] > 0 and "workload" in j["config"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reads_profile_summaries_with_commas_in_kernel_names(tmp_path, monkeypatch):
    """bench.py quotes ncu figures from the newest summary under profiles/ at run time: kernel names hold commas
    ("<Roberts, 0, 0, 0, 1>") and a summary may hold the several segment launches of one solve"""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    bench = importlib.import_module("bench")
    prof = tmp_path / "profiles"
    prof.mkdir()
    rows = ["# note ; source x.ncu-rep", "kernel,metric,unit,value"]
    for ms, rd, lanes in ((10.0, 1.0, 28.0), (30.0, 2.0, 20.0)):
        k = "void ida_solve_thread_kernel<Roberts, 0, 0, 0, 1>"
        rows += ["%s,gpu__time_duration.sum,ms,%f" % (k, ms), "%s,dram__bytes_read.sum,Gbyte,%f" % (k, rd),
                 "%s,dram__bytes_write.sum,Gbyte,0.5" % k,
                 "%s,smsp__thread_inst_executed_per_inst_executed.ratio,,%f" % (k, lanes)]
    (prof / "r9_ida_solve_thread_x_ncu_summary.csv").write_text("\n".join(rows) + "\n")
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    q = bench.quoted_profile()
    assert q["launches_in_summary"] == 2 and q["traffic_bytes_per_launch"] == 4_000_000_000
    assert abs(q["kernel_ms_under_ncu"] - 40.0) < 1e-9 and abs(q["active_lanes_per_instruction"] - 22.0) < 1e-9
