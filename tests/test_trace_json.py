"""rust_ida_b200/trace_json.py: per-attempt trace rows -> the reference's data_trace JSON (src/lib.rs:297-301, 635-639).
CPU test on rows produced by the oracle (the device rows are bit-identical to them: tests/test_gpu_ida.py)."""
import json
import re

import numpy as np

from helpers import ROBERTS_ATOL, ROBERTS_RTOL, ROBERTS_TOUTS


def _rows():
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts
    p = np.array([[0.04], [1.0e4], [3.0e7]])
    yy0, yp0 = Roberts.consistent_ic(p)
    d, q, st = ob.attempt_trace(ob.PK_ROBERTS, p[:, 0], yy0[:, 0], yp0[:, 0], ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS)
    assert (st == 0).all()
    return d, q


def test_frames_follow_the_reference_schema(tmp_path):
    from rust_ida_b200 import trace_json
    d, q = _rows()
    path = tmp_path / "roberts_b200.json"
    nfr = trace_json.write(str(path), d, q, 3, ROBERTS_RTOL, ROBERTS_ATOL)
    assert nfr == 377  # the frame count of the reference author's notebook
    text = path.read_text()
    assert text.startswith('{"data":[\n')  # src/lib.rs:300
    data = json.loads(text)["data"]
    assert len(data) == 377
    # what scripts/data_trace.ipynb reads, exactly as it reads it
    phi = np.array([s["ida_phi"]["data"] for s in data]).reshape((len(data), 6, 3))
    yy = np.array([s["nlp"]["ida_yy"]["data"] for s in data])
    kk = np.array([s["ida_kk"] for s in data])
    nst = np.array([s["counters"]["ida_nst"] for s in data])
    assert np.array_equal(phi[:, 0, :].view(np.uint64), np.ascontiguousarray(d[:, 0:3]).view(np.uint64))  # floats round-trip
    assert np.array_equal(yy.view(np.uint64), np.ascontiguousarray(d[:, 18 + 30 + 6:18 + 30 + 9]).view(np.uint64))
    assert np.array_equal(kk, q[:, 0]) and np.array_equal(nst, q[:, 7])
    assert data[0]["ida_phi"] == {"v": 1, "dim": [6, 3], "data": data[0]["ida_phi"]["data"]}  # ndarray's serde layout
    assert data[0]["ida_kk"] == 1 and nst[0] == 0 and nst[-1] == 361  # frame k = the state at the top of attempt k


def test_key_order_is_the_struct_declaration_order():
    """serde writes fields in declaration order: compare the frame's keys with `pub struct Ida` as the reference
    declares it (field names copied from src/lib.rs:89-244, less the cfg-gated file handle)"""
    from rust_ida_b200 import trace_json
    d, q = _rows()
    fr = trace_json.frame(d[5], q[5], 3, ROBERTS_RTOL, ROBERTS_ATOL)
    want = """ida_setup_done tol_control ida_constraints_set ida_suppressalg ida_phi ida_psi ida_alpha ida_beta ida_sigma
    ida_gamma ida_delta ida_id ida_constraints ida_ee ida_tstop ida_kk ida_kused ida_knew ida_phase ida_ns ida_hin ida_h0u
    ida_hh ida_hused ida_rr ida_tretlast ida_cjlast ida_eps_newt ida_epcon ida_maxncf ida_maxnef ida_maxord ida_mxstep
    ida_hmax_inv counters ida_cvals ida_dvals ida_tolsf ida_nrtfn ida_iroots ida_rootdir ida_tlo ida_thi ida_trout ida_glo
    ida_ghi ida_grout ida_toutc ida_ttol ida_taskc ida_irfnd ida_nge ida_gactive ida_mxgnull ida_zvecs nls nlp""".split()
    assert list(fr.keys()) == want
    assert list(fr["counters"].keys()) == ["ida_nst", "ida_nre", "ida_ncfn", "ida_netf", "ida_nni"]  # src/lib.rs:72-84
    assert list(fr["nlp"].keys())[:6] == ["ida_yy", "ida_yp", "ida_yypredict", "ida_yppredict", "ida_ewt", "ida_savres"]
    for name in trace_json.UNTRACED:  # every field declared untraced exists in the frame
        node = fr
        for part in name.split("."):
            assert part in node, name
            node = node[part]
    assert re.fullmatch(r"Normal", fr["ida_taskc"])
