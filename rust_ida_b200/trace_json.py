"""Per-attempt trace -> the reference's `data_trace` JSON.

With the `data_trace` cargo feature the reference serialises the whole `Ida` struct with serde at the top of every step
attempt (src/lib.rs:297-301 opens `{"data":[`, :635-639 writes one frame + `,\\n`, Drop :246-258 closes `]}`), and the
author's notebook scripts/data_trace.ipynb reads `json["data"]` frame by frame (`step['ida_phi']['data']`,
`step['nlp']['ida_yy']['data']`, `step['ida_kk']`, ...).  This module turns the rows of `Ida.trace()` -- the device
kernel's per-attempt state dump, bit-identical to the oracle's -- into that schema: same key names, same nesting
(`counters`, `tol_control`, `nls`, `nlp`, `nlp.lp`), in the struct's declaration order, arrays in ndarray's serde
format `{"v":1,"dim":[...],"data":[...]}`.  Floats are written with repr() (17 significant digits: they round-trip).

Fields of the reference struct that are scratch vectors of one attempt (ida_delta, nlp.ida_yypredict / ida_yppredict /
ida_savres, ida_cvals / ida_dvals, nlp.lp.x) are not part of the persistent state the engine keeps per instance; they
are emitted as zero arrays of the right shape and listed under the top-level key "untraced_fields", so that the
notebook's loaders work and nobody mistakes them for data.
"""
import json

import numpy as np

MXORDP1 = 6
UNTRACED = ["ida_delta", "ida_cvals", "ida_dvals", "ida_zvecs", "nlp.ida_yypredict", "nlp.ida_yppredict", "nlp.ida_savres",
            "nlp.ida_oldnrm", "nlp.lp.x", "ida_tlo", "ida_thi", "ida_trout", "ida_glo", "ida_ghi", "ida_grout", "ida_toutc",
            "ida_ttol", "ida_irfnd", "ida_nge", "ida_gactive", "ida_iroots"]


def _nd(a):
    a = np.asarray(a)
    return {"v": 1, "dim": list(a.shape), "data": [x.item() for x in a.reshape(-1)]}


def frame(d, q, n, rtol, atol, nrtfn=0):
    """one data_trace frame from one trace row: d = 10n+45+n^2 doubles, q = 16+n ints (layout: b200ida_set_trace)"""
    d = np.asarray(d, dtype=np.float64)
    q = np.asarray(q)
    o = 6 * n
    phi = d[:o].reshape(MXORDP1, n)
    psi, alpha, beta, sigma, gamma = (d[o + 6 * k:o + 6 * k + 6] for k in range(5))
    o += 30
    ee, ewt, yy, yp = (d[o + n * k:o + n * k + n] for k in range(4))
    o += 4 * n
    (hh, hused, rr, cj, cjold, cjlast, cjratio, ss, toldel, eps_newt, tn, tretlast, h0u, tolsf, _tret) = (float(v) for v in d[o:o + 15])
    o += 15
    mat_j = d[o:o + n * n].reshape(n, n).T  # stored column-major; the reference's Array2 is mathematical (i, j)
    (kk, kused, knew, phase, ns, setup_done, jcur, nst, nre, nje, nni, netf, ncfn, nsetups, nlscf, _natt) = (int(v) for v in q[:16])
    zeros = lambda *shape: _nd(np.zeros(shape))
    lp = {"ls": {"pivots": [int(v) for v in q[16:16 + n]]}, "mat_j": _nd(mat_j), "x": zeros(n), "sqrt_n": float(np.sqrt(float(n))),
          "eplifac": 0.05, "dqincfac": 1.0, "nje": nje, "npe": 0, "nli": 0, "nps": 0, "ncfl": 0, "nre_dq": 0, "njtsetup": 0,
          "njtimes": 0, "nst0": 0, "nni0": 0, "ncfn0": 0, "ncfl0": 0, "nwarn": 0, "ida_cj": cj, "ida_cjold": cjold,
          "ida_cjratio": cjratio}
    nlp = {"ida_yy": _nd(yy), "ida_yp": _nd(yp), "ida_yypredict": zeros(n), "ida_yppredict": zeros(n), "ida_ewt": _nd(ewt),
           "ida_savres": zeros(n), "ida_tn": tn, "ida_ss": ss, "ida_oldnrm": 0.0, "ida_toldel": toldel, "ida_nre": nre,
           "ida_nsetups": nsetups, "lp": lp}
    return {
        "ida_setup_done": bool(setup_done),
        "tol_control": {"ida_rtol": float(rtol), "ida_atol": _nd(np.asarray(atol, dtype=np.float64))},
        "ida_constraints_set": False, "ida_suppressalg": False,
        "ida_phi": _nd(phi), "ida_psi": _nd(psi), "ida_alpha": _nd(alpha), "ida_beta": _nd(beta), "ida_sigma": _nd(sigma),
        "ida_gamma": _nd(gamma),
        "ida_delta": zeros(n), "ida_id": _nd(np.zeros(n, dtype=bool)), "ida_constraints": zeros(n), "ida_ee": _nd(ee),
        "ida_tstop": None,
        "ida_kk": kk, "ida_kused": kused, "ida_knew": knew, "ida_phase": phase, "ida_ns": ns,
        "ida_hin": 0.0, "ida_h0u": h0u, "ida_hh": hh, "ida_hused": hused, "ida_rr": rr,
        "ida_tretlast": tretlast, "ida_cjlast": cjlast, "ida_eps_newt": eps_newt, "ida_epcon": 0.33,
        "ida_maxncf": 10, "ida_maxnef": 10, "ida_maxord": 5, "ida_mxstep": 500, "ida_hmax_inv": 0.0,
        "counters": {"ida_nst": nst, "ida_nre": nre, "ida_ncfn": ncfn, "ida_netf": netf, "ida_nni": nni},
        "ida_cvals": zeros(MXORDP1), "ida_dvals": zeros(5), "ida_tolsf": tolsf,
        "ida_nrtfn": nrtfn, "ida_iroots": zeros(nrtfn), "ida_rootdir": _nd(np.zeros(nrtfn, dtype=np.uint8)),
        "ida_tlo": 0.0, "ida_thi": 0.0, "ida_trout": 0.0, "ida_glo": zeros(nrtfn), "ida_ghi": zeros(nrtfn),
        "ida_grout": zeros(nrtfn), "ida_toutc": 0.0, "ida_ttol": 0.0, "ida_taskc": "Normal", "ida_irfnd": False, "ida_nge": 0,
        "ida_gactive": _nd(np.zeros(nrtfn, dtype=bool)), "ida_mxgnull": 1, "ida_zvecs": zeros(MXORDP1, n),
        "nls": {"delta": zeros(n), "jcur": bool(jcur), "curiter": 0, "maxiters": 4, "niters": nni, "nconvfails": nlscf},
        "nlp": nlp,
    }


def frames(d_rows, q_rows, n, rtol, atol):
    return [frame(d, q, n, rtol, atol) for d, q in zip(d_rows, q_rows)]


def write(path, d_rows, q_rows, n, rtol, atol):
    """write `path` in the reference's layout: `{"data":[` + one frame per line, comma separated + `]}`; one extra
    top-level key, "untraced_fields", follows "data" (the notebook reads json["data"] only)"""
    fr = frames(d_rows, q_rows, n, rtol, atol)
    with open(path, "w") as f:
        f.write('{"data":[\n')
        f.write(",\n".join(json.dumps(x) for x in fr))
        f.write('\n],\n"untraced_fields":%s}\n' % json.dumps(UNTRACED))
    return len(fr)
