#!/usr/bin/env python3
"""Extract the golden vectors held by the reference's own unit tests into tests/golden/*.json.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tools/extract_golden.py [/root/reference]

Only literal test DATA is extracted (the numbers inside `let x = ...;`, `array![...]`,
`vector![...]`, `matrix![...]` and `ida.field = ...;` statements), in source order, per `#[test]`
function.  No reference code is copied.  The JSON is what tests/test_oracle_golden.py replays against
the oracle (SURVEY.md section 8c "Pins").

Sources (reference file -> golden file):
    src/tests/{complete_step,set_coeffs,predict,restore,test_error,get_solution,nonlinear_solve}.rs
    crates/linear/src/dense.rs (mod tests), crates/nonlinear/src/newton.rs (test_newton)
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)"


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def split_tests(src):
    """yield (fn_name, body) for every fn following a #[test] attribute."""
    for m in re.finditer(r"#\[test\]\s*(?:#\[[^\]]*\]\s*)*fn\s+(\w+)\s*\(\)\s*\{", src):
        i = m.end()
        depth = 1
        while depth and i < len(src):
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        yield m.group(1), src[m.end():i - 1]


def parse_array(text, kind):
    text = text.strip()
    if kind == "matrix":  # rows separated by ';'
        rows = [r for r in text.split(";") if r.strip()]
        return [[float(x) for x in re.findall(NUM, r)] for r in rows]
    if text.startswith("["):  # 2-D array![[..],[..]]
        rows = re.findall(r"\[([^\[\]]*)\]", text)
        return [[float(x) for x in re.findall(NUM, r)] for r in rows]
    return [float(x) for x in re.findall(NUM, text)]


def parse_body(body):
    items = []
    pat = re.compile(
        r"(?:let\s+(?:mut\s+)?(?P<lname>\w+)(?:\s*:\s*[^=]+)?\s*=\s*|(?P<aname>ida(?:\.\w+)+)\.assign\(\s*&\s*)"
        r"(?:\w+::)*(?P<kind>array|vector|matrix)!\[(?P<arr>(?:[^\[\]]|\[[^\[\]]*\])*)\]\s*(?P<tr>\.transpose\(\))?"
        r"|let\s+(?:mut\s+)?(?P<sname>\w+)\s*=\s*(?P<sval>" + NUM + r"|true|false)\s*;"
        r"|(?P<fname>ida(?:\.\w+)+)\s*=\s*(?P<fval>" + NUM + r")\s*;",
        re.S,
    )
    for m in pat.finditer(body):
        if m.group("kind"):
            name = m.group("lname") or m.group("aname")
            items.append({"name": name, "value": parse_array(m.group("arr"), m.group("kind")),
                          "transpose": bool(m.group("tr"))})
        elif m.group("sname"):
            v = m.group("sval")
            raw = v
            val = {"true": True, "false": False}.get(v, None)
            if val is None:
                val = float(v) if re.search(r"[.eE]", v) else int(v)
            items.append({"name": m.group("sname"), "value": val, "raw": raw})
        else:
            items.append({"name": m.group("fname"), "value": float(m.group("fval")), "raw": m.group("fval")})
    return items


def main():
    os.makedirs(OUT, exist_ok=True)
    files = {
        "complete_step": "src/tests/complete_step.rs",
        "set_coeffs": "src/tests/set_coeffs.rs",
        "predict": "src/tests/predict.rs",
        "restore": "src/tests/restore.rs",
        "test_error": "src/tests/test_error.rs",
        "get_solution": "src/tests/get_solution.rs",
        "nonlinear_solve": "src/tests/nonlinear_solve.rs",
        "dense": "crates/linear/src/dense.rs",
        "newton": "crates/nonlinear/src/newton.rs",
    }
    for key, rel in files.items():
        src = strip_comments(open(os.path.join(REF, rel)).read())
        out = {"source": rel, "tests": {}}
        for name, body in split_tests(src):
            out["tests"][name] = parse_body(body)
        with open(os.path.join(OUT, key + ".json"), "w") as f:
            json.dump(out, f, indent=1)
        print(key, {k: len(v) for k, v in out["tests"].items()})


if __name__ == "__main__":
    main()
