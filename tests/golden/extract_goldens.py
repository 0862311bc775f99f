#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into a JSON fixture.

Run once in the build container (where /root/reference is mounted):

    python tests/golden/extract_goldens.py

It parses the `array![...]` literals of the reference's unit tests
(`/root/reference/src/lib.rs:880-1380`) and of its runnable examples
(`examples/fft2.rs:30-46`, `examples/rfft2.rs:36-40`) and writes
`tests/golden/reference_goldens.json`.  Only NUMBERS are extracted (test
fixtures); no reference source is copied.  The GPU box has no /root/reference,
so tests read the committed JSON only.
"""
import json
import os
import re
import sys

REF = os.environ.get("NDFB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def _rows(block):
    """Parse the body of an `array![ [..], [..] ]` or `array![a, b, c]` literal of reals."""
    inner = re.findall(r"\[([^\[\]]*)\]", block)
    if inner:
        return [[float(v) for v in re.findall(NUM, r)] for r in inner if re.search(NUM, r)]
    return [float(v) for v in re.findall(NUM, block)]


def _array_after(lines, start, name):
    """Return (values, first_line, last_line) of `let <name> ... = array![ ... ];` at/after line `start`."""
    pat = re.compile(r"let\s+(?:mut\s+)?" + re.escape(name) + r"\b")
    for i in range(start, len(lines)):
        if pat.search(lines[i]):
            j = i
            buf = []
            while True:
                buf.append(lines[j])
                if "];" in lines[j]:
                    break
                j += 1
            text = "".join(buf)
            body = text[text.index("array![") + len("array![") : text.rindex("]")]
            return _rows(body), i + 1, j + 1
    raise KeyError(name)


def _fn_line(lines, fn):
    for i, l in enumerate(lines):
        if re.search(r"fn\s+" + re.escape(fn) + r"\s*\(", l):
            return i
    raise KeyError(fn)


def _complex_rows(text):
    """Parse rows of Complex::new(a, b) literals."""
    rows = []
    for row in re.findall(r"\[((?:\s*Complex::new\([^)]*\)\s*,?\s*)+)\]", text):
        rows.append([[float(a), float(b)] for a, b in re.findall(r"Complex::new\(\s*(" + NUM + r")\s*,\s*(" + NUM + r")\s*\)", row)])
    return rows


def main():
    src = os.path.join(REF, "src", "lib.rs")
    lines = open(src).read().splitlines(keepends=True)
    g = {"_source": "preiter93/ndrustfft v0.5.0, src/lib.rs unit tests + examples", "_tolerance_abs": 1e-3}

    i = _fn_line(lines, "test_matrix")
    # test_matrix returns the literal directly (no `let`): parse from `array![` to `]` + newline
    j = i
    while "array![" not in lines[j]:
        j += 1
    k = j
    while lines[k].strip() != "]":
        k += 1
    body = "".join(lines[j : k + 1])
    g["test_matrix"] = {"values": _rows(body[body.index("array![") + 7 :]), "cite": f"src/lib.rs:{i+1}-{k+1}"}

    def grab(fn, names):
        s = _fn_line(lines, fn)
        out = {}
        for nm in names:
            vals, a, b = _array_after(lines, s, nm)
            out[nm] = vals
            out.setdefault("cite", f"src/lib.rs:{a}")
            out["cite_end"] = b
        out["cite"] = f"{out['cite']}-{out.pop('cite_end')}"
        return out

    g["test_fft"] = grab("test_fft", ["solution_re", "solution_im"])
    g["test_fft_f_layout"] = grab("test_fft_f_layout", ["solution_re", "solution_im"])
    g["test_fft_r2c"] = grab("test_fft_r2c", ["solution_re", "solution_im"])
    g["test_ifft_c2r_first_last_element"] = grab(
        "test_ifft_c2r_first_last_element", ["solution_numpy_first_elem", "solution_numpy_last_elem"]
    )
    g["test_fft_r2c_odd"] = grab("test_fft_r2c_odd", ["v"])
    for k_ in (1, 2, 3, 4):
        g[f"test_dct{k_}"] = grab(f"test_dct{k_}", ["solution"])

    ex = open(os.path.join(REF, "examples", "fft2.rs")).read()
    blk = ex[ex.index("let numpy_vhat") : ex.index("Zip::from(&vhat)")]
    g["example_fft2"] = {
        "input_real": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]],
        "numpy_vhat": _complex_rows(blk),
        "tol": 1e-4,
        "cite": "examples/fft2.rs:14-51",
    }
    ex = open(os.path.join(REF, "examples", "rfft2.rs")).read()
    blk = ex[ex.index("let numpy_vhat") : ex.index("Zip::from(&vhat)")]
    g["example_rfft2"] = {
        "input_real": [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]],
        "numpy_vhat": _complex_rows(blk),
        "tol": 1e-4,
        "cite": "examples/rfft2.rs:22-45",
    }
    # examples/fft_norm.rs:17-32 — expected printed values (comments in the example)
    g["example_fft_norm"] = {
        "input_real": [1.0, 2.0, 3.0],
        "default_roundtrip": [1.0, 2.0, 3.0],
        "none_roundtrip": [3.0, 6.0, 9.0],
        "custom_2_over_len_roundtrip": [2.0, 4.0, 6.0],
        "cite": "examples/fft_norm.rs:17-40",
    }
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT, "keys:", [k for k in g if not k.startswith("_")])


if __name__ == "__main__":
    sys.exit(main())
