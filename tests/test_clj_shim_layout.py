"""The Clojure shim (clj/.../b200_simulator.clj) cannot be compiled here (no JVM in the image), but everything in it that
would silently corrupt memory can be checked without one: the struct offsets / sizes it writes by hand, the enum values of
the gate kinds, and the C symbols it binds - against include/qcb200.h as compiled by gcc and against libqcb200.so."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "clj", "src", "org", "soulspace", "qclojure", "adapter", "backend", "b200_simulator.clj")


def _src():
    with open(SHIM) as f:
        return f.read()


def _balanced(text, start):
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "{":
            depth += 1
        elif text[i] == "}":
            depth -= 1
            if depth == 0:
                return text[start:i + 1]
    raise AssertionError("unbalanced map")


def _layouts():
    s = _src()
    m = re.search(r"\(def struct-layouts\s+\"[^\"]*\"\s*", s)
    body = _balanced(s, s.index("{", m.end() - 1))
    out = {}
    for sm in re.finditer(r":(qcb_\w+)\s+\{([^}]*)\}", body):
        out[sm.group(1)] = {k: int(v) for k, v in re.findall(r":(\w+)\s+(\d+)", sm.group(2))}
    return out


def test_struct_offsets_and_sizes_match_gcc():
    lay = _layouts()
    assert set(lay) == {"qcb_config", "qcb_op", "qcb_job_request", "qcb_job_result", "qcb_noise_entry", "qcb_noise_table"}
    lines, keys = [], []
    for st, fields in lay.items():
        for f in fields:
            keys.append((st, f))
            lines.append(f"(unsigned long)sizeof({st})" if f == "size" else f"(unsigned long)offsetof({st}, {f})")
    prog = ('#include <stdio.h>\n#include <stddef.h>\n#include "qcb200.h"\nint main(void){unsigned long v[] = {' + ", ".join(lines) +
            '}; for (unsigned i = 0; i < sizeof v / sizeof v[0]; ++i) printf("%lu\\n", v[i]); return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        with open(src, "w") as f:
            f.write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).decode().split()]
    for (st, f), g in zip(keys, got):
        assert lay[st][f] == g, f"{st}.{f}: shim says {lay[st][f]}, gcc says {g}"


def test_gate_kind_codes_match_the_header_enum():
    s = _src()
    m = re.search(r"\(def kind-code\s+\"[^\"]*\"\s*", s)
    body = _balanced(s, s.index("{", m.end() - 1))
    codes = {k: int(v) for k, v in re.findall(r":([\w-]+)\s+(\d+)", body)}
    from qclojure_b200 import ops as OPS
    assert len(codes) >= 34
    for name, code in codes.items():
        assert OPS.KIND[name] == code, name               # OPS.KIND itself is checked against the header enum by test_lib_abi
    # every gate the reference's apply-gate-to-state dispatches on is present
    for name in ("x", "h", "rz", "cnot", "crz", "iswap", "toffoli", "fredkin", "rydberg-blockade", "global-rz", "measure"):
        assert name in codes


def test_bound_symbols_are_exported_and_the_abi_version_is_current():
    from qclojure_b200 import _lib as L
    s = _src()
    names = set(re.findall(r"\(ffi \"(qcb_\w+)\"", s))
    assert {"qcb_create", "qcb_submit", "qcb_job_result_get", "qcb_job_release", "qcb_run_noisy", "qcb_noisy_set_initial_state",
            "qcb_device_count"} <= names
    lib = L.load()
    for n in names:
        assert hasattr(lib, n), n
        assert n in L.EXPORTED_SYMBOLS, n
    assert "ABI version %d" % lib.qcb_abi_version() in s


def test_parentheses_balance():
    """A cheap syntax guard for a file no compiler sees here: brackets balance outside strings, comments and char literals."""
    s = _src()
    stack, i, n = [], 0, len(s)
    pairs = {")": "(", "]": "[", "}": "{"}
    while i < n:
        c = s[i]
        if c == ";":
            while i < n and s[i] != "\n":
                i += 1
            continue
        if c == '"':
            i += 1
            while i < n and s[i] != '"':
                i += 2 if s[i] == "\\" else 1
            i += 1
            continue
        if c == "\\":
            i += 2
            continue
        if c in "([{":
            stack.append((c, i))
        elif c in ")]}":
            assert stack and stack[-1][0] == pairs[c], f"unbalanced {c!r} at offset {i} (line {s.count(chr(10), 0, i) + 1})"
            stack.pop()
        i += 1
    assert not stack, f"unclosed {stack[-1][0]!r} opened at line {s.count(chr(10), 0, stack[-1][1]) + 1}"
