"""CPU checks of the drop-in boundary: libqcb200.so loads, exports every symbol include/qcb200.h
declares, the host-only planning API works, and the product path FAILS LOUDLY without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qclojure_b200 import _lib as L
from qclojure_b200 import circuits as CIR
from qclojure_b200 import ops as OPS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "qcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint32_t\s+(qcb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    L.build()
    lib = C.CDLL(L.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 45
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/qcb200.h but not exported"
    assert set(names) == set(L.EXPORTED_SYMBOLS), "ctypes prototypes out of sync with the header"
    assert L.load().qcb_abi_version() == 2


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors must have the C compiler's sizes for every struct in the header."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "qcb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %d\\n",'
                   'sizeof(qcb_op),sizeof(qcb_config),sizeof(qcb_noise_entry),sizeof(qcb_noise_table),sizeof(qcb_stats),'
                   'sizeof(qcb_job_request),sizeof(qcb_job_result),(int)QCB_OP_KIND_COUNT);return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(OPS.QcbOp), C.sizeof(OPS.QcbConfig), C.sizeof(OPS.QcbNoiseEntry), C.sizeof(OPS.QcbNoiseTable),
            C.sizeof(OPS.QcbStats), C.sizeof(OPS.QcbJobRequest), C.sizeof(OPS.QcbJobResult), len(OPS.KIND_NAMES)]
    assert got == want


def test_plan_api_is_host_only():
    circ = CIR.random_brickwork_circuit(30, 20)
    s = L.plan_summary(30, circ["operations"])
    assert s["stages"] < len(circ["operations"]) / 4 and s["exchanges"] == 0
    u = L.plan_summary(30, circ["operations"], fusion=0)
    assert u["stages"] == len(circ["operations"])
    # errors carry the reference's message
    with pytest.raises(L.QcbError) as ei:
        L.plan_summary(2, [{"operation-type": "cy", "operation-params": {"control": 0, "target": 1}}])
    assert ei.value.code == -2 and "Unknown gate type" in ei.value.message
    with pytest.raises(L.QcbError) as ei:
        L.plan_summary(2, [{"operation-type": "h", "operation-params": {"target": 5}}])
    assert ei.value.code == -1


def test_encoder_mirrors_reference_dispatch():
    arr, n, _ = OPS.encode_ops([{"operation-type": ":cx", "operation-params": {":control": 1, ":target": 0}},
                                {"operation-type": "h", "operation-params": {}},
                                {"operation-type": "p", "operation-params": {"target": 2, "angle": 0.5}}])
    assert n == 3 and arr[0].kind == OPS.KIND["cnot"] and (arr[0].q[0], arr[0].q[1]) == (1, 0)
    assert arr[1].q[0] == 0                      # missing :target defaults to qubit 0 (circuit.clj:974-976)
    assert arr[2].kind == OPS.KIND["phase"] and arr[2].angle == 0.5
    with pytest.raises(OPS.GateError):
        OPS.encode_ops([{"operation-type": "cnot", "operation-params": {"control": 1}}])
    with pytest.raises(OPS.GateError):
        OPS.encode_ops([{"operation-type": "frobnicate", "operation-params": {}}])


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (there is no CPU path to fall back to)."""
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("CUDA device present")
    assert L.device_count() == 0
    with pytest.raises(L.QcbError) as ei:
        L.StateVector(3)
    assert ei.value.code == -3 and "no CPU fallback" in ei.value.message


def test_product_never_imports_the_oracle():
    """The oracle and the host emulator are test infrastructure: nothing under qclojure_b200/ may import,
    link or load them."""
    pkg = os.path.join(ROOT, "qclojure_b200")
    banned = re.compile(r"(import\s+oracle|from\s+oracle|qc_oracle|c_oracle|libqcoracle|libqcbemu|tests\.emu|oracle/_build)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not banned.search(src), f


def test_lowering_rejects_malformed_ops_without_crashing():
    """The C ABI never throws or crashes on malformed qcb_op records (host-only planning entry): random kinds, qubit numbers
    out of range, garbage masks -> an error code with a message, or a valid plan."""
    import ctypes as CT
    from qclojure_b200 import ops as OPS
    lib = L.load()
    rng = np.random.default_rng(0)
    n = 6
    cfg = OPS.make_config(n)
    rejected = 0
    for _ in range(600):
        k = int(rng.integers(1, 6))
        arr = (OPS.QcbOp * k)()
        for i in range(k):
            o = arr[i]
            o.kind = int(rng.integers(-2, 45))
            for j in range(3):
                o.q[j] = int(rng.integers(-3, n + 3))
            o.n_mask = int(rng.integers(-1, 9))
            o.mask = int(rng.integers(0, 1 << 12))
            o.angle = float(rng.normal())
            for j in range(8):
                o.mat[j] = float(rng.normal())
        p = CT.c_void_p()
        rc = lib.qcb_plan_create(CT.byref(cfg), arr, k, CT.byref(p))
        if rc == 0:
            lib.qcb_plan_destroy(p)
        else:
            rejected += 1
            assert L.last_error(None)
    assert rejected > 300
    assert lib.qcb_plan_create(None, None, 0, None) != 0
