"""Python driver of the HOST EMULATOR (tests/emu/emu.cpp) — test infrastructure only.

Runs the scheduler's program stage by stage on NumPy slices, one slice per simulated rank; exchanges
between ranks are done here in Python (or, in the gloo test, by torch.distributed send/recv).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from qclojure_b200 import ops as OPS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libqcbemu.so")
SRCS = [os.path.join(HERE, "emu.cpp"), os.path.join(ROOT, "qclojure_b200", "csrc", "plan.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "qclojure_b200", "csrc", "plan.h"),
               os.path.join(ROOT, "qclojure_b200", "csrc", "tile_core.h"),
               os.path.join(ROOT, "include", "qcb200.h")]

S_TILE, S_EXCHANGE, S_SUM, S_GROVER = 0, 1, 2, 3


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    cuda_inc = "/usr/local/cuda/include"
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-I" + cuda_inc, *SRCS, "-o", LIB, "-pthread"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.emu_create.restype = C.c_int
        L.emu_create.argtypes = [C.POINTER(OPS.QcbConfig), C.POINTER(OPS.QcbOp), C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
        L.emu_set_support.argtypes = [C.c_uint64]
        L.emu_set_support.restype = None
        L.emu_error.restype = C.c_char_p
        L.emu_error.argtypes = [C.c_void_p]
        L.emu_destroy.argtypes = [C.c_void_p]
        for nm in ("emu_num_stages", "emu_num_rounds", "emu_num_gates"):
            getattr(L, nm).restype = C.c_uint64
            getattr(L, nm).argtypes = [C.c_void_p]
        L.emu_stage_kind.restype = C.c_int
        L.emu_stage_kind.argtypes = [C.c_void_p, C.c_uint64]
        L.emu_stage_grover.restype = C.c_int
        L.emu_stage_grover.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.emu_stage_exchange.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.emu_algorithmic_bytes.restype = C.c_double
        L.emu_algorithmic_bytes.argtypes = [C.c_void_p]
        L.emu_unfused_bytes.restype = C.c_double
        L.emu_unfused_bytes.argtypes = [C.c_void_p]
        L.emu_perm_out.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.emu_stage_num_gates.restype = C.c_uint64
        L.emu_stage_num_gates.argtypes = [C.c_void_p, C.c_uint64]
        L.emu_stage_num_rounds.restype = C.c_uint64
        L.emu_stage_num_rounds.argtypes = [C.c_void_p, C.c_uint64]
        L.emu_stage_fraction.restype = C.c_double
        L.emu_stage_fraction.argtypes = [C.c_void_p, C.c_uint64]
        L.emu_stage_max_conflict.restype = C.c_int
        L.emu_stage_max_conflict.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.emu_run_tile_stage.restype = C.c_int
        L.emu_run_tile_stage.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_local_sum.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.emu_swz.restype = C.c_uint32
        L.emu_swz.argtypes = [C.c_uint32, C.c_uint32]
        L.emu_program_words.restype = C.c_uint64
        L.emu_program_words.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


class EmuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class EmuPlan:
    """The plan of ONE rank."""

    def __init__(self, n, ops, *, rank=0, world=1, perm_in=None, support=None, **cfgkw):
        """support: what is known about the input state (plan.h: Plan::support_in) - None = nothing, 0 = it is |0...0>."""
        self.n, self.rank, self.world = n, rank, world
        lib().emu_set_support(C.c_uint64(0xFFFFFFFFFFFFFFFF if support is None else support))
        cfg = OPS.make_config(n, rank=rank, world_size=world, **cfgkw)
        arr, cnt, self._keep = OPS.encode_ops(ops)
        h = C.c_void_p()
        perm = None
        if perm_in is not None:
            perm = (C.c_int32 * n)(*perm_in)
        rc = lib().emu_create(C.byref(cfg), arr, cnt, perm, C.byref(h))
        self.h = h
        if rc != 0:
            msg = lib().emu_error(h).decode()
            lib().emu_destroy(h)
            self.h = None
            raise EmuError(rc, msg)

    def __del__(self):
        if getattr(self, "h", None):
            lib().emu_destroy(self.h)
            self.h = None

    @property
    def num_stages(self):
        return int(lib().emu_num_stages(self.h))

    @property
    def num_rounds(self):
        return int(lib().emu_num_rounds(self.h))

    @property
    def num_gates(self):
        return int(lib().emu_num_gates(self.h))

    def stage_kind(self, i):
        return int(lib().emu_stage_kind(self.h, i))

    def stage_grover(self, i):
        marked = (C.c_uint64 * 8)()
        needs = C.c_int()
        k = int(lib().emu_stage_grover(self.h, i, marked, C.byref(needs)))
        return [int(marked[j]) for j in range(k)], bool(needs.value)

    def stage_exchange(self, i):
        g, l = C.c_int(), C.c_int()
        lib().emu_stage_exchange(self.h, i, C.byref(g), C.byref(l))
        return g.value, l.value

    def perm_out(self):
        out = (C.c_int32 * self.n)()
        lib().emu_perm_out(self.h, out)
        return list(out)

    def stage_info(self, i):
        return dict(kind=self.stage_kind(i), gates=int(lib().emu_stage_num_gates(self.h, i)),
                    rounds=int(lib().emu_stage_num_rounds(self.h, i)), fraction=float(lib().emu_stage_fraction(self.h, i)))

    def max_conflict(self, i, nthreads=512):
        return int(lib().emu_stage_max_conflict(self.h, i, nthreads))

    def algorithmic_bytes(self):
        return float(lib().emu_algorithmic_bytes(self.h))

    def unfused_bytes(self):
        return float(lib().emu_unfused_bytes(self.h))

    def run_tile_stage(self, i, local_state, dev_vals, nthreads=512):
        assert local_state.dtype == np.complex128 and local_state.flags.c_contiguous
        rc = lib().emu_run_tile_stage(self.h, i, local_state.ctypes.data, dev_vals.ctypes.data, nthreads)
        if rc != 0:
            raise EmuError(rc, "run_tile_stage")


def exchange_halves(slices, gbit, lbit, n_local):
    """Swap global physical bit `gbit` with local physical bit `lbit` across all simulated ranks."""
    j = gbit - n_local
    world = len(slices)
    idx = np.arange(slices[0].shape[0], dtype=np.int64)
    lb = (idx >> lbit) & 1
    for r in range(world):
        p = r ^ (1 << j)
        if r < p:
            # rank r has rank-bit j = 0: it gives away its lbit=1 half and receives partner's lbit=0 half
            sel_r = lb == 1
            sel_p = lb == 0
            tmp = slices[r][sel_r].copy()
            slices[r][sel_r] = slices[p][sel_p]
            slices[p][sel_p] = tmp


def run_world(n, ops, state=None, *, world=1, nthreads=512, return_plans=False, **cfgkw):
    """Emulate all ranks of a `world`-GPU run in one process; returns the full state in LOGICAL order."""
    p = world.bit_length() - 1
    n_local = n - p
    if state is None:
        state = np.zeros(1 << n, dtype=np.complex128)
        state[0] = 1.0
    slices = [np.ascontiguousarray(state[r << n_local:(r + 1) << n_local]).copy() for r in range(world)]
    plans = [EmuPlan(n, ops, rank=r, world=world, **cfgkw) for r in range(world)]
    dev_vals = np.zeros(64, dtype=np.float64)
    ns = plans[0].num_stages
    for i in range(ns):
        kind = plans[0].stage_kind(i)
        if kind == S_TILE:
            for r in range(world):
                plans[r].run_tile_stage(i, slices[r], dev_vals, nthreads)
        elif kind == S_EXCHANGE:
            g, l = plans[0].stage_exchange(i)
            exchange_halves(slices, g, l, n_local)
        elif kind == S_SUM:
            tot = sum(complex(np.sum(s)) for s in slices)
            N = float(1 << n)
            dev_vals[0:4] = [-1.0, 0.0, 2.0 * tot.real / N, 2.0 * tot.imag / N]   # a' = -a + 2*mean
        elif kind == S_GROVER:
            # kernels.cu: k_grover_step — a' = alpha a + beta, sign flips of this rank's marked states, sum for the next diffusion
            al, be = complex(dev_vals[0], dev_vals[1]), complex(dev_vals[2], dev_vals[3])
            needs = False
            for r in range(world):
                marked, needs = plans[r].stage_grover(i)
                slices[r][:] = al * slices[r] + be
                for mk in marked:
                    slices[r][mk] = -slices[r][mk]
            if needs:
                tot = sum(complex(np.sum(s)) for s in slices)
                N = float(1 << n)
                dev_vals[0:4] = [-1.0, 0.0, 2.0 * tot.real / N, 2.0 * tot.imag / N]
    full = np.concatenate(slices)
    perm = plans[0].perm_out()
    if perm != list(range(n)):
        full = unpermute(full, perm, n)
    if return_plans:
        return full, plans
    return full


def unpermute(phys_state, perm, n):
    """phys_state is indexed by physical bits; perm[logical bit] = physical bit.  Return logical order."""
    idx = np.arange(1 << n, dtype=np.int64)
    phys_idx = np.zeros_like(idx)
    for lb in range(n):
        phys_idx |= ((idx >> lb) & 1) << perm[lb]
    return phys_state[phys_idx]
