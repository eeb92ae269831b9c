#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: gates/sec and effective HBM GB/s of a random
brickwork circuit (H/RX/RZ + CNOT/CZ layers, depth 20) on an fp64 state vector.

  python bench.py --gpus N --steps K --warmup W            our arm (libqcb200.so on the B200)
  python bench.py --impl reference --gpus N --steps K ...  the reference arm: the CPU restatement of the
        reference's algorithm (oracle/qc_oracle.c, all host threads) on a bounded sample of the same workload.
        The reference itself is pure Clojure and there is no JVM on the box (probed at run time).

N = 1: 30 qubits (16 GiB state, far larger than the 126 MB L2, so no L2 flush is needed between steps).
N > 1: launched under torchrun, one rank per GPU; weak scaling with 2^QUBITS amplitudes per GPU
(QUBITS + log2 N qubits in total), global qubits swapped in through NCCL send/recv.

A step = one application of the whole circuit to |0...0>.  `value` = gates/sec over exactly K steps, timed
on the device (CUDA events on the library's stream, max over ranks).  `e2e` = the same metric through the
public backend API (submit_circuit -> job_result with 1024 measurement shots): host circuit map in, host
outcomes out, host<->device copies and the host-side scheduler inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def stop(self) -> dict:
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _probe_jvm() -> str:
    for exe in ("java", "clojure", "lein"):
        if shutil.which(exe):
            return exe
    return "absent"


def cpu_leg(n: int, depth: int, budget_s: float = 20.0):
    """Times the CPU restatement (oracle/qc_oracle.c, OpenMP on all host threads) on a bounded sample of the
    SAME circuit: the first G gates of the n-qubit brickwork circuit, G chosen to fit the time budget."""
    from oracle import c_oracle as CO
    from qclojure_b200 import circuits as C
    import psutil
    avail = psutil.virtual_memory().available
    n_cpu = min(n, 30)      # bounded sample: 16 GiB of host state at most (first touch of a larger one alone takes minutes)
    while (16 << n_cpu) * 1.25 > avail and n_cpu > 20:
        n_cpu -= 1
    circ = C.random_brickwork_circuit(n_cpu, depth)
    ops = circ["operations"]
    state = np.zeros(1 << n_cpu, dtype=np.complex128)
    state[0] = 1.0
    lib = CO.lib()
    # probe: one dense gate, to size the sample
    probe = CO.encode([{"operation-type": "rx", "operation-params": {"target": 0, "angle": 0.3}}], n_cpu)
    lib.orc_apply_ops(state.ctypes.data, n_cpu, probe, 1)            # warm-up (page faults)
    t0 = time.perf_counter()
    lib.orc_apply_ops(state.ctypes.data, n_cpu, probe, 1)
    per_gate = max(time.perf_counter() - t0, 1e-6)
    g = int(max(4, min(len(ops), budget_s / per_gate)))
    arr = CO.encode(ops[:g], n_cpu)
    state[:] = 0
    state[0] = 1.0
    t0 = time.perf_counter()
    rc = lib.orc_apply_ops(state.ctypes.data, n_cpu, arr, g)
    dt = time.perf_counter() - t0
    assert rc == 0
    return {"value": g / dt, "unit": "gates/s", "cores": CO.num_threads(), "kind": "port",
            "sample": f"first {g} of {len(ops)} gates of the {n_cpu}-qubit depth-{depth} brickwork circuit, "
                      f"oracle/qc_oracle.c in-place pairwise update with OpenMP ({dt:.1f} s); reference JVM: {_probe_jvm()}",
            "qubits": n_cpu, "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.qubits + int(math.log2(max(1, args.gpus)))
    vals = []
    leg = None
    budget = max(3.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        leg = cpu_leg(n, args.depth, budget_s=budget)
        if i >= args.warmup:
            vals.append(leg["value"])
    v = float(np.mean(vals))
    leg["value"] = v
    line = {"impl": "reference", "metric": "gates_per_sec", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * leg["seconds"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"random brickwork circuit, {n} qubits, depth {args.depth} (bounded sample per step)",
                       "qubits": n, "depth": args.depth},
            "cpu_baseline": leg,
            "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from qclojure_b200 import _lib as L
    from qclojure_b200 import backend as B
    from qclojure_b200 import circuits as C
    from qclojure_b200 import ops as OPS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(L.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    p = int(math.log2(world))
    n = args.qubits + p
    circ = C.random_brickwork_circuit(n, args.depth)
    ops = circ["operations"]
    n_gates = len(ops)
    enc = OPS.encode_ops(ops)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sv = L.StateVector(n, device=local_rank, rank=rank, world_size=world, nccl_id=nccl_id,
                       fusion=args.fusion, max_stage_cost=args.stage_cost, max_stage_rounds=args.stage_rounds,
                       tile_bits=args.tile_bits, low_bits=args.low_bits, dense_mma=args.dense_mma, tile_mover=args.tile_mover)

    def step():
        sv.set_zero()
        sv.apply_ops(enc)

    for _ in range(args.warmup):
        step()
    sv.synchronize()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    sv.timer_start()                # CUDA events on the stream the kernels are launched on
    for _ in range(args.steps):
        step()
    gpu_ms = sv.timer_stop()        # exactly K steps, device time
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1000.0
    stats = sv.stats()              # per-call statistics of the last step (sweeps, bytes, device ms of apply_ops)
    clocks = sampler.stop()
    t = torch.tensor([gpu_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_ms, wall_ms = float(t[0]), float(t[1])
    norm = sv.norm2()

    # ---- the HBM-bound configuration of the same kernel on the same circuit (at most two tensor-core rounds per sweep):
    # fewer gates/s than the default five-round sweeps, but this is where the north-star ">= 75 % of the HBM roofline"
    # is read off; reported next to the headline, never instead of it
    hbm_leg = None
    if world == 1 and args.fusion and args.stage_rounds == 0 and not args.no_hbm_leg:
        try:
            with L.StateVector(n, device=local_rank, fusion=1, max_stage_rounds=2, tile_bits=args.tile_bits,
                               low_bits=args.low_bits) as sv2:
                for _ in range(2):
                    sv2.set_zero(); sv2.apply_ops(enc)
                sv2.synchronize()
                sv2.timer_start()
                for _ in range(3):
                    sv2.set_zero(); sv2.apply_ops(enc)
                ms2 = sv2.timer_stop()
                st2 = sv2.stats()
            peak2, _src = _peaks()
            ach2 = st2["algorithmic_bytes"] / (st2["gpu_ms"] / 1000.0) / 1e9
            hbm_leg = {"max_stage_rounds": 2, "gates_per_sec": n_gates * 3 / (ms2 / 1000.0), "ms_per_step": ms2 / 3,
                       "sweeps_per_step": st2["n_sweeps"], "rounds_per_step": st2["n_rounds"], "achieved": ach2, "peak": peak2,
                       "unit": "GB/s", "frac": ach2 / peak2}
        except Exception as ex:    # noqa: BLE001 — an auxiliary leg must never take the headline down
            hbm_leg = {"error": str(ex)}

    # ---- e2e: public backend API, host circuit in -> host shots out
    e2e = None
    if world == 1 and not args.no_e2e:
        sim = B.create_simulator({"device": local_rank, "max-state-qubits": 26})
        sim._svs[n] = sv                       # reuse the resident state vector (HBM holds one 16 GiB state)
        shots = 1024
        u = np.random.default_rng(20261017).random(shots)
        opt = {"result-specs": {"measurements": {"shots": shots}}, "uniforms": u}
        for _ in range(min(2, args.warmup)):
            B.execute_circuit(sim, circ, opt, poll_s=0.001, max_polls=600000)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = B.execute_circuit(sim, circ, opt, poll_s=0.001, max_polls=600000)
            assert res["job-status"] == "completed", res
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        prog_words = L.plan_summary(n, ops, fusion=args.fusion, max_stage_cost=args.stage_cost,
                                    max_stage_rounds=args.stage_rounds, tile_bits=args.tile_bits,
                                    low_bits=args.low_bits)["program_words"]
        e2e = {"value": n_gates * args.steps / e2e_s, "unit": "gates/s",
               "h2d_bytes_per_step": int(prog_words * 8 + shots * 8), "d2h_bytes_per_step": int(shots * 8 + 8),
               "ms_per_step": 1000.0 * e2e_s / args.steps,
               "api": "backend.execute_circuit(B200Simulator, circuit, {:result-specs {:measurements {:shots 1024}}})"}
        sim._svs.pop(n, None)
    elif world > 1 and not args.no_e2e:
        # multi-GPU: the same metric through the C ABI with HOST buffers on every rank (qcb_set_zero, qcb_apply_ops on the
        # host op array, qcb_sample on host uniforms -> host outcomes), wall clock between barriers, max over ranks
        shots = 1024
        u = np.random.default_rng(20261017).random(shots)
        sv.set_zero(); sv.apply_ops(enc); sv.sample(u)          # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sv.set_zero()
            sv.apply_ops(enc)
            outcomes = sv.sample(u)
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        assert outcomes.shape[0] == shots
        prog_words = L.plan_summary(n, ops, fusion=args.fusion, max_stage_cost=args.stage_cost, max_stage_rounds=args.stage_rounds,
                                    tile_bits=args.tile_bits, low_bits=args.low_bits, rank=rank, world_size=world)["program_words"]
        e2e = {"value": n_gates * args.steps / e2e_s, "unit": "gates/s",
               "h2d_bytes_per_step": int(world * (prog_words * 8 + shots * 8)), "d2h_bytes_per_step": int(world * shots * 8),
               "ms_per_step": 1000.0 * e2e_s / args.steps,
               "api": "C ABI per rank: qcb_set_zero + qcb_apply_ops(host qcb_op[]) + qcb_sample(host uniforms -> host outcomes)"}

    if rank == 0:
        peak, peak_src = _peaks()
        sweeps = max(1, stats["n_sweeps"])
        alg_bytes = stats["algorithmic_bytes"]                 # per step, per rank
        achieved = alg_bytes / (stats["gpu_ms"] / 1000.0 - stats["exchange_ms"] / 1000.0) / 1e9 if stats["gpu_ms"] > 0 else 0.0
        traffic = None
        prof = os.path.join(ROOT, "profiles", "tile_stage_traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        tile_s = (stats["gpu_ms"] - stats["exchange_ms"]) / 1000.0
        # second roofline of the fused kernel: every tensor-core round is a dense 16x16 real block per 8 amplitudes
        # = 32 fp64 MAC per amplitude; peak = DMMA rate measured by scripts/dmma_bench2 (profiles/r1c_dmma_microbench.log)
        dmma_flops = 64.0 * stats["n_rounds"] * float(1 << args.qubits)
        dmma_peak = 37.0
        line = {
            "metric": "gates_per_sec", "value": n_gates * args.steps / (gpu_ms / 1000.0), "unit": "gates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": gpu_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"random brickwork circuit (H/RX/RZ + CNOT/CZ), {n} qubits, depth {args.depth}, fp64, gate fusion "
                                   f"{'on' if args.fusion else 'off'}", "qubits": n, "qubits_per_gpu": args.qubits, "depth": args.depth,
                       "gates": n_gates, "seed": 1000 + n, "l2": "state 16*2^n B >> 126 MB L2, no flush needed",
                       "parallelism": f"top {p} qubits global, NCCL send/recv qubit swaps" if p else "single GPU"},
            "effective_hbm_gbs": stats["unfused_bytes"] * args.steps / (gpu_ms / 1000.0) / 1e9,
            # weak scaling: the circuit grows by one qubit (twice the amplitudes, ~3 % more gates) per doubling of N, so
            # gates/s of the whole job cannot grow with N; amplitude updates per second (gates x 2^n / time, all ranks)
            # is the quantity whose per-GPU share stays constant under perfect weak scaling
            "amplitude_updates_per_sec": n_gates * float(1 << n) * args.steps / (gpu_ms / 1000.0),
            "sweeps_per_step": stats["n_sweeps"], "rounds_per_step": stats["n_rounds"],
            "gates_per_sweep": n_gates / sweeps,
            "roofline": {"bound": "hbm", "kernel": "k_tile_stage", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / sweeps, "avg_launch_ms": (stats["gpu_ms"] - stats["exchange_ms"]) / sweeps,
                         "fp64_tensor": {"achieved": dmma_flops / tile_s / 1e12 if tile_s > 0 else 0.0, "peak": dmma_peak, "unit": "TFLOP/s",
                                         "frac": dmma_flops / tile_s / 1e12 / dmma_peak if tile_s > 0 else 0.0,
                                         "peak_source": "mma.m16n8k16.f64 microbenchmark on this pool (profiles/r1c_dmma_microbench.log)",
                                         "flops_per_launch": dmma_flops / sweeps}},
            "exchange": {"count": stats["n_exchanges"], "bytes_sent_per_rank": stats["bytes_exchanged"], "ms": stats["exchange_ms"],
                         "gbs_per_direction": (stats["bytes_exchanged"] / (stats["exchange_ms"] / 1000.0) / 1e9) if stats["exchange_ms"] > 0 else None,
                         "frac_of_900": (stats["bytes_exchanged"] / (stats["exchange_ms"] / 1000.0) / 1e9 / 900.0) if stats["exchange_ms"] > 0 else None},
            "gpu_launches": int(stats["n_kernel_launches"] + 1) * args.steps,
            "wall_ms_per_step": wall_ms / args.steps, "norm": norm, "clocks": clocks,
        }
        if hbm_leg:
            line["roofline"]["hbm_bound_config"] = hbm_leg
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_leg(n, args.depth, budget_s=15.0)
            except Exception as ex:    # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    sv.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=int(os.environ.get("QCB_BENCH_QUBITS", "30")), help="qubits per GPU")
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--fusion", type=int, default=1)
    ap.add_argument("--stage-cost", type=int, default=0)
    ap.add_argument("--stage-rounds", type=int, default=0)
    ap.add_argument("--dense-mma", type=int, default=0, help="0/1 = tensor-core rounds (default), 2 = interpreter only")
    ap.add_argument("--tile-mover", type=int, default=0, help="0/1 = cp.async mover (default), 2 = TMA tensor-copy mover")
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--low-bits", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-hbm-leg", action="store_true", help="skip the auxiliary 2-rounds-per-sweep (HBM-bound) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
