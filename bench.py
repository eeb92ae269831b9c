#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: gates/sec and effective HBM GB/s of a random
brickwork circuit (H/RX/RZ + CNOT/CZ layers, depth 20) on an fp64 state vector.

  python bench.py --gpus N --steps K --warmup W            our arm (libqcb200.so on the B200)
  python bench.py --impl reference --gpus N --steps K ...  the reference arm: the CPU restatement of the
        reference's algorithm (oracle/qc_oracle.c, all host threads) on a bounded sample of the same workload.
        The reference itself is pure Clojure and there is no JVM on the box (probed at run time).

N = 1: 30 qubits (16 GiB state, far larger than the 126 MB L2, so no L2 flush is needed between steps).
N > 1: launched under torchrun, one rank per GPU; weak scaling with 2^QUBITS amplitudes per GPU
(QUBITS + log2 N qubits in total), global qubits swapped in through the in-place exchange kernel over NVLink peer memory.
The N > 1 line also carries, measured in the same run:
  parity          a 24-qubit circuit on the same N ranks, every rank's slice against the C oracle (max |error|, shots)
  weak_33q        BASELINE.json's size: 2^33 amplitudes (128 GiB) per GPU, 34 / 35 / 36 qubits on 2 / 4 / 8 GPUs
  single_process  the same circuit through ONE handle that owns all N GPUs (qcb_config.n_gpus), driven by rank 0 alone

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

DMMA_PEAK_TFLOPS = 37.0      # mma.m8n8k4 / m16n8k16 .f64 microbenchmark on this pool (profiles/r1c_dmma_microbench.log)


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


def _oracle_all_threads():
    """The C oracle with every host core (torchrun exports OMP_NUM_THREADS=1 to its children)."""
    from oracle import c_oracle as CO
    return CO, CO.set_num_threads(os.cpu_count() or 1)


def cpu_leg(n: int, depth: int, budget_s: float = 20.0):
    """Times the CPU restatement (oracle/qc_oracle.c, OpenMP on all host threads) on a bounded sample of the
    SAME circuit: the first G gates of the n-qubit brickwork circuit, G chosen to fit the time budget.  States above
    30 qubits (16 GiB) are sampled at 30 qubits - the sample says so."""
    from qclojure_b200 import circuits as C
    import psutil
    CO, threads = _oracle_all_threads()
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
    return {"value": g / dt, "unit": "gates/s", "cores": threads, "kind": "port",
            "sample": f"first {g} of {len(ops)} gates of the {n_cpu}-qubit depth-{depth} brickwork circuit"
                      + (f" (the {n}-qubit workload sampled at {n_cpu} qubits: host memory / first-touch time)" if n_cpu != n else "")
                      + f", oracle/qc_oracle.c in-place pairwise update with OpenMP ({dt:.1f} s); reference JVM: {_probe_jvm()}",
            "qubits": n_cpu, "seconds": dt}


def ref_faithful_leg(budget_s: float = 45.0):
    """SURVEY 8d CPU leg (i): the reference's OWN algorithm for one dense 1-qubit gate - expand it to a 2^n x 2^n matrix with
    Kronecker products and do a dense mat-vec (domain/gate.clj:346-395; oracle/qc_oracle.c: orc_apply_1q_dense_kron, one
    thread like the reference's fastmath path) - at the largest qubit count that completes inside the budget."""
    from oracle import c_oracle as CO
    rx = np.array([[math.cos(0.15), -1j * math.sin(0.15)], [-1j * math.sin(0.15), math.cos(0.15)]])
    best, spent, n = None, 0.0, 8
    while n <= 16:
        st = np.zeros(1 << n, dtype=np.complex128)
        st[0] = 1.0
        t0 = time.perf_counter()
        try:
            CO.apply_1q_dense_kron(st, n // 2, rx)
        except MemoryError:
            break
        dt = time.perf_counter() - t0
        spent += dt
        best = {"qubits": n, "seconds_per_gate": dt, "gates_per_sec": 1.0 / dt, "matrix_bytes": 16 * 4 ** n}
        if dt * 4.5 > budget_s - spent:       # one more qubit costs ~4x
            break
        n += 1
    if best:
        best.update({"cores": 1, "kind": "port of the reference's dense Kronecker-expand + mat-vec (gate.clj:346-395)",
                     "note": "O(4^n) per gate: the largest n that completes inside the budget; the in-place leg above is the kinder O(2^n) form"})
    return best


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
    try:
        leg["ref_faithful"] = ref_faithful_leg(30.0)
    except Exception as ex:      # noqa: BLE001
        leg["ref_faithful"] = {"error": str(ex)}
    line = {"impl": "reference", "metric": "gates_per_sec", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * leg["seconds"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"random brickwork circuit, {n} qubits, depth {args.depth} (bounded sample per step: "
                                   f"{leg['qubits']} qubits on the host)", "qubits": n, "sampled_qubits": leg["qubits"], "depth": args.depth},
            "cpu_baseline": leg,
            "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _roofline(stats, n_local, peak, peak_src, mma_form_flops):
    """Both roofs of the fused kernel from one step's statistics: HBM (algorithmic bytes) and the fp64 tensor pipe
    (executed DMMA flops); `bound` = whichever needs more time at its peak."""
    sweeps = max(1, stats["n_sweeps"])
    tile_s = max(1e-9, (stats["gpu_ms"] - stats["exchange_ms"]) / 1000.0)
    alg_bytes = stats["algorithmic_bytes"]
    hbm = alg_bytes / tile_s / 1e9
    flops_exec = mma_form_flops * stats["n_rounds"] * float(1 << n_local)
    flops_alg = 64.0 * stats["n_rounds"] * float(1 << n_local)       # dense 8x8 complex block: 8 complex MAC per amplitude
    tf = flops_exec / tile_s / 1e12
    t_hbm, t_mma = alg_bytes / (peak * 1e9), flops_exec / (DMMA_PEAK_TFLOPS * 1e12)
    bound = "hbm" if t_hbm >= t_mma else "tensor"
    r = {"bound": bound, "kernel": "k_tile_stage",
         "achieved": hbm if bound == "hbm" else tf, "peak": peak if bound == "hbm" else DMMA_PEAK_TFLOPS,
         "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": (hbm / peak) if bound == "hbm" else tf / DMMA_PEAK_TFLOPS,
         "traffic": None,
         "bound_note": "bound = the roof that needs more time for this plan: algorithmic bytes / HBM peak vs executed DMMA flops / "
                       "fp64-tensor peak (a sweep with r tensor-core rounds moves 32 B and executes 48 r flop per amplitude)",
         "hbm": {"achieved": hbm, "peak": peak, "unit": "GB/s", "frac": hbm / peak, "peak_source": peak_src,
                 "algorithmic_bytes_per_launch": alg_bytes / sweeps, "ideal_ms_per_step": 1e3 * t_hbm},
         "fp64_tensor": {"achieved": tf, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / DMMA_PEAK_TFLOPS,
                         "peak_source": "DMMA microbenchmark on this pool (profiles/r1c_dmma_microbench.log; 296 TF / 8 GPUs in the HGX spec)",
                         "flops_per_launch": flops_exec / sweeps, "flops_per_amplitude_per_round": mma_form_flops,
                         "algorithmic_tflops": flops_alg / tile_s / 1e12, "ideal_ms_per_step": 1e3 * t_mma,
                         # what the instruction mix of a three-product round allows on the fp64 pipe: 6 DMMA.8x8x4 (16.1 cycles each)
                         # + 4 DADD (3 cycles each, same pipe) = 108 cycles per batch against 96 for the DMMAs alone
                         "instruction_mix_bound": {"cycles_per_batch": 108, "dmma_cycles_per_batch": 96, "frac_of_peak": 96.0 / 108.0,
                                                   "frac_of_mix_bound": (tf / DMMA_PEAK_TFLOPS) * 108.0 / 96.0,
                                                   "source": "scripts/dmma_mix.cu on this pool (profiles/r2o_dmma_mix.log)"}},
         "avg_launch_ms": 1e3 * tile_s / sweeps, "launches_per_step": stats["n_sweeps"]}
    return r


def _traffic_record(n, world, args):
    """dram bytes per launch from the committed ncu capture - only for the configuration it was taken on."""
    prof = os.path.join(ROOT, "profiles", "tile_stage_traffic.json")
    if world != 1 or n != 30 or args.stage_rounds or args.stage_cost or args.tile_bits or not args.fusion or not os.path.exists(prof):
        return None, None
    try:
        with open(prof) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch"), d.get("source")
    except Exception:
        return None, None


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_nccl_id():
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(L.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    p = int(math.log2(world))
    n = args.qubits + p
    mma_flops = 64.0 if (args.dense_mma == 3 or os.environ.get("QCB_MMA_FORM") == "1") else 48.0
    peak, peak_src = _peaks()

    def make_sv(nq, **kw):
        return L.StateVector(nq, device=local_rank, rank=rank, world_size=world, nccl_id=new_nccl_id(),
                             fusion=args.fusion, max_stage_cost=args.stage_cost, max_stage_rounds=args.stage_rounds,
                             tile_bits=args.tile_bits, low_bits=args.low_bits, dense_mma=args.dense_mma, tile_mover=args.tile_mover, **kw)

    # ---- N > 1: parity of the sharded path in this very run (24 qubits, every rank's slice against the C oracle)
    parity = None
    if world > 1 and not args.no_parity:
        try:
            np_ = 24
            pc = C.random_brickwork_circuit(np_, 20)
            u = np.random.default_rng(24).random(1024)
            wt = torch.empty(2 << np_, dtype=torch.float64, device="cuda")
            ref_shots = torch.empty(1024, dtype=torch.int64, device="cuda")
            if rank == 0:
                from oracle import qc_oracle as O
                CO, _thr = _oracle_all_threads()
                want = CO.apply_circuit(pc)
                wt.copy_(torch.from_numpy(want.view(np.float64)))
                ref = CO.sample(want, u)
                bd = O.sample_boundary_distance(want, u)
                ref_shots.copy_(torch.from_numpy(np.where(bd > 1e-12, ref, -1)))     # -1: within 1e-12 of a cumulative boundary
            dist.broadcast(wt, 0)
            dist.broadcast(ref_shots, 0)
            lc = 1 << (np_ - p)
            mine = wt[2 * rank * lc: 2 * (rank + 1) * lc].cpu().numpy().view(np.complex128)
            with make_sv(np_) as svp:
                svp.apply_circuit(pc)
                pst = svp.stats()
                shots = svp.sample(u)
                got = svp.get_state()
            err = float(np.max(np.abs(got - mine)))
            rs = ref_shots.cpu().numpy()
            same = bool(np.all((rs < 0) | (rs == shots)))
            err, bad = maxr(err, 0.0 if same else 1.0)
            parity = {"n": np_, "ranks": world, "max_abs_err": err, "shots_identical": bad == 0.0, "shots": 1024,
                      "exchanges": pst["n_exchanges"], "against": "oracle/qc_oracle.c on rank 0's host cores, every rank its own slice"}
            del wt
        except Exception as ex:      # noqa: BLE001
            parity = {"error": str(ex)}

    circ = C.random_brickwork_circuit(n, args.depth)
    ops = circ["operations"]
    n_gates = len(ops)
    enc = OPS.encode_ops(ops)
    sv = make_sv(n)

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
    gpu_ms, wall_ms, xms = maxr(gpu_ms, wall_ms, stats["exchange_ms"])
    stats["exchange_ms"] = xms
    norm = sv.norm2()

    # ---- the HBM-bound configuration of the same kernel on the same circuit (at most two tensor-core rounds per sweep):
    # fewer gates/s than the default sweeps, but this is where the north-star ">= 75 % of the HBM roofline" is read off
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
            ach2 = st2["algorithmic_bytes"] / (st2["gpu_ms"] / 1000.0) / 1e9
            hbm_leg = {"max_stage_rounds": 2, "gates_per_sec": n_gates * 3 / (ms2 / 1000.0), "ms_per_step": ms2 / 3,
                       "sweeps_per_step": st2["n_sweeps"], "rounds_per_step": st2["n_rounds"], "achieved": ach2, "peak": peak,
                       "unit": "GB/s", "frac": ach2 / peak}
        except Exception as ex:    # noqa: BLE001 — an auxiliary leg must never take the headline down
            hbm_leg = {"error": str(ex)}

    # ---- e2e: public backend API, host circuit in -> host shots out
    e2e = None
    if world == 1 and not args.no_e2e:
        sim = B.create_simulator({"device": local_rank, "max-state-qubits": 26})
        sv.generation = 0
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
               "api": "backend.execute_circuit(B200Simulator, circuit, {:result-specs {:measurements {:shots 1024}}}) - the Python "
                      "mirror of the reference's blocking helper (application/backend.clj:209-256).  Two things the reference API "
                      "does not have: the option \"uniforms\" (caller-supplied draws, so that shots are reproducible) and the bench "
                      "handing its resident 16 GiB state vector to the backend instead of letting it allocate a second one"}
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
        e2e_s = maxr(time.perf_counter() - t0)[0]
        assert outcomes.shape[0] == shots
        prog_words = L.plan_summary(n, ops, fusion=args.fusion, max_stage_cost=args.stage_cost, max_stage_rounds=args.stage_rounds,
                                    tile_bits=args.tile_bits, low_bits=args.low_bits, rank=rank, world_size=world)["program_words"]
        e2e = {"value": n_gates * args.steps / e2e_s, "unit": "gates/s",
               "h2d_bytes_per_step": int(world * (prog_words * 8 + shots * 8)), "d2h_bytes_per_step": int(world * shots * 8),
               "ms_per_step": 1000.0 * e2e_s / args.steps,
               "api": "C ABI per rank: qcb_set_zero + qcb_apply_ops(host qcb_op[]) + qcb_sample(host uniforms -> host outcomes)"}

    # ---- streaming reductions on the resident state (measurement / expectation path of VQE, QAOA and shots): algorithmic
    # bytes = one 16 B read per amplitude per pass, device time from CUDA events on the library's stream around each call
    reductions = None
    if world == 1 and not args.no_other:
        try:
            nbytes = 16.0 * float(1 << n)
            def timed_gbs(f, passes=1.0, reps=3):
                f()
                sv.timer_start()
                for _ in range(reps):
                    f()
                ms = sv.timer_stop() / reps
                return {"ms": ms, "gbs": passes * nbytes / (ms / 1000.0) / 1e9, "frac_of_hbm_peak": passes * nbytes / (ms / 1000.0) / 1e9 / peak}
            zz = [{"coefficient": 1.0, "pauli-string": "".join("Z" if q in (a, a + 1) else "I" for q in range(n))} for a in range(0, 16)]
            xs = [{"coefficient": 1.0, "pauli-string": "".join("X" if q == a else "I" for q in range(n))} for a in (3,)]
            ush = np.random.default_rng(3).random(1024)
            reductions = {
                "unit": "GB/s of 16 B per amplitude per pass; peak = " + peak_src,
                "norm (k_reduce)": timed_gbs(lambda: sv.norm2()),
                "expect_1q (k_expect_1q)": timed_gbs(lambda: sv.expect_1q(np.array([[0, 1], [1, 0]]), 5)),
                "16 ZZ terms in one pass (k_expect_group, diagonal)": timed_gbs(lambda: sv.expect_hamiltonian(zz)),
                "1 X term (k_expect_group, pairs)": timed_gbs(lambda: sv.expect_hamiltonian(xs)),
                "1024 shots (k_chunk_sums + scan + k_sample)": timed_gbs(lambda: sv.sample(ush)),
                "marginal of 3 qubits (k_marginal)": timed_gbs(lambda: sv.marginal_probabilities([0, 7, n - 1])),
            }
        except Exception as ex:      # noqa: BLE001
            reductions = {"error": str(ex)}

    def exchange_block(st, step_ms):
        xs = st["exchange_ms"]
        gbs = (st["bytes_exchanged"] / (xs / 1000.0) / 1e9) if xs > 0 else None
        return {"count": st["n_exchanges"], "bytes_sent_per_rank": st["bytes_exchanged"], "ms": xs,
                "frac_of_step": (xs / step_ms) if step_ms > 0 else None,
                "gbs_per_direction": gbs, "frac_of_900": (gbs / 900.0) if gbs else None,
                "frac_of_measured_770": (gbs / 770.0) if gbs else None}

    sv.close()

    # ---- N > 1: BASELINE.json's weak-scaling size, 2^33 amplitudes (128 GiB) per GPU: 34 / 35 / 36 qubits on 2 / 4 / 8 GPUs
    weak33 = None
    if world > 1 and not args.no_weak33 and args.qubits < 33:
        try:
            n33 = 33 + p
            c33 = C.random_brickwork_circuit(n33, args.depth)
            e33 = OPS.encode_ops(c33["operations"])
            barrier()
            with make_sv(n33) as s33:
                s33.set_zero(); s33.apply_ops(e33); s33.synchronize()        # warm-up
                barrier()
                k33 = 2
                s33.timer_start()
                for _ in range(k33):
                    s33.set_zero(); s33.apply_ops(e33)
                ms33 = s33.timer_stop()
                barrier()
                st33 = s33.stats()
                nrm33 = s33.norm2()
            ms33, x33 = maxr(ms33, st33["exchange_ms"])
            st33["exchange_ms"] = x33
            g33 = len(c33["operations"])
            weak33 = {"qubits": n33, "qubits_per_gpu": 33, "state_gib_per_gpu": 128, "gates": g33, "steps": k33, "warmup": 1,
                      "gates_per_sec": g33 * k33 / (ms33 / 1000.0), "ms_per_step": ms33 / k33,
                      "amplitude_updates_per_sec_per_gpu": g33 * float(1 << 33) * k33 / (ms33 / 1000.0),
                      "sweeps_per_step": st33["n_sweeps"], "rounds_per_step": st33["n_rounds"], "norm": nrm33,
                      "exchange": exchange_block(st33, ms33 / k33),
                      "roofline": _roofline(st33, 33, peak, peak_src, mma_flops)}
        except Exception as ex:      # noqa: BLE001
            weak33 = {"error": str(ex)}
        barrier()

    # ---- N > 1: the same circuit through ONE handle that owns all N GPUs (rank 0's process alone; the other ranks have
    # released their state and wait) - the mode the Clojure host uses
    single = None
    if world > 1 and not args.no_single_process:
        barrier()
        # the other ranks wait on the HOST (a file flag): an NCCL barrier would park a spinning kernel on every GPU that the
        # single handle is about to use (first measured that way: 1783 instead of 3912 gates/s on 8 GPUs)
        flag = os.path.join("/tmp", "qcb_bench_single_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()))
        if rank != 0:
            t_wait = time.time()
            while not os.path.exists(flag) and time.time() - t_wait < 900:
                time.sleep(0.05)
        if rank == 0:
            try:
                with L.StateVector(n, n_gpus=world, fusion=args.fusion, max_stage_cost=args.stage_cost,
                                   max_stage_rounds=args.stage_rounds, tile_bits=args.tile_bits, low_bits=args.low_bits,
                                   dense_mma=args.dense_mma) as sg:
                    sg.set_zero(); sg.apply_ops(enc); sg.synchronize()
                    ks = max(2, min(args.steps, 3))
                    sg.timer_start()
                    for _ in range(ks):
                        sg.set_zero(); sg.apply_ops(enc)
                    mss = sg.timer_stop()
                    sts = sg.stats()
                    idx = np.random.default_rng(5).integers(0, 1 << n, 16)
                    amps = sg.get_amplitudes(idx)
                    nrm = sg.norm2()
                single = {"api": "one qcb_handle, qcb_config.n_gpus = %d, one host thread per device inside libqcb200.so" % world,
                          "gates_per_sec": n_gates * ks / (mss / 1000.0), "ms_per_step": mss / ks, "steps": ks,
                          "sweeps_per_step": sts["n_sweeps"], "exchanges": sts["n_exchanges"], "norm": nrm,
                          "spot_amplitude_abs_max": float(np.max(np.abs(amps)))}
            except Exception as ex:      # noqa: BLE001
                single = {"error": str(ex)}
            with open(flag, "w") as f:
                f.write("done")
        barrier()
        if rank == 0:
            try:
                os.remove(flag)
            except OSError:
                pass

    if rank == 0:
        step_ms = gpu_ms / args.steps
        plan0 = L.plan_summary(n, ops, fusion=args.fusion, max_stage_cost=args.stage_cost, max_stage_rounds=args.stage_rounds,
                               tile_bits=args.tile_bits, low_bits=args.low_bits, rank=0, world_size=world)
        roof = _roofline(stats, args.qubits, peak, peak_src, mma_flops)
        if os.environ.get("QCB_ZERO_SKIP", "0") not in ("", "0"):
            # experimental zero-state support path: the first sweeps visit a fraction of their tiles; the HBM figures
            # follow (algorithmic bytes are counted per visited tile), the tensor figures count every round at full size
            roof["zero_skip"] = {"enabled": True, "note": "QCB_ZERO_SKIP=1: fp64_tensor figures overstate the executed flops "
                                                          "(rounds of partially visited sweeps are counted in full)"}
        roof["traffic"], tsrc = _traffic_record(n, world, args)
        if tsrc:
            roof["traffic_source"] = tsrc + " (a committed capture of this configuration, not a measurement of this run)"
        line = {
            "metric": "gates_per_sec", "value": n_gates * args.steps / (gpu_ms / 1000.0), "unit": "gates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"random brickwork circuit (H/RX/RZ + CNOT/CZ), {n} qubits, depth {args.depth}, fp64, gate fusion "
                                   f"{'on' if args.fusion else 'off'}", "qubits": n, "qubits_per_gpu": args.qubits, "depth": args.depth,
                       "gates": n_gates, "seed": 1000 + n, "l2": "state 16*2^n B >> 126 MB L2, no flush needed",
                       "parallelism": f"top {p} qubits global, in-place qubit exchange kernel over NVLink peer memory" if p else "single GPU"},
            "effective_hbm_gbs": stats["unfused_bytes"] * args.steps / (gpu_ms / 1000.0) / 1e9,
            # weak scaling: the circuit grows by one qubit (twice the amplitudes, ~3 % more gates) per doubling of N, so
            # gates/s of the whole job cannot grow with N; amplitude updates per second (gates x 2^n / time, all ranks)
            # is the quantity whose per-GPU share stays constant under perfect weak scaling
            "amplitude_updates_per_sec": n_gates * float(1 << n) * args.steps / (gpu_ms / 1000.0),
            "amplitude_updates_per_sec_per_gpu": n_gates * float(1 << args.qubits) * args.steps / (gpu_ms / 1000.0),
            "sweeps_per_step": stats["n_sweeps"], "rounds_per_step": stats["n_rounds"],
            # passes over the shared-memory tile: a paired pass applies two dense 8x8 rounds to registers between one load
            # and one store of the tile (rank 0's plan)
            "passes_per_step": plan0["passes"], "paired_passes_per_step": plan0["paired_passes"],
            "gates_per_sweep": n_gates / max(1, stats["n_sweeps"]),
            "hbm_frac": roof["hbm"]["frac"], "fp64_tensor_frac": roof["fp64_tensor"]["frac"],
            "roofline": roof,
            "exchange": exchange_block(stats, step_ms),
            "gpu_launches": int(stats["n_kernel_launches"] + 1) * args.steps,
            "wall_ms_per_step": wall_ms / args.steps, "norm": norm, "clocks": clocks,
        }
        if hbm_leg:
            line["roofline"]["hbm_bound_config"] = hbm_leg
        if e2e:
            line["e2e"] = e2e
        if reductions is not None:
            line["reductions"] = reductions
        if parity is not None:
            line["parity"] = parity
        if weak33 is not None:
            line["weak_33q"] = weak33
        if single is not None:
            line["single_process"] = single
        if world == 1 and not args.no_other:
            try:
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                import bench_configs
                s2 = ClockSampler(local_rank)
                s2.start()
                oc = bench_configs.run(device=local_rank, quick=True)
                oc["clocks"] = s2.stop()
                line["other_configs"] = oc
            except Exception as ex:      # noqa: BLE001
                line["other_configs"] = {"error": str(ex)}
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = cpu_leg(n, args.depth, budget_s=15.0)
                line["cpu_baseline"]["ref_faithful"] = ref_faithful_leg(20.0)
            except Exception as ex:    # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line))
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
    ap.add_argument("--dense-mma", type=int, default=0, help="0/1 = tensor-core rounds, three-product form (default); 2 = interpreter only; 3 = 16x16 real form")
    ap.add_argument("--tile-mover", type=int, default=0, help="0/1 = cp.async mover (default), 2 = TMA tensor-copy mover")
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--low-bits", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-hbm-leg", action="store_true", help="skip the auxiliary 2-rounds-per-sweep (HBM-bound) measurement")
    ap.add_argument("--no-other", action="store_true", help="skip the other BASELINE.json configs (QFT / Grover / noisy / QAOA timings)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the 24-qubit sharded parity check against the C oracle")
    ap.add_argument("--no-weak33", action="store_true", help="N > 1: skip the 33-qubits-per-GPU block")
    ap.add_argument("--no-single-process", action="store_true", help="N > 1: skip the one-handle-for-all-GPUs leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
