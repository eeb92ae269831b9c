"""Worker of tests/test_multi_gpu_plan.py::test_gloo_two_process_exchange (CPU, gloo backend)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import qc_oracle as O  # noqa: E402
from qclojure_b200 import circuits as C  # noqa: E402
from tests.emu import emu as E  # noqa: E402


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 12
    p = world.bit_length() - 1
    nl = n - p
    circ = C.random_brickwork_circuit(n, 8)
    C.h(circ, 0); C.cnot(circ, 0, n - 1); C.swap(circ, 0, n - 1); C.ry(circ, 0, 0.3)
    # a Grover loop on the scrambled state: the diffusion needs the sum over ALL ranks (the one real collective of the path
    # besides the qubit exchange: ncclAllReduce in sim.cu, all_reduce here)
    marked = [5, (1 << n) - 3]
    for _ in range(3):
        for mk in marked:
            C.add_gate(circ, "phase-oracle", index=mk)
        C.add_gate(circ, "grover-diffusion")
    plan = E.EmuPlan(n, circ["operations"], rank=rank, world=world, tile_bits=6, low_bits=3)
    local = np.zeros(1 << nl, dtype=np.complex128)
    if rank == 0:
        local[0] = 1.0
    dev_vals = np.zeros(64)
    idx = np.arange(1 << nl, dtype=np.int64)
    n_ex = n_sum = 0
    for i in range(plan.num_stages):
        kind = plan.stage_kind(i)
        if kind == E.S_TILE:
            plan.run_tile_stage(i, local, dev_vals)
        elif kind == E.S_EXCHANGE:
            g, l = plan.stage_exchange(i)
            j = g - nl
            myb = (rank >> j) & 1
            peer = rank ^ (1 << j)
            sel = ((idx >> l) & 1) == (1 - myb)               # my moving half: bit l == !myb
            send = torch.from_numpy(np.ascontiguousarray(local[sel]).view(np.float64).copy())
            recv = torch.empty_like(send)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)])
            for r in reqs:
                r.wait()
            local[sel] = recv.numpy().view(np.complex128)
            n_ex += 1
        elif kind in (E.S_SUM, E.S_GROVER):
            needs = kind == E.S_SUM
            if kind == E.S_GROVER:
                mk, needs = plan.stage_grover(i)
                local[:] = complex(dev_vals[0], dev_vals[1]) * local + complex(dev_vals[2], dev_vals[3])
                for m in mk:
                    local[m] = -local[m]
            if needs:
                tot = torch.tensor([local.sum().real, local.sum().imag], dtype=torch.float64)
                dist.all_reduce(tot)
                dev_vals[0:4] = [-1.0, 0.0, 2.0 * float(tot[0]) / (1 << n), 2.0 * float(tot[1]) / (1 << n)]
            n_sum += int(needs)
        else:
            raise SystemExit("unexpected stage kind")
    gathered = [torch.empty(2 << nl, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(local.view(np.float64).copy()))
    if rank == 0:
        full = np.concatenate([g.numpy().view(np.complex128) for g in gathered])
        full = E.unpermute(full, plan.perm_out(), n)
        gates = [op for op in circ["operations"] if op["operation-type"] not in ("phase-oracle", "grover-diffusion")]
        want = O.execute_circuit(dict(circ, operations=gates))
        for _ in range(3):
            for mk in marked:
                want[mk] = -want[mk]
            want = 2 * np.mean(want) - want
        err = float(np.max(np.abs(full - want)))
        assert err <= 1e-10, err
        assert n_ex >= 1 and n_sum == 3
        print(f"PARITY OK err={err:.2e} exchanges={n_ex} all-reduced sums={n_sum}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
