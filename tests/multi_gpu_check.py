"""Multi-GPU parity check of the sharded path, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_check.py

Every rank holds the slice [rank * 2^(n-p), (rank+1) * 2^(n-p)) of the n-qubit state (top p = log2 N index bits = the
reference's qubits 0..p-1 are the rank id).  Checks against the oracle on the same circuit: every local amplitude
(<= 1e-10), the norm, a Hamiltonian energy (both need an all-reduce) and the shot outcomes on identical uniforms."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from oracle import c_oracle as CO
    from oracle import qc_oracle as O
    from qclojure_b200 import _lib as L
    from qclojure_b200 import circuits as C

    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_nccl_id():          # one ncclUniqueId per communicator (= per state vector), minted by rank 0
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(L.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    p = world.bit_length() - 1
    worst = 0.0
    # QCB_MGC_LOCAL="12:6,18:8,22:8" (local qubits : depth) overrides the sizes, e.g. for a short run on a tight GPU budget
    sizes = [tuple(int(v) for v in item.split(":")) for item in os.environ.get("QCB_MGC_LOCAL", "12:6,18:8,22:8").split(",")]
    for n, depth in [(nl + p, d) for nl, d in sizes]:
        circ = C.random_brickwork_circuit(n, depth)
        want = CO.apply_circuit(circ) if n > 16 else O.execute_circuit(circ)
        u = np.random.default_rng(n).random(512)
        H = C.max_cut_hamiltonian(C.random_regular_graph(n, 3 if n % 2 == 0 else 4, seed=11), n)
        HX = C.standard_mixer_hamiltonian(n) + [
            {"coefficient": 0.7, "pauli-string": "XY" + "I" * (n - 3) + "Z"}, {"coefficient": -0.4, "pauli-string": "Y" * 3 + "I" * (n - 3)},
            {"coefficient": 0.2, "pauli-string": "Z" + "X" * (n - 4) + "III"}]
        with L.StateVector(n, device=local_rank, rank=rank, world_size=world, nccl_id=new_nccl_id()) as sv:
            sv.apply_circuit(circ)
            stats = sv.stats()
            nrm = sv.norm2()
            energy = sv.expect_hamiltonian(H)
            energy_x = sv.expect_hamiltonian(HX)      # X / Y factors on the global qubits: localised by qubit exchanges
            e1 = sv.expect_1q(np.array([[0, -1j], [1j, 0]]), 0)      # 1-qubit observable on the most global qubit
            shots = sv.sample(u)
            got = sv.get_state()
        lc = 1 << (n - p)
        err = float(np.max(np.abs(got - want[rank * lc:(rank + 1) * lc])))
        worst = max(worst, err)
        assert err <= 1e-10, f"rank {rank} n={n}: amplitude mismatch {err}"
        assert abs(nrm - 1.0) <= 1e-10, f"rank {rank} n={n}: norm {nrm}"
        assert abs(energy - O.hamiltonian_expectation(H, want)) <= 1e-9, f"rank {rank} n={n}: energy"
        assert abs(energy_x - O.hamiltonian_expectation(HX, want)) <= 1e-9, f"rank {rank} n={n}: energy with X/Y on global qubits"
        want_e1 = float(np.real(np.vdot(want, O.apply_single_qubit_gate(want, np.array([[0, -1j], [1j, 0]]), 0))))
        assert abs(e1 - want_e1) <= 1e-10, f"rank {rank} n={n}: expect_1q on a global qubit {e1} vs {want_e1}"
        ref = O.sample_outcomes(want, u)
        dist_b = O.sample_boundary_distance(want, u)
        assert not ((shots != ref) & (dist_b > 1e-12)).any(), f"rank {rank} n={n}: shot outcomes differ"
        if rank == 0:
            print(f"n={n} world={world}: max|err|={err:.2e} exchanges={stats['n_exchanges']} sweeps={stats['n_sweeps']} "
                  f"exchange_ms={stats['exchange_ms']:.3f}", flush=True)
    # QFT across the ranks: the ladder's controls on rank bits and tile-id bits ride as far phases (no exchange, no condition bits)
    for nl in (13, 17):
        n = nl + p
        circ = C.quantum_fourier_transform_circuit(n)
        init = np.random.default_rng(n).standard_normal(1 << n) + 1j * np.random.default_rng(n + 1).standard_normal(1 << n)
        init /= np.linalg.norm(init)
        want = O.execute_circuit(circ, init)
        lc = 1 << nl
        with L.StateVector(n, device=local_rank, rank=rank, world_size=world, nccl_id=new_nccl_id()) as sv:
            sv.set_state(init[rank * lc:(rank + 1) * lc])
            sv.apply_circuit(circ)
            stats = sv.stats()
            got = sv.get_state()
        err = float(np.max(np.abs(got - want[rank * lc:(rank + 1) * lc])))
        worst = max(worst, err)
        assert err <= 1e-10, f"rank {rank} QFT n={n}: amplitude mismatch {err}"
        if rank == 0:
            print(f"QFT n={n} world={world}: max|err|={err:.2e} exchanges={stats['n_exchanges']} sweeps={stats['n_sweeps']}", flush=True)
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"multi-gpu ok: world={world} max|err|={float(t[0]):.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
