"""Exchange micro-benchmark (torchrun, one rank per GPU): layers of H on every qubit force one global<->local qubit
exchange per global qubit per layer; prints the NVLink rate of the exchanges (qcb_stats)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from qclojure_b200 import _lib as L
    from qclojure_b200 import circuits as C

    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(L.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    n = int(os.environ.get("XB_QUBITS", "30")) + world.bit_length() - 1
    circ = C.create_circuit(n)
    for _ in range(4):
        for q in range(n):
            C.add_gate(circ, "h", target=q)
    with L.StateVector(n, device=local_rank, rank=rank, world_size=world, nccl_id=bytes(idt.cpu().numpy().tobytes())) as sv:
        for it in range(3):
            sv.set_zero()
            sv.apply_circuit(circ)
            st = sv.stats()
            if rank == 0 and st["n_exchanges"]:
                print(f"iter {it}: exchanges {st['n_exchanges']} bytes/rank {st['bytes_exchanged']} ms {st['exchange_ms']:.2f} "
                      f"-> {st['bytes_exchanged'] / st['exchange_ms'] / 1e6:.1f} GB/s per direction; sweeps {st['n_sweeps']} gpu_ms {st['gpu_ms']:.1f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
