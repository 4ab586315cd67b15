#!/usr/bin/env python
"""BASELINE config 5's path (NPQC n qubits x p layers, expressibility + entanglement of S states
that cannot be resident together) on 1..8 GPUs at a sample count that fits the GPU budget.
Launch: python tools/c5_streamed.py n p S block [resident_blocks]      (1 GPU), or
        python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/c5_streamed.py ...
Rank 0 prints one JSON object: wall seconds (max over ranks), generations / broadcasts per rank,
the rates the full-size extrapolation of DESIGN.md uses."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyramaterised_b200 as pyqc          # noqa: E402
from pyramaterised_b200 import engine      # noqa: E402


def main():
    n, p, S, block = (int(x) for x in sys.argv[1:5])
    resident = int(sys.argv[5]) if len(sys.argv) > 5 else None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    qc = pyqc.templates.generate_circuit("NPQC", n, p)
    m = pyqc.measure.Measurements(qc)
    # warm-up: plan upload, NCCL communicator, allocator pools
    pyqc.gates.rng.bit_generator.state = np.random.default_rng(7).bit_generator.state
    m.expressibility_streamed(min(S, max(17, 2 * block * world)), block, want_Q=True, resident_blocks=1)
    pyqc.gates.rng.bit_generator.state = np.random.default_rng(1).bit_generator.state
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stats = {"profile": bool(os.environ.get("C5_PROFILE"))}
    t0 = time.perf_counter()
    e, Q = m.expressibility_streamed(S, block, want_Q=True, resident_blocks=resident, stats=stats)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    g = torch.tensor([stats["generations"], stats["broadcasts"], stats["broadcast_bytes"], len(Q)],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        gs = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(gs, g)
    else:
        gs = [g]
    if rank == 0:
        pairs = S * (S - 1) // 2
        gens = [int(x[0].item()) for x in gs]
        print(json.dumps({
            "config": f"C5 streamed NPQC {n}q x {p} layers, S={S}, block={block}, "
                      f"resident_blocks={resident}, {world} GPU(s): expressibility + entanglement",
            "seconds": float(dt.item()), "expr": e, "pairs": pairs, "bins": engine.n_bins(pairs),
            "state_bytes": 16 * 2 ** n, "rounds": stats["rounds"], "prefetched_rank0": stats["prefetched"],
            "generations_per_rank": gens, "generations_total": sum(gens),
            "broadcasts": int(gs[0][1].item()), "broadcast_bytes_total": int(gs[0][2].item()),
            "Q_values_per_rank": [int(x[3].item()) for x in gs],
            "phase_seconds_rank0": {k: round(stats[k], 3) for k in ("t_rows", "t_colgen", "t_bcast", "t_hist")}
            if stats.get("profile") else None,
            "pairs_per_s": pairs / float(dt.item()),
            "generations_per_s_all_ranks": sum(gens) / float(dt.item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
