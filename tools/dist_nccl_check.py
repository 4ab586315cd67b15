#!/usr/bin/env python
"""Multi-GPU check of pyramaterised_b200/dist.py over NCCL (run under torchrun, one rank per GPU):
the sample-sharded expressibility / Meyer-Wallach / QFIM-EQD of one global angle stream must equal
the single-GPU results computed on rank 0 from the same stream.  Prints one JSON line on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tools/dist_nccl_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyramaterised_b200 as pyqc                       # noqa: E402
from pyramaterised_b200 import dist as pdist, engine    # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = pdist.rank_world()
    out = {"world": world}

    # expressibility + entanglement, generic_HE 10q x 10 layers (config 2 shape, S = 3001)
    S = 3001
    qc = pyqc.templates.generate_circuit("generic_HE", 10, 10)
    ang = np.random.default_rng(1).random((S, qc.n_true_params)) * 2 * np.pi
    lo, hi = pdist.shard_bounds(S, rank, world)
    st = qc.run_batch(ang[lo:hi])
    e = pdist.sharded_expressibility(st, S, 2.0 ** 10)
    qm, qs = pdist.gathered_mean_std(engine.meyer_wallach(st), S)
    # the block-streamed form (column blocks broadcast over NVLink from their owner, the next
    # owned column generated ahead on a side stream): one round, and several rounds
    streamed = []
    for rb in (None, 1):
        stt = {}
        es = pdist.streamed_expressibility(
            lambda a, b: qc.program.run(ang[a:b], init=qc.initial_state.tensor), S, 211, 2.0 ** 10,
            resident_blocks=rb, stats=stt)
        streamed.append((es, stt["rounds"], stt["broadcasts"], stt["prefetched"]))
    if rank == 0:
        full = qc.run_batch(ang)
        pairs = S * (S - 1) // 2
        hist, _ = engine.fidelity_hist(full, bins=engine.n_bins(pairs))
        e1 = float(engine.kl_haar(hist, 2.0 ** 10).item())
        q1 = engine.meyer_wallach(full).cpu().numpy()
        out["expr_sharded"], out["expr_single"] = e, e1
        out["expr_equal"] = bool(e == e1)
        out["mw_equal"] = bool(qm == np.mean(q1) and qs == np.std(q1))
        out["streamed"] = [{"expr": x[0], "rounds": x[1], "broadcasts": x[2], "prefetched": x[3]}
                           for x in streamed]
        out["streamed_equal"] = bool(all(x[0] == e1 for x in streamed))

    # QFIM + EQD, TFIM 12q x 4 layers, 257 parameter sets
    S2 = 257
    qc2 = pyqc.templates.generate_circuit("TFIM", 12, 4)
    ang2 = np.random.default_rng(2).random((S2, qc2.n_true_params)) * 2 * np.pi
    eq = pdist.sharded_qfim_eqd(qc2, ang2, 1e-10)
    if rank == 0:
        F = qc2.qfim_batch(ang2)
        eq1 = engine.count_greater(engine.eigvalsh(F), 1e-10)
        out["eqd_equal"] = bool(torch.equal(eq.cpu(), eq1.cpu()))
        out["ok"] = bool(out["expr_equal"] and out["mw_equal"] and out["eqd_equal"] and
                         out["streamed_equal"])
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not out["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
