"""World-size-2 gloo (CPU) tests of the sample-sharding logic in pyramaterised_b200/dist.py.
The CUDA compute callables are replaced by numpy stand-ins built on the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pqc_oracle as orc
from pyramaterised_b200 import dist as pdist


def _np_pair_hist(A, B, triangular, bins):
    a, b = A.numpy(), B.numpy()
    F = np.abs(a.conj() @ b.T) ** 2
    if triangular:
        F = F[np.triu_indices(len(a), 1)]
    return torch.from_numpy(np.histogram(F.ravel(), bins=bins, range=(0, 1))[0].astype(np.int64))


def _np_kl(hist, N):
    return orc.expr_from_counts(hist.numpy(), N)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, S, q, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        specs, _ = orc.generate_circuit("NPQC", 4, 3)
        P = orc.n_params(specs)
        ang = np.random.default_rng(1).random((S, P)) * 2 * np.pi      # ONE global stream
        lo, hi = pdist.shard_bounds(S, rank, world)
        assert np.array_equal(pdist.shard_rows(ang), ang[lo:hi])
        local = torch.from_numpy(orc.run(specs, 4, ang[lo:hi]))
        allst = pdist.all_gather_rows(local, S)
        e = pdist.sharded_expressibility(local, S, 16, pair_hist=_np_pair_hist, kl=_np_kl)
        bins = int((75 / 10000) * (S * (S - 1) // 2))
        h = pdist.sharded_fidelity_hist(allst, bins, _np_pair_hist)
        es = pdist.streamed_expressibility(
            lambda a, b: torch.from_numpy(orc.run(specs, 4, ang[a:b])), S, 7, 16,
            pair_hist=_np_pair_hist, kl=_np_kl)
        assert abs(es - e) < 1e-15                               # block-streamed == gathered
        # several rounds (one resident block per rank), interrupted and resumed from the checkpoint
        run_block = lambda a, b: torch.from_numpy(orc.run(specs, 4, ang[a:b]))   # noqa: E731
        ck = os.path.join(tmp, "ck")

        class Stop(Exception):
            pass

        def stop_at(rnd, j, n_rounds, nb):
            if (rnd, j) == (1, 5):
                raise Stop()

        try:
            pdist.streamed_expressibility(run_block, S, 7, 16, pair_hist=_np_pair_hist, kl=_np_kl,
                                          resident_blocks=1, checkpoint=ck, progress=stop_at)
            raise AssertionError("the interruption hook did not fire")
        except Stop:
            pass
        assert os.path.exists(f"{ck}.rank{rank}")
        st = {}
        er = pdist.streamed_expressibility(run_block, S, 7, 16, pair_hist=_np_pair_hist, kl=_np_kl,
                                           resident_blocks=1, checkpoint=ck, stats=st)
        assert st["resumed"] == 1 and st["rounds"] < 5           # 9 blocks, 2 per round
        assert abs(er - e) < 1e-15                               # resumed == uninterrupted
        st1 = {}
        e1 = pdist.streamed_expressibility(run_block, S, 7, 16, pair_hist=_np_pair_hist, kl=_np_kl,
                                           stats=st1)
        assert abs(e1 - e) < 1e-15 and st1["rounds"] == 1
        # every state generated exactly once over the ranks when all row blocks are resident
        gen = torch.tensor([st1["generations"]])
        dist.all_reduce(gen)
        assert int(gen) == S
        qv = torch.tensor([orc.single_Q(s, 4) for s in local.numpy()])
        mean, std = pdist.gathered_mean_std(qv, S)
        q.put((rank, allst.numpy(), e, h.numpy(), mean, std))
    finally:
        dist.destroy_process_group()


def test_world2_matches_single_process(tmp_path):
    S, world = 61, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, q, str(tmp_path)))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    specs, _ = orc.generate_circuit("NPQC", 4, 3)
    ang = np.random.default_rng(1).random((S, orc.n_params(specs))) * 2 * np.pi
    ref = orc.run(specs, 4, ang)
    F = orc.fidelity_samples(ref)
    _, _, counts = orc.gen_histo(F)
    qs = [orc.single_Q(s, 4) for s in ref]
    for rank, allst, e, h, mean, std in res:
        assert np.array_equal(allst, ref)                     # gather keeps row order, exact
        assert np.array_equal(h, counts)                      # integer counts identical
        assert abs(e - orc.expr(F, 16)) < 1e-12
        assert mean == np.mean(qs) and std == np.std(qs)      # bit-identical statistics


def test_partitions_cover_everything_once():
    for S in (1, 2, 7, 61, 100):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                lo, hi = pdist.shard_bounds(S, r, world)
                rows += list(range(lo, hi))
            assert rows == list(range(S))
            seen = np.zeros((S, S), dtype=int)
            work = []
            for r in range(world):
                w = 0
                for lo, hi in pdist.triangle_blocks(S, r, world):
                    for i in range(lo, hi):
                        seen[i, i + 1:] += 1
                        w += S - 1 - i
                work.append(w)
            assert np.array_equal(seen, np.triu(np.ones((S, S), dtype=int), 1))
            if S >= 8 * world:                                # balanced within ~25 %
                assert max(work) <= 1.25 * (sum(work) / world) + S
            seen = np.zeros((S, S), dtype=int)                # own block first, cross blocks later
            for r in range(world):
                (lo, hi), cross = pdist.shard_pair_plan(S, r, world)
                for i in range(lo, hi):
                    seen[i, i + 1:hi] += 1
                for rlo, rhi, clo, chi in cross:
                    assert (rlo, rhi) != (clo, chi)
                    a, b = np.meshgrid(np.arange(rlo, rhi), np.arange(clo, chi), indexing="ij")
                    np.add.at(seen, (np.minimum(a, b), np.maximum(a, b)), 1)
            assert np.array_equal(seen, np.triu(np.ones((S, S), dtype=int), 1))
            for nb in (1, 5, 16):                             # streamed block rows: a partition
                rows = sorted(i for r in range(world) for i in pdist.streamed_rows(nb, r, world))
                assert rows == list(range(nb))
