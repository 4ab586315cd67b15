"""Sample-sharded multi-GPU execution: one process per GPU, torch.distributed.

The parameter-sample batch is the only sharded axis; a single statevector never spans GPUs.
  * state generation, Meyer-Wallach, magic / GKP, QFIM / EQD: rank r owns a contiguous block
    of rows of the ONE global angle array (drawn once, so the RNG stream matches the
    single-process run); per-sample results are all-gathered -- no data-path collective.
  * expressibility (measure.py:123-197) needs all unordered pairs of the whole sample set:
    states are all-gathered, every rank histograms a balanced share of the upper triangle and
    the int64 histograms are summed with one all-reduce; KL is then evaluated on every rank.

The compute callables default to the CUDA engine; the CPU (gloo) tests inject numpy stand-ins
so the partition / collective logic is covered without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_rows, rank, world):
    """Contiguous rows [lo, hi) of rank `rank`; the first n_rows % world ranks get one extra."""
    q, r = divmod(n_rows, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_rows(array, rank=None, world=None):
    if rank is None:
        rank, world = rank_world()
    lo, hi = shard_bounds(len(array), rank, world)
    return array[lo:hi]


def all_gather_rows(local, n_rows):
    """Concatenate the per-rank row blocks (shard_bounds order) on every rank."""
    rank, world = rank_world()
    if world == 1:
        return local
    sizes = [shard_bounds(n_rows, r, world)[1] - shard_bounds(n_rows, r, world)[0]
             for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    if local.is_complex():           # NCCL has no complex dtype: ship as float64 pairs
        parts = [torch.empty_like(torch.view_as_real(buf)) for _ in range(world)]
        dist.all_gather(parts, torch.view_as_real(buf).contiguous())
        parts = [torch.view_as_complex(p) for p in parts]
    else:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def triangle_blocks(n_rows, rank, world):
    """Balanced share of the upper triangle: rows are cut into 2*world blocks and rank r takes
    blocks r and 2*world-1-r (a long row block paired with a short one).  Returns a list of
    (lo, hi): this rank histograms pairs (i, j) with lo <= i < hi and j > i."""
    nb = 2 * world
    edges = [shard_bounds(n_rows, b, nb) for b in range(nb)]
    mine = [edges[rank], edges[nb - 1 - rank]]
    return [(lo, hi) for lo, hi in mine if hi > lo]


def _engine_pair_hist(A, B, triangular, bins):
    from . import engine
    hist, _ = engine.fidelity_hist(A, None if triangular else B, bins=bins)
    return hist


def sharded_fidelity_hist(all_states, bins, pair_hist=_engine_pair_hist):
    """int64 histogram of |<psi_i|psi_j>|^2 over ALL i<j of `all_states` (identical on every
    rank), each rank computing only its share of the pairs."""
    rank, world = rank_world()
    S = all_states.shape[0]
    hist = torch.zeros((bins,), dtype=torch.int64, device=all_states.device)
    for lo, hi in triangle_blocks(S, rank, world):
        own = all_states[lo:hi]
        hist += pair_hist(own, own, True, bins)
        if hi < S:
            hist += pair_hist(own, all_states[hi:], False, bins)
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def sharded_expressibility(local_states, n_total, hilbert_dim, pair_hist=_engine_pair_hist,
                           kl=None):
    """Measurements.expressibility over a sample set whose rows are sharded across ranks."""
    from . import engine
    allst = all_gather_rows(local_states, n_total)
    n_pairs = n_total * (n_total - 1) // 2
    if n_pairs == 0:
        return 0
    bins = engine.n_bins(n_pairs)
    if bins <= 0:
        raise ValueError("`bins` must be positive, when an integer")
    hist = sharded_fidelity_hist(allst, bins, pair_hist)
    if kl is None:
        return float(engine.kl_haar(hist, hilbert_dim).item())
    return kl(hist, hilbert_dim)


def streamed_rows(n_blocks, rank, world):
    """Block rows of the streamed pair sweep owned by `rank`: rows are dealt in a zig-zag
    (r, 2w-1-r, 2w+r, ...) because row I costs n_blocks - I block generations."""
    return [i for i in range(n_blocks) if (i % (2 * world)) in (rank, 2 * world - 1 - rank)]


def streamed_expressibility(run_block, n_total, block, hilbert_dim, pair_hist=_engine_pair_hist,
                            kl=None, per_block=None):
    """Measurements.expressibility for a sample set whose states cannot be resident together
    (BASELINE config 5: 4 GiB per 28-qubit state).  `run_block(lo, hi)` returns the states of
    samples [lo, hi) of the ONE global angle stream; two blocks are resident at a time and
    blocks are regenerated as needed, so every unordered pair is histogrammed exactly once
    (measure.py:133-159).  Block rows are dealt over the ranks; the int64 histograms are summed
    with one all-reduce and KL is evaluated on every rank.  `per_block(lo, hi, states)` is called
    once per owned diagonal block (e.g. to take Meyer-Wallach Q of the same states)."""
    from . import engine
    rank, world = rank_world()
    n_pairs = n_total * (n_total - 1) // 2
    if n_pairs == 0:
        return 0
    bins = engine.n_bins(n_pairs)
    if bins <= 0:
        raise ValueError("`bins` must be positive, when an integer")
    nb = (n_total + block - 1) // block
    hist = None
    for i in streamed_rows(nb, rank, world):
        lo, hi = i * block, min(n_total, (i + 1) * block)
        A = run_block(lo, hi)
        if hist is None:
            hist = torch.zeros((bins,), dtype=torch.int64, device=A.device)
        if per_block is not None:
            per_block(lo, hi, A)
        hist += pair_hist(A, A, True, bins)
        for j in range(i + 1, nb):
            B = run_block(j * block, min(n_total, (j + 1) * block))
            hist += pair_hist(A, B, False, bins)
            del B
        del A
    if hist is None:                      # a rank without rows still joins the all-reduce
        hist = torch.zeros((bins,), dtype=torch.int64,
                           device=engine.device() if torch.cuda.is_available() else "cpu")
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    if kl is None:
        return float(engine.kl_haar(hist, hilbert_dim).item())
    return kl(hist, hilbert_dim)


def sharded_qfim_eqd(circuit, global_angles, cutoff, want_qfim=False):
    """update_state + get_effective_quantum_dimension for every row of `global_angles`;
    each rank simulates its block, EQDs (int32) are gathered on every rank."""
    from . import engine
    rank, world = rank_world()
    lo, hi = shard_bounds(len(global_angles), rank, world)
    F = circuit.qfim_batch(global_angles[lo:hi])
    eq = engine.count_greater(engine.eigvalsh(F), cutoff)
    eq_all = all_gather_rows(eq.reshape(-1, 1), len(global_angles)).reshape(-1)
    return (eq_all, F) if want_qfim else eq_all


def gathered_mean_std(local_values, n_total):
    """np.mean / np.std (measure.py:431,439,447) of per-sample values sharded by rows: the
    values are gathered so the result is bit-identical to the single-process one."""
    allv = all_gather_rows(local_values.reshape(-1, 1), n_total).reshape(-1).cpu().numpy()
    return np.mean(allv), np.std(allv)
