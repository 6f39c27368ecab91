"""Per-GPU work split for the burst-DSP path: bursts are independent, so ranks share nothing.

* wideband batches (configs 4/5): ARFCN a belongs to rank a mod G -- `arfcn_shard`;
* one continuous stream (config 2): contiguous ranges of whole resampler chunks per rank, each range
  starting on a 157-sample slot boundary so the slot cutting needs no cross-rank state --
  `stream_shard`.  lcm(585-sample chunks, 625-sample slot groups) = 117 frames = 250 chunks = 936 bursts.
* optional gather of soft bits to every rank (`gather_soft`): the only collective, off the timed path.
"""
import numpy as np

CHUNKS_PER_BLOCK = 250     # 250 chunks * 585 = 146250 samples = 234 slot groups of 625 = 936 bursts = 117 frames
BURSTS_PER_BLOCK = 936


def arfcn_shard(n_arfcn, rank, world):
    """ARFCN indices owned by `rank` (a mod world == rank)."""
    return np.arange(rank, n_arfcn, world, dtype=np.int64)


def burst_rows_of_arfcns(arfcns, slots=8):
    """row indices (arfcn*slots + tn) of the bursts of the given ARFCNs in an (ARFCN, TN)-ordered batch"""
    return (np.asarray(arfcns, np.int64)[:, None] * slots + np.arange(slots)[None, :]).reshape(-1)


def stream_shard(n_blocks, rank, world):
    """[lo, hi) in units of 117-frame blocks for `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(n_blocks, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_soft(soft, group=None):
    """all_gather of per-rank soft-bit tensors (torch.distributed; NCCL over NVLink on GPUs, gloo on CPU).
    Returns the list of every rank's tensor.  Ranks may hold different burst counts."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([soft.shape[0]], dtype=torch.int64, device=soft.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    nmax = int(max(int(c.item()) for c in counts))
    pad = torch.zeros((nmax,) + tuple(soft.shape[1:]), dtype=soft.dtype, device=soft.device)
    pad[:soft.shape[0]] = soft
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:int(c.item())] for o, c in zip(outs, counts)]
