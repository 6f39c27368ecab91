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


def gather_soft(soft, group=None, counts=None):
    """all_gather of per-rank soft-bit tensors (torch.distributed; NCCL over NVLink on GPUs, gloo on CPU).
    Returns the list of every rank's tensor.
    * ranks hold the same burst count (the ARFCN split of 1024 ARFCNs over 2/4/8 ranks): pass counts="equal" -- ONE
      collective straight into one output tensor, no padding, no host synchronisation;
    * counts = list of every rank's burst count, known to the caller: one padded collective, no host synchronisation;
    * counts = None: the counts are exchanged first (a small collective and a host read)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if isinstance(counts, str) and counts == "equal":
        out = torch.empty((world,) + tuple(soft.shape), dtype=soft.dtype, device=soft.device)
        if soft.is_cuda:
            dist.all_gather_into_tensor(out.view((world * soft.shape[0],) + tuple(soft.shape[1:])), soft.contiguous(), group=group)
        else:
            dist.all_gather(list(out.unbind(0)), soft.contiguous(), group=group)
        return list(out.unbind(0))
    if counts is None:
        n = torch.tensor([soft.shape[0]], dtype=torch.int64, device=soft.device)
        cl = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(cl, n, group=group)
        counts = [int(c) for c in torch.cat(cl).tolist()]        # one host read for all ranks' counts
    nmax = max(counts)
    if all(c == nmax for c in counts):
        return gather_soft(soft, group, "equal")
    pad = torch.zeros((nmax,) + tuple(soft.shape[1:]), dtype=soft.dtype, device=soft.device)
    pad[:soft.shape[0]] = soft
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:c] for o, c in zip(outs, counts)]
