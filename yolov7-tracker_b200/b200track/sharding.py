"""Multi-GPU bookkeeping.

Video sequences are independent units: sequence s lives on rank ``s % world`` (or, for a
streaming service, each rank owns its own S sequences).  The data path has NO collective.  The only
coupling is the reference's process-global track-id counter (``BaseTrack._count``,
tracker/basetrack.py:22,43-46, never reset between sequences -- q8): processed one after another,
sequence s sees ids offset by the births of all earlier sequences.  With per-sequence local
counters on the device the same ids are recovered by ONE all-gather of the per-sequence birth
counts followed by an exclusive scan.
"""
import torch
import torch.distributed as dist


def shard_sequences(n_sequences, rank, world):
    """Sequence indices owned by ``rank`` (round-robin, as SURVEY 8e proposes)."""
    return list(range(rank, n_sequences, world))


def global_id_offsets(local_births, seq_ids, n_sequences, group=None):
    """local_births[k] = tracks born so far in sequence seq_ids[k] (device or CPU int64 tensor).
    Returns an int64 tensor of length n_sequences: offset[s] = sum of births of sequences < s.
    One all_gather of (n_sequences) int64 -- O(100 B), latency only."""
    local_births = torch.as_tensor(local_births, dtype=torch.int64)
    dev = local_births.device
    mine = torch.zeros(n_sequences, dtype=torch.int64, device=dev)
    if len(seq_ids):
        mine[torch.as_tensor(seq_ids, dtype=torch.int64, device=dev)] = local_births
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [torch.zeros_like(mine) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, mine, group=group)
        total = torch.stack(parts).sum(0)          # every sequence is owned by exactly one rank
    else:
        total = mine
    return torch.cumsum(total, 0) - total


def to_global_ids(local_ids, seq, offsets):
    """Local (per-sequence, 1-based) track ids -> the ids the reference would have produced."""
    return local_ids + int(offsets[seq])
