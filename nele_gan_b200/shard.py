"""Multi-GPU data parallelism of the labelling path: one process per GPU, pairs
sharded by index, one small gather of the per-pair records.

The path has no exchange step (every (clean, degraded) pair is scored
independently -- the reference's own parallelism is a process pool over files,
audio_util.py:146,174,202), so the only collective is the gather of
``[SIIB, HASPI, ESTOI, raw[10], status]`` = 14 float64 per pair (112 B; 7 MB at
65 536 pairs) over NCCL / NVLink.  ``backend='gloo'`` runs the same code on CPU
tensors (used by the CPU tests with a stub scorer).
"""
import numpy as np

RECORD = 14  # 3 scores + 10 HASPI raw + status


def partition(lengths, world_size):
    """Length-sorted round-robin deal: balances ragged batches (3-10 s
    utterances) across ranks.  Returns one int64 index array per rank; their
    concatenation is a permutation of ``range(len(lengths))``."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    return [order[r::world_size].astype(np.int64) for r in range(world_size)]


def pack_records(result):
    rec = np.empty((result.scores.shape[0], RECORD), dtype=np.float64)
    rec[:, 0:3] = result.scores
    rec[:, 3:13] = result.haspi_raw
    rec[:, 13] = result.status
    return rec


def gather_records(local_rec, local_idx, n_total, group=None, device=None):
    """all_gather the per-rank records and put them back in input order.
    Every rank returns the full ``[n_total, 14]`` array."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    cap = (n_total + world - 1) // world
    buf = torch.zeros((cap, RECORD + 1), dtype=torch.float64, device=device)
    k = local_rec.shape[0]
    if k:
        buf[:k, :RECORD] = torch.from_numpy(local_rec).to(buf.device)
        buf[:k, RECORD] = torch.from_numpy(local_idx.astype(np.float64)).to(buf.device)
    buf[k:, RECORD] = -1.0
    out = torch.empty((world * cap, RECORD + 1), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy()
    keep = out[:, RECORD] >= 0
    full = np.full((n_total, RECORD), np.nan)
    full[out[keep, RECORD].astype(np.int64)] = out[keep, :RECORD]
    return full


def score_sharded(score_fn, refs, degs, group=None, device=None):
    """``score_fn(refs, degs) -> BatchResult`` is called on this rank's shard
    (``Engine.score_batch`` bound to this rank's GPU); returns the records of
    the whole batch, in input order, on every rank."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lens = [min(len(r), len(d)) for r, d in zip(refs, degs)]
    idx = partition(lens, world)[rank]
    if len(idx):
        res = score_fn([refs[i] for i in idx], [degs[i] for i in idx])
        rec = pack_records(res)
    else:
        rec = np.zeros((0, RECORD))
    return gather_records(rec, idx, len(refs), group=group, device=device)
