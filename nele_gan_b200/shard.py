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


def gather_records_start(local_rec, local_idx, n_total, group=None, device=None):
    """Start the all_gather of the per-rank records (asynchronous collective) and return a handle for
    :func:`gather_records_finish`.  A caller that streams many batches finishes the gather of batch k after it has
    queued batch k + 1, so the ranks meet in the collective without waiting for each other on the critical path."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    cap = (n_total + world - 1) // world
    k = local_rec.shape[0]
    host = np.full((cap, RECORD + 1), -1.0)
    host[:, :RECORD] = 0.0
    if k:
        host[:k, :RECORD] = local_rec
        host[:k, RECORD] = local_idx
    buf = torch.from_numpy(host)
    if device is not None:
        buf = buf.to(device, non_blocking=False)
    out = torch.empty((world * cap, RECORD + 1), dtype=torch.float64, device=device)
    work = dist.all_gather_into_tensor(out, buf, group=group, async_op=True)
    return {"work": work, "out": out, "buf": buf, "n_total": n_total}


def gather_records_finish(handle):
    """Wait for a gather started by :func:`gather_records_start`; every rank returns the full ``[n_total, 14]``
    array in input order."""
    handle["work"].wait()
    out = handle["out"].cpu().numpy()
    keep = out[:, RECORD] >= 0
    full = np.full((handle["n_total"], RECORD), np.nan)
    full[out[keep, RECORD].astype(np.int64)] = out[keep, :RECORD]
    return full


def gather_records(local_rec, local_idx, n_total, group=None, device=None):
    """all_gather the per-rank records and put them back in input order.
    Every rank returns the full ``[n_total, 14]`` array."""
    return gather_records_finish(gather_records_start(local_rec, local_idx, n_total, group=group, device=device))


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
