"""Point sharding across GPUs -- the collective-library path.

The product path exchanges scan summaries through NVLink peer inboxes inside / behind the scan
kernels (csrc/peer_exchange.cu; mc_scan_sharded_*, mc_accumulate_step_sharded).  This module is what
`bench.py --gpus N` falls back to when peer memory cannot be opened between the processes, and what
the CPU tests run over gloo: each rank holds a block of the rows, centers are replicated, and a
sharded get_close needs one tiny exchange per scan:

    positives / evaluated : SUM
    arg-max of f0         : the reference keeps the FIRST maximum in iteration order
                            (Trainer.cpp:99), i.e. lexicographic (f0 descending, global row ascending)

as three all-reduces on a handful of scalars, or one all-gather of the per-rank records folded on the
host (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

NO_ROW = 2 ** 62


def combine_scan_results(results, row_offset: int, device):
    """results: list of (n_eval, n_pos, best_row_local, best_f0) of this rank's shard, one per scan.
    Returns the same tuples for the whole (sharded) range with global row numbers."""
    counts = torch.tensor([[r[0], r[1]] for r in results], dtype=torch.int64, device=device)
    f0 = torch.tensor([r[3] if r[2] >= 0 else -1.0 for r in results], dtype=torch.float64, device=device)
    rows = torch.tensor([r[2] + row_offset if r[2] >= 0 else NO_ROW for r in results], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        fmax = f0.clone()
        dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
        rows = torch.where(f0 == fmax, rows, torch.full_like(rows, NO_ROW))
        dist.all_reduce(rows, op=dist.ReduceOp.MIN)
        f0 = fmax
    counts, f0, rows = counts.cpu(), f0.cpu(), rows.cpu()
    out = []
    for i in range(len(results)):
        row = int(rows[i])
        out.append((int(counts[i, 0]), int(counts[i, 1]), row if row != NO_ROW else -1, float(f0[i]) if row != NO_ROW else -1.0))
    return out


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank `rank` when n rows are split as evenly as possible."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def combine_scan_records(records: torch.Tensor, rows_per_rank: int):
    """Device path: `records` is an int64 tensor [S, 4] holding S mc_scan_result records
    (n_eval, n_pos, best_row, bits of best_f0) of this rank.  One all-gather, then the fold on the
    host.  Returns global tuples like combine_scan_results."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world > 1:
        if dist.get_backend() == "nccl":
            gathered = torch.empty((world,) + tuple(records.shape), dtype=records.dtype, device=records.device)
            dist.all_gather_into_tensor(gathered, records)
        else:   # gloo (CPU tests)
            parts = [torch.empty_like(records) for _ in range(world)]
            dist.all_gather(parts, records)
            gathered = torch.stack(parts)
    else:
        gathered = records.unsqueeze(0)
    g = gathered.cpu()
    f0 = g[:, :, 3].contiguous().view(torch.float64)
    out = []
    for s in range(g.shape[1]):
        n_eval = int(g[:, s, 0].sum())
        n_pos = int(g[:, s, 1].sum())
        best_row, best_f0 = -1, -1.0
        for w in range(world):                      # ranks hold ascending row blocks: first max wins
            r = int(g[w, s, 2])
            v = float(f0[w, s])
            if r >= 0 and v > best_f0:
                best_f0, best_row = v, r + w * rows_per_rank
        out.append((n_eval, n_pos, best_row, best_f0))
    return out
