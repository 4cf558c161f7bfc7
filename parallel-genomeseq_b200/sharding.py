"""Sharding of independent alignments over the GPUs of one box, and the single collective of the path.

Alignments are independent, so the path shards with no data-path exchange: reads / DB
entries are block-partitioned over ranks exactly like the reference's MPI drivers partition them over
worker ranks (contiguous blocks of floor(n / workers), the last worker takes the remainder:
mpi_sw_solve_small.cpp:52-55, mpi_sw_solve_uniprot.cpp:71), the reference sequence is replicated, and
the per-read (score, pos) results are funnelled with ONE all-gather (the reference funnels 136-byte
structs with MPI_Send to a writer rank, mpi_sw_solve_small.cpp:142,169).  Backend-agnostic: NCCL over
NVLink on the GPU box, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def block_partition(n_items, world):
    """[(begin, end)] per rank: floor(n/world) each, remainder to the last rank."""
    q = n_items // world
    out = [(r * q, (r + 1) * q) for r in range(world)]
    out[-1] = ((world - 1) * q, n_items)
    return out


def gather_score_pos(score, pos, counts=None, group=None):
    """All-gather per-rank (score, pos) int32 tensors into global read order.

    score/pos: 1-D int32 tensors of this rank's shard (CUDA for nccl, CPU for gloo).
    counts   : shard sizes per rank when they differ (ragged last shard); None = all equal.
    Returns (score_all, pos_all) on every rank.  8 bytes per read cross the fabric."""
    world = dist.get_world_size(group)
    n = score.numel()
    if counts is None:
        out_s = torch.empty(world * n, dtype=score.dtype, device=score.device)
        out_p = torch.empty(world * n, dtype=pos.dtype, device=pos.device)
        dist.all_gather_into_tensor(out_s, score.contiguous(), group=group)
        dist.all_gather_into_tensor(out_p, pos.contiguous(), group=group)
        return out_s, out_p
    mx = max(counts)
    pad_s = torch.zeros(mx, dtype=score.dtype, device=score.device)
    pad_p = torch.zeros(mx, dtype=pos.dtype, device=pos.device)
    pad_s[:n] = score
    pad_p[:n] = pos
    out_s = torch.empty(world * mx, dtype=score.dtype, device=score.device)
    out_p = torch.empty(world * mx, dtype=pos.dtype, device=pos.device)
    dist.all_gather_into_tensor(out_s, pad_s, group=group)
    dist.all_gather_into_tensor(out_p, pad_p, group=group)
    keep = torch.cat([torch.arange(r * mx, r * mx + c, device=score.device) for r, c in enumerate(counts)])
    return out_s[keep], out_p[keep]
