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


def balanced_partition(lengths, world):
    """Database search: entries sorted by length, each assigned to the rank with the
    least total residues so far (greedy LPT), so every GPU gets the same number of cells although lengths are
    ragged.  Returns one ascending int64 index tensor per rank; deterministic, so every rank computes the same."""
    lengths = torch.as_tensor(lengths, dtype=torch.int64)
    order = torch.argsort(lengths, descending=True, stable=True).tolist()
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    return [torch.tensor(sorted(ix), dtype=torch.int64) for ix in out]


def gather_by_index(values, parts, group=None):
    """All-gather this rank's per-entry results (1-D tensor, one value per index of parts[rank]) and put them
    back into global entry order.  `parts` = balanced_partition(...) (identical on every rank)."""
    world = dist.get_world_size(group)
    counts = [int(p.numel()) for p in parts]
    mx = max(counts)
    pad = torch.zeros(mx, dtype=values.dtype, device=values.device)
    pad[: values.numel()] = values
    out = torch.empty(world * mx, dtype=values.dtype, device=values.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    total = sum(counts)
    res = torch.empty(total, dtype=values.dtype, device=values.device)
    for r, p in enumerate(parts):
        res[p.to(values.device)] = out[r * mx: r * mx + counts[r]]
    return res


def make_index_gather(parts, device, group=None):
    """gather_by_index for MANY rounds over the same partition (a database search runs one round per query): the padded
    receive buffer and the permutation back into entry order are built once, every round is one all-gather of the packed
    (score, pos) words and one index_select on the device.  Returns fn(score_i32, pos_i32) -> (score_all, pos_all)."""
    world = dist.get_world_size(group)
    counts = [int(p.numel()) for p in parts]
    mx = max(counts)
    total = sum(counts)
    src = torch.empty(total, dtype=torch.int64)          # position in the gathered buffer of every entry
    for r, p in enumerate(parts):
        src[p] = torch.arange(r * mx, r * mx + counts[r], dtype=torch.int64)
    src = src.to(device)
    pad = torch.zeros(mx, dtype=torch.int64, device=device)
    out = torch.empty(world * mx, dtype=torch.int64, device=device)

    def fn(score, pos):
        n = score.numel()
        pad[:n] = (score.to(torch.int64) << 32) | (pos.to(torch.int64) & 0xFFFFFFFF)
        dist.all_gather_into_tensor(out, pad, group=group)
        both = out.index_select(0, src)
        return (both >> 32).to(torch.int32), (both & 0xFFFFFFFF).to(torch.int32)
    return fn


def reference_sharded_align(align_piece, reads, y, ratio, make_string_range, group=None, realign=None, device="cpu"):
    """Long-pair config: the REFERENCE is split over the ranks into `world` overlapping
    ranges with the reference's own rule (_make_string_range, plocalaligner.cpp:44-67: halo = floor(ratio * m)),
    every rank aligns all reads against its range, and one all-reduce(max) picks, per read, the lowest-index
    range with the strictly greatest score — the result of the serial
    OMPParallelLocalAligner(x, y, npiece = world, ratio) (plocalaligner.cpp:106-143).

    align_piece(reads, y_piece) -> (scores, positions): this rank's aligner (the engine with y_piece as its
    reference; the oracle in the CPU tests).  `realign` (same signature), when given, re-aligns the reads a rank
    won — the reference constructs the final aligner with the DEFAULT scoring (plocalaligner.cpp:135),
    so callers with a custom scoring pass their default-scoring aligner here.
    Returns (score, pos, winner_rank) int64 tensors in read order on every rank; pos is global (left edge added,
    plocalaligner.cpp:137).  The consensus of read i lives on rank winner_rank[i]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_reads = len(reads)
    score = torch.zeros(n_reads, dtype=torch.int64)
    pos = torch.zeros(n_reads, dtype=torch.int64)
    left_of = torch.zeros(n_reads, dtype=torch.int64)
    by_len = {}
    for i, x in enumerate(reads):
        by_len.setdefault(len(x), []).append(i)
    pieces = {}
    for m, idx in sorted(by_len.items()):
        rng = make_string_range(world, m, len(y), ratio)           # raises when the reference's asserts would fire
        left, right = rng[rank]
        pieces[m] = (left, right)
        s, p = align_piece([reads[i] for i in idx], y[left:right])
        score[idx] = torch.as_tensor(s, dtype=torch.int64)
        pos[idx] = torch.as_tensor(p, dtype=torch.int64)
        left_of[idx] = left
    # strictly greatest score wins, ties go to the lowest rank (= lowest piece index, plocalaligner.cpp:122-129)
    packed = (score * world + (world - 1 - rank)).to(device)
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
    packed = packed.cpu()
    winner = (world - 1) - (packed % world)
    mine = winner == rank
    if realign is not None and bool(mine.any()):
        for m, idx in sorted(by_len.items()):
            sel = [i for i in idx if bool(mine[i])]
            if not sel:
                continue
            left, right = pieces[m]
            s, p = realign([reads[i] for i in sel], y[left:right])
            score[sel] = torch.as_tensor(s, dtype=torch.int64)
            pos[sel] = torch.as_tensor(p, dtype=torch.int64)
    out = torch.stack([torch.where(mine, score, torch.zeros_like(score)),
                       torch.where(mine, pos + left_of, torch.zeros_like(pos))]).to(device)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)        # exactly one rank contributes per read
    out = out.cpu()
    return out[0], out[1], winner


def engine_aligner(engine, **align_kw):
    """align_piece / realign callback for reference_sharded_align backed by an Engine (the CUDA path): the piece
    becomes the engine's reference, the reads are aligned as one batch.  The last batch's full result (consensus,
    arg-max cell, ...) stays available as `fn.last`."""
    def fn(reads, y_piece):
        engine.set_reference(y_piece)
        fn.last = engine.align(reads, **align_kw)
        return fn.last["score"], fn.last["pos"]
    fn.last = None
    return fn
