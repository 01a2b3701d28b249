"""Multi-GPU plumbing for libpre3: one process per GPU, torch.distributed (NCCL on
the GPUs of one NVSwitch box, gloo in the CPU tests) -- no compute lives here.

Two partitionings, as the reference's structure allows:

* independent units (frame pairs of a sequence, descriptor-matching problems, EKF frames): every
  pair is a pure function of its own cached inputs (M/find_consistent_sift_matches.m:22-32,
  M/Calculate_V_Omega_RANSAC_my_version.m:7-20), so rank r takes the contiguous block
  [P r / G, P (r+1) / G) and there is NO data-path collective; `gather_records` returns the
  per-pair result records (240 B each) to every rank when the caller wants them in one place.

* one large pair split by hypothesis block (BASELINE.json config 5): correspondences are
  replicated (one broadcast), rank r evaluates sample sets [H r / G, H (r+1) / G) and the ranks
  agree on the winner with ONE small collective:
    - "first" mode   all_reduce(MAX) of one 64-bit key (count << 32 | 0xFFFFFFFF - global id):
                     max cardinality, lowest hypothesis id (what north_star specifies);
    - "reference" mode all_gather of 16 bytes per rank (count, global id, ErrorSum), then the same
                     local pick on every rank with the full rule of RANSAC_CALC_VER2.m:165-175
                     (max cardinality, then min ErrorSum, then first index).
  The owning rank already holds (R, T, mask) of the winner and broadcasts the 240-byte record.
"""
from __future__ import annotations

import numpy as np

KEY_ID_MASK = 0xFFFFFFFF


def split_range(total: int, rank: int, world: int):
    """Contiguous block of unit indices of `rank`: unit u -> rank floor(u * world / total)'s inverse."""
    return (total * rank) // world, (total * (rank + 1)) // world


def pack_key(count: int, global_id: int) -> int:
    """(count << 32) | (0xFFFFFFFF - id): a MAX-reduce picks max count, then the LOWEST id."""
    return (int(count) << 32) | (KEY_ID_MASK - int(global_id))


def unpack_key(key: int):
    key = int(key)
    return key >> 32, KEY_ID_MASK - (key & KEY_ID_MASK)


def pick_reference(counts, ids, errsums) -> int:
    """Index (rank) of the global winner among the ranks' local winners under the reference's rule
    (M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:165-175): max cardinality, then min
    ErrorSum, then first (lowest) hypothesis id.  Ranks with no recorded hypothesis carry count < 0."""
    counts = np.asarray(counts, np.int64)
    ids = np.asarray(ids, np.int64)
    errsums = np.asarray(errsums, np.float64)
    valid = counts >= 0
    if not valid.any():
        return -1
    cmax = counts[valid].max()
    cand = np.flatnonzero(valid & (counts == cmax))
    emin = errsums[cand].min()
    cand = cand[errsums[cand] == emin]
    return int(cand[np.argmin(ids[cand])])


def _dist():
    import torch.distributed as dist
    return dist


def world(group=None):
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def allreduce_max_key(key, group=None):
    """In-place MAX all-reduce of a (1,) int64 tensor holding a packed key (counts < 2^31, so the
    signed maximum is the unsigned one).  The only collective of the "first" mode."""
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
    return key


def gather_records(local, total: int, group=None):
    """All-gather of per-unit records.  local: uint8 tensor (n_local, rec_bytes) holding the records of
    this rank's split_range block.  Returns (total, rec_bytes) on every rank.  Blocks may differ by
    one unit, so every rank pads to the largest block."""
    import torch
    dist = _dist()
    rank, ws = world(group)
    if ws == 1:
        return local
    rec = local.shape[1]
    biggest = max(split_range(total, r, ws)[1] - split_range(total, r, ws)[0] for r in range(ws))
    pad = torch.zeros(biggest, rec, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(ws):
        lo, hi = split_range(total, r, ws)
        parts.append(out[r][: hi - lo])
    return torch.cat(parts)


def agree_on_winner(count: int, global_id: int, errsum: float, mode: str, device, group=None):
    """The one collective of the hypothesis-block split.  Every rank passes its local winner
    (count < 0: none) and gets back (winner_rank, count, global_id, errsum) -- identical on all ranks.
    mode "first": one 8-byte MAX all-reduce (errsum is not exchanged: returned as None unless this
    rank owns the winner).  mode "reference": one 16-byte-per-rank all-gather + pick_reference."""
    import torch
    dist = _dist()
    rank, ws = world(group)
    if mode == "first":
        key = torch.tensor([pack_key(count, global_id) if count >= 0 else 0], dtype=torch.int64, device=device)
        allreduce_max_key(key, group)
        k = int(key.item())
        if k == 0:
            return -1, -1, -1, None
        c, gid = unpack_key(k)
        return None, c, gid, (errsum if (c == count and gid == global_id) else None)
    if mode != "reference":
        raise ValueError("mode must be 'first' or 'reference'")
    # 16 bytes per rank: int64 key (count, id) + float64 ErrorSum, reinterpreted as two int64
    mine = torch.zeros(2, dtype=torch.int64, device=device)
    mine[0] = pack_key(count, global_id) if count >= 0 else -1
    mine[1] = int(np.float64(errsum if count >= 0 else np.inf).view(np.int64))
    if ws > 1:
        allv = [torch.empty_like(mine) for _ in range(ws)]
        dist.all_gather(allv, mine, group=group)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = mine.cpu().numpy()[None]
    counts, ids = [], []
    for k in allv[:, 0]:
        if k < 0:
            counts.append(-1)
            ids.append(0)
        else:
            c, g = unpack_key(int(k))
            counts.append(c)
            ids.append(g)
    es = allv[:, 1].copy().view(np.float64)
    w = pick_reference(counts, ids, es)
    if w < 0:
        return -1, -1, -1, None
    return w, counts[w], ids[w], float(es[w])


def ransac_hypothesis_split(ctx, Ya, Yb, opts, samples=None, mode="first", group=None, want_mask=True):
    """BASELINE.json config 5: ONE pair (Ya, Yb: (N,3) float64 CUDA tensors, identical on every rank),
    opts.H hypotheses split into contiguous blocks over the ranks, fixed H (no adaptive stop).
    samples: this rank's OWN slice (Hloc, k) int32 CUDA tensor of explicit sample sets, or None
    (seeded: hypothesis id h uses the same set on any rank count).  Returns (record, mask) with
    record a numpy RESULT_DTYPE scalar (best_sample = GLOBAL hypothesis id) and mask (N,) uint8 numpy
    (None unless want_mask) -- identical on every rank.

    STREAM-ORDERED: threshold, keys and winner stay in device memory; the sequence on the context's stream (which
    must be torch's current stream: ctx.use_torch_stream()) is
        split_local kernels -> ONE collective (8-byte MAX all-reduce, or 16-byte-per-rank all-gather)
        -> split_finish kernel -> SUM all-reduce of the 240-byte record (+ mask on request)
    and the only host synchronisation is the final read of the record."""
    import torch
    from .api import RESULT_DTYPE
    dist = _dist()
    rank, ws = world(group)
    dev = Ya.device
    N, H = Ya.shape[0], int(opts.H)
    h0, h1 = split_range(H, rank, ws)
    m = 0 if mode == "first" else 1
    if mode not in ("first", "reference"):
        raise ValueError("mode must be 'first' or 'reference'")
    res = torch.zeros(240, dtype=torch.uint8, device=dev)
    mask = torch.zeros(N, dtype=torch.uint8, device=dev) if (want_mask or m == 1) else None
    key = torch.zeros(2, dtype=torch.int64, device=dev)
    ctx.ransac_split_local_dev(Ya, Yb, opts, h0, h1 - h0, m, key, res if m == 1 else None, mask if m == 1 else None,
                               samples=samples)
    if m == 0:
        exchanged = key[:1]
        if ws > 1:
            dist.all_reduce(exchanged, op=dist.ReduceOp.MAX, group=group)          # <- the one collective
    else:
        if ws > 1:
            exchanged = torch.empty(2 * ws, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(exchanged, key, group=group)               # <- the one collective
        else:
            exchanged = key
    ctx.ransac_split_finish_dev(Ya, Yb, opts, h0, h1 - h0, m, exchanged, ws, rank, res, mask if (want_mask or m == 1) else None,
                                samples=samples)
    if ws > 1:  # result delivery (not part of the selection): every rank but the owner holds zeros
        dist.all_reduce(res, op=dist.ReduceOp.SUM, group=group)
        if want_mask:
            dist.all_reduce(mask, op=dist.ReduceOp.SUM, group=group)
    ctx.sync()  # the only host synchronisation: the record is read
    rec = np.frombuffer(res.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)[0].copy()
    if m == 0 and int(key[0].item()) == 0:
        return None, None
    if m == 1 and rec["best_fit"] == 0 and rec["n_matches"] == 0:
        return None, None
    return rec, (mask.cpu().numpy() if want_mask else None)
