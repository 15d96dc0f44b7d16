"""Multi-GPU partitioning of the hot path -- host-side logic only, no collective on the data path.

  * independent proofs (BASELINE config 4): job i -> rank i mod world_size; every rank owns one GPU and returns
    1-1.5 kB proofs, gathered by the host (torch.distributed `all_gather_object` in the tests / bench).
  * one large MSM (config 5, the verifier's mega-MSM): contiguous point ranges; every GPU returns ONE compressed
    partial point and the host adds the <= 8 points (bpg_point_sum).  NCCL is not used: 8 x 32 bytes.
The IPP of a single proof does not shard (lg n' dependent rounds) -- replicas only.
"""
import ctypes

from ._capi import check, lib


def shard_jobs(n_jobs, world_size, rank):
    """Indices of the jobs rank `rank` owns (round robin, as SURVEY.md config 4 prescribes)."""
    return list(range(rank, n_jobs, world_size))


def point_ranges(n, parts):
    """`parts` contiguous [start, stop) ranges covering [0, n), sizes differing by at most one."""
    base, extra = divmod(n, parts)
    out, start = [], 0
    for r in range(parts):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def point_sum(points):
    """Host-side sum of compressed ristretto points (the combine step of a point-range sharded MSM)."""
    out = ctypes.create_string_buffer(32)
    check(lib().bpg_point_sum(b"".join(points), len(points), out))
    return out.raw


def msm_gens_partial(ctx, sG, sH, rank, world_size, sB=None, sBb=None):
    """This rank's share of  sum sG_i G_i + sum sH_i H_i (+ sB B + sBb B_blinding, added by rank 0 only).
    sG / sH: packed 32-byte scalars for the WHOLE MSM (every rank holds them; only its slice is uploaded)."""
    nG, nH = len(sG) // 32, len(sH) // 32
    g0, g1 = point_ranges(nG, world_size)[rank]
    h0, h1 = point_ranges(nH, world_size)[rank]
    out = ctypes.create_string_buffer(32)
    check(lib().bpg_msm_gens_range(ctx._h, sG[32 * g0: 32 * g1] or None, g0, g1 - g0, sH[32 * h0: 32 * h1] or None, h0,
                                   h1 - h0, sB if rank == 0 else None, sBb if rank == 0 else None, out))
    return out.raw


def msm_gens_sharded(ctxs, sG, sH, sB=None, sBb=None):
    """Single-process form over several contexts (one per GPU, or several on one GPU): returns the compressed MSM."""
    ws = len(ctxs)
    return point_sum([msm_gens_partial(c, sG, sH, r, ws, sB, sBb) for r, c in enumerate(ctxs)])
