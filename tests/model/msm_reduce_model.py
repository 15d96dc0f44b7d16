"""Integer model of the index logic of msm.cu stages 4-6 (chunked accumulation with collapsed CTAs, the per-bucket slot
lists and the bit-marginal butterfly): every "point" is a Python int and point addition is integer addition, so the
expected result is simply sum_b (b + 1) * S_b.  Host-only; mirrors the kernels line by line (no product code is run)."""
import random

ACC_THREADS, BR_THREADS, BR_CAP, MAXV = 128, 128, 8, 16


def bucket_upper(off, lo, hi, e):
    while lo < hi:
        mid = (lo + hi) >> 1
        if off[mid] > e:
            hi = mid
        else:
            lo = mid + 1
    return lo


def accumulate(entries, off, G, CL):
    """k_accumulate: returns {slot: value}."""
    E = off[G]
    partials = {}
    nblocks = (E + ACC_THREADS * CL - 1) // (ACC_THREADS * CL)

    def put(slot, v):
        assert slot not in partials
        partials[slot] = v

    for blk in range(nblocks + 1):           # one block beyond the end: must exit
        blk_e0 = blk * ACC_THREADS * CL
        if blk_e0 >= E:
            continue
        blk_e1 = min(blk_e0 + ACC_THREADS * CL, E)
        lanes = []                            # per thread: (active, acc, b_final, b_start)
        for tid in range(ACC_THREADS):
            t = blk * ACC_THREADS + tid
            e0 = t * CL
            if e0 >= E:
                lanes.append((False, 0, 0, 0))
                continue
            e1 = min(e0 + CL, E)
            j = bucket_upper(off, 0, G, e0)
            b, nxt = j - 1, off[j]
            b_start, acc = b, 0
            for e in range(e0, e1):
                if e == nxt:
                    put(t + b, acc)
                    acc = 0
                    b += 1
                    nxt = off[b + 1]
                    if nxt == e:
                        j = bucket_upper(off, b + 1, G, e)
                        b, nxt = j - 1, off[j]
                acc += entries[e]
            lanes.append((True, acc, b, b_start))
        b0 = lanes[0][3]
        uniform = blk_e1 <= off[b0 + 1]
        warp_sums = []
        for w in range(ACC_THREADS // 32):
            t0 = blk * ACC_THREADS + 32 * w
            w_e0 = t0 * CL
            w_e1 = min(w_e0 + 32 * CL, E)
            bw0 = lanes[32 * w][3]
            w_uniform = w_e0 < E and w_e1 <= off[bw0 + 1]
            if not w_uniform:
                assert not (uniform and w_e0 < E)
                for k in range(32):
                    act, acc, b, _ = lanes[32 * w + k]
                    if act:
                        put(t0 + k + b, acc)
                continue
            tot = sum(lanes[32 * w + k][1] for k in range(32))
            if not uniform:
                put(t0 + bw0, tot)
            else:
                warp_sums.append(tot)
        if uniform:
            nw = (blk_e1 - blk_e0 + 32 * CL - 1) // (32 * CL)
            assert nw == len(warp_sums)
            put(blk * ACC_THREADS + b0, sum(warp_sums))
    return partials


def whole_units(lo, hi, E, span):
    return (lo + span - 1) // span, ((E + span - 1) // span if hi == E else hi // span)


def slot_list(gb, lo, hi, E, CL):
    if hi <= lo:
        return []
    t_lo, t_hi = lo // CL, (hi - 1) // CL
    w_lo, w_hi = whole_units(lo, hi, E, 32 * CL)
    if w_hi <= w_lo:
        return [t + gb for t in range(t_lo, t_hi + 1)]
    out = [t + gb for t in range(t_lo, 32 * w_lo)]
    k_lo, k_hi = whole_units(lo, hi, E, ACC_THREADS * CL)
    wpb = ACC_THREADS // 32
    if k_hi <= k_lo:
        out += [32 * w + gb for w in range(w_lo, w_hi)]
    else:
        out += [32 * w + gb for w in range(w_lo, wpb * k_lo)]
        out += [ACC_THREADS * k + gb for k in range(k_lo, k_hi)]
        out += [32 * w + gb for w in range(wpb * k_hi, w_hi)]
    out += [t + gb for t in range(32 * w_hi, t_hi + 1)]
    return out


def bucket_reduce(partials, off, G, nb, CL):
    """k_bucket_reduce + k_reduce_final for every set; returns the list of results."""
    E = off[G]
    bpb = min(nb, BR_THREADS)
    lv = bpb.bit_length() - 1
    nblk = nb // bpb
    used = set()
    blockres = []
    for blk in range(G // bpb):
        val = [0] * max(bpb, 32)
        for tid in range(bpb):
            gb = blk * bpb + tid
            sl = slot_list(gb, off[gb], off[gb + 1], E, CL)
            for s in sl:
                assert s not in used
                used.add(s)
            val[tid] = sum(partials[s] for s in sl)
        for k in range(lv):
            h = 1 << k
            new = list(val)
            for tid in range(len(val)):
                p = tid & (2 * h - 1)
                if p <= k:
                    new[tid] = val[tid] + val[tid + h]
                elif p == k + 1 and k >= 2:
                    new[tid] = val[tid - p + h]
            val = new
        blockres.append(val[:lv + 1] + [None] * (MAXV - lv - 1))
    assert used == set(partials), "slot lists do not cover the partials exactly"
    results = []
    for s in range(G // nb):
        cur = blockres[s * nblk:(s + 1) * nblk]
        L = lv
        while len(cur) > 1:
            nxt = []
            for q in range(len(cur) // 2):
                A, B = cur[2 * q], cur[2 * q + 1]
                nxt.append([A[v] + B[v] for v in range(L + 1)] + [B[0]] + [None] * (MAXV - L - 2))
            cur, L = nxt, L + 1
        vals = cur[0]
        total = 0
        for i in range(L + 1):
            total += vals[i] << max(i - 1, 0) if i else vals[0]
        results.append(total)
    return results


def run_case(rnd, nsets, c, npoints_entries, CL, dist):
    nb = 1 << (c - 1)
    G = nsets * nb
    buckets = []
    for _ in range(npoints_entries):
        if dist == "uniform":
            buckets.append(rnd.randrange(G))
        elif dist == "heavy":
            buckets.append(rnd.choice([5 % G, 5 % G, 5 % G, G - 1, rnd.randrange(G)]))
        elif dist == "one":
            buckets.append(3 % G)
        else:
            buckets.append(rnd.randrange(min(G, 40)))
    buckets.sort()
    entries = [rnd.randrange(1, 1 << 30) for _ in buckets]
    off = [0] * (G + 1)
    for b in buckets:
        off[b + 1] += 1
    for i in range(G):
        off[i + 1] += off[i]
    want = [0] * nsets
    for b, v in zip(buckets, entries):
        want[b // nb] += (b % nb + 1) * v
    partials = accumulate(entries, off, G, CL)
    got = bucket_reduce(partials, off, G, nb, CL)
    assert got == want, (nsets, c, npoints_entries, CL, dist)


def main(cases=200, seed=1):
    rnd = random.Random(seed)
    for i in range(cases):
        nsets = rnd.choice([1, 2, 3])
        c = rnd.choice([4, 5, 6, 8, 9, 10, 12])
        n = rnd.choice([0, 1, 5, 100, 1000, 5000, 20000])
        CL = rnd.choice([1, 2, 3, 8, 13, 32, 56])
        dist = rnd.choice(["uniform", "heavy", "one", "low"])
        run_case(rnd, nsets, c, n, CL, dist)
    return cases


if __name__ == "__main__":
    print("ok", main())
