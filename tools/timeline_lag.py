"""Reads a TIMELINE_DUMP of tools/gpu_timeline.py: submission lag (GPU start - max(host launch, previous op of the stream
done)) per kernel, and what is ready-but-waiting while no k_accumulate runs."""
import gzip, collections, statistics as st, random, sys, json
G = []; H = {}
for l in gzip.open(sys.argv[1], 'rt'):
    t, s, e, a, cid, nm = l.split(None, 5); nm = nm.strip()
    if t == 'G': G.append((float(s), float(e), int(a), int(cid), nm))
    else: H[int(cid)] = (float(s), float(e), int(a), nm)
bys = collections.defaultdict(list)
for x in G: bys[x[2]].append(x)
recs = []
for sid, L in bys.items():
    L.sort()
    for i, x in enumerate(L):
        h = H.get(x[3])
        if not h: continue
        prev_end = L[i - 1][1] if i else 0
        recs.append((x[4], max(h[0], prev_end), x[0], x[1], sid, h[0] > prev_end))
span = max(r[3] for r in recs); lo, hi = 0.15 * span, 0.85 * span
w = collections.defaultdict(list)
fresh = []; chained = []
for nm, ready, s, e, sid, host_bound in recs:
    if lo <= s <= hi:
        w[nm].append(s - ready)
        (fresh if host_bound else chained).append(s - ready)
out = {"fresh_submission_lag_us": {"n": len(fresh), "median": st.median(fresh), "mean": sum(fresh) / len(fresh), "p90": sorted(fresh)[int(.9 * len(fresh))]},
       "in_stream_successor_lag_us": {"n": len(chained), "median": st.median(chained), "mean": sum(chained) / len(chained), "p90": sorted(chained)[int(.9 * len(chained))]}}
out["by_kernel_mean_us"] = {nm: round(sum(v) / len(v)) for nm, v in sorted(w.items(), key=lambda kv: -sum(kv[1]))[:8]}
acc = [(s, e) for nm, r, s, e, sid, hb in recs if nm.startswith('k_accumulate')]
random.seed(1)
n_no = 0; ra0 = 0; ro = []
for _ in range(3000):
    t = random.uniform(lo, hi)
    if any(s <= t < e for s, e in acc): continue
    n_no += 1
    ra0 += not any(nm.startswith('k_accumulate') and r <= t < s for nm, r, s, e, sid, hb in recs)
    ro.append(sum(1 for nm, r, s, e, sid, hb in recs if r <= t < s))
out["no_acc_frac"] = n_no / 3000; out["no_acc_and_none_ready_frac"] = ra0 / max(n_no, 1); out["ready_waiting_kernels_median"] = st.median(ro) if ro else 0
print(json.dumps(out))
