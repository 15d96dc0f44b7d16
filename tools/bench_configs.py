#!/usr/bin/env python3
"""Measures the BASELINE.json configs other than the headline one (bench.py = config 2) on ONE B200:
  1  example.{gadgets,inst,wtns}                      (tests/golden/fixtures copy)
  3  Merkle membership with MiMC, depth 32            (instance siblings: n'=2^16; witness siblings: n'=2^17)
  4  batch of independent LESS_THAN / SET_MEMBER proofs (default 4096)
  5  raw fixed-base MSM sweep 2^10 .. 2^22 points     (uniform scalars = verifier shape, 0/1 scalars = a_L/a_R shape)
One JSON line per measurement on stdout.  Product code only (statements are flattened by the library's own front end).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bulletproof_gadgets_b200 as bpg  # noqa: E402
from bulletproof_gadgets_b200 import workloads as W  # noqa: E402

IMAD_WIDE_PEAK = 8.157e12
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")


def emit(**kw):
    print(json.dumps(kw), flush=True)


def run_threads(fn, ctxs, jobs):
    lock, nxt, errs = threading.Lock(), [0], []

    def work(c):
        try:
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= jobs:
                    return
                fn(c, i)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(c,)) for c in ctxs]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return time.perf_counter() - t0


def r1cs_config(name, ctxs, text, steps):
    gad, inst, wtns = text
    ctx = ctxs[0]
    t0 = time.perf_counter()
    st = bpg.flatten_prover(name, inst, wtns, gad, b"\x11" * 32)
    t_front = time.perf_counter() - t0
    ctx.gens_ensure(max(st.n, 1))
    circ = bpg.Circuit(ctx, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)

    def step(c, i, times=None):
        seed = (i + 1).to_bytes(32, "little")
        ta = time.perf_counter()
        p = bpg.Prover(c, bpg.Transcript(st.label))
        coms = p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
        p.attach(circ)
        proof = p.prove(seed)
        tb = time.perf_counter()
        vf = bpg.Verifier(c, bpg.Transcript(st.label))
        vf.commit_batch(coms)
        vf.attach(circ)
        assert vf.verify(proof, seed)
        if times is not None:
            times.append((tb - ta, time.perf_counter() - tb))
        return proof

    for i in range(3):
        step(ctx, i)
    times = []
    for i in range(5):
        step(ctx, i, times)
    prove_ms = 1e3 * min(t[0] for t in times)
    verify_ms = 1e3 * min(t[1] for t in times)
    run_threads(step, ctxs, len(ctxs))
    dt = run_threads(step, ctxs, steps)
    npad = 1
    while npad < st.n:
        npad *= 2
    emit(config=name, n=st.n, n_pad=npad, m=st.m, q=st.q, nnz=int(st.row_start[-1]), proof_bytes=len(step(ctx, 0)),
         frontend_ms=1e3 * t_front, prove_latency_ms=prove_ms, verify_latency_ms=verify_ms,
         prove_verify_per_s=steps / dt, inflight=len(ctxs), steps=steps)


def batch_config(ctxs, count):
    texts = W.batch_texts(count)
    t0 = time.perf_counter()
    sts = [bpg.flatten_prover("batch-%d" % i, inst, wtns, gad, bytes([i % 256]) * 32) for i, (gad, inst, wtns) in enumerate(texts)]
    t_front = time.perf_counter() - t0
    ctxs[0].gens_ensure(512)
    ok = [0]

    def job(c, i):
        st = sts[i]
        proof, coms = W.prove_statement(bpg, c, st, (i + 1).to_bytes(32, "little"))
        if W.verify_statement(bpg, c, st, proof, coms):
            ok[0] += 1

    run_threads(job, ctxs, 4 * len(ctxs))
    ok[0] = 0
    dt = run_threads(job, ctxs, count)
    assert ok[0] == count
    emit(config="batch of %d independent LESS_THAN (n=379) / SET_MEMBER k=16 (n=32) proofs" % count, proofs=count,
         prove_verify_per_s=count / dt, frontend_ms_per_statement=1e3 * t_front / count, inflight=len(ctxs),
         note="host-buffer path per proof (commit + load_cs + prove + load_cs + verify); launch/latency bound at these sizes")


def msm_sweep(ctx, lgs):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.get("stream"), device=dev)
    L = 2**252 + 27742317777372353535851937790883648493
    for lg in lgs:
        N = 1 << lg
        half = N // 2
        ctx.gens_ensure(half)
        rng = np.random.default_rng(lg)
        for shape in ("uniform", "bits"):
            if shape == "uniform":
                a = rng.integers(0, 256, size=(N, 32), dtype=np.uint8)
                a[:, 31] &= 0x0F      # < 2^252 < l: canonical
            else:
                a = np.zeros((N, 32), dtype=np.uint8)
                a[:, 0] = rng.integers(0, 2, size=N, dtype=np.uint8)
            d = torch.from_numpy(a).to(dev)
            pG, pH = d.data_ptr(), d.data_ptr() + 32 * half
            for _ in range(3):
                out = ctx.msm_gens_dev(pG, half, pH, half)
            reps = 10 if lg <= 18 else 4
            stream.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                out2 = ctx.msm_gens_dev(pG, half, pH, half)
            dt = (time.perf_counter() - t0) / reps
            assert out2 == out
            ctx.set("time_accum", 1)
            ctx.msm_gens_dev(pG, half, pH, half)
            acc_us, ent = ctx.get("accum_us"), ctx.get("accum_entries")
            ctx.set("time_accum", 0)
            emit(config="msm", lg_points=lg, scalars=shape, ms=1e3 * dt, mpoints_per_s=N / dt / 1e6, entries=ent,
                 accumulate_us=acc_us, accumulate_frac_of_imad_peak=(ent * 504 / (acc_us * 1e-6) / IMAD_WIDE_PEAK) if acc_us else None,
                 whole_msm_frac_of_imad_peak=ent * 504 / dt / IMAD_WIDE_PEAK,
                 note="includes the 128-byte result read-back and host ristretto compression")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--inflight", type=int, default=16)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--max-lg", type=int, default=22)
    ap.add_argument("--only", default="1,3,4,5")
    args = ap.parse_args()
    only = set(args.only.split(","))
    ctx0 = bpg.Context(0)
    ctxs = [ctx0] + [ctx0.shared() for _ in range(args.inflight - 1)]
    if "1" in only:
        text = tuple(open(os.path.join(FIX, "example" + e)).read() for e in (".gadgets", ".inst", ".wtns"))
        r1cs_config("example", ctxs, text, args.steps * 4)
    if "3" in only:
        r1cs_config("merkle depth 32 (instance siblings)", ctxs, W.merkle_text(32), args.steps * 2)
        r1cs_config("merkle depth 32 (witness siblings)", ctxs, W.merkle_text(32, witness_siblings=True), args.steps)
    if "4" in only:
        batch_config(ctxs, args.batch)
    if "5" in only:
        msm_sweep(ctx0, range(10, args.max_lg + 1))


if __name__ == "__main__":
    main()
