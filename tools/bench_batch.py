#!/usr/bin/env python3
"""BASELINE config 4 across GPUs: a batch of independent LESS_THAN / SET_MEMBER proofs (default 4096), statement i on
rank i mod world_size (bulletproof_gadgets_b200.sharding.shard_jobs), every rank proves AND verifies its share through
the statement-level C ABI path (front end flattening is done before the timed region; commit + load_cs + prove +
load_cs + verify are timed).  No collective on the data path: the ranks only meet at a gloo barrier; rank 0 prints one
JSON line with the whole-job proofs/s (max over ranks of the elapsed time).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_batch.py
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bulletproof_gadgets_b200 as bpg  # noqa: E402
from bulletproof_gadgets_b200 import sharding, workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=4096)
    ap.add_argument("--inflight", type=int, default=24)
    args = ap.parse_args()
    ws, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if ws > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("gloo")
    ctx0 = bpg.Context(local)
    ctxs = [ctx0] + [ctx0.shared() for _ in range(args.inflight - 1)]
    texts = W.batch_texts(args.count)
    mine = sharding.shard_jobs(args.count, ws, rank)
    sts = {i: bpg.flatten_prover("batch-%d" % i, texts[i][1], texts[i][2], texts[i][0], bytes([i % 256]) * 32) for i in mine}
    ctx0.gens_ensure(512)
    ok = [0]
    lock = threading.Lock()

    def run(jobs):
        nxt = [0]

        def work(c):
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= len(jobs):
                    return
                st = sts[jobs[k]]
                proof, coms = W.prove_statement(bpg, c, st, (jobs[k] + 1).to_bytes(32, "little"))
                if W.verify_statement(bpg, c, st, proof, coms):
                    with lock:
                        ok[0] += 1

        ts = [threading.Thread(target=work, args=(c,)) for c in ctxs]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    run(mine[: 4 * args.inflight])          # warm-up
    ok[0] = 0
    if ws > 1:
        dist.barrier()
    t0 = time.perf_counter()
    run(mine)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert ok[0] == len(mine)
    if ws > 1:
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"config": "batch of %d independent LESS_THAN (n=379) / SET_MEMBER k=16 (n=32) proofs" % args.count,
                          "n_gpus": ws, "proofs": args.count, "prove_verify_per_s": args.count / dt, "seconds": dt,
                          "inflight_per_gpu": args.inflight, "host_cores": os.cpu_count(), "scaling": "strong (fixed batch)"}), flush=True)


if __name__ == "__main__":
    main()
