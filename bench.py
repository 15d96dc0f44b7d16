#!/usr/bin/env python3
"""bench.py -- BASELINE.json headline: R1CS prove & verify per second on B200 (config 2:
`BOUND` 64-bit range statements x1024 in ONE R1CS proof, n = 2^17 multipliers, m = 3072 commitments).

One "step" = one pass of the hot path over one BATCH of PROOFS_PER_STEP = 32 independent statements (four rounds of the
eight-wide lockstep in which the library runs the provers' Merlin rng streams); each statement costs m Pedersen
commitments + Prover::prove + Verifier::verify (accepting).  value = steps x 32 / time, in proofs per second.  With
the default 8 steps the timed region holds 256 statements, so the drain of the last <= 48 in flight weighs little
(measured: 64 statements in the timed region read 71/s where 128 read 94/s).  Legs:
  value  constraint system and witness already resident in HBM (bpg_circuit); proofs verified inside
         the timed region.  `inflight` host threads (one bpg context = one stream each, generator
         tables shared) keep several independent steps in flight on the GPU, because the prover's
         Merlin TranscriptRng stream (2n dependent Keccak-f permutations) is a sequential host job.
  e2e    the same steps through the C ABI with HOST buffers in pinned memory (bpg_prover_load_cs /
         bpg_verifier_load_cs): every host->device copy of witness / constraints and the
         device->host proof are inside the timer.
  latency  one step at a time on one context (no overlap), resident circuit.
  cpu_baseline  the CPU restatement of dalek's algorithms (oracle/c, 1 core) on the same statement.
`--impl reference` times that CPU restatement alone with all host cores (the real reference is pure
Rust and cannot be built in this image: no cargo/rustc, crates not vendored).

N > 1 (torchrun): every rank proves/verifies its own independent statements ("weak" scaling, no
data-path collective); value = all ranks' statements / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "r1cs_prove_verify_per_sec"
UNIT = "proof+verify/s"
WORKLOAD = "bounds_check 64-bit x1024 in one R1CS proof (n=2^17 multipliers, m=3072, q=265216)"
# measured on this pool's B200 by tools/imad_peak.cu (profiles/r01_imad_peak.jsonl): sustained
# IMAD.WIDE.U32 issue rate, the instruction the field multiplication is built from
IMAD_WIDE_PEAK_TOPS = 8.157
IMAD_PER_MADD = 504  # 7 field muls x (64 + 8) 32x32->64 multiply-adds, SURVEY.md 8(d)
PROOFS_PER_STEP = 32  # statements per step (one batch); every statement is proven and verified
try:
    HBM_PEAK_GBS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    HBM_PEAK_GBS = 6537.6   # this pool's measured copy bandwidth (MEASURED_PEAKS.json of round 1)
SORT_TRAFFIC_BYTES = 13.76e6   # dram read+write per launch of k_digits<1> (IPP-round MSM), profiles/r01_digits_ncu_details.csv


def _dist():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    return ws, int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.samples, self.reasons, self.stop_flag, self.max_mhz = gpu_index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _thread_cpu_by_name():
    """CPU seconds of the threads still alive at the end of the run, grouped by thread name (driver / runtime threads;
    the per-statement host threads have exited and are accounted by cpu_s_per_proof_worker_threads)."""
    out, tick = {}, os.sysconf("SC_CLK_TCK")
    try:
        for tid in os.listdir("/proc/self/task"):
            with open("/proc/self/task/%s/stat" % tid) as f:
                raw = f.read()
            name = raw[raw.index("(") + 1: raw.rindex(")")]
            rest = raw[raw.rindex(")") + 2:].split()
            out[name] = out.get(name, 0.0) + (int(rest[11]) + int(rest[12])) / tick
    except OSError:
        pass
    return {k: round(v, 3) for k, v in sorted(out.items(), key=lambda kv: -kv[1])[:8]}


def run_reference(args, ws, rank):
    """CPU restatement (oracle/c) on all host cores; bounded sample: BOUND x128 per worker per step."""
    if rank != 0:
        return
    import multiprocessing as mp
    from bulletproof_gadgets_b200 import workloads as W
    from oracle import coracle
    coracle.lib()
    sample_count, scale = 128, 8.0  # 1024 / 128; cost is linear in n (Straus / folds / Pippenger per point)
    cores = os.cpu_count() or 1
    st = W.bounds_check_statement(sample_count)

    def worker(q_in, q_out):
        from oracle import coracle as co
        while True:
            job = q_in.get()
            if job is None:
                return
            proof, coms = co.prove_flat(st, b"\x07" * 32, cache_gens=False)  # the reference rebuilds generators per run
            ok = co.verify_flat(st, coms, proof, b"\x09" * 32, cache_gens=False)
            q_out.put(ok is True)

    q_in, q_out = mp.Queue(), mp.Queue()
    procs = [mp.Process(target=worker, args=(q_in, q_out)) for _ in range(cores)]
    for p in procs:
        p.start()

    def step():
        for _ in range(cores):
            q_in.put(1)
        assert all(q_out.get() for _ in range(cores))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    for _ in procs:
        q_in.put(None)
    for p in procs:
        p.join()
    value = cores * args.steps / (dt * scale)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "backend": "CPU restatement of dalek's u64 serial algorithms (oracle/c), "
                       "not the Rust binary (no cargo/rustc in the image)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "BOUND x128 (n=2^14) prove+verify per core per step, generators rebuilt per proof "
                                       "as the reference does; scaled x1/8 to the x1024 statement (cost linear in n)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8, help="batches of %d statements" % PROOFS_PER_STEP)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--count", type=int, default=1024, help="BOUND statements per proof (1024 = BASELINE config 2)")
    ap.add_argument("--inflight", type=int, default=0, help="steps in flight per GPU (host threads); 0 = auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    ws, rank, local_rank = _dist()
    if args.impl == "reference":
        return run_reference(args, ws, rank)

    import torch
    import torch.distributed as dist
    if ws > 1:
        torch.cuda.set_device(local_rank)
        # The data path has no collective (independent proofs per rank): torch.distributed only provides the barrier
        # and one max-reduce of the timing.  Measured on the 8-GPU box (32 vCPUs): with an NCCL process group alive
        # every rank burns ~23 ms of host CPU per step in NCCL's service threads (479 vs 559 prove+verify/s,
        # profiles/r01_bench_8gpu_{nccl,gloo}.json), and this workload is host-CPU bound there -- so gloo by default.
        backend = os.environ.get("BPG_DIST_BACKEND", "gloo")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    import bulletproof_gadgets_b200 as bpg
    from bulletproof_gadgets_b200 import build, workloads as W
    build.build_lib()
    warmup = max(args.warmup, 3)
    # Host threads mostly block (GPU waits, rng batcher).  Measured on one B200 with 16 vCPUs: 24 / 32 / 48 / 64 in flight
    # = 72 / 79 / 92 / 92 prove+verify/s (profiles/r01_summary.md); the 8-GPU box has 4 vCPUs per rank and was measured
    # at 32, so the larger default is used only where a rank has the cores for it.
    inflight = args.inflight or (48 if (os.cpu_count() or 1) // ws >= 12 else 32)
    ctx0 = bpg.Context(local_rank)  # raises loudly without an sm_100a device: there is no CPU path
    ctxs = [ctx0] + [ctx0.shared() for _ in range(inflight - 1)]
    st = W.bounds_check_statement(args.count, seed=20261018 + rank, label=b"bench-bound-%d" % rank).pin(bpg)
    ctx0.gens_ensure(st.n)
    circuit = bpg.Circuit(ctx0, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)
    dev = torch.device("cuda", local_rank)
    streams = [torch.cuda.ExternalStream(c.get("stream"), device=dev) for c in ctxs]

    def step_resident(ctx, i):
        seed = (i + 1).to_bytes(32, "little")
        p = bpg.Prover(ctx, bpg.Transcript(st.label))
        coms = p.commit_batch_packed(st.v_bytes, st.vbl_bytes)
        p.attach(circuit)
        proof = p.prove(seed)
        vf = bpg.Verifier(ctx, bpg.Transcript(st.label))
        vf.commit_batch(coms)
        vf.attach(circuit)
        if not vf.verify(proof, seed):
            raise SystemExit("GPU proof did not verify")
        return proof

    def step_e2e(ctx, i):
        seed = (i + 1).to_bytes(32, "little")
        proof, coms = W.prove_statement(bpg, ctx, st, seed)
        if not W.verify_statement(bpg, ctx, st, proof, coms, seed):
            raise SystemExit("GPU proof did not verify")
        return proof

    gad_txt, inst_txt, wtns_txt = W.bounds_check_text(args.count, seed=20261018 + rank)
    name_txt = "bench-bound-%d" % rank

    def step_statement(ctx, i):
        """Text in, proof out: the c_prove / c_verify mirror (bpg_prove / bpg_verify), front end included on both sides."""
        seed = (i + 1).to_bytes(32, "little")
        proof, coms_txt, _ = bpg.prove(ctx, name_txt, inst_txt, wtns_txt, gad_txt, seed, seed)
        if not bpg.verify(ctx, name_txt, inst_txt, proof, coms_txt, gad_txt, seed):
            raise SystemExit("GPU proof did not verify")
        return proof

    def barrier():
        for s in streams:
            s.synchronize()
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()

    worker_cpu = [0.0]

    def run_steps(fn, first, count, use_ctxs):
        """`count` steps spread over the contexts: each host thread pulls the next step index."""
        if len(use_ctxs) == 1:
            for i in range(first, first + count):
                fn(use_ctxs[0], i)
            return
        lock, nxt, errs = threading.Lock(), [first], []

        def work(c):
            try:
                while True:
                    with lock:
                        i = nxt[0]
                        nxt[0] += 1
                    if i >= first + count:
                        return
                    fn(c, i)
            except BaseException as e:  # noqa: BLE001
                errs.append(e)
            finally:
                with lock:
                    worker_cpu[0] += time.thread_time()   # CPU of this host thread (python + C ABI inside it)

        ts = [threading.Thread(target=work, args=(c,)) for c in use_ctxs]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]

    def timed(fn, steps, warm, use_ctxs):   # `steps` here = number of statements
        run_steps(fn, 0, max(warm * PROOFS_PER_STEP, len(use_ctxs)) if len(use_ctxs) > 1 else warm, use_ctxs)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        run_steps(fn, 1000, steps, use_ctxs)
        for s in streams:
            s.synchronize()
        e1.record(streams[0])
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if ws > 1:
            t = torch.tensor([ms], device="cuda" if dist.get_backend() == "nccl" else "cpu", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(c.get("launches") for c in ctxs)
    stat0 = [bpg.lib().bpg_rng_batcher_stat(k) for k in range(3)]
    cpu0 = time.process_time()
    nproofs = args.steps * PROOFS_PER_STEP
    ms_res = timed(step_resident, nproofs, warmup, ctxs)
    cpu_res = time.process_time() - cpu0
    cpu_res_workers = worker_cpu[0]
    launches = sum(c.get("launches") for c in ctxs) - launches0
    ms_e2e = timed(step_e2e, nproofs, warmup, ctxs)
    ms_stmt = timed(step_statement, nproofs, warmup, ctxs)
    stat1 = [bpg.lib().bpg_rng_batcher_stat(k) for k in range(3)]
    cpu_parts = {k: sum(c.get("cpu_%s_ns" % k) for c in ctxs) * 1e-9 for k in ("sync", "commit", "prove", "verify", "rng")}
    lat_steps = 10
    ms_lat = timed(step_resident, lat_steps, warmup, ctxs[:1])
    sampler.stop_flag = True

    # dominant kernel (MSM bucket accumulation): one instrumented step, CUDA events around every launch
    ctx0.set("time_accum", 1)
    step_resident(ctx0, 10 ** 6)
    acc_ns, acc_entries = ctx0.get("sum_accum_ns"), ctx0.get("sum_entries")
    sct_ns, sct_points = ctx0.get("sum_scatter_ns"), ctx0.get("sum_points")
    ctx0.set("time_accum", 0)

    # second half of the BASELINE metric: raw fixed-base MSM (verifier mega-MSM shape), 2n' = 2^18 points, uniform
    # scalars resident in HBM; whole MSM (all stages + 128-byte read-back + host ristretto compression)
    import numpy as np
    npts = 2 * st.n
    sc_np = np.random.default_rng(7).integers(0, 256, size=(npts, 32), dtype=np.uint8)
    sc_np[:, 31] &= 0x0F
    d_sc = torch.from_numpy(sc_np).to(dev)
    for _ in range(3):
        ctx0.msm_gens_dev(d_sc.data_ptr(), st.n, d_sc.data_ptr() + 32 * st.n, st.n)
    streams[0].synchronize()
    msm_runs = []
    for _ in range(5):      # wall clock around synchronous calls: median of five batches of ten
        t0 = time.perf_counter()
        for _ in range(10):
            ctx0.msm_gens_dev(d_sc.data_ptr(), st.n, d_sc.data_ptr() + 32 * st.n, st.n)
        msm_runs.append((time.perf_counter() - t0) / 10)
    msm_s = sorted(msm_runs)[len(msm_runs) // 2]

    dist_backend = dist.get_backend() if ws > 1 else None
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    per_step_ms = ms_res / args.steps
    value = ws * nproofs / (ms_res * 1e-3)
    e2e_value = ws * nproofs / (ms_e2e * 1e-3)
    achieved = acc_entries * IMAD_PER_MADD / (acc_ns * 1e-9) / 1e12 if acc_ns else None
    lg = max(st.n - 1, 0).bit_length()
    # bytes crossing PCIe per e2e step, counted from the buffers handed to the C ABI plus the library's own uploads
    csr = 4 * (st.q + 1) + 36 * st.nnz
    h2d = 2 * 32 * st.n + csr            # bpg_prover_load_cs: a_L, a_R, CSR constraints
    h2d += csr                           # bpg_verifier_load_cs
    h2d += 64 * st.m + 32 * st.m         # commit_batch (v, blinding) ; v_blinding vector of the prover
    h2d += 128 * st.n                    # 64-byte TranscriptRng draws for s_L, s_R (reduced mod l on the device)
    h2d += 32 * (6 + st.m + 5 + 2 * lg) * 2  # verifier: compressed points + head scalars
    d2h = 32 * st.m + 128 * (3 + 2 * lg) + 9 * 32 + 64 + 128 + 16  # commitments, A/S/L/R points, t scalars, a/b, check point, flags
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": warmup,
        "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if args.count == 1024 else "bounds_check 64-bit x%d" % args.count,
                   "proofs_per_step": PROOFS_PER_STEP, "inflight": inflight, "dist_backend": (dist_backend if ws > 1 else None),
                   "l2": "per-step working set (fixed-base tables 403 MB + entries) exceeds the 126 MB L2",
                   "rng": "transcript rng seeded per step; proofs byte-identical to the CPU oracle",
                   "window_bits": ctx0.get("window_bits"), "task_len": 32},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d * PROOFS_PER_STEP, "d2h_bytes_per_step": d2h * PROOFS_PER_STEP},
        "e2e_statement": {"value": ws * nproofs / (ms_stmt * 1e-3), "unit": UNIT, "ms_per_step": ms_stmt / args.steps,
                          "note": "text formats in, proof + commitments text out and back in: bpg_prove / bpg_verify (the "
                                  "c_prove / c_verify mirror), host front end (parse, gadgets, witness assignment, "
                                  "flattening) inside the timed region on both sides"},
        "latency": {"ms_per_proof": ms_lat / lat_steps, "proofs": lat_steps,
                    "note": "one proof+verify at a time on one context; the prover waits on the host for the sequential "
                            "Merlin TranscriptRng stream (2n Keccak-f permutations)"},
        "msm": {"points": npts, "mpoints_per_s": npts / msm_s / 1e6, "ms": msm_s * 1e3, "ms_batches": [round(x * 1e3, 4) for x in msm_runs], "scalars": "uniform mod l",
                "frac_of_imad_peak_whole_msm": 16 * npts * IMAD_PER_MADD / msm_s / (IMAD_WIDE_PEAK_TOPS * 1e12),
                "note": "one GPU; sweep 2^10..2^22 in profiles/r01_configs_1gpu.jsonl (tools/bench_configs.py)"},
        "gpu_launches": launches,
        "host": {"cores": os.cpu_count(), "cpu_s_per_proof_rank0": cpu_res / (nproofs + max(warmup * PROOFS_PER_STEP, inflight)),
                 "cpu_s_per_proof_worker_threads": cpu_res_workers / (nproofs + max(warmup * PROOFS_PER_STEP, inflight)),
                 "live_threads_cpu_s": _thread_cpu_by_name(),
                 "rng_streams": stat1[0] - stat0[0], "rng_vector_batches": stat1[1] - stat0[1],
                 "rng_streams_alone": stat1[2] - stat0[2],
                 "thread_cpu_s_total_all_legs": cpu_parts,
                 "note": "process CPU time of rank 0 over the resident leg (incl. its warm-up steps); the Merlin rng "
                         "streams of proofs in flight run eight at a time (AVX-512) when batches form"},
        "roofline": {"kernel": "k_accumulate (MSM bucket accumulation, mixed Edwards adds)", "bound": "imad",
                     "achieved": achieved, "peak": IMAD_WIDE_PEAK_TOPS, "unit": "T IMAD.WIDE/s",
                     "frac": achieved / IMAD_WIDE_PEAK_TOPS if achieved else None, "traffic": 719.5e6,
                     "traffic_note": "dram bytes read+write per launch from profiles/r01_accumulate_ncu_details_v2.csv "
                                     "(IPP-round MSM, 4.19 M entries x 96 B = 403 MB algorithmic gather)",
                     "peak_source": "tools/imad_peak.cu on this pool (profiles/r01_imad_peak.jsonl); IMAD.WIDE.U32 issues at "
                                    "28/clk/SM vs 64 for 32-bit IMAD; not in MEASURED_PEAKS.json",
                     "work": "%d mixed adds x %d IMAD.WIDE" % (acc_entries, IMAD_PER_MADD),
                     "kernel_ms_per_proof": acc_ns * 1e-6, "share_of_step": acc_ns * 1e-6 / (ms_lat / lat_steps),
                     "share_note": "share of one un-overlapped proof+verify (the latency leg)"},
        # the sort stage north_star asks to see against HBM: digit decomposition + scatter by bucket.  Algorithmic bytes
        # per launch: 32 B per scalar read + 4 B per entry written + 4 B per entry of offset reads.
        "roofline_sort": {"kernel": "k_digits<1> (signed-digit decomposition + counting-sort scatter)", "bound": "hbm",
                          "achieved": (32 * sct_points + 8 * acc_entries) / (sct_ns * 1e-9) / 1e9 if sct_ns else None,
                          "peak": HBM_PEAK_GBS, "unit": "GB/s",
                          "frac": (32 * sct_points + 8 * acc_entries) / (sct_ns * 1e-9) / 1e9 / HBM_PEAK_GBS if sct_ns else None,
                          "traffic": SORT_TRAFFIC_BYTES, "kernel_ms_per_proof": sct_ns * 1e-6,
                          "note": "the scatter pass places entries at bucket offset + the rank the histogram pass's atomicAdd "
                                  "returned (no atomics of its own); bound by 4-byte scattered stores through L2, not by HBM bytes. "
                                  "`traffic` is the ncu capture of the earlier atomic version (profiles/r01_digits_ncu_details.csv); "
                                  "the rank array adds 4 B per entry each way, L2-resident"},
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        from oracle import coracle
        cst = W.bounds_check_statement(args.count, seed=20261018 + rank, label=b"bench-bound-%d" % rank)
        t0 = time.perf_counter()
        p_c, coms_c = coracle.prove_flat(cst, (1).to_bytes(32, "little"), cache_gens=False)
        ok = coracle.verify_flat(cst, coms_c, p_c, (1).to_bytes(32, "little"), cache_gens=False)
        dt = time.perf_counter() - t0
        same = p_c == step_resident(ctx0, 0)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "the full statement once (prove+verify, generators rebuilt per run as the reference "
                                          "does): %.1f s; CPU proof accepted=%s, byte-identical to the GPU proof=%s" % (dt, ok, same)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
