#!/usr/bin/env python3
"""bench.py -- BASELINE.json headline: R1CS prove & verify per second on B200 (config 2:
`BOUND` 64-bit range statements x1024 in ONE R1CS proof, n = 2^17 multipliers, m = 3072 commitments).

One "step" = one pass of the hot path over one BATCH of PROOFS_PER_STEP = 32 independent statements (four rounds of the
eight-wide lockstep in which the library runs the provers' Merlin rng streams); each statement costs m Pedersen
commitments + Prover::prove + Verifier::verify (accepting).  value = steps x 32 / time, in proofs per second.  With
the default 8 steps the timed region holds 256 statements, so the drain of the last <= 48 in flight weighs little.
Every throughput leg is ONE call of the library's batch entry point (bpg_r1cs_prove_batch with BPG_JOB_VERIFY, or
bpg_prove_batch for text): the library owns the `inflight` host threads, one bpg context (= one stream) each.  Legs:
  value  constraint system and witness already resident in HBM (bpg_circuit); proofs verified inside the timed region.
  e2e    the same statements with HOST buffers in pinned memory (bpg_prover_load_cs / bpg_verifier_load_cs inside the
         library): every host->device copy of witness / constraints and the device->host proof are inside the timer.
  e2e_statement  text formats in (bpg_prove_batch): the host front end is inside the timer too.
  latency  one statement at a time on one context (no overlap), resident circuit.
  msm / msm_stages  raw fixed-base MSM of 2^18 points and the CUDA-event time of each of its stages.
  config4  BASELINE config 4: 4096 independent LESS_THAN / SET_MEMBER proofs, statement i on rank i mod N.
  msm_sharded  BASELINE config 5 at N GPUs: one MSM of 2^22 points split by point range, partial points added by rank 0.
  cpu_baseline  the CPU restatement of dalek's algorithms (oracle/c, 1 core) on the same statement.
`--impl reference` times that CPU restatement alone with all host cores (the real reference is pure
Rust and cannot be built in this image: no cargo/rustc, crates not vendored).

N > 1 (torchrun): every rank proves/verifies its own independent statements ("weak" scaling, no
data-path collective); value = all ranks' statements / max-over-ranks time.
"""
import argparse
import json
import os
# 48-96 statements in flight = as many streams: with the default of 8 hardware work queues every stream waits behind the
# unfinished kernel chains of the streams that share its queue (profiles/r02_summary.md, r02_ab5.jsonl).  Read by the CUDA
# driver when the process creates its context, so it is set before anything touches CUDA (the library does the same).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "r1cs_prove_verify_per_sec"
UNIT = "proof+verify/s"
WORKLOAD = "bounds_check 64-bit x1024 in one R1CS proof (n=2^17 multipliers, m=3072, q=265216)"
# fallback only: the peak is measured inside every run (bpg_measure_imad_peak, a 50 ms register-only kernel);
# tools/imad_peak.cu measured 8.157-8.167 T IMAD.WIDE.U32/s on this pool (profiles/r01_imad_peak.jsonl, r02_imad_peak.jsonl)
IMAD_WIDE_PEAK_TOPS = 8.157
IMAD_PER_MADD = 504  # 7 field muls x (64 + 8) 32x32->64 multiply-adds, SURVEY.md 8(d)
PROOFS_PER_STEP = 32  # statements per step (one batch); every statement is proven and verified
try:
    HBM_PEAK_GBS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    HBM_PEAK_GBS = 6537.6   # this pool's measured copy bandwidth (MEASURED_PEAKS.json of round 1)
# dram read+write bytes per launch from the ncu captures of the CURRENT kernels (profiles/r02_msm_kernels_details.csv,
# raw MSM of 2^18 points = the shape of an IPP-round MSM): bucket accumulation and the scatter pass of the sort
ACC_TRAFFIC_BYTES = 780.2e6    # k_accumulate: 759.2 MB read + 21.0 MB written
SORT_TRAFFIC_BYTES = 38.4e6    # k_sort_count 8.4 + k_sort_scan 0.3 + k_sort_scatter 8.7 + k_sort_bins 21.0 MB (the rest stays in L2)
MSM_STAGES = ["sort_count", "sort_scan", "sort_scatter_bins", "accumulate", "bucket_reduce", "unused", "total"]


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


def measure_imad_peak(gpu_index):
    """(IMAD.WIDE.U32 per second, 32-bit IMAD per second) from the register-only issue-rate microbenchmark
    (bulletproof_gadgets_b200/bin/imad_peak --quick, ~0.4 s); falls back to the pool's recorded figure if it cannot run."""
    exe = os.path.join(ROOT, "bulletproof_gadgets_b200", "bin", "imad_peak")
    try:
        out = subprocess.run([exe, "--quick", str(gpu_index)], capture_output=True, text=True, timeout=60).stdout
        vals = {}
        for ln in out.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                vals[d["op"]] = d["Tops"] * 1e12
        return vals["IMAD.WIDE.U32"], vals["IMAD"]
    except Exception:  # noqa: BLE001
        return IMAD_WIDE_PEAK_TOPS * 1e12, 18.5e12


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
    ap.add_argument("--steps", type=int, default=16, help="batches of %d statements" % PROOFS_PER_STEP)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--count", type=int, default=1024, help="BOUND statements per proof (1024 = BASELINE config 2)")
    ap.add_argument("--inflight", type=int, default=0, help="statements in flight per GPU (library host threads); 0 = auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip config4 / msm_sharded / e2e_statement")
    ap.add_argument("--config4-count", type=int, default=4096)
    ap.add_argument("--sharded-lg", type=int, default=22)
    args = ap.parse_args()
    ws, rank, local_rank = _dist()
    if args.impl == "reference":
        return run_reference(args, ws, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    if ws > 1:
        torch.cuda.set_device(local_rank)
        # The data path has no collective (independent proofs per rank): torch.distributed only provides the barrier,
        # the max-reduce of the timings and the gather of <= 8 x 32-byte partial points.  Measured on the 8-GPU box
        # (32 vCPUs): with an NCCL process group alive every rank burns ~23 ms of host CPU per step in NCCL's service
        # threads (profiles/r01_bench_8gpu_{nccl,gloo}.json), and this workload is host-CPU bound there -- so gloo.
        backend = os.environ.get("BPG_DIST_BACKEND", "gloo")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    import bulletproof_gadgets_b200 as bpg
    from bulletproof_gadgets_b200 import build, sharding, workloads as W
    build.build_lib()
    warmup = max(args.warmup, 3)
    cores_per_rank = (os.cpu_count() or 1) // ws
    # (one rank of an 8-GPU box emulated with `taskset -c 0-3`: 16 / 32 / 48 / 64 in flight = 81 / 113 / 131 / 128 per second,
    # profiles/r02_ab7.jsonl -- 48 even where a rank has four cores)
    # and with `taskset -c 0-11` (a rank of the 2-GPU box): 48 / 64 / 96 in flight = 124 / 132 / 137 per second (r02_ab8.jsonl)
    inflight = args.inflight or (96 if cores_per_rank >= 12 else 64 if cores_per_rank >= 6 else 48)
    ctx0 = bpg.Context(local_rank)  # raises loudly without an sm_100a device: there is no CPU path
    # roofline denominator, measured in this run BEFORE the load (a kernel timed alone sees these clocks: the "burst" figure):
    # the issue-rate microbenchmark of round 1 (tools/imad_peak.cu, built with the library), rank 0's GPU
    imad_wide_peak, imad32_peak = measure_imad_peak(local_rank)
    ctxs = [ctx0] + [ctx0.shared() for _ in range(inflight - 1)]
    text_inflight = min(inflight, int(os.environ.get("BPG_BENCH_TEXT_INFLIGHT", "48")))
    st = W.bounds_check_statement(args.count, seed=20261018 + rank, label=b"bench-bound-%d" % rank).pin(bpg)
    ctx0.gens_ensure(st.n)
    circuit = bpg.Circuit(ctx0, st.n, st.m, st.row_start, st.term_var, st.term_coef, st.q).set_witness(st.aL, st.aR)
    dev = torch.device("cuda", local_rank)
    streams = [torch.cuda.ExternalStream(c.get("stream"), device=dev) for c in ctxs]
    gad_txt, inst_txt, wtns_txt = W.bounds_check_text(args.count, seed=20261018 + rank)
    name_txt = "bench-bound-%d" % rank

    def seeds(first, n):
        return [(first + i + 1).to_bytes(32, "little") for i in range(n)]

    def run_resident(first, n, use=None):
        sd = seeds(first, n)
        out = bpg.prove_batch(use or ctxs, [st] * n, sd, circuits=circuit, verify=True, verify_seeds=sd)
        if any(o[0] != 0 for o in out):
            raise SystemExit("GPU proof did not verify: %r" % [o[0] for o in out if o[0]][:4])
        return out

    def run_e2e(first, n, use=None):
        sd = seeds(first, n)
        out = bpg.prove_batch(use or ctxs, [st] * n, sd, verify=True, verify_seeds=sd)   # host buffers: uploads inside
        if any(o[0] != 0 for o in out):
            raise SystemExit("GPU proof did not verify")
        return out

    def run_statement(first, n, use=None):
        sd = seeds(first, n)
        # (the text front end costs ~35 ms of host CPU per side: more than 48 threads only oversubscribe a 16-core host)
        out = bpg.prove_text_batch(use or ctxs[:text_inflight], [(name_txt, inst_txt, wtns_txt, gad_txt)] * n, sd, sd, verify=True)
        if not all(o[0] == 0 and o[3] for o in out):
            raise SystemExit("GPU proof did not verify")
        return out

    def barrier():
        for s in streams:
            s.synchronize()
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()

    def max_over_ranks(x):
        if ws > 1:
            t = torch.tensor([x], device="cuda" if dist.get_backend() == "nccl" else "cpu", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            x = float(t.item())
        return x

    def timed(fn, nstat, warm_stat, use=None):
        fn(0, warm_stat, use)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        fn(1000, nstat, use)
        for s in streams:
            s.synchronize()
        e1.record(streams[0])
        e1.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    nproofs = args.steps * PROOFS_PER_STEP
    nwarm = max(warmup * PROOFS_PER_STEP, inflight)
    launches0 = sum(c.get("launches") for c in ctxs)
    stat0 = [bpg.lib().bpg_rng_batcher_stat(k) for k in range(3)]
    cpu0 = time.process_time()
    ms_res = timed(run_resident, nproofs, nwarm)
    cpu_res = time.process_time() - cpu0
    launches = sum(c.get("launches") for c in ctxs) - launches0
    ms_e2e = timed(run_e2e, nproofs, nwarm)
    ms_stmt = None if args.no_extra_legs else timed(run_statement, nproofs, nwarm)
    stat1 = [bpg.lib().bpg_rng_batcher_stat(k) for k in range(3)]
    cpu_parts = {k: sum(c.get("cpu_%s_ns" % k) for c in ctxs) * 1e-9 for k in ("sync", "commit", "prove", "verify", "rng")}
    lat_n = 10
    ms_lat = timed(run_resident, lat_n, 3, ctxs[:1])
    sampler.stop_flag = True

    # one instrumented statement: CUDA events around every MSM stage (synchronous mode of the library)
    ctx0.set("time_accum", 1)
    run_resident(10 ** 6, 1, ctxs[:1])
    acc_ns, acc_entries = ctx0.get("sum_accum_ns"), ctx0.get("sum_entries")
    sct_ns, sct_points = ctx0.get("sum_scatter_ns"), ctx0.get("sum_points")
    step_stage_ms = {nm: ctx0.get("stage_ns_%d" % i) * 1e-6 for i, nm in enumerate(MSM_STAGES)}
    sort_ns = sum(ctx0.get("stage_ns_%d" % i) for i in range(3))
    step_msms = ctx0.get("timed_msms")
    ctx0.set("time_accum", 0)
    imad_wide_sustained, _ = measure_imad_peak(local_rank)   # ... and again after the legs (warm part)
    peak_tops = max(imad_wide_peak, imad_wide_sustained) / 1e12
    imad_wide_peak = peak_tops * 1e12

    # second half of the BASELINE metric: raw fixed-base MSM (verifier mega-MSM shape), 2n' = 2^18 points, uniform
    # scalars resident in HBM; whole MSM (all stages + 128-byte read-back + host ristretto compression)
    npts = 2 * st.n
    sc_np = np.random.default_rng(7).integers(0, 256, size=(npts, 32), dtype=np.uint8)
    sc_np[:, 31] &= 0x0F
    d_sc = torch.from_numpy(sc_np).to(dev)
    msm_call = lambda: ctx0.msm_gens_dev(d_sc.data_ptr(), st.n, d_sc.data_ptr() + 32 * st.n, st.n)
    for _ in range(3):
        msm_call()
    streams[0].synchronize()
    msm_runs = []
    for _ in range(5):      # wall clock around synchronous calls: median of five batches of ten
        t0 = time.perf_counter()
        for _ in range(10):
            msm_call()
        msm_runs.append((time.perf_counter() - t0) / 10)
    msm_s = sorted(msm_runs)[len(msm_runs) // 2]
    # the same MSM as a THROUGHPUT (the unit of the metric is points per second): four contexts, one host thread each, issue
    # their synchronous calls side by side, so one MSM's read-back / host compression and the narrow tail of its bucket
    # reduction overlap another one's sort and bucket kernels -- how the proofs above use the GPU
    import threading
    msm_k, msm_reps = 4, 12

    def msm_worker(c):
        for _ in range(msm_reps):
            c.msm_gens_dev(d_sc.data_ptr(), st.n, d_sc.data_ptr() + 32 * st.n, st.n)

    msm_par = []
    for _ in range(4):
        th = [threading.Thread(target=msm_worker, args=(c,)) for c in ctxs[:msm_k]]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        msm_par.append((time.perf_counter() - t0) / (msm_k * msm_reps))
    msm_par_s = sorted(msm_par[1:])[1]   # first round warms the other contexts' buffers
    ctx0.set("time_accum", 1)
    for _ in range(5):
        msm_call()
    msm_stage_us = {nm: ctx0.get("stage_ns_%d" % i) / ctx0.get("timed_msms") / 1e3 for i, nm in enumerate(MSM_STAGES)}
    msm_cl = ctx0.get("last_chunk_len")
    ctx0.set("time_accum", 0)
    del d_sc

    config4 = msm_sharded = None
    if not args.no_extra_legs:
        # ---- BASELINE config 4: a batch of independent small proofs, statement i on rank i mod N (strong scaling) ----
        texts = W.batch_texts(args.config4_count)
        mine = sharding.shard_jobs(args.config4_count, ws, rank)
        jobs = [("batch-%d" % i, texts[i][1], texts[i][2], texts[i][0]) for i in mine]
        sd = [(i + 1).to_bytes(32, "little") for i in mine]

        def run_c4(first, n, use=None):
            out = bpg.prove_text_batch(ctxs[:48], jobs[:n], sd[:n], sd[:n], verify=True)   # small statements: bound by driver calls, not by streams
            if not all(o[0] == 0 and o[3] for o in out):
                raise SystemExit("config 4: a proof did not verify")

        ms_c4 = timed(run_c4, len(jobs), min(len(jobs), 4 * inflight))
        config4 = {"proofs": args.config4_count, "value": args.config4_count / (ms_c4 * 1e-3), "unit": UNIT, "scaling": "strong",
                   "workload": "LESS_THAN (n=379) / SET_MEMBER k=16 (n=32) alternating, statement i on rank i mod N; text in, "
                               "proof out and verified (bpg_prove_batch with BPG_JOB_VERIFY), front end inside the timer",
                   "ms": ms_c4}
        # ---- BASELINE config 5 at N GPUs: ONE MSM of 2^lg points split by contiguous point range ----
        lg = args.sharded_lg
        half = 1 << (lg - 1)
        g0, g1 = sharding.point_ranges(half, ws)[rank]
        ctx0.gens_ensure(half)
        rng_s = np.random.default_rng(99)          # every rank draws the same scalars and keeps its slice
        allsc = rng_s.integers(0, 256, size=(2 * half, 32), dtype=np.uint8)
        allsc[:, 31] &= 0x0F
        dG = torch.from_numpy(allsc[g0:g1].copy()).to(dev)
        dH = torch.from_numpy(allsc[half + g0: half + g1].copy()).to(dev)
        part_call = lambda: ctx0.msm_gens_range_dev(dG.data_ptr(), g0, g1 - g0, dH.data_ptr(), g0, g1 - g0)
        for _ in range(3):
            part = part_call()
        barrier()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for _ in range(reps):
            part = part_call()
            if ws > 1:   # 32 bytes per rank, host side (gloo): one small tensor all-gather
                mine_t = torch.frombuffer(bytearray(part), dtype=torch.uint8)
                if dist.get_backend() == "nccl":
                    mine_t = mine_t.to(dev)
                got = [torch.empty_like(mine_t) for _ in range(ws)]
                dist.all_gather(got, mine_t)
                parts = [bytes(g.cpu().numpy().tobytes()) for g in got]
            else:
                parts = [part]
            total_pt = sharding.point_sum(parts) if rank == 0 else None
        e1.record(streams[0])
        e1.synchronize()
        ms_sh = max_over_ranks(e0.elapsed_time(e1)) / reps
        msm_sharded = {"points": 2 * half, "mpoints_per_s": 2 * half / (ms_sh * 1e-3) / 1e6, "ms": ms_sh, "scaling": "strong",
                       "result": total_pt.hex() if total_pt else None,
                       "frac_of_imad_peak": 16 * 2 * half * IMAD_PER_MADD / (ms_sh * 1e-3) / (ws * imad_wide_peak),
                       "note": "uniform scalars resident in HBM, every rank multiplies its contiguous range of G and H, the "
                               "N compressed partial points are gathered on the host (gloo) and added by bpg_point_sum; "
                               "the gather and the sum are inside the timed region"}
        del dG, dH

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
    # bytes crossing PCIe per e2e statement, counted from the buffers handed to the C ABI plus the library's own uploads
    csr = 4 * (st.q + 1) + 36 * st.nnz
    h2d = 2 * 32 * st.n + csr            # bpg_prover_load_cs: a_L, a_R, CSR constraints
    h2d += csr                           # bpg_verifier_load_cs
    h2d += 64 * st.m + 32 * st.m         # commit_batch (v, blinding) ; v_blinding vector of the prover
    h2d += 128 * st.n                    # 64-byte TranscriptRng draws for s_L, s_R (reduced mod l on the device)
    h2d += 32 * (6 + st.m + 5 + 2 * lg) * 2  # verifier: compressed points + head scalars
    d2h = 32 * st.m + 128 * (3 + 2 * lg) + 9 * 32 + 64 + 128 + 16  # commitments, A/S/L/R points, t scalars, a/b, check point, flags
    kernel_ms_per_statement = acc_ns * 1e-6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps, "warmup": warmup,
        "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if args.count == 1024 else "bounds_check 64-bit x%d" % args.count,
                   "proofs_per_step": PROOFS_PER_STEP, "inflight": inflight, "inflight_text_legs": text_inflight,
                   "hardware_queues": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "dist_backend": (dist_backend if ws > 1 else None),
                   "api": "bpg_r1cs_prove_batch (BPG_JOB_VERIFY): library-owned host threads, one context per statement in flight",
                   "l2": "per-step working set (fixed-base tables 403 MB + entries) exceeds the 126 MB L2",
                   "rng": "transcript rng seeded per statement; proofs byte-identical to the CPU oracle",
                   "window_bits": ctx0.get("window_bits"), "chunk_len": "device-derived (one wave of equal chunks)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": h2d * PROOFS_PER_STEP, "d2h_bytes_per_step": d2h * PROOFS_PER_STEP},
        "latency": {"ms_per_proof": ms_lat / lat_n, "proofs": lat_n,
                    "note": "one proof+verify at a time on one context; the prover waits on the host for the sequential "
                            "Merlin TranscriptRng stream (2n Keccak-f permutations)"},
        "msm": {"points": npts, "mpoints_per_s": npts / msm_s / 1e6, "ms": msm_s * 1e3, "ms_batches": [round(x * 1e3, 4) for x in msm_runs],
                "scalars": "uniform mod l", "chunk_len": msm_cl,
                "frac_of_imad_peak_whole_msm": 16 * npts * IMAD_PER_MADD / msm_s / imad_wide_peak,
                "in_flight4": {"ms_per_msm": msm_par_s * 1e3, "mpoints_per_s": npts / msm_par_s / 1e6,
                               "frac_of_imad_peak_whole_msm": 16 * npts * IMAD_PER_MADD / msm_par_s / imad_wide_peak,
                               "note": "four contexts (streams) issuing the same MSM side by side, wall clock / MSMs"},
                "stage_us": {k: round(v, 1) for k, v in msm_stage_us.items()},
                "stage_note": "CUDA events between the launches of one MSM in the library's synchronous timing mode "
                              "(each figure carries ~3 us of event / launch gap); `ms` is wall clock without events",
                "note": "one GPU; sweep 2^12..2^22 in profiles/r02_msm_stages.jsonl (tools/gpu_msm_stages.py)"},
        "msm_stages_per_statement_ms": {"msms": step_msms, **{k: round(v, 4) for k, v in step_stage_ms.items()}},
        "gpu_launches": launches,
        "host": {"cores": os.cpu_count(), "cpu_s_per_proof_rank0": cpu_res / (nproofs + nwarm),
                 "rng_streams": stat1[0] - stat0[0], "rng_vector_batches": stat1[1] - stat0[1],
                 "rng_streams_alone": stat1[2] - stat0[2],
                 "thread_cpu_s_total_all_legs": cpu_parts,
                 "note": "process CPU time of rank 0 over the resident leg (incl. its warm-up statements); the Merlin rng "
                         "streams of proofs in flight run eight at a time (AVX-512) when batches form"},
        "roofline": {"kernel": "k_accumulate (MSM bucket accumulation, mixed Edwards adds)", "bound": "imad",
                     "achieved": achieved, "peak": peak_tops, "unit": "T IMAD.WIDE/s",
                     "frac": achieved / peak_tops if achieved else None, "traffic": ACC_TRAFFIC_BYTES,
                     "traffic_note": "dram bytes read+write per launch, ncu --set full of this kernel inside a raw MSM of 2^18 "
                                     "points (profiles/r02_msm_kernels_details.csv); algorithmic gather 4.19 M entries x 96 B = 403 MB",
                     "peak_source": "measured in this run by tools/imad_peak.cu --quick (register-only IMAD.WIDE.U32 issue-rate kernel, "
                                    "the microbenchmark of round 1) before the legs; not in MEASURED_PEAKS.json",
                     "peak_sustained": imad_wide_sustained / 1e12,
                     "peak_sustained_note": "the same kernel after the legs (warm part, power-capped clocks)",
                     "frac_of_int32_imad_peak": achieved * 1e12 / imad32_peak if achieved else None,
                     "int32_imad_peak_tops": imad32_peak / 1e12,
                     "int32_note": "the same work against the 32-bit IMAD issue rate north_star literally names: one IMAD.WIDE "
                                   "counted as one instruction (count it as two 32-bit multiplies and the fraction doubles)",
                     "work": "%d mixed adds x %d IMAD.WIDE" % (acc_entries, IMAD_PER_MADD),
                     "kernel_ms_per_proof": kernel_ms_per_statement,
                     "share_of_step": kernel_ms_per_statement * PROOFS_PER_STEP / per_step_ms,
                     "share_note": "kernel time of one statement x statements per step / timed throughput step"},
        # the sort stage north_star asks to see against HBM: digit decomposition + scatter by bucket.  Algorithmic bytes
        # per launch: 32 B per scalar read + 4 B per entry written + 4 B per entry of offset reads.
        "roofline_sort": {"kernel": "k_sort_count + k_sort_scan + k_sort_scatter + k_sort_bins (signed-digit decomposition + two-level "
                                    "shared-memory counting sort by bucket; k_digits / k_scan_meta for the MSM shapes those do not serve)",
                          "bound": "hbm", "achieved": (32 * sct_points + 8 * acc_entries) / (sort_ns * 1e-9) / 1e9 if sort_ns else None,
                          "peak": HBM_PEAK_GBS, "unit": "GB/s",
                          "frac": (32 * sct_points + 8 * acc_entries) / (sort_ns * 1e-9) / 1e9 / HBM_PEAK_GBS if sort_ns else None,
                          "traffic": SORT_TRAFFIC_BYTES, "kernel_ms_per_proof": sort_ns * 1e-6,
                          "note": "algorithmic bytes = 32 B per scalar + 4 B per entry written + 4 B per entry of offsets; the stage "
                                  "is bound by shared-memory atomics and launch latency (four kernels of 7-25 us), its intermediate "
                                  "arrays stay in the 126 MB L2: DRAM traffic per launch is the `traffic` figure"},
        "clocks": sampler.summary(),
    }
    if ms_stmt is not None:
        line["e2e_statement"] = {"value": ws * nproofs / (ms_stmt * 1e-3), "unit": UNIT, "ms_per_step": ms_stmt / args.steps,
                                 "note": "text formats in, proof + commitments text out and back in: bpg_prove_batch (the batched "
                                         "c_prove / c_verify), host front end (parse, gadgets, witness assignment, flattening) "
                                         "inside the timed region on both sides"}
    if config4:
        line["config4"] = config4
    if msm_sharded:
        line["msm_sharded"] = msm_sharded
    if not args.no_cpu_baseline:
        from oracle import coracle
        cst = W.bounds_check_statement(args.count, seed=20261018 + rank, label=b"bench-bound-%d" % rank)
        t0 = time.perf_counter()
        p_c, coms_c = coracle.prove_flat(cst, (1).to_bytes(32, "little"), cache_gens=False)
        ok = coracle.verify_flat(cst, coms_c, p_c, (1).to_bytes(32, "little"), cache_gens=False)
        dt = time.perf_counter() - t0
        same = p_c == run_resident(0, 1, ctxs[:1])[0][1]
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "the full statement once (prove+verify, generators rebuilt per run as the reference "
                                          "does): %.1f s; CPU proof accepted=%s, byte-identical to the GPU proof=%s" % (dt, ok, same)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
