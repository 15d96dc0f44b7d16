"""Multi-process (N > 1) host logic on CPU: world_size-2 `gloo` runs of the two partitionings of SURVEY.md 8e.

No GPU here, so each rank's device work is played by the oracle (test infrastructure); what is under test is the
product's host side: job sharding, point-range splitting, the host combine of partial points (bpg_point_sum,
pure host code of libbpg.so) and the gather over torch.distributed."""
import os
import random
import socket

import pytest

from bulletproof_gadgets_b200 import sharding

L = 2**252 + 27742317777372353535851937790883648493


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import coracle
    from oracle.pyref import frontend as F
    from tests import frontend_glue as G

    # (1) independent proofs: job i -> rank i mod world
    jobs = [("LESS_THAN W0 W1", "", "W0 = 0x%02x\nW1 = 0x%02x" % (i, i + 1 + i % 3)) if i % 2 == 0 else
            ("SET_MEMBER W0 I0 I1 I2", "I0 = 0x01\nI1 = 0x%02x\nI2 = 0x7f" % (i + 2), "W0 = 0x%02x" % (i + 2)) for i in range(7)]
    mine = {}
    for i in sharding.shard_jobs(len(jobs), world, rank):
        gad, inst, wtns = jobs[i]
        st = F.compile_prover("job-%d" % i, inst, wtns, gad, G.blinding(b"job%d" % i))
        proof, coms = coracle.prove_flat(st, bytes([i]) * 32)
        mine[i] = (proof, st.coms_text(coms))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # (2) one MSM split by point range: every rank computes ONE partial point, the host adds them
    rnd = random.Random(5)
    nG, nH = 37, 22
    sG = b"".join(rnd.randrange(L).to_bytes(32, "little") for _ in range(nG))
    sH = b"".join(rnd.randrange(L).to_bytes(32, "little") for _ in range(nH))
    sB = rnd.randrange(L).to_bytes(32, "little")
    g0, g1 = sharding.point_ranges(nG, world)[rank]
    h0, h1 = sharding.point_ranges(nH, world)[rank]
    scal = [sG[32 * i: 32 * i + 32] for i in range(g0, g1)] + [sH[32 * i: 32 * i + 32] for i in range(h0, h1)]
    pts = coracle.gens("G", g0, g1 - g0) + coracle.gens("H", h0, h1 - h0)
    if rank == 0:
        scal.append(sB)
        pts += coracle.gens("B", 0, 1)
    partial = coracle.msm(scal, pts)
    partials = [None] * world
    dist.all_gather_object(partials, partial)
    if rank == 0:
        allp = {}
        for d in gathered:
            allp.update(d)
        assert sorted(allp) == list(range(len(jobs)))
        for i, (gad, inst, _) in enumerate(jobs):
            proof, text = allp[i]
            vs = F.compile_verifier("job-%d" % i, inst, text, gad)
            assert coracle.verify_flat(vs, vs.V, proof, b"\x09" * 32) is True
        total = sharding.point_sum(partials)
        assert total == coracle.msm_gens(sG, sH, sB, None)
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from bulletproof_gadgets_b200 import build
    build.build_lib()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_partition_helpers():
    assert sharding.shard_jobs(10, 4, 1) == [1, 5, 9]
    assert sharding.shard_jobs(3, 8, 5) == []
    for n in (0, 1, 7, 8, 9, 1 << 17):
        for parts in (1, 2, 3, 8):
            r = sharding.point_ranges(n, parts)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_point_sum_host_only():
    from tests.test_oracle_anchors import RFC9496_MULTIPLES as M
    pts = [bytes.fromhex(M[k]) for k in (1, 2, 3, 4)]
    assert sharding.point_sum(pts).hex() == M[10]
    assert sharding.point_sum([]) == bytes(32)
    assert sharding.point_sum([bytes.fromhex(M[7])]).hex() == M[7]
    import bulletproof_gadgets_b200 as bpg
    with pytest.raises(bpg.BpgError):
        sharding.point_sum([bytes.fromhex("00" + "ff" * 31)])
