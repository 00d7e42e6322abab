"""CPU test of the N>1 host logic with world_size 2 over gloo: shard ranges, the all-gather of partial points,
and that summing per-shard partials reproduces the single-context result (oracle used as the checker)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, n, seed, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    import oracle_lib as O
    from msm_blst_b200 import distributed as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(n, rank, world)
    sc = O.gen_scalars(seed, n)
    # this rank's partial: oracle Pippenger over its own shard P_lo..P_hi-1, kept Jacobian via to_affine -> (x, y, 1)
    oc = O.OracleCtx(1, "8", n=hi - lo, first=lo)
    oc.init_fix_points()
    aff = oc.msm(4, sc[lo:hi])
    one = np.frombuffer(bytes.fromhex("fdff02000000097602000cc40b00f4ebba58c7535798485f455752705358ce776dec56a2971a075c93e480fac35ef615"), dtype=np.uint8)
    jac = np.concatenate([aff, one if aff.any() else np.zeros(48, dtype=np.uint8)])
    gathered = D.all_gather_partials(torch.from_numpy(jac.copy()))
    assert gathered.shape == (world, 144)
    out = np.zeros(96, dtype=np.uint8)
    g = gathered.numpy().copy()
    O.oracle().oracle_sum_partials(1, O.ptr(g), world, O.ptr(out))
    full, _ = O.closed_form(1, sc)
    ok = bool((out == full).all()) and bool((g[rank] == jac).all())
    q.put((rank, ok, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    sys.path.insert(0, ROOT)
    from msm_blst_b200 import distributed as D

    for n in (1, 7, 1000, 1 << 21):
        for world in (1, 2, 3, 4, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert D.shard_config_name(1 << 18) == "18" and D.shard_config_name(1 << 21) == "21" and D.shard_config_name(100) == "8"
    assert D.shard_config_name(1 << 16, 2) == "17" and D.shard_config_name(1 << 16, 1) == "16" and D.shard_config_name(1 << 20, 1) == "19"


def test_two_rank_gather_and_sum_gloo():
    world, n = 2, 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 11, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
