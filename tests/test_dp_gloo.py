"""Host-side logic of the data-parallel path on CPU: world_size 2, gloo.

Covers what does not need a GPU: sample sharding, identical sampler streams on every rank, arena layout, and the
order "sum the pre-sign gradients over ranks, THEN take the sign" (sign_sgd.py:389 acts on the accumulated grad;
utils/distributed.py:44-46).  The kernels themselves are covered by the `-m gpu` tests."""
import os
import random
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from auto_round_b200.quantizer import DataParallel, IndexSampler, TuneArena, lr_schedule_table


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dp = DataParallel(rank, world, None)
        # every rank draws the same batch sequence from the same python-random seed and takes its slice
        random.seed(42)
        sampler = IndexSampler(16, 8)
        batches = [sampler.next_batch() for _ in range(5)]
        mine = [dp.shard(b) for b in batches]
        gathered = [None] * world
        dist.all_gather_object(gathered, (batches, mine))
        assert all(g[0] == batches for g in gathered)
        for it in range(5):
            union = sorted(sum((g[1][it] for g in gathered), []))
            assert union == sorted(batches[it]) and len(mine[it]) == 8 // world
        # per-sample "gradients": rank-local partial sums, all-reduced before the sign
        g = torch.Generator().manual_seed(7)
        per_sample = torch.randn(16, 64, generator=g)
        arena = torch.zeros(64)
        for s in mine[0]:
            arena += per_sample[s]
        loss = torch.tensor([float(len(mine[0]))], dtype=torch.float64)
        dp.all_reduce_(arena, loss)
        full = per_sample[batches[0]].sum(0)
        assert torch.allclose(arena, full, atol=1e-5)
        assert float(loss) == 8.0
        big = full.abs() > 1e-4
        assert torch.equal(torch.sign(arena)[big], torch.sign(full)[big])
        # identical update on every rank -> identical parameters without a broadcast
        p = torch.zeros(64) - 0.005 * torch.sign(arena)
        ps = [torch.zeros_like(p) for _ in range(world)]
        dist.all_gather(ps, p)
        assert all(torch.equal(ps[0], q) for q in ps)
        # optimized RTN under DP: each rank accumulates sum x^2 and the sample count over ITS shard of the calibration
        # forward; one all-reduce of (sum, count) gives every rank the reference's normalised importance matrix
        xs = torch.randn(16, 5, 32, generator=torch.Generator().manual_seed(11))
        per = 16 // world
        part = xs[rank * per:(rank + 1) * per]
        acc, cnt = part.reshape(-1, 32).pow(2).sum(0), torch.tensor([float(part.shape[0])])
        dp.all_reduce_(acc, cnt)
        assert float(cnt) == 16.0
        assert torch.allclose(acc / cnt, xs.reshape(-1, 32).pow(2).sum(0) / 16, rtol=1e-6)
        # the transport calls of the sharded update (gloo emulation of reduce-scatter / all-gather; NCCL on the GPU box):
        # reduce_scatter_ -> this rank's slice of the SUM; all_gather_ -> every rank's slice, also in place and for the
        # integer / byte payloads (routing ids, histograms, 4-bit wire segments)
        full = (torch.arange(8, dtype=torch.float32) + 10 * rank).to(torch.bfloat16)
        mine = torch.empty(8 // world, dtype=torch.bfloat16)
        dp.reduce_scatter_(mine, full.clone())
        want = sum((torch.arange(8, dtype=torch.float32) + 10 * r) for r in range(world))
        assert torch.equal(mine.float(), want[rank * (8 // world):(rank + 1) * (8 // world)])
        buf = torch.zeros(8, dtype=torch.bfloat16)
        buf[rank * 4:(rank + 1) * 4] = torch.tensor([1.5, -2.0, 0.25, 3.0]) * (rank + 1)
        dp.all_gather_(buf, buf[rank * 4:(rank + 1) * 4])
        assert buf.tolist() == [1.5, -2.0, 0.25, 3.0, 3.0, -4.0, 0.5, 6.0]
        for dt in (torch.uint8, torch.int64, torch.int32):
            seg = torch.full((3,), rank + 7, dtype=dt)
            allb = torch.zeros(3 * world, dtype=dt)
            dp.all_gather_(allb, seg)
            assert allb.tolist() == [7, 7, 7, 8, 8, 8]
        assert dp.row_shard(8) == (rank * 4, rank * 4 + 4) and dp.row_shard(7) is None and DataParallel().row_shard(8) is None
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_dp_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_arena_layout_and_lr_table():
    from auto_round_b200 import ops
    specs = {"q": ops.make_spec("int_sym", 4, 128, 256, 512), "k": ops.make_spec("int_sym", 4, 128, 64, 512),
             "m": ops.make_spec("mx_fp4", 4, 32, 128, 256)}
    a = TuneArena(specs, "cpu")
    nv = 256 * 512 + 64 * 512 + 128 * 256
    assert a.clamp_begin == nv and a.numel % 4 == 0 and a.clamp_begin % 4 == 0
    assert set(a.views["q"]) == {"value", "max_scale", "min_scale"} and set(a.views["m"]) == {"value", "max_scale"}
    assert float(a.params[:nv].abs().sum()) == 0.0 and bool((a.params[nv:] == 1).all())
    v = a.layer_views("k")
    v["value"].fill_(3.0)
    o, n, _ = a.views["k"]["value"]
    assert float(a.params[o:o + n].sum()) == 3.0 * n          # views alias the arena
    # fp32 LinearLR recursion == torch's scheduler on a tensor lr (quantizer.py:426-429)
    tab = lr_schedule_table(200, 1 / 200, 1 / 200)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([{"params": [p], "lr": torch.tensor(1 / 200)}], lr=torch.tensor(1 / 200))
    sch = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=200)
    for t in range(200):
        assert float(opt.param_groups[0]["lr"]) == float(tab[t, 0]), t
        opt.step()
        sch.step()
