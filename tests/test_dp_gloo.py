"""world_size-2 gloo test (CPU) of the data-parallel host logic: flat parameter/gradient views,
sample sharding by rank, SUM all-reduce + 1/world scaling == DDP's mean, parameter broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
    from joligen_b200 import dp
    from joligen_b200.trainer import FlatParams
    from oracle import palette_oracle as O
    torch.manual_seed(100 + rank)  # deliberately different initial weights per rank
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GroupNorm(4, 8), torch.nn.Conv2d(8, 3, 1))
    flat = FlatParams(net)
    dp.broadcast_(flat.data)
    w_after = [p.detach().clone() for p in net.parameters()]
    # rank-sharded synthetic batch
    data = O.synthetic_batch(2, 16, dp.shard_seed(1234, rank))
    loss = (net(data["gt"]) - data["cond"]).pow(2).mean()
    loss.backward()  # accumulates into views of flat.grad
    for p, o in zip(flat.params, flat.offsets):
        assert p.grad.data_ptr() == flat.grad.data_ptr() + 4 * o
    local = flat.grad.clone()
    scale = dp.allreduce_sum_(flat.grad)
    out[rank] = {"w": w_after, "local": local, "reduced": flat.grad.clone() * scale, "scale": scale,
                 "gt_sum": float(data["gt"].sum())}
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_mean():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert a["scale"] == 0.5
    for wa, wb in zip(a["w"], b["w"]):
        assert torch.equal(wa, wb)  # broadcast made the replicas identical
    assert a["gt_sum"] != b["gt_sum"]  # each rank saw its own shard
    mean = 0.5 * (a["local"] + b["local"])
    assert torch.allclose(a["reduced"], mean, rtol=1e-6, atol=1e-8)
    assert torch.equal(a["reduced"], b["reduced"])
    assert float(mean.abs().sum()) > 0


def test_plan_buckets_covers_the_flat_buffer_in_reverse_order():
    from joligen_b200 import dp
    sizes = [64, 640, 128, 4096, 64, 1024, 2048, 64]
    offsets, off = [], 0
    for n in sizes:
        offsets.append(off)
        off += n
    buckets = dp.plan_buckets(offsets, sizes, off, n_buckets=4, min_elems=1)
    assert buckets[0][1] == off and buckets[-1][0] == 0          # last parameters leave first
    for (lo, hi, members), (lo2, hi2, _) in zip(buckets, buckets[1:]):
        assert lo == hi2 and lo2 < hi2                            # contiguous, descending
    seen = sorted(i for _, _, m in buckets for i in m)
    assert seen == list(range(len(sizes)))                        # every parameter in exactly one bucket
    for lo, hi, members in buckets:
        assert lo == min(offsets[i] for i in members) and hi == max(offsets[i] + sizes[i] for i in members)
    assert dp.plan_buckets([0], [10], 10, n_buckets=8, min_elems=1) == [(0, 10, [0])]


def _bucket_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from joligen_b200 import dp
    sizes = [64, 640, 128, 4096, 64, 1024, 2048, 64]
    offsets, off = [], 0
    for n in sizes:
        offsets.append(off)
        off += n
    g = torch.Generator().manual_seed(7 + rank)
    grad = torch.randn(off, generator=g)
    local = grad.clone()
    comm = dp.Comm(None, "cpu")
    order = []

    def on_bucket(b):
        lo, hi, _ = gb.buckets[b]
        order.append(b)
        comm.allreduce_async(grad[lo:hi])

    gb = dp.GradBuckets(offsets, sizes, off, on_bucket, n_buckets=4, min_elems=1)
    gb.begin()
    for i in (7, 6, 5, 3, 4, 1, 2):   # backward order, slightly shuffled; parameter 0 never gets a gradient
        gb.ready(i)
    launched_during_backward = list(order)
    gb.finish()
    comm.wait()
    total = local.clone()
    dist.all_reduce(total)
    res = {"ok": bool(torch.allclose(grad, total)), "during": launched_during_backward, "order": list(order),
           "nb": len(gb.buckets)}
    # a gradient that arrives after its bucket left must be reported, not silently dropped
    gb.begin()
    gb.ready(7)
    gb.ready(6)
    while not gb.launched[gb.bucket_of[7]]:
        gb.ready(min(gb.missing[gb.bucket_of[7]]))
    gb.ready(7)
    try:
        gb.finish()
        res["late_detected"] = False
    except RuntimeError:
        res["late_detected"] = True
    comm.wait()
    out[rank] = res
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_matches_one_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        assert out[r]["ok"], "bucketed sum differs from the single all-reduce"
        assert out[r]["order"] == list(range(out[r]["nb"]))       # launch order = bucket order on every rank
        assert len(out[r]["during"]) >= out[r]["nb"] - 1          # all but the bucket with the unused parameter
        assert out[r]["late_detected"]
