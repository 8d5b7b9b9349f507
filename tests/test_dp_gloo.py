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
