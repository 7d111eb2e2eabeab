"""N>1 path on CPU: world_size-2 gloo run of the bucketed gradient all-reduce (host logic only, no kernels)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, overlap, out):
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from stinet_b200.parallel import GradAllReducer, init_distributed
    init_distributed("gloo")
    torch.manual_seed(0)                                    # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ELU(), torch.nn.Linear(16, 16), torch.nn.ELU(),
                              torch.nn.Linear(16, 3))
    red = GradAllReducer(net, bucket_bytes=600, overlap=overlap)   # several buckets
    assert len(red.buckets) > 1
    g = torch.Generator().manual_seed(100 + rank)           # different shard per rank
    x = torch.randn(32, 6, generator=g)
    for _ in range(2):                                      # second iteration checks zero_grad / hook re-arming
        red.zero_grad()
        (net(x).square().mean() * red.loss_scale).backward()
        red.finish()
    assert all(p.grad.data_ptr() == v.data_ptr() for g_, vs in zip(red._groups, red._views) for p, v in zip(g_, vs))
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    # a parameter that receives no gradient must not keep its bucket from being reduced (it contributes zeros), and a
    # disabled reducer (warm-up of a graph capture) must leave gradients and collectives alone
    red.zero_grad()
    (net[0](x).square().mean() * red.loss_scale).backward()      # only the first Linear gets gradients
    red.finish()
    assert all(p.grad is not None for p in net.parameters())
    assert float(net[4].weight.grad.abs().max()) == 0.0 and float(net[0].weight.grad.abs().max()) > 0.0
    red.enabled = False
    red.zero_grad()
    net(x).square().mean().backward()
    red.finish()
    assert all(p.grad.data_ptr() != v.data_ptr() for g_, vs in zip(red._groups, red._views) for p, v in zip(g_, vs))
    red.enabled = True
    if rank == 0:
        torch.save(flat, out)
    dist.barrier()
    dist.destroy_process_group()


def _run(overlap, tmp_path):
    world, port = 2, _free_port()
    out = str(tmp_path / f"grads_{overlap}.pt")
    mp.spawn(_worker, args=(world, port, overlap, out), nprocs=world, join=True)
    got = torch.load(out)
    # single-process reference: mean over ranks of the per-rank gradients == gradient of the mean loss
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ELU(), torch.nn.Linear(16, 16), torch.nn.ELU(),
                              torch.nn.Linear(16, 3))
    loss = 0
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        loss = loss + net(torch.randn(32, 6, generator=g)).square().mean() / world
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-7)


def test_bucketed_allreduce_overlapped(tmp_path):
    _run(True, tmp_path)


def test_bucketed_allreduce_after_backward(tmp_path):
    _run(False, tmp_path)
