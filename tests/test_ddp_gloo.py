"""CPU, world_size 2, gloo: the data-parallel plumbing of mdil_ss_b200.parallel (flat gradient buffer, ONE
all-reduce per optimiser step, fused-Adam semantics, batch sharding).  The kernels are not involved: a small torch
model stands in for the network so the N>1 host path is covered without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdil_ss_b200.parallel import FlatAdam, FlatBuffer, shard_batch


def _toy(seed=0):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 1))


def test_flat_buffer_makes_params_and_grads_views():
    m = _toy()
    ref = [p.detach().clone() for p in m.parameters()]
    buf = FlatBuffer(list(m.parameters()))
    assert buf.numel() == sum(p.numel() for p in m.parameters())
    for p, r in zip(m.parameters(), ref):
        assert torch.equal(p.detach(), r)
        assert p.data_ptr() >= buf.data.data_ptr() and p.grad.data_ptr() >= buf.grad.data_ptr()
    m(torch.rand(2, 3, 5, 5)).sum().backward()
    assert float(buf.grad.abs().sum()) > 0  # autograd accumulated straight into the flat buffer


def test_flat_adam_matches_torch_adam_two_groups():
    a, b = _toy(1), _toy(1)
    pa, pb = list(a.parameters()), list(b.parameters())
    opt_ref = torch.optim.Adam([{"params": pa[:2], "lr": 5e-6}, {"params": pa[2:]}], 5e-4, (0.9, 0.999), eps=1e-8,
                               weight_decay=1e-4)
    opt = FlatAdam([{"params": pb[:2], "lr": 5e-6}, {"params": pb[2:]}], 5e-4)
    x = torch.rand(4, 3, 6, 6)
    for it in range(5):
        opt_ref.zero_grad()
        a(x).square().mean().backward()
        opt_ref.step()
        opt.zero_grad()
        b(x).square().mean().backward()
        v0 = [p._version for p in pb]
        opt.step()
        assert all(p._version > v for p, v in zip(pb, v0))  # weight-pack caches key on the version counter
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)
    opt.set_lr_factor(0.5)
    assert abs(opt.groups[0]["lr"] - 2.5e-6) < 1e-18 and abs(opt.groups[1]["lr"] - 2.5e-4) < 1e-12


def test_shard_batch():
    assert shard_batch(24, 3, 8) == slice(9, 12)
    with pytest.raises(ValueError):
        shard_batch(6, 0, 4)


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
    torch.set_num_threads(1)
    m = _toy(3)
    opt = FlatAdam([{"params": list(m.parameters())}], 1e-3)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 3, 6, 6, generator=g)
    sl = shard_batch(4, rank, world)
    for _ in range(3):
        opt.zero_grad()
        m(x[sl]).square().mean().backward()   # a mean over equal shards averages exactly
        opt.step()
    assert opt.reducer.calls == 3              # exactly one collective per optimiser step
    if rank == 0:
        torch.save([p.detach().clone() for p in m.parameters()], out)
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_full_batch(tmp_path):
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    m = _toy(3)
    opt = FlatAdam([{"params": list(m.parameters())}], 1e-3)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 3, 6, 6, generator=g)
    for _ in range(3):
        opt.zero_grad()
        m(x).square().mean().backward()
        opt.step()
    for p, q in zip(m.parameters(), got):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-7)
