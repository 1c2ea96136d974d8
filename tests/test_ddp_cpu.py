"""N > 1 host-side logic on CPU (gloo, world_size 2): batch sharding + gradient all-reduce reproduce the 1-process gradient.

The CUDA kernels cannot run here, so the per-rank "layer" is the CPU oracle of the same KAN convolution; what is under test
is the data-parallel recipe bench.py uses (rank r takes images [r*B/N, (r+1)*B/N), DDP averages gradients), which relies on
the layer math being per-sample (InstanceNorm statistics are per (n, c)) - SURVEY 8(e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kan_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = O.OracleVGG(3, 10, arch="VGG16_kansmall", dropout_linear=0.0)
    ddp = torch.nn.parallel.DistributedDataParallel(model, gradient_as_bucket_view=True)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (4,), generator=g)
    per = 4 // world
    xs, ys = x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]
    loss = torch.nn.functional.cross_entropy(ddp(xs), ys)
    loss.backward()
    # max over ranks of a per-rank timing-like scalar, as bench.py does for ms_per_step
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({"grads": {k: p.grad.clone() for k, p in model.named_parameters()}, "tmax": float(t)}, out)
    dist.destroy_process_group()


def test_two_rank_data_parallel_matches_single_process(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["tmax"] == 2.0
    torch.manual_seed(0)
    model = O.OracleVGG(3, 10, arch="VGG16_kansmall", dropout_linear=0.0)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (4,), generator=g)
    torch.nn.functional.cross_entropy(model(x), y).backward()
    for k, p in model.named_parameters():
        assert torch.allclose(got["grads"][k], p.grad, rtol=1e-4, atol=1e-6), k
