"""Data-parallel training on the CUDA kernels over NCCL (needs >= 2 GPUs; skipped otherwise): two ranks on half the batch each
reproduce the single-GPU gradients of the full batch - per-sample layer math (InstanceNorm statistics per (n, c)) plus a
gradient all-reduce, which is the whole multi-GPU design."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev):
    import kanconv_b200 as K  # noqa: F401
    from kanconv_b200.models.kan_vgg import vggkan
    torch.manual_seed(0)
    return vggkan(3, 10, arch="VGG16_kansmall", classifier_type="Linear", expected_feature_shape=(1, 1), spline_order=3,
                  grid_size=5, dropout_linear=0.0).to(dev)


def _batch():
    g = torch.Generator().manual_seed(1234)
    return torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    dev = torch.device("cuda", rank)
    model = _build(dev)
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank], gradient_as_bucket_view=True)
    x, y = _batch()
    per = x.shape[0] // world
    xs, ys = x[rank * per:(rank + 1) * per].to(dev), y[rank * per:(rank + 1) * per].to(dev)
    torch.nn.functional.cross_entropy(ddp(xs), ys).backward()
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({k: p.grad.detach().cpu() for k, p in model.named_parameters()}, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_ddp_matches_single_gpu(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    dev = torch.device("cuda", 0)
    model = _build(dev)
    x, y = _batch()
    torch.nn.functional.cross_entropy(model(x.to(dev)), y.to(dev)).backward()
    for k, p in model.named_parameters():
        ref = p.grad.detach().cpu()
        scale = float(ref.abs().max()) or 1.0
        # same kernels, same per-sample results; only the order of the cross-sample sums differs (wgrad split-K vs all-reduce)
        assert float((got[k] - ref).abs().max()) <= 2e-2 * scale, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sharded_batch_gradients_equal_whole_batch_on_one_gpu(precision):
    """The data-parallel property without NCCL (runs on one GPU): the mean of the gradients of the shards of a batch equals the
    gradient of the whole batch, because every layer works per sample (InstanceNorm statistics per (n, c)) and the kernels'
    per-sample results do not depend on the batch they are computed in - tile geometry, CTA pairs, split-K counts and the
    dynamic tile queue only change WHICH CTA computes a value and the order of the fp32 cross-sample sums."""
    import kanconv_b200 as K
    dev = torch.device("cuda", 0)
    old = K.get_precision() if hasattr(K, "get_precision") else None
    K.set_precision(precision)
    try:
        model = _build(dev)
        model.eval()
        x, y = _batch()
        x, y = x.to(dev), y.to(dev)

        def grads(xs, ys):
            model.zero_grad(set_to_none=True)
            torch.nn.functional.cross_entropy(model(xs), ys).backward()
            return [p.grad.detach().double().clone() for p in model.parameters()]

        whole = grads(x, y)
        for shards in (2, 4):
            per = x.shape[0] // shards
            acc = None
            for s in range(shards):
                g = grads(x[s * per:(s + 1) * per], y[s * per:(s + 1) * per])
                acc = g if acc is None else [a + b for a, b in zip(acc, g)]
            num = sum(float((a / shards - w).square().sum()) for a, w in zip(acc, whole))
            den = sum(float(w.square().sum()) for w in whole)
            rel = (num / den) ** 0.5
            print(f"{precision}: {shards} shards of {per}: relative L2 of mean(shard gradients) - whole-batch gradient = {rel:.2e}")
            assert rel < 1e-4, rel
    finally:
        if old is not None:
            K.set_precision(old)
