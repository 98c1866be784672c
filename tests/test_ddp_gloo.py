"""The N > 1 path on CPU: world_size-2 gloo processes all-reducing the flat gradient buffer of a ParamStore,
plus batch sharding. (On GPUs the same class runs over NCCL; bench.py uses it for --gpus N.)"""
import json
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, golden_dir, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, ROOT)
    from vilbert_b200.config import BertConfig
    from vilbert_b200.ddp import FlatGradAllReducer, shard_batch
    from vilbert_b200.engine import ParamStore
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = BertConfig.from_dict(json.load(open(os.path.join(golden_dir, "tiny_b4.json")))["config"])
    ps = ParamStore(cfg, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    ps.grad.copy_(torch.randn(ps.numel, generator=g))
    ps.flat.copy_(torch.randn(ps.numel, generator=g))
    red = FlatGradAllReducer(ps.grad, n_buckets=5)
    assert sum(b.numel() for b in red.buckets) == ps.numel
    red.broadcast_params(ps.flat)
    red.allreduce()
    expect = sum(torch.randn(ps.numel, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    ok_grad = torch.allclose(ps.grad, expect, atol=1e-6)
    g0 = torch.Generator().manual_seed(100)
    torch.randn(ps.numel, generator=g0)
    ok_param = torch.equal(ps.flat, torch.randn(ps.numel, generator=g0))
    # a named view sees the reduced values (Parameters' .grad are views of the flat buffer)
    name = "bert.encoder.c_layer.0.biOutput.dense1.weight"
    off, shape = ps.entries[name]
    ok_view = torch.equal(ps.g(name).flatten(), ps.grad[off:off + ps.g(name).numel()])
    start, per = shard_batch(512, rank, world)
    out[rank] = (ok_grad, ok_param, ok_view, start, per)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2(golden_dir):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, golden_dir, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok_grad, ok_param, ok_view, start, per = out[rank]
        assert ok_grad and ok_param and ok_view
        assert (start, per) == (rank * 256, 256)
