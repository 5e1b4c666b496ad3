"""world_size-2 gloo tests (CPU) of the data-parallel host logic used by bench.py / training:
disjoint scene sharding and the single flat-gradient all-reduce (DDP semantics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from backtoreality_b200 import dist_utils


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv1d(4, 8, 1), torch.nn.ReLU(), torch.nn.Conv1d(8, 3, 1))


def _scene(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(4, 50, generator=g)


def _worker(rank, world, port, out, as_views):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _make_net()
        bucket = dist_utils.FlatGradBucket(net.parameters(), as_views=as_views)
        seeds = dist_utils.shard_scene_indices(8, rank, world, step=3)
        x = torch.stack([_scene(s) for s in seeds])
        net(x).square().mean().backward()   # per-rank mean over its local scenes
        if as_views:
            bucket.allreduce_mean()
        else:                                # gradients assigned by autograd, packed afterwards
            bucket.reduce_from([p.grad for p in net.parameters()])
            for p, v in zip(net.parameters(), bucket.views):
                assert p.grad is v
        torch.save({"seeds": seeds, "flat": bucket.flat.clone()}, out % rank)
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("as_views", [True, False])
def test_two_rank_flat_gradient_allreduce_matches_full_batch(tmp_path, as_views):
    world = 2
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_worker, args=(world, _free_port(), out, as_views), nprocs=world, join=True)
    r0, r1 = torch.load(out % 0), torch.load(out % 1)
    # disjoint shards that cover the global batch of that step
    assert sorted(r0["seeds"] + r1["seeds"]) == list(range(1000 + 3 * 8, 1000 + 4 * 8))
    assert not set(r0["seeds"]) & set(r1["seeds"])
    # identical replicas after the collective
    assert torch.equal(r0["flat"], r1["flat"])
    # and equal to the single-process gradient of the whole global batch
    net = _make_net()
    bucket = dist_utils.FlatGradBucket(net.parameters())
    x = torch.stack([_scene(s) for s in range(1000 + 3 * 8, 1000 + 4 * 8)])
    net(x).square().mean().backward()
    assert torch.allclose(bucket.flat, r0["flat"], rtol=1e-5, atol=1e-7)


def test_shard_validation_and_bucket_views():
    import pytest
    with pytest.raises(ValueError):
        dist_utils.shard_scene_indices(7, 0, 2)
    net = _make_net()
    bucket = dist_utils.FlatGradBucket(net.parameters())
    assert bucket.flat.numel() == sum(p.numel() for p in net.parameters())
    net(torch.randn(2, 4, 9)).sum().backward()
    assert float(bucket.flat.abs().sum()) > 0
    for p in net.parameters():   # grads are views of the flat buffer
        assert p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr()
    bucket.allreduce_mean()      # no process group: no-op
    bucket.zero()
    assert float(bucket.flat.abs().sum()) == 0


def test_pipelined_step_default_caps():
    """PipelinedTrainStep.default_caps (host logic, no GPU): the persistent-grid caps that keep
    the next batch's FPS clusters and the MLP kernels off each other's SMs."""
    from backtoreality_b200.train_step import PipelinedTrainStep
    caps, head = PipelinedTrainStep.default_caps(8, 4, None)      # pre-pass starts with the step
    assert caps == [(116, 0), (116, 0), (116, 140), (116, 140)] and head == (116, 140)
    caps, head = PipelinedTrainStep.default_caps(8, 4, 1)         # ... after SA2's forward
    assert caps == [(0, 140), (0, 140), (116, 116), (116, 116)] and head == (116, 116)
    caps, head = PipelinedTrainStep.default_caps(4, 5, 3)         # GroupFree3D: 4 scenes, 5-CTA FPS
    assert caps == [(0, 144)] * 4 and head == (128, 128)
    caps, _ = PipelinedTrainStep.default_caps(64, 4, 1)           # never below 32 CTAs
    assert caps[2] == (32, 32) and caps[0] == (0, 84)
