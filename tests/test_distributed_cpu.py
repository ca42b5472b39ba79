"""CPU, gloo, world_size 2: the N>1 host logic -- contiguous batch sharding and the flat all-reduce of the op's
projection gradients reproduce the single-process full-batch result.  The CPU stand-in for the kernels is the exported
debug function (pure PyTorch); the GPU counterpart of this test is tests/test_multi_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grit_b200.dist_utils import OP_PARAM_NAMES, OpGradBucket, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 257):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_range(n, world, r)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _CpuAttn(torch.nn.Module):
    """MSDeformAttn with the core op swapped for the pure-PyTorch debug function, so it runs without a GPU."""

    def __init__(self, **kw):
        super().__init__()
        import grit_b200
        from grit_b200.ops.modules import ms_deform_attn as modfile
        self.inner = grit_b200.MSDeformAttn(**kw)
        self._modfile = modfile

    def forward(self, *args):
        import grit_b200

        class _Fn:
            @staticmethod
            def apply(value, shapes, lsi, loc, attn, step):
                return grit_b200.ms_deform_attn_core_pytorch(value, shapes, loc, attn)
        saved = self._modfile.MSDeformAttnFunction
        self._modfile.MSDeformAttnFunction = _Fn
        try:
            return self.inner(*args)
        finally:
            self._modfile.MSDeformAttnFunction = saved


def _problem():
    torch.manual_seed(0)
    N, Lq, C, M, L, P = 4, 6, 16, 2, 2, 2
    shapes = torch.tensor([[4, 5], [2, 3]])
    lsi = torch.tensor([0, 20])
    S = 26
    mods = [_CpuAttn(d_model=C, n_levels=L, n_heads=M, n_points=P).double() for _ in range(2)]
    for m in mods:
        with torch.no_grad():
            m.inner.sampling_offsets.weight.normal_(0, 0.05)
            m.inner.attention_weights.weight.normal_(0, 0.3)
    query = torch.randn(N, Lq, C, dtype=torch.float64)
    src = torch.randn(N, S, C, dtype=torch.float64)
    ref = torch.rand(N, Lq, L, 2, dtype=torch.float64)
    return mods, query, src, ref, shapes, lsi


def _loss(mods, query, src, ref, shapes, lsi):
    x = query
    for m in mods:
        x = x + m(x, ref, src, shapes, lsi)
    return (x ** 2).sum()


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mods, query, src, ref, shapes, lsi = _problem()  # same seed on every rank: identical replicas
        mine = list(shard_range(query.shape[0], world, rank))
        loss = _loss(mods, query[mine], src[mine], ref[mine], shapes, lsi)
        loss.backward()
        bucket = OpGradBucket([m.inner for m in mods])
        work = bucket.all_reduce_async()
        bucket.finish(work)
        grads = {f"{i}.{n}": dict(m.inner.named_parameters())[n].grad.numpy() for i, m in enumerate(mods)
                 for n in OP_PARAM_NAMES}
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **grads)
    finally:
        dist.destroy_process_group()


def test_sharded_backward_plus_bucket_allreduce_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mods, query, src, ref, shapes, lsi = _problem()
    _loss(mods, query, src, ref, shapes, lsi).backward()
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{rank}.npz"))
        for i, m in enumerate(mods):
            for n in OP_PARAM_NAMES:
                full = dict(m.inner.named_parameters())[n].grad.numpy()
                # sum over shards / world  ==  full-batch gradient / world  (the loss is a sum over images)
                np.testing.assert_allclose(got[f"{i}.{n}"] * world, full, rtol=1e-10, atol=1e-12)
