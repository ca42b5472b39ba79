"""GPU, >= 2 devices, NCCL: batch-sharded replicas give bit-identical per-image results and the flat-bucket all-reduce
of the op's projection gradients equals the full-batch gradient.  Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

from . import helpers

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import grit_b200
    from grit_b200 import _lib
    from grit_b200.dist_utils import OpGradBucket, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        case = helpers.make_inputs(4, 300, 8, 32, [(20, 30), (10, 15), (5, 8), (3, 4)], 4, seed=5, dtype=np.float32)
        mine = list(shard_range(4, world, rank))
        sl = slice(mine[0], mine[-1] + 1)
        t = helpers.to_cuda({k: (v[sl] if k in ("value", "loc", "attn", "grad_out") else v) for k, v in case.items()},
                            torch.float32, device=f"cuda:{rank}")
        out = _lib.forward(t["value"], t["shapes"], t["level_start"], t["loc"], t["attn"])
        gv, gl, ga = _lib.backward(t["value"], t["shapes"], t["level_start"], t["loc"], t["attn"], t["grad_out"],
                                   _lib.FLAG_DETERMINISTIC)
        # module replicas + bucket all-reduce
        torch.manual_seed(1)
        mod = grit_b200.MSDeformAttn(256, 4, 8, 4).cuda()
        with torch.no_grad():
            mod.sampling_offsets.weight.normal_(0, 0.02)
            mod.attention_weights.weight.normal_(0, 0.2)
        gen = torch.Generator().manual_seed(9)
        query = torch.randn(4, 300, 256, generator=gen)[sl].cuda()
        src = torch.randn(4, t["value"].shape[1], 256, generator=gen)[sl].cuda()
        ref = torch.rand(4, 300, 4, 2, generator=gen)[sl].cuda()
        grit_b200.set_deterministic(True)
        y = mod(query, ref, src, t["shapes"], t["level_start"])
        (y ** 2).sum().backward()
        bucket = OpGradBucket([mod])
        bucket.finish(bucket.all_reduce_async())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), out=out.cpu().numpy(), gv=gv.cpu().numpy(),
                 gl=gl.cpu().numpy(), ga=ga.cpu().numpy(), first=mine[0], last=mine[-1],
                 **{"p." + n: p.grad.cpu().numpy() for n, p in mod.named_parameters()})
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_sharding_and_gradient_allreduce(tmp_path):
    import torch.multiprocessing as mp

    import grit_b200
    from grit_b200 import _lib
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # single-GPU full batch
    case = helpers.make_inputs(4, 300, 8, 32, [(20, 30), (10, 15), (5, 8), (3, 4)], 4, seed=5, dtype=np.float32)
    t = helpers.to_cuda(case, torch.float32)
    out = _lib.forward(t["value"], t["shapes"], t["level_start"], t["loc"], t["attn"])
    gv, gl, ga = _lib.backward(t["value"], t["shapes"], t["level_start"], t["loc"], t["attn"], t["grad_out"],
                               _lib.FLAG_DETERMINISTIC)
    torch.manual_seed(1)
    mod = grit_b200.MSDeformAttn(256, 4, 8, 4).cuda()
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.attention_weights.weight.normal_(0, 0.2)
    gen = torch.Generator().manual_seed(9)
    query = torch.randn(4, 300, 256, generator=gen).cuda()
    src = torch.randn(4, t["value"].shape[1], 256, generator=gen).cuda()
    ref = torch.rand(4, 300, 4, 2, generator=gen).cuda()
    prev = grit_b200.set_deterministic(True)
    try:
        (mod(query, ref, src, t["shapes"], t["level_start"]) ** 2).sum().backward()
    finally:
        grit_b200.set_deterministic(prev)
    for rank in range(world):
        z = np.load(os.path.join(tmp_path, f"rank{rank}.npz"))
        sl = slice(int(z["first"]), int(z["last"]) + 1)
        # images are independent: the shard's results are bit-identical to the same images inside the full batch
        assert np.array_equal(z["out"], out[sl].cpu().numpy())
        assert np.array_equal(z["gv"], gv[sl].cpu().numpy())
        assert np.array_equal(z["gl"], gl[sl].cpu().numpy()) and np.array_equal(z["ga"], ga[sl].cpu().numpy())
        for n, p in mod.named_parameters():
            full = p.grad.cpu().numpy()
            got = z["p." + n] * world  # bucket averages; the loss is a sum over images
            assert np.abs(got - full).max() <= 2e-4 * max(np.abs(full).max(), 1e-6), n
