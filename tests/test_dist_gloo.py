"""world_size-2 gloo test (CPU) of the data-parallel plumbing: the batch shards by rank, nothing is
exchanged on the data path, and ONE all-reduce of the flat weight-gradient buffer reproduces the
full-batch gradients.  Per-rank gradients come from the oracle (the CUDA path needs a GPU); what is
under test is mga_yolo_b200.dist (shard_range + FlatGradReducer)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mga_yolo_b200 import FlatGradReducer, MaskGuidedCBAM, shard_range
        from oracle import cbam_oracle as co
        from tests._golden import PARAM_KEYS

        torch.manual_seed(0)
        B, C, H, W = 6, 16, 6, 6
        x, mk, g = torch.randn(B, C, H, W), torch.randn(B, 1, H, W), torch.randn(B, C, H, W)
        levels = [MaskGuidedCBAM(C, r=4), MaskGuidedCBAM(C, r=8)]  # same seed on every rank -> identical replicas
        reducer = FlatGradReducer([p for m in levels for p in m.parameters()], average=False)
        assert reducer.flat.numel() == sum(p.numel() for m in levels for p in m.parameters())
        mine = shard_range(B, rank, world)
        sl = slice(mine.start, mine.stop)
        reducer.zero()
        for m in levels:
            p = co.CbamParams.from_state_dict(m.state_dict())
            out, sv = co.cbam_forward(x[sl], mk[sl], p)
            grads = co.cbam_backward(g[sl], p, sv)
            for name, prm in m.named_parameters():
                prm.grad.copy_(grads[name])  # .grad is a view into the flat buffer
        work = reducer.all_reduce(async_op=True)
        reducer.finish(work)
        # full-batch reference on every rank
        for m in levels:
            p = co.CbamParams.from_state_dict(m.state_dict())
            out, sv = co.cbam_forward(x, mk, p)
            full = co.cbam_backward(g, p, sv)
            for name, prm in m.named_parameters():
                assert torch.allclose(prm.grad, full[name], rtol=1e-4, atol=1e-5), (rank, name)
        # averaging variant (what DDP does) + synchronous call
        red2 = FlatGradReducer(levels[0].parameters(), average=True)
        red2.flat.fill_(float(rank + 1))
        red2.all_reduce()
        assert torch.allclose(red2.flat, torch.full_like(red2.flat, (1 + world) / 2))
        torch.save(torch.tensor(1), os.path.join(tmp, f"ok{rank}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_flat_grad_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
