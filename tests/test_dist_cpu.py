"""World-size-2 `gloo` tests (CPU) of the N > 1 path: scenes shard one per rank with no data-path collective; the only
exchange is ONE all-reduce of the flat gradient buffer (ponderv2_b200/dist.py; reference: DDP in engines/defaults.py:22-43,
train.py:212-216).  The product kernels need a GPU, so the model here is a small torch module: what is under test is the
host logic (flat re-homing of parameters, broadcast, mean all-reduce, one optimizer step keeping replicas identical).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tiny_model(seed: int) -> nn.Module:
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(6, 16), nn.ReLU(), nn.Linear(16, 16), nn.BatchNorm1d(16), nn.Linear(16, 3))


def _worker(rank: int, world: int, port: int, outdir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ponderv2_b200.dist import FlatParameters, broadcast_parameters
        torch.set_num_threads(1)
        model = _tiny_model(100 + rank)            # deliberately different replicas: the broadcast must fix that
        flat = FlatParameters(model)
        n_params = sum(p.numel() for p in model.parameters())
        assert flat.flat_param.numel() == n_params and flat.flat_grad.numel() == n_params
        for p in model.parameters():               # parameters and gradients are views of the flat buffers
            assert p.data.untyped_storage().data_ptr() == flat.flat_param.untyped_storage().data_ptr()
            assert p.grad.untyped_storage().data_ptr() == flat.flat_grad.untyped_storage().data_ptr()
        broadcast_parameters(flat, src=0)
        opt = torch.optim.SGD(flat.optimizer_params(), lr=0.1, momentum=0.9)

        torch.manual_seed(7 + rank)                # one "scene" per rank (weak scaling: per-rank work is fixed)
        x, y = torch.randn(32, 6), torch.randn(32, 3)
        for _ in range(2):
            flat.zero_grad()
            loss = (model(x) - y).square().mean()
            loss.backward()
            local = flat.flat_grad.clone()
            flat.all_reduce_mean()
            opt.step()
        torch.save((flat.flat_param.clone(), local, flat.flat_grad.clone(), x, y), os.path.join(outdir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_allreduce_world2(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=100)
        assert p.exitcode == 0
    got = {r: torch.load(tmp_path / f"rank{r}.pt") for r in range(world)}
    # replicas identical after two steps, reduced gradient = mean of the two local gradients (last step)
    assert torch.equal(got[0][0], got[1][0])
    mean = (got[0][1] + got[1][1]) / 2
    assert torch.allclose(got[0][2], mean, rtol=0, atol=1e-7) and torch.equal(got[0][2], got[1][2])

    # single-process restatement of the same two steps: per-rank BN statistics (sync_bn=False in the reference), averaged
    # gradients, one SGD step on the shared weights
    from ponderv2_b200.dist import FlatParameters
    models = [_tiny_model(100), _tiny_model(100)]
    flats = [FlatParameters(m) for m in models]
    opts = [torch.optim.SGD(f.optimizer_params(), lr=0.1, momentum=0.9) for f in flats]
    for _ in range(2):
        for r in range(2):
            flats[r].zero_grad()
            (models[r](got[r][3]) - got[r][4]).square().mean().backward()
        g = (flats[0].flat_grad + flats[1].flat_grad) / 2
        for r in range(2):
            flats[r].flat_grad.copy_(g)
            opts[r].step()
    assert torch.allclose(flats[0].flat_param, got[0][0], rtol=0, atol=1e-6)


def test_flat_parameters_single_process_noop_collective():
    """Without an initialised process group (N = 1) the collective is a no-op and the optimizer sees one tensor."""
    from ponderv2_b200.dist import FlatParameters, broadcast_parameters
    m = _tiny_model(0)
    ref = [p.detach().clone() for p in m.parameters()]
    flat = FlatParameters(m)
    broadcast_parameters(flat)
    for a, b in zip(ref, m.parameters()):
        assert torch.equal(a, b.detach())
    m(torch.randn(8, 6)).sum().backward()
    g = flat.flat_grad.clone()
    flat.all_reduce_mean()
    assert torch.equal(g, flat.flat_grad)
    assert len(list(flat.optimizer_params())) == 1
