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


def _worker(rank: int, world: int, port: int, outdir: str, overlap: bool = False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ponderv2_b200.dist import FlatParameters, broadcast_parameters
        torch.set_num_threads(1)
        model = _tiny_model(100 + rank)            # deliberately different replicas: the broadcast must fix that
        # buffer laid out in backward-completion order (last layer first), cut into 3 slices for the overlapped reduce
        flat = FlatParameters(model, order=list(model.parameters())[::-1], num_chunks=3)
        n_params = sum(p.numel() for p in model.parameters())
        assert flat.flat_param.numel() >= n_params and flat.flat_grad.numel() == flat.flat_param.numel()
        assert 2 <= len(flat.chunks) <= 3 and flat.chunks[0][0] == 0 and flat.chunks[-1][1] == flat.flat_grad.numel()
        assert all(flat.chunks[i][1] == flat.chunks[i + 1][0] for i in range(len(flat.chunks) - 1))
        for p in model.parameters():               # parameters and gradients are views of the flat buffers, 16 B aligned
            assert p.data.untyped_storage().data_ptr() == flat.flat_param.untyped_storage().data_ptr()
            assert p.grad.untyped_storage().data_ptr() == flat.flat_grad.untyped_storage().data_ptr()
            assert p.data.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0
        broadcast_parameters(flat, src=0)
        for b0 in model.buffers():                 # BatchNorm buffers are broadcast too
            t = b0.detach().clone().float()
            dist.all_reduce(t)
            assert torch.allclose(t, b0.float() * world)
        opt = flat.make_optimizer(torch.optim.SGD, lr=0.1, momentum=0.9)
        if overlap:
            flat.enable_overlap()

        torch.manual_seed(7 + rank)                # one "scene" per rank (weak scaling: per-rank work is fixed)
        x, y = torch.randn(32, 6), torch.randn(32, 3)
        for _ in range(2):
            opt.zero_grad()                        # the reference trainer's idiom (engines/train.py:186)
            loss = (model(x) - y).square().mean()
            if overlap:
                # local gradient for the check: a second, hook-free backward on a copy
                import copy
                m2 = copy.deepcopy(model)
                for q in m2.parameters():
                    q.grad = None
                (m2(x) - y).square().mean().backward()
                local_list = [q.grad.detach().clone() for q in m2.parameters()]
                # m2's BatchNorm update must not count twice: restore is not needed, m2 is a throwaway copy
                loss.backward()                    # slices go out from the gradient hooks while backward runs
                assert all(flat._launched), flat._launched
                flat.all_reduce_mean()
                local = torch.zeros_like(flat.flat_grad)
                for p0, g0 in zip(model.parameters(), local_list):
                    i = next(j for j, q in enumerate(flat.params) if q is p0)
                    local[flat.offsets[i]:flat.offsets[i] + g0.numel()] = g0.reshape(-1)
            else:
                loss.backward()
                local = flat.flat_grad.clone()
                flat.all_reduce_mean()
            opt.step()
        torch.save((flat.flat_param.clone(), local, flat.flat_grad.clone(), x, y), os.path.join(outdir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("overlap", [False, True])
def test_flat_gradient_allreduce_world2(tmp_path, overlap):
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), overlap)) for r in range(world)]
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
    flats = [FlatParameters(m, order=list(m.parameters())[::-1], num_chunks=3) for m in models]
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


def test_flat_parameters_survive_zero_grad_idioms():
    """optimizer.zero_grad(set_to_none=True) and model.zero_grad() (engines/train.py:186) must neither drop the master
    gradient nor let gradients accumulate across steps; the flat step equals a plain per-parameter SGD step."""
    from ponderv2_b200.dist import FlatParameters
    torch.manual_seed(3)
    x, y = torch.randn(16, 6), torch.randn(16, 3)
    ref = _tiny_model(5)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-2)
    m = _tiny_model(5)
    flat = FlatParameters(m)
    opt = flat.make_optimizer(torch.optim.SGD, lr=0.05, momentum=0.9, weight_decay=1e-2)
    for step in range(3):
        ref_opt.zero_grad()
        (ref(x) - y).square().mean().backward()
        ref_opt.step()
        if step == 1:
            m.zero_grad()                 # drops the per-parameter views (set_to_none=True)
        else:
            opt.zero_grad(set_to_none=True)
        (m(x) - y).square().mean().backward()
        flat.all_reduce_mean()
        opt.step()
        assert flat.master.grad is flat.flat_grad
        for p in m.parameters():
            assert p.grad.untyped_storage().data_ptr() == flat.flat_grad.untyped_storage().data_ptr()
    for a, b in zip(ref.parameters(), m.parameters()):
        assert torch.allclose(a, b, rtol=0, atol=1e-6)


def test_flat_parameters_freeze_untouched():
    """A parameter no loss reaches keeps its value (reference: grad None under find_unused_parameters -> no decay)."""
    from ponderv2_b200.dist import FlatParameters

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.used = nn.Linear(4, 2)
            self.unused = nn.Parameter(torch.ones(3))

        def forward(self, x):
            return self.used(x)

    m = M()
    flat = FlatParameters(m)
    flat.freeze_untouched(["unused"])
    opt = flat.make_optimizer(torch.optim.SGD, lr=0.1, momentum=0.9, weight_decay=0.5)
    w0 = m.used.weight.detach().clone()
    for _ in range(2):
        opt.zero_grad()
        m(torch.randn(5, 4)).sum().backward()
        opt.step()
    assert torch.equal(m.unused.detach(), torch.ones(3)) and not torch.equal(m.used.weight.detach(), w0)
