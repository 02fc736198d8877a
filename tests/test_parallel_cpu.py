"""CPU, world_size 2 over gloo: the host-side multi-GPU logic (tile sharding, bucketed gradient averaging)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_everything():
    from probabilistic_domain_adaptation_b200.parallel import shard_range, shard_tiles
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_tiles(list(range(10)), rank=1, world=4) == [3, 4, 5]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from probabilistic_domain_adaptation_b200.parallel import GradAllReducer, broadcast_parameters, shard_tiles
    torch.manual_seed(100 + rank)  # different initial weights per rank: broadcast must fix that
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 1))
    broadcast_parameters(net, 0)
    red = GradAllReducer(net, bucket_mb=0.0005)  # ~500 B buckets -> several buckets
    torch.manual_seed(0)
    data = torch.randn(8, 8)
    target = torch.randn(8, 1)
    mine = shard_tiles(list(range(8)), rank, world)
    for it in range(2):  # two iterations: the hooks re-arm
        net.zero_grad()
        loss = ((net(data[mine]) - target[mine]) ** 2).mean()
        loss.backward()
        red.finish()
    out[rank] = {"grads": [p.grad.clone() for p in net.parameters()], "n_buckets": len(red.buckets),
                 "weights": [p.detach().clone() for p in net.parameters()]}
    dist.destroy_process_group()


def test_bucketed_gradient_average_world2():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0]["n_buckets"] > 1
    for a, b in zip(out[0]["weights"], out[1]["weights"]):
        assert torch.equal(a, b)
    for a, b in zip(out[0]["grads"], out[1]["grads"]):
        assert torch.equal(a, b)
    # equals the single-process gradient of the mean over per-rank losses
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 1))
    with torch.no_grad():
        for p, w in zip(net.parameters(), out[0]["weights"]):
            p.copy_(w)
    torch.manual_seed(0)
    data = torch.randn(8, 8)
    target = torch.randn(8, 1)
    loss = 0.5 * (((net(data[:4]) - target[:4]) ** 2).mean() + ((net(data[4:]) - target[4:]) ** 2).mean())
    loss.backward()
    for p, g in zip(net.parameters(), out[0]["grads"]):
        assert torch.allclose(p.grad, g, atol=1e-6)


class _ParkedReg(torch.autograd.Function):
    """A regulariser stand-in: value sum_t c_t * sum(p_t); its backward parks the gradients (training._defer_reg_grads)
    instead of returning them, like training.L2NormSumFn does on the GPU."""

    @staticmethod
    def forward(ctx, coef, *params):
        ctx.coef, ctx.params = coef, params
        return sum(c * p.sum() for c, p in zip(coef, params))

    @staticmethod
    def backward(ctx, g):
        from probabilistic_domain_adaptation_b200 import training
        training._defer_reg_grads(ctx.params, [torch.full_like(p, c) * g for c, p in zip(ctx.coef, ctx.params)])
        return (None,) * (1 + len(ctx.params))


def test_parked_regulariser_gradients_reach_param_grad():
    """Host logic of the bulk regulariser-gradient path (CPU tensors): the parked gradients are added to param.grad at
    the end of backward, or by a GradAllReducer bucket before it is copied -- in both cases param.grad equals data
    gradient + regulariser gradient, repeatedly, and nothing stays parked."""
    from probabilistic_domain_adaptation_b200 import training
    from probabilistic_domain_adaptation_b200.parallel import GradAllReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    params = list(net.parameters())
    coef = [0.5, -1.0, 2.0, 0.25]
    x = torch.randn(4, 6)

    def loss_fn():
        return net(x).square().sum() + 3.0 * _ParkedReg.apply(coef, *params)

    for p in params:
        p.grad = None
    net(x).square().sum().backward()
    want = [p.grad.clone() + 3.0 * c for p, c in zip(params, coef)]
    for with_reducer in (False, True):
        red = GradAllReducer(net, bucket_mb=1e-5) if with_reducer else None
        try:
            for _ in range(2):
                for p in params:
                    p.grad = None
                loss_fn().backward()
                if red is not None:
                    red.finish()
                assert not training._PENDING_REG
                for w, p in zip(want, params):
                    assert torch.allclose(p.grad, w, rtol=1e-6, atol=1e-6)
        finally:
            if red is not None:
                red.remove()
    # a parameter that receives ONLY the regulariser gradient
    extra = torch.nn.Parameter(torch.ones(3))
    (net(x).sum() + _ParkedReg.apply([4.0], extra)).backward()
    assert torch.equal(extra.grad, torch.full((3,), 4.0)) and not training._PENDING_REG
