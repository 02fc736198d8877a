"""torchrun --nproc-per-node 2 tools/ddp_check.py
Data-parallel check on real GPUs (NCCL): with the BCE reconstruction loss (a SUM over pixels) the bucket-averaged
gradients of two ranks, each on half of a batch, equal 0.5 x the single-process gradient of the whole batch (the KL
mean and the L2 term are per-rank means / replicated, which the same identity covers up to the KL's 1/B factor, so
the comparison uses beta = 0 and drops the regulariser).  Also checks that parameters stay identical after 3 steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilistic_domain_adaptation_b200 import ProbabilisticUnet, steps  # noqa: E402
from probabilistic_domain_adaptation_b200.optim import FusedAdam  # noqa: E402
from probabilistic_domain_adaptation_b200.parallel import GradAllReducer, broadcast_parameters, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    model = ProbabilisticUnet(1, 1, [64, 128, 256, 512], 6, 3, 0.0, rl_swap=False).to(dev).train()
    broadcast_parameters(model, 0)
    g = torch.Generator().manual_seed(0)
    B = max(4, 2 * world)     # at least two images per rank
    x = torch.randn(B, 1, 64, 64, generator=g).to(dev)
    y = (torch.rand(B, 1, 64, 64, generator=g) > 0.5).float().to(dev)
    eps = torch.randn(B, 6, generator=g).to(dev)
    a, b = shard_range(B, rank, world)

    def loss_of(m, xs, ys, es):
        m.forward(xs, ys, training=True)
        d = m.posterior_latent_space
        z = d.base_dist.loc + d.base_dist.scale * es
        m.posterior_latent_space.rsample = lambda *aa, **kk: z
        return -m.elbo(ys)

    red = GradAllReducer(model, bucket_mb=8.0)
    loss_of(model, x[a:b], y[a:b], eps[a:b]).backward()
    red.finish()
    mine = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    red.remove()
    for p in model.parameters():
        p.grad = None
    loss_of(model, x, y, eps).backward()
    worst = 0.0
    for k, p in model.named_parameters():
        if p.grad is None or k.startswith("prior."):
            continue  # beta = 0: the prior receives no gradient
        ref = p.grad / world
        err = (mine[k] - ref).norm() / (ref.norm() + 1e-30)
        worst = max(worst, err.item())
    # three optimizer steps keep the replicas bit-identical
    for p in model.parameters():
        p.grad = None
    opt = FusedAdam(model.parameters(), lr=1e-4)
    red = GradAllReducer(model)
    bp = steps.default_backprop(opt, red, model)
    for it in range(3):
        steps.punet_step(model, opt, x[a:b], y[a:b], backprop=bp)
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    same = all(torch.equal(other[0], o) for o in other)
    if rank == 0:
        print(f"ddp_check world={world}: worst relative gradient deviation vs single-process whole batch = {worst:.3e}; "
              f"replicas identical after 3 steps: {same}")
        assert worst < 2e-2 and same
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
