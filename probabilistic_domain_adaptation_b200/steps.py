"""Step bodies of the reference trainers on the fused kernels.

Each function is the per-batch body of one `_train_epoch_impl`, with the same order of operations, the same
loss and the same in-place effects (optimizer step, EMA), minus logging / progress / checkpointing (torch_em):

  punet_step         prob_utils/my_trainer/punet_trainer.py:24-36
  mean_teacher_step  prob_utils/my_trainer/mean_teacher_trainer.py:101-131
  fixmatch_step      prob_utils/my_trainer/fixmatch_trainer.py:67-95
  adamt_step         prob_utils/my_trainer/adamt_trainer.py:89-128
  adamatch_step      prob_utils/my_trainer/adamatch_trainer.py:62-102

`backprop(loss)` is torch_em's callable (backward + optimizer.step, with GradScaler under AMP); the default does
loss.backward(); [grad all-reduce]; optimizer.step().
"""
import os

import torch

from . import consensus
from .autograd_ops import refresh_packed
from .my_models.utils import l2_regularisation


def default_backprop(optimizer, reducer=None, model=None):
    """backward + [gradient all-reduce] + optimizer step (+ one-launch refresh of the packed bf16 conv operands)."""
    def backprop(loss):
        loss.backward()
        if reducer is not None:
            reducer.finish()
        optimizer.step()
        if model is not None:
            refresh_packed(model, rot180=True)
            if hasattr(model, "release_graph"):
                model.release_graph()  # no cached tensor keeps this step's autograd graph alive (see GraphedStep)
    return backprop


def punet_loss(model, x, y, consm=None, use_consm=False):
    """forward + elbo + 1e-5 * L2(posterior, prior, fcomb.layers)  (punet_trainer.py:30-34)."""
    model.forward(x, y, training=True)
    elbo = model.elbo(y, consm) if use_consm else model.elbo(y)
    reg_loss = l2_regularisation(model.posterior) + l2_regularisation(model.prior) + \
        l2_regularisation(model.fcomb.layers)
    return -elbo + 1e-5 * reg_loss


def punet_step(model, optimizer, x, y, backprop=None):
    backprop = backprop or default_backprop(optimizer)
    optimizer.zero_grad()
    loss = punet_loss(model, x, y)
    backprop(loss)
    return loss.detach()  # detached: a live loss keeps the autograd graph (and its AccumulateGrad nodes) alive


def mean_teacher_step(model, teacher, optimizer, ema, x1, x2, n_samples=16, do_consensus_masking=False,
                      momentum=0.999, backprop=None, eps=None):
    """Teacher MC pseudo-label + consensus on the weak view, student ELBO step on the strong view, EMA.
    `ema` is a consensus.MomentumUpdater(model, teacher)."""
    backprop = backprop or default_backprop(optimizer)
    with torch.no_grad():
        y, z = consensus.sample_from_teacher(teacher, x1, n_samples, do_consensus_masking=do_consensus_masking,
                                             eps=eps)
    optimizer.zero_grad()
    loss = punet_loss(model, x2, y, z, use_consm=True)
    backprop(loss)
    lr = optimizer.param_groups[0]["lr"]
    if lr:  # mean_teacher_trainer.py:126 -- always true for a positive learning rate
        ema.step(momentum)
    return loss.detach(), y, z


@torch.no_grad()
def validation_step(model, x, y, n_samples=8):
    """Per-batch body of PUNetTrainer._validate_impl (punet_trainer.py:62-86): forward(training=True) + ELBO + L2,
    mean of n_samples sigmoid(sample(testing=False)), dice_score(mean, gt).  Returns (loss, dice, 1 - dice) as 0-dim
    DEVICE tensors: the reference copies the full-resolution prediction to the host for every validation batch; here
    nothing synchronises until the caller reads the three scalars."""
    from . import ops
    model.forward(x, y, training=True)
    elbo = model.elbo(y)
    reg_loss = l2_regularisation(model.posterior) + l2_regularisation(model.prior) + \
        l2_regularisation(model.fcomb.layers)
    loss = -elbo + 1e-5 * reg_loss
    pred = consensus.sample_from_model(model, n_samples)
    dice = ops.dice_score(pred, y)
    return loss, dice, 1.0 - dice


def distribution_alignment(y, source_distribution):
    """fixmatch_trainer.py:77-84 on the device: class frequencies of the binarised pseudo-label (one counting pass instead
    of torch.unique's sort + host sync), ratio = source / target, rescaled and clipped pseudo-label.  Returns
    (y_aligned, ratio (2,) device tensor).  `source_distribution`: [background, foreground] frequencies (list or tensor)."""
    from . import _lib, ops
    ops._need_cuda(y)
    y = y.contiguous()
    if not torch.is_tensor(source_distribution) or not source_distribution.is_cuda:
        source_distribution = torch.as_tensor(source_distribution, dtype=torch.float32).to(y.device)
    source_distribution = source_distribution.to(torch.float32).contiguous()
    out = torch.empty_like(y)
    ratio = torch.empty(2, dtype=torch.float32, device=y.device)
    scratch = torch.empty(1, dtype=torch.int64, device=y.device)
    _lib.check(_lib.load().pda_distribution_alignment(y.data_ptr(), y.numel(), source_distribution.data_ptr(),
                                                      scratch.data_ptr(), out.data_ptr(), ratio.data_ptr(),
                                                      ops._stream()), "distribution_alignment")
    return out, ratio


def fixmatch_step(model, optimizer, x1, x2, n_samples=16, do_consensus_masking=False, source_distribution=None,
                  backprop=None, eps=None):
    """Weak-view MC pseudo-label + consensus from the model itself, ELBO step on the strong view."""
    backprop = backprop or default_backprop(optimizer)
    with torch.no_grad():
        y, z = consensus.sample_from_weak_model(model, x1, n_samples, do_consensus_masking=do_consensus_masking,
                                                eps=eps)
    y, z = y.detach(), z.detach()
    ratio = None
    if source_distribution is not None:
        y, ratio = distribution_alignment(y, source_distribution)
    optimizer.zero_grad()
    loss = punet_loss(model, x2, y, z, use_consm=True)
    backprop(loss)
    return loss.detach(), y, z, ratio


# PDA_CONCURRENT_BRANCHES (default 1; 0 = everything on one stream): the source branch of the joint steps (forward + ELBO of the labelled batch) is enqueued on
# a side stream and runs NEXT TO the target branch (weak-view MC pseudo-labels + strong-view forward); autograd then runs
# each branch's backward on its own stream.  For small batches (LIVECell: 2 + 2 images of 256 x 256) most layers fill
# only a fraction of the 148 SMs, so the two branches overlap on the device: LIVECell joint step 8.45 -> 6.5 ms, MitoEM
# (4 + 4 x 512^2) 25.8 -> 24.7 ms as CUDA-graph replays.
CONCURRENT_BRANCHES = os.environ.get("PDA_CONCURRENT_BRANCHES", "1") == "1"
_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
        # the parameters' AccumulateGrad nodes now receive gradients from two streams: intended (the engine inserts the
        # synchronisation), so torch's once-per-process warning about it is switched off
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
    return _SIDE_STREAMS[key]


def _joint_step(model, pseudo_net, optimizer, xs, ys, xt1, xt2, n_samples, do_consensus_masking, backprop, eps):
    optimizer.zero_grad()
    side = None
    if CONCURRENT_BRANCHES and xs.is_cuda:
        main = torch.cuda.current_stream(xs.device)
        side = _side_stream(xs.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            supervised_loss = punet_loss(model, xs, ys)
    else:
        supervised_loss = punet_loss(model, xs, ys)
    with torch.no_grad():
        y, z = consensus.sample_from_teacher(pseudo_net, xt1, n_samples, do_consensus_masking=do_consensus_masking,
                                             eps=eps)
    y, z = y.detach(), z.detach()
    target_loss = punet_loss(model, xt2, y, z, use_consm=True)
    if side is not None:
        main.wait_stream(side)
        supervised_loss.record_stream(main)
    loss = (supervised_loss + target_loss) / 2
    backprop(loss)
    return loss.detach(), y, z


def adamatch_step(model, optimizer, xs, ys, xt1, xt2, n_samples=16, do_consensus_masking=False, backprop=None,
                  eps=None):
    """Joint FixMatch: source ELBO + weak-view pseudo-labelled target ELBO, one backward."""
    backprop = backprop or default_backprop(optimizer)
    return _joint_step(model, model, optimizer, xs, ys, xt1, xt2, n_samples, do_consensus_masking, backprop, eps)


def adamt_step(model, teacher, optimizer, ema, iteration, xs, ys, xt1, xt2, n_samples=16,
               do_consensus_masking=False, momentum=0.999, backprop=None, eps=None):
    """Joint mean teacher: source ELBO + teacher pseudo-labelled target ELBO, one backward, warm-up EMA."""
    backprop = backprop or default_backprop(optimizer)
    out = _joint_step(model, teacher, optimizer, xs, ys, xt1, xt2, n_samples, do_consensus_masking, backprop, eps)
    if torch.is_tensor(iteration):  # device-side iteration counter (advanced by the kernel): graph-capturable
        ema.step(momentum, iteration_dev=iteration)
    else:
        ema.step(consensus.adamt_momentum(iteration, momentum))
    return out


class GraphedStep:
    """One whole step body captured into a CUDA graph and replayed: no Python, ctypes, autograd or allocator work per
    step, which is what bounds the small-batch configurations (LIVECell joint training feeds 2 + 2 images of 256 x 256:
    ~500 launches for ~4 ms of device work).  The reference has nothing comparable (eager PyTorch, `compile_model=False`
    in every script); this is the B200-side answer to launch-bound inner loops.

        opt = FusedAdam(model.parameters(), lr=1e-5, capturable=True)
        reducer = GradAllReducer(model)                       # also for one GPU: keeps p.grad storage fixed
        bp = default_backprop(opt, reducer, model)
        step = GraphedStep(lambda x1, x2: mean_teacher_step(model, teacher, opt, ema, x1, x2, backprop=bp)[0],
                           (x1, x2), optimizer=opt)
        loss = step(x1, x2)                                   # copies the batch into the static inputs, replays

    Rules (torch.cuda.graph's): fixed shapes; no tensor of an earlier eager step may still hold that step's autograd
    graph (its AccumulateGrad nodes are tied to the stream they were built on; the step bodies here return detached
    losses and call `model.release_graph()` for this reason -- pass models stepped otherwise as `modules=`); `fn` must
    not synchronise with the host (no .item(); FixMatch's distribution alignment runs as a device kernel here for that
    reason, not through torch.unique); host-side scalars are frozen at capture (beta, a constant EMA momentum) except
    the Adam step count / learning rate (FusedAdam(capturable=True)) and AdaMT's warm-up momentum (pass the iteration
    to adamt_step as an int64 device tensor), which live on the device.  Random draws inside `fn` (latent samples) advance per replay (torch's graph-safe generator).
    The `warmup` eager calls on the first batch are REAL training steps."""

    def __init__(self, fn, example_inputs, optimizer=None, warmup=3, modules=()):
        self.fn, self.optimizer = fn, optimizer
        if optimizer is not None and not getattr(optimizer, "capturable", False):
            raise ValueError("GraphedStep needs FusedAdam(capturable=True): a replay must not bake in the step count")
        # (a step body whose optimizer.zero_grad() drops the gradients needs a GradAllReducer -- it keeps p.grad storage
        # fixed; FusedAdam raises a PdaError naming this if its pointer table would have to be rebuilt under capture)
        self.static_in = [x.clone() for x in example_inputs]
        for m in modules:
            m.release_graph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst is not src:
                dst.copy_(src, non_blocking=True)
        if self.optimizer is not None:
            self.optimizer.sync_lr()
        self.graph.replay()
        return self.static_out
