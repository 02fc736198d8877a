"""Consensus weighting / masking and mean-teacher entry points on the fused kernels.

Drop-ins for the helper bodies that the reference duplicates in its trainers and prediction functions:
  sample_from_teacher / sample_from_weak_model   mean_teacher_trainer.py:72-88, adamt_trainer.py:60-76,
                                                 fixmatch_trainer.py:37-54, adamatch_trainer.py:33-49
  sample_from_model                              mean_teacher_trainer.py:90-93
  _momentum_update                               mean_teacher_trainer.py:52-55, adamt_trainer.py:40-43
  _custom_punet_prediction                       punet_predictions.py:29-33
  pseudo-label + consensus mask                  punet_predictions.py:104-124
(paths relative to /root/reference/prob_utils).
"""
import torch

from . import ops


@torch.no_grad()
def sample_from_teacher(net, inputs, n_samples=16, upper_thres=0.9, lower_thres=0.1, do_consensus_masking=False,
                        eps=None):
    """net.forward(inputs, None, training=False); n_samples x sigmoid(net.sample()); mean and consensus.
    Returns (samples, consensus): (B,1,H,W) fp32 pseudo-label and fp32 k/n weight, or int64 {0,1} mask."""
    net.forward(inputs, None, training=False)
    return net.mc_consensus(n_samples, eps=eps, upper_thres=upper_thres, lower_thres=lower_thres,
                            do_consensus_masking=do_consensus_masking)


sample_from_weak_model = sample_from_teacher


@torch.no_grad()
def sample_from_model(net, n_samples=16, eps=None):
    """Mean of n_samples sigmoid samples of an already-forwarded net (logging / validation helper)."""
    mean, _ = net.mc_consensus(n_samples, eps=eps, want_consensus=False)
    return mean


@torch.no_grad()
def punet_mc_prediction(model, raw, prior_samples=8, eps=None):
    """The per-tile closure of punet_prediction: forward, prior_samples x sigmoid(sample(testing=True)), mean."""
    model.forward(raw, None, training=False)
    mean, _ = model.mc_consensus(prior_samples, eps=eps, testing=True, want_consensus=False)
    return mean


@torch.no_grad()
def punet_pseudo_labels(model, patch, prior_samples=8, upper_threshold=0.9, lower_threshold=0.1, eps=None):
    """Whole-image pseudo-label + consensus MASK (punet_predictions.py:104-124) -> (mean fp32, mask uint8)."""
    model.forward(patch, None, training=False)
    mean, mask = model.mc_consensus(prior_samples, eps=eps, testing=True, upper_thres=upper_threshold,
                                    lower_thres=lower_threshold, do_consensus_masking=True)
    return mean, mask.to(torch.uint8)


class MomentumUpdater:
    """teacher = teacher * m + student * (1 - m) over all parameters in ONE kernel launch.
    Keeps a device-side pointer table; rebuilt if any parameter storage moved."""

    def __init__(self, model, teacher):
        self.model, self.teacher = model, teacher
        self._table, self._key = None, None

    def _refresh(self):
        tp = [p.data for p in self.teacher.parameters()]
        sp = [p.data for p in self.model.parameters()]
        key = tuple(t.data_ptr() for t in tp) + tuple(s.data_ptr() for s in sp)
        if key != self._key:
            self._table = ops.build_ema_table(tp, sp)
            self._key = key

    @torch.no_grad()
    def step(self, momentum, iteration_dev=None):
        """iteration_dev: int64 device scalar -> AdaMT warm-up momentum computed (and the counter advanced) on the device,
        which keeps a graph-captured AdaMT step replayable."""
        self._refresh()
        ops.multi_tensor_ema(self._table, momentum, iteration_dev)
        # parameter memory changed behind autograd's back: advance the version counters (packed-weight cache key)
        from .optim import bump_versions
        from .autograd_ops import refresh_packed
        bump_versions(self.teacher.parameters())
        refresh_packed(self.teacher, rot180=False, bf16=False, f16=True)  # the teacher only runs the no-grad forward


def adamt_momentum(iteration, momentum=0.999):
    """adamt_trainer.py:41 warm-up schedule."""
    return min(1 - 1 / (iteration + 1), momentum)


class HostPredictor:
    """End-to-end Monte-Carlo prediction with HOST buffers, pipelined: the host->device copy of step i+1 and the
    device->host copy of step i run on their own streams while step i / i+1 computes (the prediction driver of
    punet_predictions.py:35-63 feeds one tile batch after the other).  `submit` enqueues one batch; `flush` makes the
    current stream wait for every outstanding copy."""

    def __init__(self, model, n_samples, do_consensus_masking, depth=2):
        self.model, self.n_samples, self.masking, self.depth = model, n_samples, do_consensus_masking, depth
        dev = next(model.parameters()).device
        self.dev = dev
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.xbuf = [None] * depth
        self.in_done = [torch.cuda.Event() for _ in range(depth)]
        self.compute_done = [torch.cuda.Event() for _ in range(depth)]
        self.i = 0

    @torch.no_grad()
    def submit(self, host_images, out_mean, out_cons, eps=None):
        slot = self.i % self.depth
        self.i += 1
        cur = torch.cuda.current_stream(self.dev)
        if self.xbuf[slot] is None or self.xbuf[slot].shape != host_images.shape:
            self.xbuf[slot] = torch.empty(host_images.shape, dtype=torch.float32, device=self.dev)
            self.compute_done[slot].record(cur)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.compute_done[slot])  # the step that last read this input slot is done
            self.xbuf[slot].copy_(host_images, non_blocking=True)
            self.in_done[slot].record(self.s_in)
        cur.wait_event(self.in_done[slot])
        mean, cons = sample_from_teacher(self.model, self.xbuf[slot], self.n_samples,
                                         do_consensus_masking=self.masking, eps=eps)
        self.compute_done[slot].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.compute_done[slot])
            out_mean.copy_(mean, non_blocking=True)
            out_cons.copy_(cons, non_blocking=True)
        mean.record_stream(self.s_out)
        cons.record_stream(self.s_out)
        return out_mean, out_cons

    def flush(self):
        torch.cuda.current_stream(self.dev).wait_stream(self.s_out)


@torch.no_grad()
def predict_host(model, host_images, n_samples, do_consensus_masking, out_mean, out_cons, eps=None):
    """End-to-end call with HOST buffers (pinned): H2D of the image batch, forward + fused MC consensus,
    D2H of the mean probability and the consensus.  Everything is enqueued on the current stream."""
    dev = next(model.parameters()).device
    x = host_images.to(dev, non_blocking=True)
    mean, cons = sample_from_teacher(model, x, n_samples, do_consensus_masking=do_consensus_masking, eps=eps)
    out_mean.copy_(mean, non_blocking=True)
    out_cons.copy_(cons, non_blocking=True)
    return out_mean, out_cons


class GraphedMCPredictor:
    """`sample_from_teacher` for one fixed input shape, captured in a CUDA graph: forward + n_samples fused samples +
    consensus replay as one graph launch.  For small tiles (the 256 x 256 Lung-XRay images of BASELINE config 1) the
    ~45 launches of the eager path cost more host time than the device needs for the arithmetic.  The latent draws
    happen inside the graph (torch's graph-safe generator: fresh samples on every replay) unless `eps` is passed to a
    call.  Weights are read at replay time (an optimizer / EMA step between calls is picked up after
    `refresh_packed`, which the step bodies already run)."""

    def __init__(self, net, example_inputs, n_samples=16, upper_thres=0.9, lower_thres=0.1, do_consensus_masking=False):
        self.net = net
        self.x = example_inputs.clone()
        self.eps = torch.empty(n_samples, self.x.shape[0], net.latent_dim, device=self.x.device)
        self._fresh = True

        def run():
            if self._fresh:
                self.eps.normal_()
            return sample_from_teacher(net, self.x, n_samples, upper_thres, lower_thres, do_consensus_masking,
                                       eps=self.eps)
        self.graphs = {}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        for fresh in (True, False):
            self._fresh = fresh
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = run()
            self.graphs[fresh] = (g, out)

    @torch.no_grad()
    def __call__(self, inputs, eps=None):
        if inputs is not self.x:
            self.x.copy_(inputs, non_blocking=True)
        if eps is not None:
            self.eps.copy_(eps, non_blocking=True)
        g, out = self.graphs[eps is None]
        g.replay()
        return out
