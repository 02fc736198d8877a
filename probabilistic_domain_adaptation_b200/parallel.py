"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on the GPU box,
gloo in the CPU tests).

* Monte-Carlo inference shards TILES across ranks: no data-path collective.
* Training is data parallel: ONE exchange step per iteration, the gradient all-reduce (average), issued per
  bucket from post-accumulate-grad hooks so that it overlaps the rest of backward.  The reference itself has no
  distributed code; the wrapper works on the bare module because ProbabilisticUnet.forward returns None and the
  loss comes from methods (elbo), which DistributedDataParallel would hide.
"""
import weakref

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [start, stop) share of n_items for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_tiles(tiles, rank=None, world=None):
    """This rank's share of a sequence of independent tiles / images."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    a, b = shard_range(len(tiles), rank, world)
    return tiles[a:b]


def broadcast_parameters(module, src=0):
    """Identical initial student / teacher weights on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for p in module.parameters():
        dist.broadcast(p.data, src)


class GradAllReducer:
    """Bucketed gradient averaging for a bare nn.Module.

    Parameters are grouped into buckets of ~bucket_mb in REVERSE registration order (the order in which backward
    produces their gradients).  When the last gradient of a bucket has been accumulated, the bucket is packed into a
    flat buffer and an asynchronous all-reduce (AVG on NCCL, SUM + scale elsewhere) is launched; `finish()` waits for
    all buckets and rebinds p.grad to its averaged view inside the flat buffer.  Call `finish()` after backward and
    before optimizer.step().

    reserve_sms (default 0): with more than one rank, the persistent tensor-core kernels leave this many SMs free
    (`pda_set_sm_budget`) so that NCCL's CTAs can run NEXT TO the backward kernels -- those otherwise occupy all 148 SMs
    with one ~200 KB CTA each and the all-reduce only progresses between kernels (round 1: +0.57 ms exposed at 8 GPUs).
    comm_dtype: torch.bfloat16 halves the bytes on the wire (gradients are averaged in bf16, then widened back into the
    fp32 buckets the optimizer reads); default fp32 = exact averaging.
    Measured at 2 GPUs (profiles/r02_2gpu_allreduce_options.md): neither option changes the step time (15.22 ms with
    4 SMs reserved, 15.23 ms without, 15.28 ms with bf16 on the wire), so both stay off by default."""

    def __init__(self, module, bucket_mb=25.0, process_group=None, reserve_sms=0, comm_dtype=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.comm_dtype = None if comm_dtype in (None, torch.float32) else comm_dtype
        self._prev_budget = None
        if self.world > 1 and reserve_sms > 0 and next(module.parameters()).is_cuda:
            from . import _lib
            self._prev_budget = _lib.load().pda_set_sm_budget(148 - int(reserve_sms))
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.buckets = []  # list of dict(params, flat, views, pending, work)
        limit = int(bucket_mb * 1024 * 1024)
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= limit:
                self._add_bucket(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._add_bucket(cur)
        self._bucket_of = {}
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._bucket_of[p] = bi
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        from . import training
        ref = weakref.WeakMethod(self._late_target)           # (a forgotten reducer must not be kept alive by the registry)
        self._probe = lambda p: (ref() or (lambda _p: None))(p)
        training._LAUNCHED.append(self._probe)
        self._avg = dist.is_initialized() and dist.get_backend(process_group) == "nccl"

    def _add_bucket(self, params):
        total = sum(p.numel() for p in params)
        flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        views, off = [], 0
        for p in params:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        wire = wire_views = None
        if self.comm_dtype is not None:
            wire = torch.zeros(total, dtype=self.comm_dtype, device=params[0].device)
            wire_views, off = [], 0
            for p in params:
                wire_views.append(wire[off:off + p.numel()].view_as(p))
                off += p.numel()
        self.buckets.append({"params": list(params), "flat": flat, "views": views, "pending": len(params),
                             "work": None, "wire": wire, "wire_views": wire_views})

    def _on_grad(self, p):
        b = self.buckets[self._bucket_of[p]]
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    def _launch(self, b):
        # regulariser gradients parked by training.L2NormSumFn: one multi-tensor add per bucket instead of one ATen add
        # per parameter inside the autograd engine
        from . import training
        reg_p, reg_g = training.take_reg_grads(b["params"])
        if reg_p:
            with torch.no_grad():
                torch._foreach_add_([p.grad for p in reg_p], reg_g)
        b["launched"] = True
        grads = [p.grad for p in b["params"]]
        if b["wire"] is not None and self.world > 1:
            torch._foreach_copy_(b["wire_views"], grads)       # fp32 -> bf16 while packing
            buf = b["wire"]
        else:
            torch._foreach_copy_(b["views"], grads)
            buf = b["flat"]
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            b["work"] = dist.all_reduce(buf, op=op, group=self.group, async_op=True)

    def _late_target(self, p):
        """The flat-gradient view of p if p's bucket has already been copied / handed to the all-reduce in this backward
        pass (training.flush_reg_grads adds a late regulariser gradient there, after the reduction), else None."""
        bi = self._bucket_of.get(p)
        if bi is None:
            return None
        b = self.buckets[bi]
        if not b.get("launched"):
            return None
        if b["work"] is not None:
            b["work"].wait()
            b["work"] = None
            if b["wire"] is not None:
                b["flat"].copy_(b["wire"])                     # widen the averaged bf16 gradients
            if not self._avg:
                b["flat"].div_(self.world)
        return b["views"][b["params"].index(p)]

    def finish(self):
        for b in self.buckets:
            if b["pending"] != 0:  # parameters that received no gradient this iteration
                missing = [p for p in b["params"] if p.grad is None]
                if len(missing) == len(b["params"]):
                    b["pending"] = len(b["params"])
                    continue
                for p, v in zip(b["params"], b["views"]):
                    if p.grad is None:
                        v.zero_()
                        p.grad = v
                self._launch(b)
            if b["work"] is not None:
                b["work"].wait()
                b["work"] = None
                if b["wire"] is not None:
                    b["flat"].copy_(b["wire"])                     # widen the averaged bf16 gradients
                if not self._avg:
                    b["flat"].div_(self.world)
            for p, v in zip(b["params"], b["views"]):
                p.grad = v
            b["pending"] = len(b["params"])
            b["launched"] = False

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        from . import training
        if self._probe in training._LAUNCHED:
            training._LAUNCHED.remove(self._probe)
        if self._prev_budget is not None:
            from . import _lib
            _lib.load().pda_set_sm_budget(self._prev_budget)
            self._prev_budget = None
