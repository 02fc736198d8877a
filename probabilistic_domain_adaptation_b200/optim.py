"""Adam on the multi-tensor kernel: same update rule and state layout as torch.optim.Adam
(the optimizer every reference script builds, e.g. LIVECell/livecell_punet.py:58), one launch per step."""
import torch

from . import _lib, ops

_CHUNK = 16384  # elements per 256-thread block of the multi-tensor kernel


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for torch.optim.Adam(params, lr, betas, eps, weight_decay) (amsgrad / maximize not supported).
    state_dict() has torch's per-parameter layout (`step`, `exp_avg`, `exp_avg_sq`).

    capturable=True (torch.optim.Adam(capturable=True) semantics): the step count and the learning rate are read from
    device memory, so a CUDA graph that captured `step()` can be replayed (steps.GraphedStep).  `state[p]["step"]` is
    then one shared int64 device scalar; `sync_lr()` pushes a changed `param_groups[i]["lr"]` (LR schedulers) to the
    device and must run OUTSIDE the captured region (GraphedStep does it before every replay)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}
        self.capturable = capturable
        self._step_dev, self._lr_dev, self._lr_host = None, {}, {}

    def sync_lr(self):
        """capturable mode: copy changed learning rates to their device scalars (outside any graph capture)."""
        for gi, group in enumerate(self.param_groups):
            t = self._lr_dev.get(gi)
            if t is not None and self._lr_host.get(gi) != float(group["lr"]):
                t.fill_(float(group["lr"]))
                self._lr_host[gi] = float(group["lr"])

    def _table(self, gi, params):
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1]
        if torch.cuda.is_current_stream_capturing():
            # building the table needs a host -> device copy, which is illegal under stream capture
            raise _lib.PdaError(
                "FusedAdam: parameter / gradient storage changed inside a CUDA-graph capture.  Keep p.grad storage "
                "fixed (attach a parallel.GradAllReducer, or zero gradients in place with zero_grad(set_to_none="
                "False)) and run one eager step before capturing")
        rows = []
        for p in params:
            st = self.state[p]
            for t in (p, p.grad, st["exp_avg"], st["exp_avg_sq"]):
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.PdaError("FusedAdam needs contiguous fp32 parameters, gradients and state")
            n = p.numel()
            for off in range(0, n, _CHUNK):
                rows.append((p.data_ptr() + 4 * off, p.grad.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off,
                             st["exp_avg_sq"].data_ptr() + 4 * off, min(_CHUNK, n - off)))
        table = torch.tensor(rows, dtype=torch.int64).to(params[0].device)
        self._tables[gi] = (key, table)
        return table

    @torch.no_grad()
    def step(self, closure=None, inv_scale=None, found_inf=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            if self.capturable and self._step_dev is None:
                self._step_dev = torch.zeros((), dtype=torch.int64, device=params[0].device)
            for p in params:
                st = self.state[p]
                if not st:
                    st["step"] = self._step_dev if self.capturable else torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            table = self._table(gi, params)
            b1, b2 = group["betas"]
            if self.capturable:
                if gi not in self._lr_dev:
                    self._lr_dev[gi] = torch.full((), float(group["lr"]), dtype=torch.float32, device=params[0].device)
                    self._lr_host[gi] = float(group["lr"])
                if not torch.cuda.is_current_stream_capturing():
                    self.sync_lr()
                if gi > 0:
                    raise _lib.PdaError("capturable FusedAdam supports one parameter group (shared step counter)")
                _lib.check(lib.pda_multi_tensor_adam_capturable(
                    table.data_ptr(), table.shape[0], self._lr_dev[gi].data_ptr(), float(b1), float(b2),
                    float(group["eps"]), float(group["weight_decay"]), self._step_dev.data_ptr(), ops._ptr(inv_scale),
                    ops._ptr(found_inf), ops._stream()), "multi_tensor_adam_capturable")
                bump_versions(params)
                continue
            step = int(self.state[params[0]]["step"].item()) + 1
            _lib.check(lib.pda_multi_tensor_adam(table.data_ptr(), table.shape[0], float(group["lr"]), float(b1),
                                                 float(b2), float(group["eps"]), float(group["weight_decay"]), step,
                                                 ops._ptr(inv_scale), ops._ptr(found_inf),
                                                 ops._stream()), "multi_tensor_adam")
            for p in params:
                self.state[p]["step"] += 1
            bump_versions(params)
        return loss


def bump_versions(tensors):
    """Kernels wrote these tensors behind autograd's back: advance their version counters so that saved-tensor
    checks and the packed-weight cache (autograd_ops.packed_weight) see the change."""
    tensors = list(tensors)
    torch._C._autograd._unsafe_set_version_counter(tensors, [t._version + 1 for t in tensors])
