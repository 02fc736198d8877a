"""Adam on the multi-tensor kernel: same update rule and state layout as torch.optim.Adam
(the optimizer every reference script builds, e.g. LIVECell/livecell_punet.py:58), one launch per step."""
import torch

from . import _lib, ops

_CHUNK = 16384  # elements per 256-thread block of the multi-tensor kernel


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for torch.optim.Adam(params, lr, betas, eps, weight_decay) (amsgrad / maximize not supported).
    state_dict() has torch's per-parameter layout (`step`, `exp_avg`, `exp_avg_sq`)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    def _table(self, gi, params):
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1]
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
            for p in params:
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            step = int(self.state[params[0]]["step"].item()) + 1
            table = self._table(gi, params)
            b1, b2 = group["betas"]
            _lib.check(lib.pda_multi_tensor_adam(table.data_ptr(), table.shape[0], float(group["lr"]), float(b1),
                                                 float(b2), float(group["eps"]), float(group["weight_decay"]), step,
                                                 ops._ptr(inv_scale), ops._ptr(found_inf),
                                                 torch.cuda.current_stream().cuda_stream), "multi_tensor_adam")
            for p in params:
                self.state[p]["step"] += 1
            bump_versions(params)
        return loss


def bump_versions(tensors):
    """Kernels wrote these tensors behind autograd's back: advance their version counters so that saved-tensor
    checks and the packed-weight cache (autograd_ops.packed_weight) see the change."""
    tensors = list(tensors)
    torch._C._autograd._unsafe_set_version_counter(tensors, [t._version + 1 for t in tensors])
