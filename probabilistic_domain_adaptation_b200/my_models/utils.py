"""Host-side mirror of prob_utils/my_models/utils.py (inits + L2 regulariser).

Reference: /root/reference/prob_utils/my_models/utils.py:8-40.  Plotting helpers of the reference are
out of scope (component 4); `clean_folder` is kept because punet_predictions.py:12 imports it.
"""
import os

import torch
import torch.nn as nn


def truncated_normal_(tensor, mean=0, std=1):
    """utils.py:8-14: per element, the first of four N(0,1) draws that lies in (-2, 2), scaled by std."""
    with torch.no_grad():
        draws = torch.randn(tuple(tensor.shape) + (4,), dtype=tensor.dtype, device=tensor.device)
        inside = (draws < 2) & (draws > -2)
        first = inside.max(-1, keepdim=True)[1]
        tensor.copy_(draws.gather(-1, first).squeeze(-1))
        tensor.mul_(std).add_(mean)
    return tensor


def init_weights(m):
    """utils.py:17-22: He-normal (fan_in, relu) weights, truncated-normal(0, 1e-3) bias."""
    if type(m) in (nn.Conv2d, nn.ConvTranspose2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_in", nonlinearity="relu")
        truncated_normal_(m.bias, mean=0, std=0.001)


def init_weights_orthogonal_normal(m):
    """utils.py:25-29: orthogonal weights, truncated-normal(0, 1e-3) bias."""
    if type(m) in (nn.Conv2d, nn.ConvTranspose2d):
        nn.init.orthogonal_(m.weight)
        truncated_normal_(m.bias, mean=0, std=0.001)


def l2_regularisation(m):
    """utils.py:32-40: sum over the parameter TENSORS of m of ||W||_2 (not squared, biases included)."""
    from ..autograd_ops import l2_norm_sum
    params = [p for p in m.parameters()]
    if not params:
        return None
    return l2_norm_sum(params)


def clean_folder(folder_path):
    """utils.py:50-55."""
    for name in os.listdir(folder_path):
        path = folder_path + name
        if os.path.isfile(path):
            os.remove(path)
