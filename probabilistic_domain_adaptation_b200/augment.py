"""On-device weak / strong view augmentation.

The reference builds the two target-domain views per sample on the CPU inside the DataLoader workers
(prob_utils/my_datasets/my_image_collection_dataset.py:349-357) with

    get_raw_transform(normalizer=my_standardize_torch, augmentation1=Compose([
        my_standardize_torch,
        RandomApply([GaussianBlur(sigma=...)], p),
        RandomApply([AdditiveGaussianNoise(scale=..., clip_kwargs=False)], p),
        RandomApply([RandomContrast(alpha=..., mean=0.0, clip_kwargs=False)], p)]))

(MitoEM/common.py:50-68, LIVECell/livecell_fm.py:43-67, LIVECell/livecell_adamatch.py:16-38, ...).  Here the random
DECISIONS are drawn on the host, in the reference's order (torchvision RandomApply: one torch.rand(1) per transform;
torch_em transforms: numpy draws for kernel size / sigma / scale / alpha), and the pixel work of a whole batch runs in
one statistics pass + one fused kernel per view (csrc/augment_kernels.cu).  torch_em is not vendored: the parameter
ranges / draw order of its three transforms are restated from its published source (parity pinned only against
torchvision's GaussianBlur / RandomApply and the reference's own my_standardize_torch, see tests/golden/augment.pt).
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream

NPARAM = 8
MAX_KSIZE = 31


@dataclass
class ViewSpec:
    """One `my_*_augmentations(p)` recipe.  A probability of 0 (or None) removes the transform from the Compose."""
    blur_p: Optional[float] = None                        # None: no RandomApply(GaussianBlur) in the Compose
    blur_sigma: Tuple[float, float] = (0.0, 5.0)          # torch_em GaussianBlur default
    blur_kernel_size: Tuple[int, int] = (2, 24)           # torch_em GaussianBlur default
    noise_p: Optional[float] = None
    noise_scale: Tuple[float, float] = (0.0, 0.3)
    contrast_p: Optional[float] = None                    # None: no RandomContrast in the Compose
    contrast_alpha: Tuple[float, float] = (0.5, 2.0)
    contrast_mean: float = 0.0                            # every script passes mean=0.0
    n_standardize: int = 2                                # normalizer + first entry of the Compose


# --- the recipes of the experiment scripts --------------------------------------------------------------------------
def weak_view(p=0.25):
    """my_weak_augmentations of every script (e.g. MitoEM/common.py:50-57)."""
    return ViewSpec(blur_p=p, noise_p=p, noise_scale=(0.0, 0.15))


def mitoem_strong_view(p=0.5):
    """MitoEM/common.py:60-68, LIVECell/livecell_adamatch.py:29-38."""
    return ViewSpec(blur_p=p, blur_sigma=(0.6, 3.0), noise_p=p / 2, noise_scale=(0.05, 0.25), contrast_p=p,
                    contrast_alpha=(0.33, 3.0))


def livecell_fixmatch_strong_view(p=0.9):
    """LIVECell/livecell_fm.py:56-68."""
    return ViewSpec(blur_p=p, blur_sigma=(1.0, 4.0), noise_p=p, noise_scale=(0.1, 0.35), contrast_p=p,
                    contrast_alpha=(0.33, 3.0))


def sample_view_params(spec, n_images, image_shape=None, torch_gen=None, np_rng=None, cpu_noise=False):
    """Draws the random decisions of `n_images` applications of one view recipe, in the reference's order.

    Returns (params (n_images, 8) float32 CPU tensor, noise or None, max_ksize).  With `cpu_noise` the additive noise
    is drawn like the reference does (torch.normal(0, scale, image_shape) from the CPU generator, between the
    RandomApply draws) and returned pre-scaled with a scale parameter of 1; otherwise the caller draws a unit-normal
    field on the device."""
    np_rng = np.random if np_rng is None else np_rng
    params = torch.zeros(n_images, NPARAM, dtype=torch.float32)
    noise = None
    max_ksize = 1
    for b in range(n_images):
        ksize, sigma, scale, alpha = 1, 0.0, 0.0, 1.0
        # torchvision.transforms.RandomApply.forward: `if self.p < torch.rand(1): return img`
        if spec.blur_p is not None:
            applied = not (spec.blur_p < float(torch.rand(1, generator=torch_gen)))
            if applied:
                lo, hi = spec.blur_kernel_size
                ksize = 2 * (int(np_rng.randint(lo, hi)) // 2) + 1
                sigma = float(np_rng.uniform(spec.blur_sigma[1], spec.blur_sigma[0]))
                # torchvision.transforms.GaussianBlur.forward draws its sigma with torch.empty(1).uniform_(lo, hi)
                # even for lo == hi: one draw of the torch stream, and sigma becomes a float32
                sigma = float(torch.empty(1).uniform_(sigma, sigma, generator=torch_gen))
        if spec.noise_p is not None:
            applied = not (spec.noise_p < float(torch.rand(1, generator=torch_gen)))
            if applied:
                scale = float(np_rng.uniform(spec.noise_scale[0], spec.noise_scale[1]))
                if cpu_noise:
                    if noise is None:
                        noise = torch.zeros((n_images,) + tuple(image_shape), dtype=torch.float32)
                    noise[b] = torch.normal(0, scale, tuple(image_shape), generator=torch_gen)
                    scale = 1.0
        if spec.contrast_p is not None:
            applied = not (spec.contrast_p < float(torch.rand(1, generator=torch_gen)))
            if applied:
                alpha = float(np_rng.uniform(spec.contrast_alpha[0], spec.contrast_alpha[1]))
        if ksize > MAX_KSIZE:
            raise ValueError(f"blur kernel size {ksize} > {MAX_KSIZE}")
        max_ksize = max(max_ksize, ksize)
        params[b, 0], params[b, 1], params[b, 2] = ksize, sigma, scale
        params[b, 3], params[b, 4], params[b, 5] = alpha, spec.contrast_mean, spec.n_standardize
    return params, noise, max_ksize


def image_stats(raw):
    """(B, 2) float64: per-image sum and sum of squares (one pass; shared by every view of the batch)."""
    _need_cuda(raw)
    assert raw.dtype == torch.float32 and raw.is_contiguous()
    B = raw.shape[0]
    stats = torch.empty(B, 2, dtype=torch.float64, device=raw.device)
    lib = _lib.load()
    _lib.check(lib.pda_image_stats(raw.data_ptr(), B, raw[0].numel(), stats.data_ptr(), _stream()), "image_stats")
    return stats


def augment_view(raw, params, max_ksize, noise=None, stats=None, eps=1e-7):
    """One augmented view of a batch: raw (B, 1, H, W) fp32 CUDA, params from `sample_view_params` (CPU or CUDA)."""
    _need_cuda(raw)
    assert raw.dim() == 4 and raw.shape[1] == 1 and raw.dtype == torch.float32 and raw.is_contiguous()
    B, _, H, W = raw.shape
    assert params.shape == (B, NPARAM)
    if stats is None:
        stats = image_stats(raw)
    if not params.is_cuda:
        params = params.pin_memory().to(raw.device, non_blocking=True) if torch.cuda.is_available() else params
    if noise is not None:
        noise = noise.to(raw.device, non_blocking=True).contiguous()
        assert noise.shape == raw.shape and noise.dtype == torch.float32
    out = torch.empty_like(raw)
    lib = _lib.load()
    _lib.check(lib.pda_augment_view(raw.data_ptr(), _ptr(noise), out.data_ptr(), B, H, W, stats.data_ptr(),
                                    params.data_ptr(), float(eps), int(max_ksize), _stream()), "augment_view")
    return out


class DualViewAugmenter:
    """raw (B,1,H,W) on the device -> (raw, raw1, raw2): the tuple contract of DualImageCollectionDataset /
    DualSegmentationDataset (my_image_collection_dataset.py:369-372) with augmentation1 = `weak`,
    augmentation2 = `strong`, for a whole batch at once."""

    def __init__(self, weak=None, strong=None, torch_gen=None, np_rng=None):
        self.weak = weak_view() if weak is None else weak
        self.strong = mitoem_strong_view() if strong is None else strong
        self.torch_gen, self.np_rng = torch_gen, np_rng

    @torch.no_grad()
    def __call__(self, raw):
        B = raw.shape[0]
        # per sample: augmentation1 then augmentation2 (my_image_collection_dataset.py:353-357)
        rows = [[], []]
        for _ in range(B):
            for v, spec in enumerate((self.weak, self.strong)):
                rows[v].append(sample_view_params(spec, 1, torch_gen=self.torch_gen, np_rng=self.np_rng))
        stats = image_stats(raw)
        views = []
        for v in range(2):
            params = torch.cat([r[0] for r in rows[v]], 0)
            max_ksize = max(r[2] for r in rows[v])
            noise = torch.randn_like(raw) if bool((params[:, 2] != 0).any()) else None
            views.append(augment_view(raw, params, max_ksize, noise=noise, stats=stats))
        return raw, views[0], views[1]
