"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the per-sample view augmentation of the target-domain loaders
(/root/reference/MitoEM/common.py:50-68, LIVECell/livecell_fm.py:43-67, applied at
prob_utils/my_datasets/my_image_collection_dataset.py:349-357).

What is pinned and what is not:
  * `my_standardize_torch` restates prob_utils/my_utils/util.py:9-14 and is checked against the reference function
    itself through tests/golden/augment.pt (oracle/make_golden.py runs the reference's own function).
  * `transforms.RandomApply` / `transforms.GaussianBlur` are torchvision's own classes (installed in this image).
  * `GaussianBlur`, `AdditiveGaussianNoise`, `RandomContrast`, `get_raw_transform` restate torch_em.transform.raw
    (unpinned third-party dependency, not vendored, not installed): PARITY UNPINNED for their parameter draws
    (kernel size 2*(randint(2,24)//2)+1, sigma uniform(hi, lo), scale / alpha uniform, torch.normal noise).
"""
import numpy as np
import torch
from torchvision import transforms


def my_standardize_torch(tensor, mean=None, std=None, axis=None, eps=1e-7):
    mean = tensor.mean() if mean is None else mean
    tensor -= mean
    std = tensor.std() if std is None else std
    tensor /= (std + eps)
    return tensor


class GaussianBlur:
    def __init__(self, kernel_size=(2, 24), sigma=(0, 5)):
        self.kernel_size, self.sigma = kernel_size, sigma

    def __call__(self, img):
        kernel_size = 2 * (np.random.randint(self.kernel_size[0], self.kernel_size[1]) // 2) + 1
        sigma = np.random.uniform(self.sigma[1], self.sigma[0])
        return transforms.GaussianBlur(kernel_size, sigma=sigma)(img)


class AdditiveGaussianNoise:
    def __init__(self, scale=(0.0, 0.3), clip_kwargs={"a_min": 0, "a_max": 1}):
        self.scale, self.clip_kwargs = scale, clip_kwargs

    def __call__(self, img):
        scale = np.random.uniform(self.scale[0], self.scale[1])
        img = img + torch.normal(0, scale, img.shape)
        return torch.clip(img, **self.clip_kwargs) if self.clip_kwargs else img


class RandomContrast:
    def __init__(self, alpha=(0.5, 2), mean=0.5, clip_kwargs={"a_min": 0, "a_max": 1}):
        self.alpha, self.mean, self.clip_kwargs = alpha, mean, clip_kwargs

    def __call__(self, img):
        alpha = np.random.uniform(self.alpha[0], self.alpha[1])
        mean = img.mean() if self.mean is None else self.mean
        img = mean + alpha * (img - mean)
        return torch.clip(img, **self.clip_kwargs) if self.clip_kwargs else img


def get_raw_transform(normalizer, augmentation1=None):
    def raw_transform(raw):
        raw = normalizer(raw)
        if augmentation1 is not None:
            raw = augmentation1(raw)
        return raw
    return raw_transform


def make_view(standardize, blur_p=None, blur_sigma=(0, 5), noise_p=None, noise_scale=(0.0, 0.3), contrast_p=None,
              contrast_alpha=(0.5, 2), contrast_mean=0.0):
    """The `my_*_augmentations` recipe of the scripts with the given standardize function."""
    steps = [standardize]
    if blur_p is not None:
        steps.append(transforms.RandomApply([GaussianBlur(sigma=blur_sigma)], p=blur_p))
    if noise_p is not None:
        steps.append(transforms.RandomApply([AdditiveGaussianNoise(scale=noise_scale, clip_kwargs=False)], p=noise_p))
    if contrast_p is not None:
        steps.append(transforms.RandomApply([RandomContrast(alpha=contrast_alpha, mean=contrast_mean,
                                                            clip_kwargs=False)], p=contrast_p))
    return get_raw_transform(normalizer=standardize, augmentation1=transforms.Compose(steps))


WEAK = dict(blur_p=0.25, noise_p=0.25, noise_scale=(0, 0.15))                                  # MitoEM/common.py:50-57
STRONG = dict(blur_p=0.5, blur_sigma=(0.6, 3.0), noise_p=0.25, noise_scale=(0.05, 0.25), contrast_p=0.5,
              contrast_alpha=(0.33, 3.0))                                                       # MitoEM/common.py:60-68
# probabilities pushed to 1 so that one small fixture exercises every transform
ALL_ON = dict(blur_p=1.0, blur_sigma=(0.6, 3.0), noise_p=1.0, noise_scale=(0.05, 0.25), contrast_p=1.0,
              contrast_alpha=(0.33, 3.0))


def dual_views(raw_batch, weak_kw, strong_kw, seed, standardize=my_standardize_torch):
    """Per sample: raw1 = augmentation1(copy), raw2 = augmentation2(copy)  (my_image_collection_dataset.py:349-357),
    with both global RNGs seeded once."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    weak, strong = make_view(standardize, **weak_kw), make_view(standardize, **strong_kw)
    out1, out2 = [], []
    for b in range(raw_batch.shape[0]):
        out1.append(weak(raw_batch[b].clone()))
        out2.append(strong(raw_batch[b].clone()))
    return torch.stack(out1), torch.stack(out2)
