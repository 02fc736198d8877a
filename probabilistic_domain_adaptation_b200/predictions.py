"""Drop-ins for `prob_utils.my_predictions.punet_prediction` / `punet_pseudo_prediction` on the fused kernels.

Reference: /root/reference/prob_utils/my_predictions/punet_predictions.py:15-63 (tiled MC prediction -> mean
probability TIFF) and :66-136 (whole-image pseudo-label + consensus mask TIFFs).  Same signatures, same glob patterns,
same output paths / file names / dtypes; the file I/O stays `imageio.v3` exactly as in the reference.  What changes is
everything between `imread` and `imwrite`:

  punet_prediction         torch_em.predict_with_halo (one block at a time, S x sample(), ATen stack/sum)
                           -> tiled.predict_with_halo: blocks gathered + standardised in batches on the device, one
                              forward + ONE pda_fcomb_mc_consensus launch per block batch, one scatter kernel
  punet_pseudo_prediction  S x sample() + numpy thresholds on the host
                           -> consensus.punet_pseudo_labels: one forward + ONE pda_fcomb_mc_consensus launch returning
                              the mean probability and the consensus mask

`run.install()` binds these names into `prob_utils.my_predictions`, so an unchanged `--predict` /
`--get_pseudo_labels` script reaches them.
"""
import os
from glob import glob

import numpy as np
import torch

from . import _lib, consensus, ops, tiled
from .my_models.utils import clean_folder


def _imageio():
    import imageio.v3 as imageio  # the reference's reader / writer (punet_predictions.py:5)
    return imageio


def standardize_image(img, device, eps=1e-7):
    """torch_em.transform.raw.standardize on a whole image ((x - mean) / (std + eps), population std; third party,
    restated -- see tiled.py) as ONE device kernel.  -> (1, 1, H, W) fp32 on `device`."""
    x = torch.as_tensor(np.ascontiguousarray(img)).to(device=device, dtype=torch.float32, non_blocking=True)
    if x.dim() != 2:
        raise ValueError(f"expected a 2-D image, got shape {tuple(x.shape)}")
    H, W = x.shape
    roi = torch.tensor([[0, 0, H, W]], dtype=torch.int32).to(device, non_blocking=True)
    stats = torch.empty(2, dtype=torch.float64, device=device)
    out = torch.empty((1, 1, H, W), dtype=torch.float32, device=device)
    _lib.check(_lib.load().pda_tile_gather_standardize(x.data_ptr(), H, W, roi.data_ptr(), 1, H, W, stats.data_ptr(),
                                                       out.data_ptr(), ops._stream()), "standardize")
    return out


@torch.no_grad()
def punet_prediction(input_image_path, output_pred_path, model, prior_samples=8, device="cpu", mysig=None,
                     block_shape=(384, 384), halo=(64, 64), batch_tiles=8):
    """punet_predictions.py:15-63.  `device` and `mysig` are accepted for signature compatibility: the model's own
    device is used (the kernels have no CPU path) and the sigmoid is part of the fused kernel."""
    imageio = _imageio()
    model.eval()
    os.makedirs(output_pred_path, exist_ok=True)
    for img_path in glob(input_image_path):
        img_name = os.path.basename(img_path)
        input_img = imageio.imread(img_path)
        pred = tiled.predict_with_halo(torch.as_tensor(np.ascontiguousarray(input_img)), model,
                                       prior_samples=prior_samples, block_shape=block_shape, halo=halo,
                                       batch_tiles=batch_tiles)
        # the reference hands predict_with_halo a float64 output array (np.zeros(input_img.shape), :48)
        pred = pred.cpu().numpy().astype(np.float64)
        output_path = os.path.join(output_pred_path, f"{img_name[:-4]}.tif")
        imageio.imwrite(output_path, pred, compression="zlib")
        print(f"Saved image at '{output_path}")


def punet_pseudo_prediction(input_image_path, output_pred_path, model, prior_samples=8, device="cpu", cellname_=None,
                            split_name=None):
    """punet_predictions.py:66-136: pseudo-label (mean of `prior_samples` sigmoid samples, fp32) under
    annotations/<split>/<cell>/ and the consensus mask (uint8 {0,1}) under consensus/<split>/<cell>/."""
    imageio = _imageio()
    os.makedirs(output_pred_path, exist_ok=True)
    clean_folder(output_pred_path)
    upper_threshold, lower_threshold = 0.9, 0.1
    model.eval()
    dev = next(model.parameters()).device
    with torch.no_grad():
        my_data_dir = input_image_path + f"{cellname_}*.tif"
        for i in glob(my_data_dir):
            my_image_name = i.split("/")[-1]
            my_patch = standardize_image(imageio.imread(i), dev)
            mean, mask = consensus.punet_pseudo_labels(model, my_patch, prior_samples, upper_threshold,
                                                       lower_threshold)
            mypred = mean.cpu().numpy().squeeze()
            consensus_mask = mask.cpu().numpy().squeeze()
            dir1 = output_pred_path + f"annotations/{split_name}/{cellname_}/"
            dir2 = output_pred_path + f"consensus/{split_name}/{cellname_}/"
            os.makedirs(dir1, exist_ok=True)
            os.makedirs(dir2, exist_ok=True)
            imageio.imwrite(dir1 + f"{my_image_name}", mypred)
            imageio.imwrite(dir2 + f"{my_image_name}", consensus_mask.astype("uint8"))
            print(f"{my_image_name}'s predictions saved")


def punet_trainer_sample(self, n_samples=16):
    """PUNetTrainer._sample (punet_trainer.py:15-17): n_samples x model.sample() -> list of logits (B,1,H,W), from ONE
    fused launch (same RNG stream as n_samples successive sample() calls)."""
    with torch.no_grad():
        _, _, logits, _ = self.model.mc_consensus(n_samples, want_consensus=False, return_samples=True)
    return list(logits.unbind(0))
