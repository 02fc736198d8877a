"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the tiled prediction the reference delegates to
torch_em.util.prediction.predict_with_halo (prob_utils/my_predictions/punet_predictions.py:41-49).

PARITY UNPINNED: torch_em is an unpinned third-party dependency that is not vendored and not installed in the build
container; its blocking and its default per-block preprocessing (torch_em.transform.raw.standardize:
(x - mean) / (std + 1e-7), population std) are restated here from its published behaviour.
"""
import numpy as np


def standardize(raw, eps=1e-7):
    raw = raw.astype("float32")
    mean = raw.mean()
    std = raw.std()
    return (raw - mean) / (std + eps)


def predict_with_halo(image, predict_fn, block_shape=(384, 384), halo=(64, 64)):
    """predict_fn(tile (h, w) float32) -> (h, w) prediction.  Inner blocks on a block_shape grid; the outer block adds
    the halo on every side, clipped to the image; the prediction of the outer block is cropped back to the inner one."""
    H, W = image.shape
    out = np.zeros((H, W), dtype="float32")
    for by in range(0, H, block_shape[0]):
        for bx in range(0, W, block_shape[1]):
            ih, iw = min(block_shape[0], H - by), min(block_shape[1], W - bx)
            oy0, ox0 = max(0, by - halo[0]), max(0, bx - halo[1])
            oy1, ox1 = min(H, by + ih + halo[0]), min(W, bx + iw + halo[1])
            tile = standardize(image[oy0:oy1, ox0:ox1])
            pred = predict_fn(tile)
            out[by:by + ih, bx:bx + iw] = pred[by - oy0:by - oy0 + ih, bx - ox0:bx - ox0 + iw]
    return out
