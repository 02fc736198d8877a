"""Tiled Monte-Carlo prediction on the device.

The reference hands each image to torch_em.util.prediction.predict_with_halo(block_shape=(384, 384), halo=(64, 64))
(prob_utils/my_predictions/punet_predictions.py:41-49), which cuts it into blocks on the host, standardises each
outer block, runs `_custom_punet_prediction` on ONE block at a time and writes the halo-cropped result back.  Here the
image is uploaded once; equally sized outer blocks are gathered + standardised in batches by one kernel, go through
forward + fused MC mean as a batch, and are scattered back by one kernel.  Blocks are independent, so ranks simply
take disjoint shares of the block list (no data-path collective; the stitched image is the sum of the partial ones).

torch_em is a third-party dependency that is not vendored: the blocking / standardisation rules are restated from its
published behaviour (oracle/tiled_oracle.py says so as well): inner blocks tile the image on a block_shape grid, the
outer block adds `halo` on every side clipped to the image, standardisation is (x - mean) / (std + 1e-7) per outer
block with the population std.
"""
import torch

from . import _lib, consensus, ops
from .parallel import shard_range


def blocking(shape, block_shape=(384, 384), halo=(64, 64)):
    """-> list of (outer (y0, x0, h, w), inner (y0, x0, h, w)) in image coordinates, row-major block order."""
    H, W = shape
    out = []
    for by in range(0, H, block_shape[0]):
        for bx in range(0, W, block_shape[1]):
            ih, iw = min(block_shape[0], H - by), min(block_shape[1], W - bx)
            oy0, ox0 = max(0, by - halo[0]), max(0, bx - halo[1])
            oy1, ox1 = min(H, by + ih + halo[0]), min(W, bx + iw + halo[1])
            out.append(((oy0, ox0, oy1 - oy0, ox1 - ox0), (by, bx, ih, iw)))
    return out


@torch.no_grad()
def predict_with_halo(image, model, prior_samples=8, block_shape=(384, 384), halo=(64, 64), batch_tiles=8,
                      rank=0, world=1, eps_fn=None, output=None):
    """image: 2-D fp32 tensor (host or device).  Returns the (H, W) fp32 mean-probability image on the device; with
    world > 1 only this rank's blocks are filled (zeros elsewhere)."""
    dev = next(model.parameters()).device
    lib = _lib.load()
    img = image.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
    H, W = img.shape
    out = torch.zeros((H, W), dtype=torch.float32, device=dev) if output is None else output
    blocks = blocking((H, W), block_shape, halo)
    a, b = shard_range(len(blocks), rank, world)
    groups = {}
    for outer, inner in blocks[a:b]:
        groups.setdefault((outer[2], outer[3]), []).append((outer, inner))
    stream = torch.cuda.current_stream().cuda_stream
    for (th, tw), items in groups.items():
        for i in range(0, len(items), batch_tiles):
            chunk = items[i:i + batch_tiles]
            T = len(chunk)
            rois = torch.tensor([o for o, _ in chunk], dtype=torch.int32).to(dev, non_blocking=True)
            inner = torch.tensor([n for _, n in chunk], dtype=torch.int32).to(dev, non_blocking=True)
            stats = torch.empty(2 * T, dtype=torch.float64, device=dev)
            tiles = torch.empty((T, 1, th, tw), dtype=torch.float32, device=dev)
            _lib.check(lib.pda_tile_gather_standardize(img.data_ptr(), H, W, rois.data_ptr(), T, th, tw,
                                                       stats.data_ptr(), tiles.data_ptr(), stream), "tile_gather")
            eps = None if eps_fn is None else eps_fn(chunk)
            pred = consensus.punet_mc_prediction(model, tiles, prior_samples, eps=eps)
            _lib.check(lib.pda_tile_scatter(pred.data_ptr(), T, th, tw, rois.data_ptr(), inner.data_ptr(),
                                            out.data_ptr(), H, W, stream), "tile_scatter")
    return out
