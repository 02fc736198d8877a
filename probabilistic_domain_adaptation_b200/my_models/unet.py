"""Host-side mirror of prob_utils/my_models/unet.py (4-level U-Net trunk) on the sm_100a kernels.

Reference: /root/reference/prob_utils/my_models/unet.py:6-69.  Channel wiring is the INTENDED one
(block i consumes num_filters[i-1] channels); the vendored reference assigns `output` before `input`
(unet.py:27-28) and cannot run -- see SURVEY.md section 0 fact 3.
"""
import torch.nn as nn

from .unet_blocks import DownConvBlock, UpConvBlock, as_nchw, run_conv_stack, _check_spatial


class Unet(nn.Module):
    def __init__(self, input_channels, num_classes, num_filters, initializers, apply_last_layer=True, padding=True):
        super().__init__()
        self.input_channels = input_channels
        self.num_classes = num_classes
        self.num_filters = list(num_filters)
        self.padding = padding
        self.activation_maps = []
        self.apply_last_layer = apply_last_layer
        if apply_last_layer:
            raise NotImplementedError("the PUNet path builds the trunk with apply_last_layer=False "
                                      "(probabilistic_unet.py:256); the plain-UNet head is out of scope")
        if input_channels != 1:
            raise NotImplementedError("input_channels must be 1 (every reference script)")
        if self.num_filters[0] > 64 or max(self.num_filters) > 512 or min(self.num_filters) < 1:
            raise NotImplementedError("num_filters[0] <= 64 (the fused Fcomb kernels are 64 wide) and every width <= 512; "
                                      "widths that are not multiples of 64 run zero-padded to the 64-channel tiles")

        self.contracting_path = nn.ModuleList()
        prev = input_channels
        for i, f in enumerate(self.num_filters):
            self.contracting_path.append(DownConvBlock(prev, f, initializers, padding, pool=(i != 0)))
            prev = f
        self.upsampling_path = nn.ModuleList()
        for i in range(len(self.num_filters) - 2, -1, -1):
            self.upsampling_path.append(UpConvBlock(prev + self.num_filters[i], self.num_filters[i], initializers,
                                                    padding))
            prev = self.num_filters[i]

    def forward_nhwc(self, patch):
        """patch: fp32 (B,1,H,W).  Returns the (B,H,W,num_filters[0]) bf16 feature map."""
        nlev = len(self.contracting_path)
        _check_spatial(patch.shape[2], patch.shape[3], nlev)
        skips = []
        x = None
        for i, down in enumerate(self.contracting_path):
            last = i == nlev - 1
            # the 2x2 average pool that opens block i+1 (unet_blocks.py:17) runs in block i's last epilogue
            full, pooled = run_conv_stack(down.convs(), x, first_input=(patch, None) if i == 0 else None,
                                          pool_last=not last, keep_full=True)
            if not last:
                skips.append(full)
                x = pooled
            else:
                x = full
        nf = self.num_filters
        for i, up in enumerate(self.upsampling_path):
            lvl = nlev - 2 - i  # level of the bridge: real channels (nf[lvl + 1] from below, nf[lvl] from the skip)
            x = up.forward_nhwc(x, skips[-i - 1], seg_real=(nf[lvl + 1], nf[lvl]))
        return x

    def forward(self, x, val):
        feat = as_nchw(self.forward_nhwc(x))[:, :self.num_filters[0]]
        if val:
            self.activation_maps.append(feat)
        return feat
