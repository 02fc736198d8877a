"""ctypes binding of libpda_b200.so (the C ABI declared in include/pda_b200.h).

There is no fallback: if the shared library is missing or fails to load, importing the ops raises.
"""
import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDA_B200_LIB") or os.path.join(PKG, "libpda_b200.so")  # override: A/B experiments only

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_F = _c.c_float

# name -> argtypes; every function returns int (status) unless listed in _RESTYPES
SIGNATURES = {
    "pda_abi_version": [],
    "pda_error_string": [_I],
    "pda_launch_count": [],
    "pda_reset_launch_count": [],
    "pda_pack_conv3x3_weights": [_P, _P, _I, _I, _I, _I, _P],
    "pda_pack_conv3x3_weights_multi": [_P, _I, _P],
    "pda_conv3x3_first": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pda_conv3x3_tc": [_P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "pda_conv3x3_up_tc": [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "pda_set_conv_pair": [_I],
    "pda_set_first_conv_tc": [_I],
    "pda_set_sm_budget": [_I],
    "pda_conv3x3_bf16_simt": [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pda_avgpool2": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pda_upsample2x_bilinear": [_P, _P, _I, _I, _I, _I, _I, _P],
    "pda_gauss_head_scratch_rows": [_I],
    "pda_gauss_head": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pda_latent_samples": [_P, _P, _P, _I, _I, _I, _P],
    "pda_kl_diag_gauss": [_P, _P, _P, _I, _I, _P],
    "pda_fcomb_scratch_floats": [_I, _I],
    "pda_fcomb_mc_consensus": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P],
    "pda_fcomb_mc_consensus_fp32": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _I, _P],
    "pda_fcomb_mc_consensus_deep": [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _I,
                                    _P],
    "pda_tile_gather_standardize": [_P, _I, _I, _P, _I, _I, _I, _P, _P, _P],
    "pda_tile_scatter": [_P, _I, _I, _I, _P, _P, _P, _I, _I, _P],
    "pda_image_stats": [_P, _I, _c.c_longlong, _P, _P],
    "pda_augment_view": [_P, _P, _P, _I, _I, _I, _P, _P, _F, _I, _P],
    "pda_multi_tensor_ema": [_P, _I, _c.c_double, _P],
    "pda_multi_tensor_ema_warmup": [_P, _I, _c.c_double, _P, _P],
    "pda_conv3x3_wgrad_scratch_floats": [_I, _I],
    "pda_conv3x3_wgrad_bf16": [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pda_conv3x3_wgrad_det_scratch_floats": [_I, _I, _I, _I, _I],
    "pda_conv3x3_wgrad_bf16_det": [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pda_relu_pool_bwd_bf16": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "pda_upsample2x_bilinear_bwd_bf16": [_P, _P, _I, _I, _I, _I, _P],
    "pda_conv3x3_first_bwd_scratch_floats": [_I, _I, _I, _I, _I],
    "pda_conv3x3_first_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P],
    "pda_gauss_head_mean": [_P, _P, _I, _I, _I, _P],
    "pda_gauss_head_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "pda_kl_diag_gauss_bwd": [_P, _P, _P, _P, _P, _I, _I, _P],
    "pda_recon_loss_blocks": [_c.c_longlong],
    "pda_recon_loss_fwd": [_P, _P, _P, _P, _c.c_longlong, _I, _P, _P, _P, _P],
    "pda_recon_loss_bwd": [_P, _P, _P, _P, _c.c_longlong, _I, _P, _P, _P, _P],
    "pda_dice_score": [_P, _P, _c.c_longlong, _F, _F, _P, _P, _P],
    "pda_multi_tensor_l2norm_fwd": [_P, _I, _I, _P, _P, _P, _P],
    "pda_multi_tensor_l2norm_bwd": [_P, _I, _P, _P, _P, _P],
    "pda_multi_tensor_adam": [_P, _I, _c.c_double, _c.c_double, _c.c_double, _c.c_double, _c.c_double, _c.c_longlong,
                              _P, _P, _P],
    "pda_distribution_alignment": [_P, _c.c_longlong, _P, _P, _P, _P, _P],
    "pda_multi_tensor_adam_capturable": [_P, _I, _P, _c.c_double, _c.c_double, _c.c_double, _c.c_double, _P, _P, _P, _P],
    "pda_fcomb_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "pda_fcomb_bwd_fp32": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
}
_RESTYPES = {"pda_error_string": _c.c_char_p, "pda_launch_count": _c.c_longlong, "pda_reset_launch_count": None,
             "pda_fcomb_scratch_floats": _c.c_longlong, "pda_conv3x3_wgrad_scratch_floats": _c.c_longlong, "pda_conv3x3_wgrad_det_scratch_floats": _c.c_longlong, "pda_conv3x3_first_bwd_scratch_floats": _c.c_longlong}

_lib = None


class PdaError(RuntimeError):
    pass


def load():
    """Loads the library once; raises if it is not built (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PdaError(
            f"{LIB_PATH} not found: build it with `python -m probabilistic_domain_adaptation_b200.build` "
            "(the package has no CPU or eager-PyTorch fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _I)
    if lib.pda_abi_version() != 1:
        raise PdaError("libpda_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().pda_error_string(code).decode()
        raise PdaError(f"{what or 'libpda_b200'} failed: {msg} ({code})")
