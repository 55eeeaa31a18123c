"""ctypes binding of libvince_b200.so (the C ABI declared in include/vince_b200.h).

There is no CPU or PyTorch fallback: if the library cannot be loaded `lib()` raises, and every op in
vince_b200.ops refuses non-CUDA tensors.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VINCE_B200_LIB", os.path.join(_HERE, "csrc", "libvince_b200.so"))


class ConvDesc(Structure):
    _fields_ = [
        ("a_hi", c_void_p), ("a_lo", c_void_p), ("w_hi", c_void_p), ("w_lo", c_void_p), ("out", c_void_p),
        ("M", c_int32), ("N", c_int32), ("K", c_int32), ("im2col", c_int32),
        ("batch", c_int32), ("H", c_int32), ("W", c_int32), ("Cin", c_int32), ("R", c_int32), ("S", c_int32),
        ("stride", c_int32), ("pad_lo_h", c_int32), ("pad_lo_w", c_int32), ("pad_hi_h", c_int32),
        ("pad_hi_w", c_int32), ("passes", c_int32), ("block_n", c_int32),
        ("scale", c_void_p), ("bias", c_void_p), ("relu", c_int32), ("halo_mode", c_int32), ("stats", c_void_p),
        ("bn_gamma", c_void_p), ("bn_beta", c_void_p), ("bn_running_mean", c_void_p), ("bn_running_var", c_void_p),
        ("bn_num_batches_tracked", c_void_p), ("bn_coef", c_void_p), ("bn_counter", c_void_p),
        ("bn_momentum", c_float), ("bn_eps", c_float),
        ("a_pixel_stride", c_int64), ("a_row_stride", c_int64), ("a_img_stride", c_int64),
        ("alpha", c_float), ("stats_only", c_int32),
        ("out_hi", c_void_p), ("out_lo", c_void_p), ("ep_coef", c_void_p), ("res_kind", c_int32), ("reserved", c_int32),
        ("res_hi", c_void_p), ("res_lo", c_void_p), ("res_raw", c_void_p), ("res_coef", c_void_p),
        ("bn_save", c_int32), ("kchunk", c_int32), ("taps", c_int32), ("shift_w", c_int32), ("alpha_dev", c_void_p),
    ]


class BnSide(Structure):
    _fields_ = [("raw", c_void_p), ("coef", c_void_p)]


class BnBwdDesc(Structure):
    _fields_ = [
        ("dA", c_void_p), ("dB", c_void_p), ("bcast_hw", c_int32), ("mask_kind", c_int32),
        ("out_hi", c_void_p), ("out_lo", c_void_p), ("raw", c_void_p), ("coef", c_void_p),
        ("M", c_int64), ("C", c_int32), ("work", c_void_p), ("dgamma", c_void_p), ("dbeta", c_void_p),
        ("accumulate", c_int32), ("d_hi", c_void_p), ("d_lo", c_void_p), ("d_f32", c_void_p), ("dz_out", c_void_p),
        ("dil", c_int32), ("P", c_int32), ("Q", c_int32), ("Hd", c_int32), ("Wd", c_int32),
    ]


class InfoNceDesc(Structure):
    _fields_ = [
        ("q", c_void_p), ("keys", c_void_p), ("queue_tf32", c_void_p),
        ("B", c_int32), ("Bk", c_int32), ("K", c_int32), ("D", c_int32), ("num_frames", c_int32),
        ("temperature", c_float),
        ("dists", c_void_p), ("weights", c_void_p), ("pos_sim", c_void_p), ("neg_max", c_void_p),
        ("row_lse", c_void_p), ("scalars", c_void_p), ("workspace", c_void_p),
    ]


# name -> (restype, argtypes); mirrors include/vince_b200.h one to one
SIGNATURES = {
    "vince_last_error": (c_char_p, []),
    "vince_abi_version": (c_int32, []),
    "vince_conv_fwd": (c_int32, [POINTER(ConvDesc), c_void_p]),
    "vince_stem_pack": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "vince_stem_pack_u8": (c_int32, [c_void_p, c_void_p, POINTER(c_float), POINTER(c_float), c_void_p, c_void_p, c_int32,
                                     c_int32, c_int32, c_void_p]),
    "vince_stem_pack_grid": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                       c_void_p]),
    "vince_stem_pack_u8_grid": (c_int32, [c_void_p, c_void_p, POINTER(c_float), POINTER(c_float), c_void_p, c_void_p,
                                          c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vince_weight_prep": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "vince_bn_eval_coef": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int32, c_void_p]),
    "vince_bn_apply": (c_int32, [POINTER(BnSide), c_int32, c_void_p, c_void_p, POINTER(BnSide), c_int32, c_void_p,
                                 c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "vince_bn_relu_maxpool": (c_int32, [POINTER(BnSide), c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                        c_void_p]),
    "vince_bn_final_pool": (c_int32, [POINTER(BnSide), c_int32, c_void_p, c_void_p, POINTER(BnSide), c_void_p,
                                      c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "vince_split_f16": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vince_count_saturated": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "vince_round_tf32": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vince_l2_normalize": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "vince_jigsaw_patchify": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vince_jigsaw_gather": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "vince_infonce_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "vince_infonce_fwd": (c_int32, [POINTER(InfoNceDesc), c_void_p]),
    "vince_infonce_bwd_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "vince_infonce_bwd": (c_int32, [POINTER(InfoNceDesc), c_float, c_int32, c_int32, c_void_p, c_void_p]),
    "vince_masked_ce_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vince_ema_enqueue": (c_int32, [c_void_p, c_int32, c_float, c_float, c_void_p, c_void_p, c_void_p, c_int64,
                                    c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "vince_bn_bwd": (c_int32, [POINTER(BnBwdDesc), c_void_p]),
    "vince_transpose_pad": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32,
                                      c_int32, c_int32, c_int32, c_int64, c_int32, c_void_p]),
    "vince_wgrad_reduce": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_float, c_void_p,
                                     c_int32, c_void_p]),
    "vince_maxpool_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                    c_void_p]),
    "vince_stem_wgrad": (c_int32, [c_void_p, c_void_p, c_void_p, POINTER(c_float), POINTER(c_float), c_void_p, c_void_p,
                                   c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vince_sgemm": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                              c_int32, c_int32, c_void_p, c_void_p]),
    "vince_colsum": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "vince_normalize_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_float, c_void_p]),
    "vince_sgd_step": (c_int32, [c_void_p, c_int32, c_float, c_float, c_float, c_float, c_int32, c_void_p]),
    "vince_allreduce_sum": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vince_knn_classify": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "vince_comm_unique_id": (c_int32, [c_void_p]),
    "vince_comm_init": (c_int32, [POINTER(c_void_p), c_void_p, c_int32, c_int32]),
    "vince_comm_destroy": (c_int32, [c_void_p]),
    "vince_allgather_enqueue": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_int64,
                                          c_void_p, c_void_p]),
    "vince_allgather_enqueue_ema": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_int64,
                                              c_void_p, c_void_p, c_int32, c_float, c_float, c_void_p]),
}

_dll = None


def lib():
    """The loaded shared library; raises (never falls back) if it is missing."""
    global _dll
    if _dll is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libvince_b200.so is not built (%s). Run `python -m vince_b200.build`; there is no CPU / PyTorch "
                "fallback for the vince_b200 hot path." % LIB_PATH)
        dll = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(dll, name)
            fn.restype = res
            fn.argtypes = args
        _dll = dll
    return _dll


def last_error():
    msg = lib().vince_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, last_error()))
