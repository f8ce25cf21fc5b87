"""ctypes binding of libvqacore_sm100a.so (include/vqacore.h).

Loading never falls back to anything: if the shared library is missing the import of the
package's compute paths raises, and every compute call on a machine without an sm_100 GPU
returns VQA_ENODEVICE / a CUDA error that `check()` turns into a RuntimeError.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libvqacore_sm100a.so")

ABI_VERSION = 2
MAXG = 8
GLIMPSES = 4
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
MATH_FP32_SIMT, MATH_TF32X3, MATH_TF32, MATH_BF16, MATH_BF16X3 = 0, 1, 2, 3, 4
MATH_BY_NAME = {"fp32": MATH_FP32_SIMT, "fp32_simt": MATH_FP32_SIMT, "tf32x3": MATH_TF32X3, "tf32": MATH_TF32,
                "bf16": MATH_BF16, "bf16x3": MATH_BF16X3}
VQA_OK, VQA_EINVAL, VQA_ECUDA, VQA_ENODEVICE, VQA_EWORKSPACE = 0, -1, -2, -3, -4

fp = C.c_void_p          # device pointers travel as integers
i64 = C.c_int64
PA = fp * MAXG
IA = i64 * MAXG
U32A = C.c_uint32 * MAXG
U64A = C.c_uint64 * MAXG


class Dropout(C.Structure):
    _fields_ = [("p", C.c_float), ("layer", C.c_uint32), ("seed", C.c_uint64), ("seed_dev", fp)]


class LinearFwd(C.Structure):
    _fields_ = [("groups", C.c_int), ("M", i64), ("K", i64), ("N", i64), ("act", C.c_int), ("math", C.c_int),
                ("p", C.c_float), ("seed", C.c_uint64), ("seed_dev", fp),
                ("X", PA), ("ldx", IA), ("W", PA), ("b", PA), ("Y", PA), ("ldy", IA),
                ("layer", U32A), ("drop_index_base", U64A), ("drop_bits", PA), ("Wp", PA), ("workspace", fp),
                ("workspace_bytes", C.c_size_t)]


class LinearBwd(C.Structure):
    _fields_ = [("groups", C.c_int), ("M", i64), ("K", i64), ("N", i64), ("act", C.c_int), ("math", C.c_int),
                ("p", C.c_float), ("seed", C.c_uint64), ("seed_dev", fp), ("accumulate_w", C.c_int),
                ("accumulate_x", C.c_int),
                ("X", PA), ("ldx", IA), ("W", PA), ("Y", PA), ("ldy", IA), ("dY", PA), ("lddy", IA),
                ("dW", PA), ("db", PA), ("dX", PA), ("lddx", IA), ("layer", U32A), ("drop_index_base", U64A),
                ("drop_bits", PA), ("Wp", PA), ("workspace", fp), ("workspace_bytes", C.c_size_t),
                ("pool_alpha", fp), ("pool_dpooled", fp), ("pool_regions", i64)]


class MutanFwd(C.Structure):
    _fields_ = [("R", C.c_int), ("M", i64), ("K1", i64), ("K2", i64), ("F", i64), ("rows_per_h2", i64),
                ("math", C.c_int), ("X1", fp), ("ldx1", i64), ("X2", fp), ("ldx2", i64),
                ("W1", PA), ("b1", PA), ("W2", PA), ("b2", PA), ("H1", fp), ("H2", fp), ("Y", fp), ("ldy", i64),
                ("W1p", fp), ("W2p", fp), ("workspace", fp), ("workspace_bytes", C.c_size_t)]


class MutanBwd(C.Structure):
    _fields_ = [("R", C.c_int), ("M", i64), ("K1", i64), ("K2", i64), ("F", i64), ("rows_per_h2", i64),
                ("math", C.c_int), ("accumulate_w", C.c_int), ("accumulate_x1", C.c_int), ("accumulate_x2", C.c_int),
                ("X1", fp), ("ldx1", i64), ("X2", fp), ("ldx2", i64), ("W1", PA), ("W2", PA),
                ("H1", fp), ("H2", fp), ("dY", fp), ("lddy", i64), ("dH2", fp),
                ("dW1", PA), ("db1", PA), ("dW2", PA), ("db2", PA),
                ("dX1", fp), ("lddx1", i64), ("dX2", fp), ("lddx2", i64),
                ("W1p", fp), ("W2p", fp), ("workspace", fp), ("workspace_bytes", C.c_size_t)]


class BitsSegment(C.Structure):
    _fields_ = [("layer", C.c_uint32), ("n", C.c_uint64), ("out", fp)]


class ParamSegment(C.Structure):
    _fields_ = [("param", fp), ("offset", i64), ("numel", i64)]


class ClipAdam(C.Structure):
    _fields_ = [("nsegs", C.c_int), ("segs", C.POINTER(ParamSegment)), ("grads_flat", fp), ("exp_avg", fp),
                ("exp_avg_sq", fp), ("total", i64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("step", i64), ("max_norm", C.c_float), ("write_clipped_grads", C.c_int),
                ("scratch", fp), ("step_dev", fp), ("lr_dev", fp), ("lr_gamma", C.c_double)]


class PackSegment(C.Structure):
    _fields_ = [("src", fp), ("dst", fp), ("rows", i64), ("rows_pad", i64), ("K", i64)]


class PoolFwd(C.Structure):
    _fields_ = [("B", i64), ("N", i64), ("Ff", i64), ("D", i64), ("drop", Dropout),
                ("fuse", fp), ("Wc", fp), ("bc", fp), ("x", fp), ("alpha", fp), ("pooled", fp),
                ("drop_bits", fp)]


class PoolBwd(C.Structure):
    _fields_ = [("B", i64), ("N", i64), ("Ff", i64), ("D", i64), ("drop", Dropout),
                ("accumulate_w", C.c_int), ("accumulate_x", C.c_int),
                ("fuse", fp), ("Wc", fp), ("x", fp), ("alpha", fp), ("dpooled", fp), ("dalpha0_ext", fp),
                ("dalpha", fp), ("dz", fp), ("dWc", fp), ("dbc", fp), ("dfuse", fp), ("dx", fp),
                ("drop_bits", fp), ("dalpha_ext", fp)]


class CompoundFwd(C.Structure):
    _fields_ = [("B", i64), ("N", i64), ("D", i64), ("x", fp), ("pooled", fp), ("alpha", fp), ("g1", fp), ("g2", fp),
                ("v2", fp), ("v2_planes", fp), ("v2_nplanes", C.c_int), ("v2_plane_stride", i64), ("v2_keep_bits", fp),
                ("v2_keep_scale", C.c_float)]


class CompoundBwd(C.Structure):
    _fields_ = [("B", i64), ("N", i64), ("D", i64), ("x", fp), ("pooled", fp), ("alpha", fp), ("g1", fp), ("g2", fp),
                ("dv2", fp), ("dg1", fp), ("dg2", fp), ("dpooled", fp), ("dalpha0_ext", fp),
                ("dv2_keep_bits", fp), ("dv2_keep_scale", C.c_float), ("dv2_pool_alpha", fp), ("dv2_pool_dpooled", fp)]


class OdaFwd(C.Structure):
    _fields_ = [("B", i64), ("N", i64), ("H", i64), ("D", i64), ("train", C.c_int), ("drop", Dropout),
                ("vl", fp), ("ql", fp), ("W", fp), ("bc", fp), ("x", fp), ("wsum", fp), ("alpha", fp), ("pooled", fp),
                ("workspace", fp), ("workspace_bytes", C.c_size_t), ("keep_bits_ready", C.c_int)]


class OdaBwd(C.Structure):
    _fields_ = [("B", i64), ("N", i64), ("H", i64), ("D", i64), ("train", C.c_int), ("drop", Dropout),
                ("accumulate_w", C.c_int), ("vl", fp), ("ql", fp), ("W", fp), ("x", fp), ("alpha", fp), ("wsum", fp),
                ("dpooled", fp), ("dalpha", fp), ("dz", fp), ("dwsum", fp), ("dW", fp), ("dbc", fp), ("dvl", fp),
                ("dql", fp), ("workspace", fp), ("workspace_bytes", C.c_size_t)]


class KldParams(C.Structure):
    _fields_ = [("B", i64), ("C", i64), ("grad_scale", C.c_float), ("logits", fp), ("target", fp),
                ("loss_rows", fp), ("dlogits", fp)]


class PeerAllreduce(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("buffers", fp * 8), ("signals", fp * 8), ("offset", i64),
                ("count", i64), ("max_ctas", C.c_int), ("spin_limit_ms", C.c_int), ("multicast", fp), ("cta_threads", C.c_int)]


class GruGateFwd(C.Structure):
    _fields_ = [("B", i64), ("H", i64), ("act", C.c_int), ("gi", fp * 3), ("gh", fp * 3), ("h_prev", fp),
                ("hmask", fp * 3), ("h", fp), ("r", fp), ("i", fp), ("n", fp), ("hm", fp * 3)]


class GruGateBwd(C.Structure):
    _fields_ = [("B", i64), ("H", i64), ("act", C.c_int), ("t", i64), ("dh_partial", fp), ("dhm", fp * 3),
                ("hmask", fp * 3), ("dx_last", fp), ("last_pos", fp), ("r", fp), ("i", fp), ("n", fp), ("gh_n", fp),
                ("h_prev", fp), ("da", fp * 3), ("dgh_n", fp), ("dh_partial_out", fp)]


class ModelFwd(C.Structure):
    _fields_ = [("B", i64), ("N", i64), ("C", i64), ("train", C.c_int), ("math", C.c_int), ("seed", C.c_uint64),
                ("seed_dev", fp), ("v", fp), ("q", fp), ("params", C.POINTER(fp)), ("logits", fp), ("alpha1", fp), ("alpha2", fp),
                ("v2", fp), ("workspace", fp), ("workspace_bytes", C.c_size_t)]


class ModelBwd(C.Structure):
    _fields_ = [("fwd", ModelFwd), ("dlogits", fp), ("grads", C.POINTER(fp)), ("accumulate", C.c_int),
                ("grads_flat", fp), ("grads_flat_bytes", C.c_size_t), ("group_events", fp * 16), ("dq", fp)]


STRUCTS = {
    "vqa_dropout": Dropout, "vqa_pack_segment": PackSegment, "vqa_bits_segment": BitsSegment, "vqa_param_segment": ParamSegment, "vqa_clip_adam_params": ClipAdam, "vqa_linear_fwd_params": LinearFwd, "vqa_linear_bwd_params": LinearBwd,
    "vqa_mutan_fwd_params": MutanFwd, "vqa_mutan_bwd_params": MutanBwd,
    "vqa_region_softmax_pool_fwd_params": PoolFwd, "vqa_region_softmax_pool_bwd_params": PoolBwd,
    "vqa_cor_compound_fwd_params": CompoundFwd, "vqa_cor_compound_bwd_params": CompoundBwd,
    "vqa_oda_pair_attn_fwd_params": OdaFwd, "vqa_oda_pair_attn_bwd_params": OdaBwd,
    "vqa_kld_logsoftmax_params": KldParams, "vqa_peer_allreduce_params": PeerAllreduce, "vqa_gru_gate_fwd_params": GruGateFwd, "vqa_gru_gate_bwd_params": GruGateBwd,
    "vqa_model_fwd_params": ModelFwd, "vqa_model_bwd_params": ModelBwd,
}

# every symbol include/vqacore.h declares: name -> (restype, argtypes)
_OP = lambda s: (C.c_int, [C.POINTER(s), C.c_void_p])
SYMBOLS = {
    "vqa_abi_version": (C.c_int, []),
    "vqa_last_error": (C.c_char_p, []),
    "vqa_device_check": (C.c_int, []),
    "vqa_sizeof": (C.c_size_t, [C.c_char_p]),
    "vqa_launch_count": (C.c_ulonglong, []),
    "vqa_profile_begin": (C.c_int, []),
    "vqa_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
    "vqa_pack_weights": (C.c_int, [C.POINTER(PackSegment), C.c_int, C.c_void_p]),
    "vqa_dropout_bits": (C.c_int, [C.c_float, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p,
                                   C.c_void_p]),
    "vqa_dropout_bits_batch": (C.c_int, [C.c_float, C.c_uint64, C.c_void_p, C.POINTER(BitsSegment), C.c_int,
                                         C.c_void_p]),
    "vqa_seed_advance": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vqa_grad_groups": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int]),
    "vqa_clip_adam_step": _OP(ClipAdam),
    "vqa_linear_fwd": _OP(LinearFwd), "vqa_linear_bwd": _OP(LinearBwd),
    "vqa_linear_fwd_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, i64, i64, i64]),
    "vqa_linear_bwd_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, i64, i64, i64]),
    "vqa_mutan_fwd": _OP(MutanFwd), "vqa_mutan_bwd": _OP(MutanBwd),
    "vqa_mutan_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, i64, i64, i64, i64, i64, C.c_int]),
    "vqa_region_softmax_pool_fwd": _OP(PoolFwd), "vqa_region_softmax_pool_bwd": _OP(PoolBwd),
    "vqa_cor_compound_fwd": _OP(CompoundFwd), "vqa_cor_compound_bwd": _OP(CompoundBwd),
    "vqa_oda_pair_attn_fwd": _OP(OdaFwd), "vqa_oda_pair_attn_bwd": _OP(OdaBwd),
    "vqa_kld_logsoftmax_fwd_bwd": _OP(KldParams),
    "vqa_peer_allreduce_signal_bytes": (C.c_size_t, []),
    "vqa_peer_allreduce_f32": _OP(PeerAllreduce),
    "vqa_seq_dropout_masks": (C.c_int, [C.c_float, C.c_uint64, C.c_void_p, C.c_uint32, i64, i64, C.c_int, C.c_void_p,
                                        C.c_void_p]),
    "vqa_gru_embed_fwd": (C.c_int, [i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_gru_embed_bwd": (C.c_int, [i64, i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_gru_gate_fwd": _OP(GruGateFwd), "vqa_gru_gate_bwd": _OP(GruGateBwd),
    "vqa_gru_last_pos": (C.c_int, [i64, i64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_gru_select_last": (C.c_int, [i64, i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_cast_bf16_f32": (C.c_int, [i64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_argmax_rows": (C.c_int, [i64, i64, C.c_void_p, C.c_void_p, i64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_sum_rows": (C.c_int, [i64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_scale_by_device_scalar": (C.c_int, [i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vqa_cor2_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "vqa_cor2_fwd": _OP(ModelFwd), "vqa_cor2_bwd": _OP(ModelBwd),
    "vqa_oda_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "vqa_oda_pair_attn_workspace_bytes": (C.c_size_t, [i64, i64, i64]),
    "vqa_oda_fwd": _OP(ModelFwd), "vqa_oda_bwd": _OP(ModelBwd),
    "vqa_stash_info": (C.c_int, [C.c_int, C.c_char_p, i64, i64, i64, C.POINTER(C.c_size_t), C.POINTER(i64),
                                 C.POINTER(i64), C.POINTER(i64)]),
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s not found: build it with `python vqa-playground-pytorch_b200/build.py` "
                "(or __graft_entry__.build()). There is no CPU or PyTorch fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)           # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if L.vqa_abi_version() != ABI_VERSION:
            raise RuntimeError("libvqacore ABI version %d, binding expects %d" % (L.vqa_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != VQA_OK:
        msg = lib().vqa_last_error().decode("utf-8", "replace")
        kind = {VQA_EINVAL: ValueError}.get(rc, RuntimeError)
        raise kind("libvqacore %s failed (%d): %s" % (what, rc, msg))
