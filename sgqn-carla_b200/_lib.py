"""ctypes binding of libsgqn_b200.so (include/sgqn_b200.h).  There is no CPU fallback: a missing or
stale library is a hard error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGQN_LIB") or os.path.join(_HERE, "libsgqn_b200.so")      # SGQN_LIB: an instrumented debug build
ABI_VERSION = 2

_p, _i, _ll, _f, _d, _ull = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_ulonglong

# name -> argtypes, exactly the prototypes of include/sgqn_b200.h
SIGNATURES = {
    "sgqn_replay_gather": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "sgqn_frames_copy": [_p, _p, _p, _p, _i, _i, _p],
    "sgqn_take_rows": [_p, _p, _p, _i, _i, _p],
    "sgqn_crop_shift": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_zero": [_p, _ll, _p],
    "sgqn_linear_fwd": [_p, _i, _ll, _p, _ll, _p, _ll, _p, _i, _ll, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_linear_dgrad": [_p, _i, _ll, _p, _ll, _p, _i, _ll, _p, _i, _ll, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_linear_wgrad": [_p, _i, _ll, _p, _i, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _i, _p],
    "sgqn_colsum": [_p, _i, _i, _i, _p, _p],
    "sgqn_linear_fwd_tc": [_p, _i, _ll, _p, _ll, _p, _ll, _p, _i, _ll, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_linear_dgrad_tc": [_p, _i, _ll, _p, _ll, _p, _i, _ll, _p, _i, _ll, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_linear_wgrad_tc": [_p, _i, _ll, _p, _i, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _i, _p],
    "sgqn_conv_fwd": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_conv_dgrad": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_conv_wgrad": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_conv1_fwd": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "sgqn_conv_tc": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_conv_chain": [_p, _i, _p, _ll, _p],
    "sgqn_conv_wgrad_tc": [_p, _p, _p, _i, _i, _i, _p],
    "sgqn_conv_weights_prep": [_p, _ll, _p, _p, _i, _p],
    "sgqn_pad_copy": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_conv1_wgrad": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "sgqn_conv1_dgrad": [_p, _p, _p, _i, _i, _i, _p],
    "sgqn_conv1_im2col": [_p, _p, _i, _i, _p],
    "sgqn_conv1_fwd_col": [_p, _p, _p, _p, _i, _i, _p],
    "sgqn_conv1_wgrad_col": [_p, _p, _p, _p, _i, _p],
    "sgqn_conv1_dgrad_col": [_p, _p, _p, _p, _i, _p],
    "sgqn_conv_tcg": [_p, _p, _p, _p, _p] + [_i] * 15 + [_p],
    "sgqn_conv_tcg_taps": [_p, _p, _p, _p, _p] + [_i] * 16 + [_p],
    "sgqn_gemm_wgrad_tcg": [_p, _p, _p] + [_i] * 9 + [_p],
    "sgqn_conv1_im2col96": [_p, _p, _i, _i, _p],
    "sgqn_conv1_weights_prep": [_p, _p, _p, _p],
    "sgqn_conv1_col2im": [_p, _i, _p, _i, _p],
    "sgqn_conv1_fused_tc": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "sgqn_conv1_dgrad_fused_tc": [_p, _p, _p, _i, _p],
    "sgqn_conv_weights_prep_g": [_p, _p, _p, _i, _i, _i, _p],
    "sgqn_conv_wgrad_tcg": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_conv_weights_prep_phase": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "sgqn_conv_phase_fold": [_p, _p, _p, _p, _i, _i, _i, _p],
    "sgqn_bce_phase": [_p, _p, _p, _p] + [_i] * 9 + [_p],
    "sgqn_conv_wgrad_tcg_ld": [_p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "sgqn_pool2_bwd": [_p, _p, _p, _i, _i, _i, _i, _p],
    "sgqn_upsample2_bwd": [_p, _p, _p, _i, _i, _i, _i, _p],
    "sgqn_p2p_layout": [_p],
    "sgqn_p2p_allreduce_sum": [_p, _i, _i, _ll, _ll, _i, _ll, _ll, _i, _ll, _ll, _p],
    "sgqn_p2p_small": [_p, _i, _i, _ll, _ll, _ll, _i, _p, _p, _i, _i, _p],
    "sgqn_minmax": [_p, _ll, _p, _p, _p],
    "sgqn_attribution_mask": [_p, _p, _p, _p, _f, _p, _p, _i, _i, _i, _p],
    "sgqn_overlay_u8": [_p, _p, _p, _f, _f, _p, _i, _i, _p],
    "sgqn_overlay_f32": [_p, _p, _p, _f, _f, _p, _i, _i, _p],
    "sgqn_ln_tanh_fwd": [_p, _p, _p, _p, _i, _i, _i, _p],
    "sgqn_ln_tanh_bwd": [_p, _i, _p, _p, _i, _p, _p, _p, _p, _i, _i, _p],
    "sgqn_set_cols": [_p, _i, _i, _p, _i, _i, _i, _p],
    "sgqn_actor_head_fwd": [_p, _p, _f, _f, _p, _p, _i, _p, _p, _i, _i, _p],
    "sgqn_actor_head_bwd": [_p, _p, _p, _i, _p, _f, _f, _p, _i, _i, _i, _p],
    "sgqn_critic_loss": [_p, _ll, _p, _p, _p, _p, _p, _p, _f, _i, _f, _f, _p, _p, _p, _i, _i, _p],
    "sgqn_actor_loss": [_p, _ll, _p, _p, _f, _p, _p, _p, _i, _i, _p],
    "sgqn_bce": [_p, _p, _p, _p] + [_i] * 10 + [_p],
    "sgqn_ce_diag": [_p, _i, _p, _p, _i, _i, _i, _p],
    "sgqn_mse_loss": [_p, _p, _p, _p, _i, _i, _i, _p],
    "sgqn_bn_relu_fwd": [_p, _p, _p, _p, _p, _i, _i, _p],
    "sgqn_bn_relu_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "sgqn_soda_loss": [_p, _p, _p, _p, _i, _i, _i, _p],
    "sgqn_adam_prep": [_p, _p, _d, _d, _p],
    "sgqn_adam": [_p, _p, _p, _p, _ll, _p, _f, _f, _f, _f, _f, _p, _ll, _f, _f, _f, _p],
    "sgqn_ema": [_p, _p, _ll, _ll, _f, _f, _p],
    "sgqn_alpha_adam": [_p, _p, _p, _p, _d, _d, _d, _d, _p],
    "sgqn_rng_step": [_ull, _p, _p, _p, _p, _i, _p, _i, _p, _p, _p, _i, _i, _ull, _p],
}



class ConvLayer(C.Structure):
    """sgqn_conv_layer (include/sgqn_b200.h): one layer of a sgqn_conv_chain call = the arguments of one sgqn_conv_tc call."""
    _fields_ = [(n, _p) for n in ("x", "w", "bias", "mask", "out", "dbias")] + \
               [(n, _i) for n in ("B", "Hr", "Wp", "Hv", "Wv", "shift", "Hq", "Wq", "oy", "ox", "Hm", "Wm", "flags")]


def conv_layers(rows):
    """rows: tuples in sgqn_conv_tc argument order (x, w, bias, mask, out, dbias, B, Hr, Wp, Hv, Wv, shift, Hq, Wq, oy, ox, Hm, Wm,
    flags) -> (ctypes array, its address, n)."""
    arr = (ConvLayer * len(rows))(*[ConvLayer(*r) for r in rows])
    return arr, C.addressof(arr), len(rows)


_lib = None
launch_count = 0          # kernels-launching ABI calls made (bench.py's `gpu_launches` bookkeeping)


class KernelError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KernelError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(sgqn-carla_b200/csrc/build.sh). There is no CPU / PyTorch fallback for the SGSAC update path.")
    lib = C.CDLL(LIB_PATH)
    lib.sgqn_abi_version.restype = _i
    if lib.sgqn_abi_version() != ABI_VERSION:
        raise KernelError("libsgqn_b200.so ABI version mismatch: rebuild")
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _i
    _lib = lib
    return lib


class _Api:
    """`K.<name>(*args)`: calls sgqn_<name>, raises on a non-zero cudaError_t."""

    def __getattr__(self, name):
        fn = getattr(load(), "sgqn_" + name)

        def call(*args):
            global launch_count
            launch_count += 1
            rc = fn(*args)
            if rc != 0:
                raise KernelError(f"sgqn_{name} failed with cudaError {rc}")
        setattr(self, name, call)
        return call


K = _Api()
