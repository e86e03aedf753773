"""ctypes binding of the C-ABI in ``include/sudormrf_b200.h``.

The shared library is built in-tree (``sudo_rm_rf_b200/libsudormrf_b200.so``)
by ``build()`` / ``__graft_entry__.build()``.  There is NO fallback: if the
library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDR_B200_LIB") or os.path.join(_HERE, "libsudormrf_b200.so")   # env override: A/B-testing kernel builds
CSRC = os.path.join(_HERE, "csrc")

ABI_VERSION = 2


class SdrConfig(C.Structure):
    """``sdr_config`` (include/sudormrf_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "variant", "in_audio_channels", "out_channels", "in_channels", "num_blocks",
        "upsampling_depth", "enc_kernel_size", "enc_num_basis", "num_sources", "group_size")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class SdrNormIn(C.Structure):
    """``sdr_norm_in``."""
    _fields_ = [("stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("prelu", C.c_void_p), ("count", C.c_double), ("prelu_per_channel", C.c_int32)]


class NativeError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

_SIGNATURES = {
    "sdr_abi_version": (C.c_int, []),
    "sdr_error_string": (C.c_char_p, [C.c_int]),
    "sdr_num_params": (C.c_int, [C.POINTER(SdrConfig)]),
    "sdr_param_numel": (C.c_int64, [C.POINTER(SdrConfig), C.c_int]),
    "sdr_padded_length": (C.c_int64, [C.POINTER(SdrConfig), C.c_int64]),
    "sdr_packed_weight_bytes": (C.c_size_t, [C.POINTER(SdrConfig)]),
    "sdr_pack_weights": (C.c_int, [C.POINTER(SdrConfig), C.POINTER(C.c_void_p), C.c_int,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "sdr_workspace_bytes": (C.c_size_t, [C.POINTER(SdrConfig), C.c_int, C.c_int64]),
    "sdr_forward": (C.c_int, [C.POINTER(SdrConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                              C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sdr_forward_launch_count": (C.c_int, [C.POINTER(SdrConfig)]),
    "sdr_forward_launch_count_at": (C.c_int, [C.POINTER(SdrConfig), C.c_int64]),
    "sdr_host_staging_bytes": (C.c_size_t, [C.POINTER(SdrConfig), C.c_int, C.c_int64]),
    "sdr_forward_host": (C.c_int, [C.POINTER(SdrConfig), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "sdr_mixture_consistency": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "sdr_encoder": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_encoder_mma_packed_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "sdr_encoder_mma_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sdr_encoder_mma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_pointwise": (C.c_int, [C.c_void_p, C.POINTER(SdrNormIn), C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_pointwise_mma_packed_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "sdr_pointwise_mma_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "sdr_pointwise_mma": (C.c_int, [C.c_void_p, C.POINTER(SdrNormIn), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_depthwise": (C.c_int, [C.c_void_p, C.POINTER(SdrNormIn), C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p]),
    "sdr_pyramid_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "sdr_depthwise_pyramid": (C.c_int, [C.c_void_p, C.POINTER(SdrNormIn), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_merge_pyramid": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_causal_pyramid": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_merge": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(SdrNormIn), C.c_int, C.c_void_p,
                            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_tac": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_int,
                          C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_residual_norm": (C.c_int, [C.c_void_p, C.POINTER(SdrNormIn), C.c_void_p, C.POINTER(SdrNormIn), C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_softmax_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "sdr_overlap_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int64, C.c_void_p]),
    "sdr_utterance_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "sdr_separate_workspace_bytes": (C.c_size_t, [C.POINTER(SdrConfig), C.c_int, C.c_int64]),
    "sdr_separate": (C.c_int, [C.POINTER(SdrConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                               C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sdr_separate_ragged": (C.c_int, [C.POINTER(SdrConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sdr_pairwise_neg_sdr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p]),
    "sdr_pit_sisdr_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "sdr_pit_sisdr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "sdr_stabilized_sisdr_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "sdr_stabilized_sisdr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into the in-tree shared library."""
    proc = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise NativeError("building libsudormrf_b200.so failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout)
    return LIB_PATH


def lib():
    """The loaded library (loads on first use; raises if it was never built)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                    "g.build()'` (there is no CPU / eager fallback).")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            if handle.sdr_abi_version() != ABI_VERSION:
                raise NativeError("libsudormrf_b200.so ABI version mismatch; rebuild")
            _lib = handle
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().sdr_error_string(code).decode()
        raise NativeError(f"{what or 'native call'} failed: {msg} (code {code})")
