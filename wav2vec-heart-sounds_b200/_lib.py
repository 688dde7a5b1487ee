"""ctypes binding of ``libmpcg_b200.so`` (C ABI declared in ``include/mpcg_b200.h``).

The shared object is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc`` with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no fallback of any kind: if the
library is missing, or a tensor is not a CUDA float32 tensor, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess

import torch

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = _HERE / "libmpcg_b200.so"
_lib = None

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int

_SIGNATURES = {
    "mpcg_abi_version": (c_int, []),
    "mpcg_error_string": (ctypes.c_char_p, [c_int]),
    "mpcg_biquad_cascade_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, ctypes.c_void_p, c_int, ctypes.c_void_p]),
    "mpcg_biquad_cascade_masked_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, ctypes.c_void_p, c_int, c_f32p,
                                               ctypes.c_void_p]),
    "mpcg_resample_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_i64, ctypes.c_void_p, c_int, c_int, c_int, c_i64,
                                  ctypes.c_void_p]),
    "mpcg_despike_f32": (c_int, [c_f32p, c_i64, c_i64, c_i64, ctypes.c_double, c_int, c_int, ctypes.c_void_p,
                                 ctypes.c_void_p, c_int, ctypes.c_void_p]),
    "mpcg_fill_nans_f32": (c_int, [c_f32p, c_i64, c_int, c_i64, ctypes.c_void_p, ctypes.c_void_p]),
    "mpcg_absmax_norm_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, ctypes.c_void_p]),
    "mpcg_segment_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_int,
                                 ctypes.c_void_p]),
    "mpcg_window_count": (c_i64, [c_i64, c_i64, c_i64, c_i64]),
    "mpcg_preprocess_segment_f32": (c_int, [c_f32p, c_f32p, c_i64, c_int, ctypes.c_void_p, ctypes.c_void_p, c_i64,
                                            ctypes.c_void_p, ctypes.c_void_p, c_int, ctypes.c_void_p]),
    "mpcg_preprocess_segment_work_bytes": (c_i64, [c_i64]),
    "mpcg_debug_set_phase_clock_buffer": (None, [ctypes.c_void_p]),
    "mpcg_mel_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.c_void_p,
                             c_int, c_f32p, c_int, c_i64, c_int, ctypes.c_void_p]),
    "mpcg_aug_chain_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, ctypes.c_float, c_f32p, c_f32p, c_f32p, ctypes.c_uint64,
                                   ctypes.c_uint64, c_f32p, c_f32p, ctypes.c_void_p, c_int, c_f32p, c_f32p, c_f32p, c_f32p,
                                   ctypes.c_uint64, ctypes.c_uint64, c_int, ctypes.c_void_p, c_i64, ctypes.c_void_p]),
    "mpcg_aug_chain_work_bytes": (c_i64, []),
    "mpcg_aug_draw_f32": (c_int, [c_f32p, c_f32p, c_i64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                  ctypes.c_uint64, ctypes.c_void_p]),
    "mpcg_sosfiltfilt_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_i64, ctypes.c_void_p, c_int, ctypes.c_void_p, c_i64,
                                     ctypes.c_void_p]),
    "mpcg_sosfiltfilt_epi_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_i64, ctypes.c_void_p, c_int, ctypes.c_void_p, c_i64,
                                         c_int, ctypes.c_void_p]),
    "mpcg_hilbert_work_bytes": (c_i64, [c_i64, c_i64]),
    "mpcg_hilbert_envelope_f32": (c_int, [c_f32p, c_f32p, ctypes.c_void_p, c_i64, c_i64, c_i64, c_int, ctypes.c_void_p]),
    "mpcg_gen_condition_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_i64, c_i64, c_int, ctypes.c_double, c_int,
                                       ctypes.c_void_p]),
    "mpcg_row_normalise_f32": (c_int, [c_f32p, c_f32p, ctypes.c_void_p, c_i64, c_i64, c_int, c_int, ctypes.c_double,
                                       ctypes.c_double, c_int, ctypes.c_void_p]),
    "mpcg_gen_condition_rows_f32": (c_int, [c_f32p, c_f32p, c_f32p, ctypes.c_void_p, c_i64, c_i64, c_i64, c_int,
                                            ctypes.c_double, c_int, ctypes.c_void_p]),
    "mpcg_cycle_rebuild_f32": (c_int, [c_f32p, c_f32p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       c_i64, c_i64, c_i64, c_int, c_i64, c_int, ctypes.c_void_p]),
    "mpcg_mel_dm_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.c_void_p,
                                c_f32p, ctypes.c_void_p, c_int, c_i64, c_int, ctypes.c_void_p]),
    "mpcg_mel_tc_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, ctypes.c_void_p, c_f32p,
                                ctypes.c_float, c_int, c_i64, c_int, ctypes.c_void_p]),
    "mpcg_logmap_f32": (c_int, [c_f32p, c_f32p, c_i64, ctypes.c_void_p]),
    "mpcg_hpss_stft_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, c_int, c_i64, c_f32p, c_f32p, ctypes.c_void_p]),
    "mpcg_hpss_median_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, c_int, c_int, ctypes.c_void_p]),
    "mpcg_hpss_istft_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_int, c_int, c_i64, ctypes.c_float,
                                    ctypes.c_float, c_f32p, c_f32p, ctypes.c_void_p]),
    "mpcg_hpss_finish_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_int, c_int, c_i64, ctypes.c_void_p]),
    "mpcg_hpss_finish3_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_i64, c_int, c_int, c_i64, ctypes.c_void_p]),
    "mpcg_hpss_mix_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                  ctypes.c_void_p]),
    "mpcg_time_warp_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_i64, ctypes.c_double, ctypes.c_void_p]),
    "mpcg_mix_noise_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_i64, c_i64, c_i64, ctypes.c_void_p, ctypes.c_void_p,
                                   c_f32p, ctypes.c_void_p]),
    "mpcg_beamform_fwd_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_int, c_i64, ctypes.c_void_p, c_int, ctypes.c_void_p]),
    "mpcg_beamform_bwd_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_i64, c_int, c_i64, ctypes.c_void_p, c_int,
                                      ctypes.c_void_p]),
    "mpcg_noise_combine_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_i64, c_i64, c_int, ctypes.c_void_p, ctypes.c_void_p, c_f32p,
                                       c_int, ctypes.c_void_p]),
    "mpcg_aug_stage_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_int, ctypes.c_float, c_f32p, c_f32p, c_f32p, c_int,
                                   ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]),
    "mpcg_aug_warp_f32": (c_int, [c_f32p, c_f32p, c_i64, c_i64, c_f32p, c_int, ctypes.c_void_p]),
    "mpcg_aug_eq_mix_f32": (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_i64, c_f32p, c_int, ctypes.c_void_p]),
}
EUNSUPPORTED = -3
ABI_VERSION = 2

class ChainKind(ctypes.Structure):
    """mpcg_chain_kind (include/mpcg_b200.h)."""
    _fields_ = [("despike", c_int), ("n_sections", c_int), ("sos", (ctypes.c_double * 6) * 2)]


class ChainDesc(ctypes.Structure):
    """mpcg_chain_desc (include/mpcg_b200.h)."""
    _fields_ = [("t_in", c_i64), ("t_out", c_i64), ("up", c_int), ("down", c_int), ("taps_per_phase", c_int),
                ("offset", c_i64), ("taps", ctypes.c_void_p), ("despike_win", c_i64),
                ("despike_threshold", ctypes.c_double), ("despike_max_iterations", c_int), ("median_mode", c_int),
                ("norm_flags", c_int), ("seg_start", c_i64), ("seg_win", c_i64), ("seg_hop", c_i64),
                ("seg_n", c_i64), ("channels_last", c_int), ("n_kinds", c_int), ("kinds", ChainKind * 2),
                ("kind_of_channel", ctypes.c_uint8 * 8), ("row_t_in", ctypes.c_void_p), ("row_t_out", ctypes.c_void_p),
                ("row_out_offset", ctypes.c_void_p), ("plane_elems", c_i64)]


MEDIAN_LOWER, MEDIAN_MEAN = 0, 1
NORM_NAN_TO_NUM, NORM_PEAK_GT0 = 1, 2
RN_MINMAX, RN_ZSCORE, RN_KPEAK = 0, 1, 2
RN_EPS, RN_GLOBAL = 1, 2
EPI_NONE, EPI_EXP = 0, 1
ENV_LOG = 1
GEN_NO_NORM = 4
AUG_IDENTITY, AUG_NOISE, AUG_SINE_MUL, AUG_SINE_ADD, AUG_SELECT = 0, 1, 2, 3, 4


def build(verbose: bool = False) -> pathlib.Path:
    """Compile every ``csrc/*.cu`` for sm_100a into ``libmpcg_b200.so`` (idempotent: make decides)."""
    proc = subprocess.run(["make", "-C", str(_HERE / "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("building libmpcg_b200.so failed (see output above)")
    return LIB_PATH


def declared_symbols() -> list[str]:
    return list(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for this path)")
        # MPCG_B200_LIB: tools/ only -- load an experimental build of the same ABI (A/B timing of kernel variants)
        handle = ctypes.CDLL(os.environ.get("MPCG_B200_LIB", str(LIB_PATH)))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)           # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        if handle.mpcg_abi_version() != ABI_VERSION:
            raise RuntimeError("libmpcg_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code == 0:
        return
    msg = lib().mpcg_error_string(code).decode()
    if code < 0:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: CUDA error {code}: {msg}")


def require_cuda_f32(x: torch.Tensor, name: str = "x") -> torch.Tensor:
    if not torch.is_tensor(x):
        raise TypeError(f"{name} must be a torch.Tensor on a CUDA device (got {type(x).__name__})")
    if not x.is_cuda:
        raise ValueError(f"{name} is on {x.device}: this path runs on B200 only and has no CPU fallback")
    if x.dtype != torch.float32:
        raise ValueError(f"{name} has dtype {x.dtype}: the CUDA path is float32 in / float32 out")
    return x.contiguous()


def stream_ptr(x: torch.Tensor) -> int:
    return torch.cuda.current_stream(x.device).cuda_stream


def ptr(x) -> int:
    return 0 if x is None else x.data_ptr()


_workspaces: dict = {}


def aug_workspace(x: torch.Tensor) -> torch.Tensor:
    """Scratch of the fused augmentation chain (its EQ recipe), per (device, stream) like every workspace here."""
    return workspace(x, int(lib().mpcg_aug_chain_work_bytes()), tag="aug")


def workspace(x: torch.Tensor, nbytes: int, tag: str = "pre") -> torch.Tensor:
    """Device scratch for entry points that take a caller-provided workspace (the C ABI never allocates).  One grow-only
    buffer per (device, current stream): launches that may overlap on different streams never share scratch."""
    key = (tag, x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=x.device)
        _workspaces[key] = buf
    return buf
