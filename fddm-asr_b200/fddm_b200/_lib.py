"""ctypes binding of libfddm_b200.so -- the C-ABI declared in include/fddm_b200.h.

The library is the product: every compute entry point of this package goes through it, on CUDA
tensors only.  There is no CPU path and no fallback; if the shared library is missing the import
of this module raises, and passing a CPU tensor to any op raises ValueError.

PyTorch is only the host here: it owns device memory (tensors), streams
(`torch.cuda.current_stream()`) and the process group.  The signatures below carry raw pointers
(`tensor.data_ptr()`), int64 sizes and the raw `cudaStream_t`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDDM_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libfddm_b200.so")

F32, BF16, F16 = 0, 1, 2
JUMP_EXACT, JUMP_SAMPLE, JUMP_WRITE_P, JUMP_DEBUG_W = 0x1, 0x2, 0x4, 0x8
LFD_PLANES_VALID = 0x2
KL_MASK_F32, KL_CLAMP_T = 0x1, 0x2
XGPU_P2P, XGPU_NVLS = 1, 2
MAX_VOCAB = 49152

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"libfddm_b200.so not found at {LIB_PATH}; build it with `make -C fddm-asr_b200/csrc -j8` "
        "(or python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i64, _i32, _f32, _f64, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_uint64

# name -> (restype, argtypes); mirrors include/fddm_b200.h one to one
SIGNATURES = {
    "fddm_version": (_i32, []),
    "fddm_last_error": (C.c_char_p, []),
    "fddm_launch_count": (_i64, []),
    "fddm_set_sm_reserve": (_i32, [_i32]),
    "fddm_profile_enable": (_i32, [_i32]),
    "fddm_profile_read": (_i64, [_vp, _i64]),
    "fddm_q_sample_dense": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp, _vp]),
    "fddm_sample_q_ids": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp, _u64, _u64, _vp, _vp, _vp]),
    "fddm_q_posterior_dense": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp, _vp]),
    "fddm_multistep_coeffs": (_i32, [_vp, _i64, _i64, _vp, _vp, _i64, _i64, _vp, _vp]),
    "fddm_q_posterior_multi_dense": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _f32, _vp, _vp]),
    "fddm_kl_workspace_bytes": (C.c_size_t, [_i64, _i64]),
    "fddm_kl_forward": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _f64, _vp, _vp, _vp]),
    "fddm_kl_forward_backward": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _f64, _vp, _vp,
                                        _vp, _vp, _vp]),
    "fddm_scale_inplace": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp]),
    "fddm_jump_workspace_bytes": (C.c_size_t, [_i64, _i64]),
    "fddm_jump_step": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _f32, _f32, _vp, _u64, _u64,
                              _vp, _vp, _vp, _vp, _vp, _vp]),
    "fddm_edit_distance_workspace_bytes": (C.c_size_t, [_i64, _i64]),
    "fddm_edit_distance": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "fddm_xgpu_signal_pad_bytes": (_i64, []),
    "fddm_xgpu_allreduce": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32, _i32, _vp]),
    "fddm_lfd_workspace_bytes": (C.c_size_t, [_i64, _i64, _i64]),
    "fddm_lfd_stats": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp]),
    "fddm_lfd_xcov": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _f64, _f32, _vp, _vp, _vp]),
    "fddm_lfd_loss": (_i32, [_vp, _i64, _f64, _f32, _vp, _vp, _vp, _vp]),
    "fddm_lfd_bn_parts": (_i64, [_i64, _i64, _i64]),
    "fddm_lfd_backward": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _f64, _f32, _vp, _f64, _vp, _vp, _vp, _i64,
                                 _i32, _vp, _vp, _vp]),
}
MISSING = []                               # symbols of the ABI the loaded library does not export
for _name, (_res, _args) in SIGNATURES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)              # calling it later raises AttributeError (no fallback)
        continue
    _fn.restype = _res
    _fn.argtypes = _args


class FddmError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.fddm_last_error().decode("utf-8", "replace")
        kind = {-1: "EINVAL", -2: "ECUDA", -3: "EUNSUPPORTED"}.get(status, str(status))
        raise FddmError(f"{what} failed: {kind}: {msg}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; expected float32, bfloat16 or float16") from None


def require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    """All given tensors must live on one CUDA device (no CPU fallback: CPU tensors are an error)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("fddm_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on "
                             f"{t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise ValueError("no tensor argument")
    return dev


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(lib.fddm_launch_count())


def set_sm_reserve(n: int) -> None:
    """SMs the persistent row kernels leave free for a collective running concurrently on another stream."""
    check(lib.fddm_set_sm_reserve(int(n)), "set_sm_reserve")


def profile_enable(on: bool) -> None:
    """Per-kernel CUDA-event timing inside the library (measurement aid, see include/fddm_b200.h)."""
    lib.fddm_profile_enable(1 if on else 0)


def profile_read() -> dict:
    """{kernel name: (launches, total milliseconds)} of the launches recorded since profile_enable(True)."""
    need = int(lib.fddm_profile_read(None, 0))
    buf = C.create_string_buffer(need + 16)
    lib.fddm_profile_read(C.cast(buf, C.c_void_p), need + 16)
    out = {}
    for line in buf.value.decode("utf-8", "replace").splitlines():
        name, n, ms = line.split("\t")
        out[name] = (int(n), float(ms))
    return out


_zero_ws = {}


def zeroed_workspace(device: torch.device, key: str, nbytes: int) -> torch.Tensor:
    """A per-(device, stream, key) uint8 buffer that is zero-initialised once.  Kernels that use
    such a buffer as a row scheduler restore it to zero before they exit, so it is reusable by
    stream-ordered calls; distinct streams get distinct buffers."""
    k = (device.index, torch.cuda.current_stream(device).cuda_stream, key)
    buf = _zero_ws.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(nbytes, 256), dtype=torch.uint8, device=device)
        _zero_ws[k] = buf
    return buf
