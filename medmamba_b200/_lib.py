"""ctypes loader of the in-tree CUDA library (libmedmamba_b200.so).

There is no fallback: if the shared object is missing and cannot be built, or a call returns a
non-zero status, this raises.  PyTorch is only the owner of device memory and streams here.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

from . import build as _build

MMB_F32, MMB_BF16, MMB_F16 = 0, 1, 2
_DTYPES = {torch.float32: MMB_F32, torch.bfloat16: MMB_BF16, torch.float16: MMB_F16}
ABI_VERSION = 4

_lock = threading.Lock()
_lib = None


class MedMambaLibraryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """The loaded library (built on first use if the in-tree .so is absent or stale).

    No silent fallback: when the sources changed and the rebuild fails, this raises even if an older .so is lying
    around (ctypes checks neither arity nor types, so stale kernels behind new call sites corrupt memory).  The
    library carries the digest of the sources it was built from (``mmb_source_digest``); a mismatch raises too.
    ``MMB_ALLOW_PREBUILT=1`` is the explicit opt-in for running a prebuilt library without a toolchain."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = _build.LIB_PATH
                allow_prebuilt = os.environ.get("MMB_ALLOW_PREBUILT", "0") == "1"
                if not _build.up_to_date():
                    try:
                        path = _build.build()
                    except Exception as e:
                        if not (allow_prebuilt and os.path.exists(path)):
                            raise MedMambaLibraryError(
                                f"libmedmamba_b200.so is missing or stale and could not be rebuilt: {e} "
                                "(set MMB_ALLOW_PREBUILT=1 to load an existing library anyway)") from e
                        import warnings
                        warnings.warn(f"medmamba_b200: rebuild failed ({e}); loading the PREBUILT library at {path} "
                                      "because MMB_ALLOW_PREBUILT=1", RuntimeWarning)
                handle = ctypes.CDLL(path)
                handle.mmb_status_string.restype = ctypes.c_char_p
                if handle.mmb_abi_version() != ABI_VERSION:
                    raise MedMambaLibraryError("libmedmamba_b200.so has a different ABI version; rebuild it")
                try:
                    handle.mmb_source_digest.restype = ctypes.c_char_p
                    built_from = handle.mmb_source_digest().decode()
                except AttributeError:
                    built_from = "absent"
                if built_from != _build._digest() and not allow_prebuilt:
                    raise MedMambaLibraryError(
                        f"libmedmamba_b200.so was built from other sources (digest {built_from[:12]}...); "
                        "run python -m medmamba_b200.build --force")
                _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().mmb_status_string(ctypes.c_int(status)).decode()
        raise MedMambaLibraryError(f"{what} failed with status {status}: {msg}")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; expected float32, bfloat16 or float16") from None


def ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def i64(v) -> ctypes.c_int64:
    return ctypes.c_int64(int(v))


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("medmamba_b200 kernels run on CUDA tensors only (there is no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("all tensors must be on the same CUDA device")
    return dev


# ------------------------------------------------------------------------ launch accounting
class KernelTimer:
    """Counts the launches of this library's kernels and, with CUDA events recorded on the launch
    stream either side of each one, their device time (bench.py's live roofline measurement)."""

    def __init__(self, timing: bool = True):
        self.timing = timing
        self.launches = 0
        self._events = []

    def add(self, key, start, end):
        self._events.append((key, start, end))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for key, a, b in self._events:
            d = out.setdefault(key, {"count": 0, "total_ms": 0.0})
            d["count"] += 1
            d["total_ms"] += a.elapsed_time(b)
        for d in out.values():
            d["avg_ms"] = d["total_ms"] / d["count"]
        return out


_timer = None


def set_kernel_timer(timer) -> None:
    global _timer
    _timer = timer


class timed_launch:
    """with timed_launch("kernel", "shape key"): <one C-ABI call that launches one kernel>"""

    __slots__ = ("key", "start")

    def __init__(self, name: str, shape: str = ""):
        self.key = f"{name}[{shape}]" if shape else name
        self.start = None

    def __enter__(self):
        t = _timer
        if t is not None:
            t.launches += 1
            if t.timing:
                self.start = torch.cuda.Event(enable_timing=True)
                self.start.record()
        return self

    def __exit__(self, *exc):
        if self.start is not None and _timer is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _timer.add(self.key, self.start, end)
        return False
