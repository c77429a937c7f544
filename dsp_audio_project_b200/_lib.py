"""ctypes binding of libdspb200.so (the C ABI declared in include/dspb200.h).

The library is the product: if it is missing, fails to load, or a call fails,
this module raises -- there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdspb200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_ALLOC = range(6)
F32, F64 = 0, 1
FFT_MAX = 1 << 17

c_i64 = C.c_int64
c_p = C.c_void_p
_pi = C.POINTER(C.c_int)
_pi64 = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)

# name -> (restype, argtypes): every symbol include/dspb200.h declares
SIGNATURES = {
    "dspb200_version": (C.c_int, []),
    "dspb200_last_error_string": (C.c_char_p, []),
    "dspb200_device_count": (C.c_int, [_pi]),
    "dspb200_current_device": (C.c_int, [C.POINTER(C.c_int)]),
    "dspb200_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, _pi, _pi, _pi, C.POINTER(C.c_size_t)]),
    "dspb200_design_sinc_taps": (C.c_int, [C.c_double, C.c_int, _pd, C.c_int, _pi]),
    "dspb200_design_src_filter": (C.c_int, [C.c_int, C.c_int, _pd, C.c_int, _pi]),
    "dspb200_design_peaking_biquad": (C.c_int, [C.c_double, C.c_double, C.c_double, _pd, _pd]),
    "dspb200_eq_select_sections": (C.c_int, [C.c_double, _pd, _pd, C.c_int, _pd, _pd, _pi, _pi]),
    "dspb200_src_geometry": (C.c_int, [C.c_int, C.c_int, c_i64, _pi, _pi64, _pi64]),
    "dspb200_src_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(c_p)]),
    "dspb200_src_plan_destroy": (C.c_int, [c_p]),
    "dspb200_src_run_f32": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_src_run_f64": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_src_plan_kernel_kind": (C.c_int, [c_p, c_i64, c_i64, c_i64, _pi]),
    "dspb200_src_host_f32": (C.c_int, [C.c_int, C.c_int, c_p, c_i64, c_i64, c_p, c_i64, _pi64]),
    "dspb200_src_host_f64": (C.c_int, [C.c_int, C.c_int, c_p, c_i64, c_i64, c_p, c_i64, _pi64]),
    "dspb200_src_plan_host_f32": (C.c_int, [c_p, c_p, c_i64, c_i64, c_p, c_i64, _pi64]),
    "dspb200_src_plan_host_f64": (C.c_int, [c_p, c_p, c_i64, c_i64, c_p, c_i64, _pi64]),
    "dspb200_eq_plan_create": (C.c_int, [C.c_double, _pd, _pd, C.c_int, C.c_int, C.c_int, C.POINTER(c_p)]),
    "dspb200_eq_plan_create_raw": (C.c_int, [_pd, C.c_int, C.c_int, C.c_int, C.POINTER(c_p)]),
    "dspb200_eq_plan_create_bands": (C.c_int, [C.c_double, _pd, C.c_int, C.POINTER(c_p), _pi]),
    "dspb200_eq_plan_destroy": (C.c_int, [c_p]),
    "dspb200_eq_run_f32": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_eq_run_f64": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_eq_stream_chunk": (C.c_int, []),
    "dspb200_eq_run_stream_f32": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p, C.c_int, c_p]),
    "dspb200_eq_host_f32": (C.c_int, [c_p, c_p, c_p, c_i64, c_i64]),
    "dspb200_eq_host_f64": (C.c_int, [c_p, c_p, c_p, c_i64, c_i64]),
    "dspb200_eq_plan_kernel_kind": (C.c_int, [c_p, c_i64, c_i64, c_i64, _pi]),
    "dspb200_eq_plan_chunk_system": (C.c_int, [c_p, _pi, _pi, _pd, _pd, _pd]),
    "dspb200_eq_plan_warm_chunks": (C.c_int, [c_p, C.POINTER(C.c_int)]),
    "dspb200_eq_plan_describe": (C.c_int, [c_p, _pi, _pd, C.c_int]),
    "dspb200_fft_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(c_p)]),
    "dspb200_fft_plan_destroy": (C.c_int, [c_p]),
    "dspb200_fft_workspace_bytes": (C.c_int, [c_p, c_i64, C.POINTER(C.c_size_t)]),
    "dspb200_fftmag_run_f32": (C.c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_p, c_i64, c_i64, c_i64, c_p, C.c_size_t, c_p]),
    "dspb200_fftmag_run_f64": (C.c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_p, c_i64, c_i64, c_i64, c_p, C.c_size_t, c_p]),
    "dspb200_fft_c2c_run_f32": (C.c_int, [c_p, c_p, c_p, c_i64, c_p, C.c_size_t, c_p]),
    "dspb200_fft_c2c_run_f64": (C.c_int, [c_p, c_p, c_p, c_i64, c_p, C.c_size_t, c_p]),
    "dspb200_fftmag_host_f32": (C.c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_p]),
    "dspb200_fftmag_host_f64": (C.c_int, [c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_p]),
    "dspb200_fft_c2c_host_f64": (C.c_int, [c_p, c_p, c_p, c_i64]),
    "dspb200_chain_workspace_bytes": (C.c_int, [c_p, c_p, c_p, c_i64, c_i64, C.c_int, C.POINTER(C.c_size_t)]),
    "dspb200_chain_run_f32": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_p, C.c_size_t, c_p]),
    "dspb200_chain_run_f64": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_p, c_p, c_p, C.c_size_t, c_p]),
    "dspb200_chain_kernel_kind": (C.c_int, [c_p, c_p, c_i64, c_i64, c_i64, _pi]),
    "dspb200_chain_host_f32": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "dspb200_chain_host_f64": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "dspb200_chain_host_pcm16_f32": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "dspb200_chain_host_pcm16_f64": (C.c_int, [c_p, c_p, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "dspb200_host_release": (C.c_int, []),
    "dspb200_pcm16_run_f32": (C.c_int, [c_p, c_i64, c_p, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_pcm16_run_f64": (C.c_int, [c_p, c_i64, c_p, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_mono_normalize_run_f64": (C.c_int, [c_p, c_i64, c_i64, C.c_int, c_p, c_i64, c_p, c_p]),
    "dspb200_mono_normalize_run_f32": (C.c_int, [c_p, c_i64, c_i64, C.c_int, c_p, c_i64, c_p, c_p]),
    "dspb200_generate_uniform_f32": (C.c_int, [c_p, c_i64, c_i64, c_i64, c_i64, C.c_uint64, C.c_double, C.c_double, c_p]),
    "dspb200_generate_uniform_f64": (C.c_int, [c_p, c_i64, c_i64, c_i64, c_i64, C.c_uint64, C.c_double, C.c_double, c_p]),
    "dspb200_launch_count": (C.c_int64, []),
}
# test hooks exported by the library but not part of the public header
EXTRA_SIGNATURES = {
    "dspb200_src_run_generic_f32": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_src_run_generic_f64": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_src_run_tiled_f32": (C.c_int, [c_p, c_p, c_i64, c_p, c_i64, c_i64, c_i64, c_p]),
    "dspb200_chain_fused_f32": (C.c_int, [c_p, c_p, c_p, c_i64, c_i64, c_i64, c_p, c_i64, c_p]),
}

_lib = None
_lock = threading.Lock()


class Dspb200Error(RuntimeError):
    """A libdspb200 call failed with a CUDA / device / allocation error."""


def build_library(verbose: bool = False) -> str:
    """Compile libdspb200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libdspb200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise Dspb200Error(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C dsp_audio_project_b200/csrc`).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + list(EXTRA_SIGNATURES.items()):
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().dspb200_last_error_string().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a status code to the exception dsp_core's callers would see."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_ALLOC:
        raise MemoryError(msg)
    raise Dspb200Error(msg)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().dspb200_device_count(C.byref(n))
    return n.value if rc == OK else 0


def launch_count() -> int:
    return int(load().dspb200_launch_count())
