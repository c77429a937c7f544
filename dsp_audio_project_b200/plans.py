"""Plan objects over the C ABI: SRC, EQ and FFT.

A plan owns device-side tables (polyphase tap rows, biquad state-space tables,
twiddles/window) and is immutable after creation.  Two call styles:

* ``run(...)`` -- batched ``[channels, time]`` torch CUDA tensors, enqueued on
  torch's current stream (torch is only the allocator / stream provider);
* ``run_host(...)`` -- numpy arrays in host memory; the library does the copies.
"""
from __future__ import annotations

import ctypes as C
import math
import functools

import numpy as np

from . import _lib
from ._lib import F32, F64, check

BAND_CENTRES_HZ = {  # dsp_core.py:225-228
    "Sub-Bass": 40, "Bass": 150, "Low Mids": 1000,
    "High Mids": 3000, "Presence": 5000, "Brilliance": 10000,
}
BAND_ORDER = tuple(BAND_CENTRES_HZ)
UNKNOWN_BAND_HZ = 1000  # dsp_core.py:235


def _dtype_id(dtype) -> int:
    name = str(dtype)
    if name.startswith("torch."):
        if name == "torch.float32":
            return F32
        if name == "torch.float64":
            return F64
        raise TypeError(f"unsupported dtype {dtype}; use float32 or float64")
    dt = np.dtype(dtype)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise TypeError(f"unsupported dtype {dt}; use float32 or float64")


def _np_dtype(dtype_id: int):
    return np.float32 if dtype_id == F32 else np.float64


def _torch():
    import torch
    return torch


def _stream_ptr(t) -> int:
    torch = _torch()
    return int(torch.cuda.current_stream(t.device).cuda_stream)


def _check_tensor(t, dtype_id, name):
    torch = _torch()
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name} must be a CUDA torch tensor (no CPU fallback)")
    want = torch.float32 if dtype_id == F32 else torch.float64
    if t.dtype != want:
        raise TypeError(f"{name} has dtype {t.dtype}, plan wants {want}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be [channels, time] with unit time stride")


def _row_stride(t) -> int:
    """Element stride between channels (a single-row tensor may carry any stride)."""
    return int(t.stride(0)) if t.shape[0] > 1 else max(int(t.stride(0)), int(t.shape[1]))


def _as_host(x, dtype_id, name="x"):
    a = np.ascontiguousarray(x, dtype=_np_dtype(dtype_id))
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise ValueError(f"{name} must be [time] or [channels, time]")
    return a


def src_geometry(L: int, M: int, n_in: int):
    """(taps, centre offset P, n_out) -- SURVEY.md 8a row a1 / dsp_core.py:155-170."""
    t = C.c_int()
    p = C.c_int64()
    n = C.c_int64()
    check(_lib.load().dspb200_src_geometry(int(L), int(M), int(n_in), C.byref(t), C.byref(p), C.byref(n)))
    return t.value, p.value, n.value


class SrcPlan:
    """L/M polyphase resampler (replaces conversion_tasa_muestreo, dsp_core.py:133-173)."""

    def __init__(self, L: int, M: int, dtype=np.float32):
        self.L, self.M = int(L), int(M)
        self.dtype_id = _dtype_id(dtype)
        self._h = C.c_void_p()
        check(_lib.load().dspb200_src_plan_create(self.L, self.M, self.dtype_id, C.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        try:
            if h:
                _lib.load().dspb200_src_plan_destroy(h)
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def out_len(self, n_in: int) -> int:
        return src_geometry(self.L, self.M, n_in)[2]

    def kernel_kind(self, channels: int, n_in: int) -> str:
        k = C.c_int()
        check(_lib.load().dspb200_src_plan_kernel_kind(self._h, channels, n_in, n_in, C.byref(k)))
        return {0: "generic", 1: "tiled", 2: "tensor"}[k.value]

    def run(self, x, out=None, *, force_generic=False, force_tiled=False):
        torch = _torch()
        _check_tensor(x, self.dtype_id, "x")
        ch, n_in = x.shape
        n_out = self.out_len(n_in)
        if out is None:
            pitch = -(-n_out // 4) * 4          # keep rows 16-byte aligned for vector stores
            out = torch.empty((ch, pitch), dtype=x.dtype, device=x.device)[:, :n_out]
        _check_tensor(out, self.dtype_id, "out")
        if out.shape != (ch, n_out):
            raise ValueError(f"out must be [{ch}, {n_out}]")
        lib = _lib.load()
        suffix = "f32" if self.dtype_id == F32 else "f64"
        if force_tiled and self.dtype_id == F32:      # the FFMA kernel, bypassing the tensor-core form
            fn = lib.dspb200_src_run_tiled_f32
        else:
            fn = getattr(lib, ("dspb200_src_run_generic_" if force_generic else "dspb200_src_run_") + suffix)
        with torch.cuda.device(x.device):
            check(fn(self._h, x.data_ptr(), _row_stride(x), out.data_ptr(), _row_stride(out), ch, n_in,
                     _stream_ptr(x)))
        return out

    def run_host(self, x):
        a = _as_host(x, self.dtype_id)
        ch, n_in = a.shape
        n_out = self.out_len(n_in)
        y = np.empty((ch, n_out), dtype=a.dtype)
        got = C.c_int64()
        # through THIS plan's tables (dspb200_src_host_* would build and drop a plan per call)
        fn = _lib.load().dspb200_src_plan_host_f32 if self.dtype_id == F32 else _lib.load().dspb200_src_plan_host_f64
        check(fn(self._h, a.ctypes.data, ch, n_in, y.ctypes.data, n_out, C.byref(got)))
        return y


def select_sections(fs: float, gains: dict):
    """The cascade's band rules (dsp_core.py:222-251) in the caller's dict
    order.  Returns (bypass, [(fc_eff, gain_db), ...])."""
    names = list(gains.keys())
    n = len(names)
    fc = (C.c_double * max(n, 1))(*[float(BAND_CENTRES_HZ.get(k, UNKNOWN_BAND_HZ)) for k in names])
    g = (C.c_double * max(n, 1))(*[float(gains[k]) for k in names])
    fce = (C.c_double * max(n, 1))()
    ge = (C.c_double * max(n, 1))()
    na = C.c_int()
    byp = C.c_int()
    check(_lib.load().dspb200_eq_select_sections(float(fs), fc, g, n, fce, ge, C.byref(na), C.byref(byp)))
    return bool(byp.value), [(fce[i], ge[i]) for i in range(na.value)]


class EqPlan:
    """Biquad cascade (replaces sistema_ecualizador / aplicar_ecuacion_diferencias,
    dsp_core.py:205-254).  ``sections`` = [(fc_eff, gain_db), ...]."""

    def __init__(self, fs: float, sections, dtype=np.float32, clip: bool = True, raw_ba=None):
        self.dtype_id = _dtype_id(dtype)
        self._h = C.c_void_p()
        lib = _lib.load()
        if raw_ba is not None:
            ba = np.ascontiguousarray(raw_ba, dtype=np.float64).reshape(-1, 6)
            self.n_sections = ba.shape[0]
            check(lib.dspb200_eq_plan_create_raw(ba.ctypes.data_as(C.POINTER(C.c_double)), self.n_sections,
                                                 int(bool(clip)), self.dtype_id, C.byref(self._h)))
        else:
            sections = list(sections)
            self.n_sections = len(sections)
            n = max(self.n_sections, 1)
            fc = (C.c_double * n)(*[float(s[0]) for s in sections])
            g = (C.c_double * n)(*[float(s[1]) for s in sections])
            check(lib.dspb200_eq_plan_create(float(fs), fc, g, self.n_sections, int(bool(clip)),
                                             self.dtype_id, C.byref(self._h)))

    @classmethod
    def from_gains(cls, fs: float, gains: dict, dtype=np.float32):
        """Plan for a gains dict under the reference's rules; None when the
        reference would bypass (return its input object untouched)."""
        bypass, sections = select_sections(fs, gains)
        if bypass:
            return None
        return cls(fs, sections, dtype=dtype, clip=True)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        try:
            if h:
                _lib.load().dspb200_eq_plan_destroy(h)
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def kernel_kind(self, channels: int, n: int) -> str:
        k = C.c_int()
        check(_lib.load().dspb200_eq_plan_kernel_kind(self._h, channels, n, n, C.byref(k)))
        return {0: "scan", 1: "tensor"}[k.value]

    def chunk_system(self):
        """(T, K, O, Phi) of the chunk system the tensor-core form multiplies, float64, host only:
        z = T x + O s, s' = Phi s + K x over T.shape[0] samples.  None when the plan has no tensor form."""
        rows, states = C.c_int(), C.c_int()
        lib = _lib.load()
        check(lib.dspb200_eq_plan_chunk_system(self._h, C.byref(rows), C.byref(states), None, None, None))
        if rows.value == 0:
            return None
        r, n = rows.value, states.value
        tk, o, phi = np.zeros((r + 16, r)), np.zeros((r, 16)), np.zeros((16, 16))
        pd = C.POINTER(C.c_double)
        check(lib.dspb200_eq_plan_chunk_system(self._h, C.byref(rows), C.byref(states), tk.ctypes.data_as(pd),
                                               o.ctypes.data_as(pd), phi.ctypes.data_as(pd)))
        return tk[:r], tk[r:r + n], o[:, :n], phi[:n, :n]

    def warm_chunks(self) -> int:
        """96-sample chunks after which the cascade started from zero is within 2^-24 of max|x| of its true output
        (the overlap of the tensor-core form's independent time slices on narrow batches); 0: no tensor form."""
        k = C.c_int()
        check(_lib.load().dspb200_eq_plan_warm_chunks(self._h, C.byref(k)))
        return int(k.value)

    def describe(self) -> np.ndarray:
        n = C.c_int()
        buf = np.zeros((16, 9))
        check(_lib.load().dspb200_eq_plan_describe(self._h, C.byref(n), buf.ctypes.data_as(C.POINTER(C.c_double)), 16))
        return buf[:n.value]

    def run(self, x, out=None):
        torch = _torch()
        _check_tensor(x, self.dtype_id, "x")
        if out is None:
            out = torch.empty_like(x)
        _check_tensor(out, self.dtype_id, "out")
        if out.shape != x.shape:
            raise ValueError("out must have x's shape")
        ch, n = x.shape
        fn = _lib.load().dspb200_eq_run_f32 if self.dtype_id == F32 else _lib.load().dspb200_eq_run_f64
        with torch.cuda.device(x.device):
            check(fn(self._h, x.data_ptr(), _row_stride(x), out.data_ptr(), _row_stride(out), ch, n, _stream_ptr(x)))
        return out

    def run_stream(self, x, state=None, out=None):
        """One time block of a longer signal (float32): `state` is None for the first block, afterwards the
        tensor this call returned for the previous block.  Blocks other than the last must be multiples of
        `stream_chunk()` samples.  Returns (z, state)."""
        torch = _torch()
        _check_tensor(x, self.dtype_id, "x")
        if self.dtype_id != F32:
            raise ValueError("the streaming form is float32 only")
        if out is None:
            out = torch.empty_like(x)
        _check_tensor(out, self.dtype_id, "out")
        if out.shape != x.shape:
            raise ValueError("out must have x's shape")
        ch, n = x.shape
        first = state is None
        if first:
            state = torch.zeros((ch, 16), dtype=torch.float32, device=x.device)
        elif state.shape != (ch, 16) or state.dtype != torch.float32 or not state.is_contiguous() or state.device != x.device:
            raise ValueError("state must be the [channels, 16] float32 tensor a previous run_stream returned")
        with torch.cuda.device(x.device):
            check(_lib.load().dspb200_eq_run_stream_f32(self._h, x.data_ptr(), _row_stride(x), out.data_ptr(),
                                                        _row_stride(out), ch, n, state.data_ptr(), int(first),
                                                        _stream_ptr(x)))
        return out, state

    @staticmethod
    def stream_chunk() -> int:
        return int(_lib.load().dspb200_eq_stream_chunk())

    def run_host(self, x):
        a = _as_host(x, self.dtype_id)
        z = np.empty_like(a)
        fn = _lib.load().dspb200_eq_host_f32 if self.dtype_id == F32 else _lib.load().dspb200_eq_host_f64
        check(fn(self._h, a.ctypes.data, z.ctypes.data, a.shape[0], a.shape[1]))
        return z


class FftPlan:
    """Radix-2 FFT / Hann magnitude spectrum (replaces fft_diezmado_en_tiempo and
    calcular_espectro_magnitud's arithmetic, dsp_core.py:41-98)."""

    def __init__(self, n_fft: int, dtype=np.float32, hann: bool = True, db: bool = False):
        self.n_fft = int(n_fft)
        self.bins = self.n_fft // 2 + 1
        self.dtype_id = _dtype_id(dtype)
        self._h = C.c_void_p()
        flags = (1 if hann else 0) | (2 if db else 0)   # DSPB200_FFT_HANN | DSPB200_FFT_DB
        check(_lib.load().dspb200_fft_plan_create(self.n_fft, flags, self.dtype_id, C.byref(self._h)))
        self._ws = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        try:
            if h:
                _lib.load().dspb200_fft_plan_destroy(h)
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def workspace_bytes(self, n_transforms: int) -> int:
        b = C.c_size_t()
        check(_lib.load().dspb200_fft_workspace_bytes(self._h, int(n_transforms), C.byref(b)))
        return int(b.value)

    def _workspace(self, n_transforms, device):
        """Scratch of a long transform, taken per call from torch's stream-ordered caching allocator on the launch
        stream: plans are shared between threads and streams (cached_fft_plan), a block kept on the plan would be."""
        torch = _torch()
        need = self.workspace_bytes(n_transforms)
        if need == 0:
            return None, 0, 0
        with torch.cuda.device(device):
            ws = torch.empty(need, dtype=torch.uint8, device=device)
        return ws, ws.data_ptr(), need

    def n_frames(self, n: int, hop=None, offset: int = 0) -> int:
        hop = self.n_fft if hop is None else int(hop)
        return 0 if n - offset < self.n_fft else (n - offset - self.n_fft) // hop + 1

    def first_bin_above(self, fs: float, f_min: float = 0.5) -> int:
        """app.py:207 keeps the bins with rfftfreq(n_fft, 1/fs) > 0.5 before the dB conversion.  The bin frequencies
        k*fs/n_fft rise with k, so that mask is the suffix starting at the bin returned here (1 for any practical
        fs/n_fft: only DC goes): ``spectra[..., plan.first_bin_above(fs):]`` is the masked spectrum as a view."""
        k = int(math.floor(f_min * self.n_fft / float(fs))) + 1 if f_min >= 0 else 0
        while k > 0 and (k - 1) * float(fs) / self.n_fft > f_min:
            k -= 1
        while k < self.bins and k * float(fs) / self.n_fft <= f_min:
            k += 1
        return min(k, self.bins)

    def magnitudes(self, x, *, hop=None, offset: int = 0, n_frames=None, n_valid=None, out=None):
        """|FFT(hann * frame)|[:n_fft/2+1] for frames of x [channels, time] ->
        [channels, n_frames, bins].  Samples at or beyond n_valid read as zero."""
        torch = _torch()
        _check_tensor(x, self.dtype_id, "x")
        ch, n = x.shape
        hop = self.n_fft if hop is None else int(hop)
        n_valid = n if n_valid is None else int(n_valid)
        if n_frames is None:
            n_frames = self.n_frames(n, hop, offset)
        if out is None:
            out = torch.empty((ch, n_frames, self.bins), dtype=x.dtype, device=x.device)
        if tuple(out.shape) != (ch, n_frames, self.bins) or not out.is_contiguous():
            raise ValueError(f"out must be contiguous [{ch}, {n_frames}, {self.bins}]")
        ws_keep, ws, ws_bytes = self._workspace(ch * n_frames, x.device)
        fn = _lib.load().dspb200_fftmag_run_f32 if self.dtype_id == F32 else _lib.load().dspb200_fftmag_run_f64
        with torch.cuda.device(x.device):
            check(fn(self._h, x.data_ptr(), _row_stride(x), n_valid, int(offset), hop, n_frames, out.data_ptr(),
                     self.bins, n_frames * self.bins, ch, ws, ws_bytes, _stream_ptr(x)))
        del ws_keep          # back to the allocator, ordered after the launch on this stream
        return out

    def magnitudes_host(self, x, *, hop=None, offset: int = 0, n_frames=None):
        a = _as_host(x, self.dtype_id)
        ch, n = a.shape
        hop = self.n_fft if hop is None else int(hop)
        if n_frames is None:
            n_frames = self.n_frames(n, hop, offset)
        mag = np.empty((ch, n_frames, self.bins), dtype=a.dtype)
        fn = _lib.load().dspb200_fftmag_host_f32 if self.dtype_id == F32 else _lib.load().dspb200_fftmag_host_f64
        check(fn(self._h, a.ctypes.data if a.size else None, ch, n, int(offset), hop, n_frames, mag.ctypes.data))
        return mag

    def c2c(self, x, out=None):
        """Complex transform of [batch, n_fft] complex tensors (natural order)."""
        torch = _torch()
        want = torch.complex64 if self.dtype_id == F32 else torch.complex128
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == want and x.is_contiguous()):
            raise TypeError(f"x must be a contiguous CUDA {want} tensor")
        if x.shape[-1] != self.n_fft:
            raise ValueError("last dimension must equal n_fft")
        batch = x.numel() // self.n_fft
        if out is None:
            out = torch.empty_like(x)
        ws_keep, ws, ws_bytes = self._workspace(batch, x.device)
        fn = _lib.load().dspb200_fft_c2c_run_f32 if self.dtype_id == F32 else _lib.load().dspb200_fft_c2c_run_f64
        with torch.cuda.device(x.device):
            check(fn(self._h, x.data_ptr(), out.data_ptr(), batch, ws, ws_bytes, _stream_ptr(x)))
        del ws_keep
        return out

    def c2c_host_f64(self, x):
        a = np.ascontiguousarray(x, dtype=np.complex128)
        if a.shape[-1] != self.n_fft:
            raise ValueError("last dimension must equal n_fft")
        if self.dtype_id != F64:
            raise TypeError("c2c_host_f64 needs a float64 plan")
        out = np.empty_like(a)
        check(_lib.load().dspb200_fft_c2c_host_f64(self._h, a.ctypes.data, out.ctypes.data, a.size // self.n_fft))
        return out


class Chain:
    """SRC -> EQ -> framed magnitude spectra (app.py:161-167, :202-205)."""

    def __init__(self, L: int, M: int, fs_in: float, gains: dict, n_fft: int = 4096, dtype=np.float32,
                 db: bool = False):
        self.dtype_id = _dtype_id(dtype)
        self.L, self.M = int(L), int(M)
        self.src = None if (self.L == 1 and self.M == 1) else SrcPlan(L, M, dtype)
        self.fs_out = int(fs_in * self.L / self.M)  # dsp_core.py:172
        self.eq = EqPlan.from_gains(self.fs_out, gains, dtype)
        self.fft = FftPlan(n_fft, dtype, hann=True, db=db)   # db: spectra leave as 20 log10(|X| + 1e-12), app.py:207-210
        self._ws = None

    def out_len(self, n_in):
        return self.src.out_len(n_in) if self.src else n_in

    def kernel_kind(self, channels: int, n_in: int) -> str:
        """"fused" when run() without keep_y would run SRC and EQ as the one tensor-core kernel that never writes y
        (wide float32 batches of a 160/147-shaped ratio), else "cascade" (three kernels)."""
        if self.src is None or self.eq is None or self.dtype_id != F32:
            return "cascade"
        k = C.c_int()
        check(_lib.load().dspb200_chain_kernel_kind(self.src._h, self.eq._h, channels, n_in, -(-n_in // 4) * 4, C.byref(k)))
        return "fused" if k.value == 1 else "cascade"

    def run_fused(self, x, out=None):
        """Test hook: SRC->EQ through the fused kernel at any batch width (float32).  x [channels, n_in] -> z."""
        torch = _torch()
        _check_tensor(x, self.dtype_id, "x")
        if self.src is None or self.eq is None:
            raise ValueError("the fused form needs both a resampler and an equaliser")
        ch, n_in = x.shape
        n_out = self.out_len(n_in)
        if out is None:
            pitch = -(-n_out // 4) * 4
            out = torch.empty((ch, pitch), dtype=x.dtype, device=x.device)[:, :n_out]
        _check_tensor(out, self.dtype_id, "out")
        with torch.cuda.device(x.device):
            check(_lib.load().dspb200_chain_fused_f32(self.src._h, self.eq._h, x.data_ptr(), _row_stride(x), ch, n_in,
                                                      out.data_ptr(), _row_stride(out), _stream_ptr(x)))
        return out

    def run(self, x, *, keep_y=False, z=None, mag=None):
        """x [channels, n_in] CUDA tensor -> (y or None, z, mag).  z / mag: optional preallocated contiguous outputs."""
        torch = _torch()
        _check_tensor(x, self.dtype_id, "x")
        ch, n_in = x.shape
        n_out = self.out_len(n_in)
        n_frames = n_out // self.fft.n_fft
        if z is None:
            z = torch.empty((ch, n_out), dtype=x.dtype, device=x.device)
        elif tuple(z.shape) != (ch, n_out) or not z.is_contiguous() or z.dtype != x.dtype:
            raise ValueError(f"z must be a contiguous [{ch}, {n_out}] tensor of x's dtype")
        y = torch.empty((ch, n_out), dtype=x.dtype, device=x.device) if (keep_y and self.src) else None
        if mag is None:
            mag = torch.empty((ch, n_frames, self.fft.bins), dtype=x.dtype, device=x.device)
        elif tuple(mag.shape) != (ch, n_frames, self.fft.bins) or not mag.is_contiguous() or mag.dtype != x.dtype:
            raise ValueError(f"mag must be a contiguous [{ch}, {n_frames}, {self.fft.bins}] tensor of x's dtype")
        need = C.c_size_t()
        lib = _lib.load()
        check(lib.dspb200_chain_workspace_bytes(self.src._h if self.src else None, self.eq._h if self.eq else None,
                                                self.fft._h, ch, n_in, int(keep_y), C.byref(need)))
        fn = lib.dspb200_chain_run_f32 if self.dtype_id == F32 else lib.dspb200_chain_run_f64
        with torch.cuda.device(x.device):
            # scratch of a long FFT: per call, from the stream-ordered allocator (a chain may be shared between streams)
            ws = torch.empty(need.value, dtype=torch.uint8, device=x.device) if need.value else None
            check(fn(self.src._h if self.src else None, self.eq._h if self.eq else None, self.fft._h,
                     x.data_ptr(), _row_stride(x), ch, n_in, y.data_ptr() if y is not None else None,
                     z.data_ptr(), mag.data_ptr(), ws.data_ptr() if ws is not None else None,
                     need.value, _stream_ptr(x)))
        del ws
        return y, z, mag

    def run_host(self, x, z=None, mag=None):
        """Host numpy (ideally pinned) in, host arrays out; copies pipelined in slabs."""
        a = x if (isinstance(x, np.ndarray) and x.flags.c_contiguous and x.dtype == _np_dtype(self.dtype_id)
                  and x.ndim == 2) else _as_host(x, self.dtype_id)
        ch, n_in = a.shape
        n_out = self.out_len(n_in)
        n_frames = n_out // self.fft.n_fft
        if z is None:
            z = np.empty((ch, n_out), dtype=a.dtype)
        if mag is None:
            mag = np.empty((ch, n_frames, self.fft.bins), dtype=a.dtype)
        lib = _lib.load()
        fn = lib.dspb200_chain_host_f32 if self.dtype_id == F32 else lib.dspb200_chain_host_f64
        check(fn(self.src._h if self.src else None, self.eq._h if self.eq else None, self.fft._h,
                 a.ctypes.data, ch, n_in, z.ctypes.data, mag.ctypes.data))
        return z, mag

    def run_host_pcm16(self, x, z_pcm=None, mag=None, peaks=None):
        """Export form of run_host: z comes back as the int16 signal app.py:349-354 hands to the WAV writer (peak
        normalised per clip, truncated), half the PCIe bytes of float32 z; returns (z_pcm, peaks, mag).  Build the
        chain with db=True for the dB spectra of app.py:207-210."""
        a = x if (isinstance(x, np.ndarray) and x.flags.c_contiguous and x.dtype == _np_dtype(self.dtype_id)
                  and x.ndim == 2) else _as_host(x, self.dtype_id)
        ch, n_in = a.shape
        n_out = self.out_len(n_in)
        n_frames = n_out // self.fft.n_fft
        if z_pcm is None:
            z_pcm = np.empty((ch, n_out), dtype=np.int16)
        if peaks is None:
            peaks = np.empty((ch,), dtype=a.dtype)
        if mag is None:
            mag = np.empty((ch, n_frames, self.fft.bins), dtype=a.dtype)
        if z_pcm.dtype != np.int16 or z_pcm.shape != (ch, n_out) or not z_pcm.flags.c_contiguous:
            raise ValueError(f"z_pcm must be a contiguous int16 [{ch}, {n_out}] array")
        lib = _lib.load()
        fn = lib.dspb200_chain_host_pcm16_f32 if self.dtype_id == F32 else lib.dspb200_chain_host_pcm16_f64
        check(fn(self.src._h if self.src else None, self.eq._h if self.eq else None, self.fft._h,
                 a.ctypes.data, ch, n_in, z_pcm.ctypes.data, peaks.ctypes.data, mag.ctypes.data))
        return z_pcm, peaks, mag


class WaveScheduler:
    """The C5 job as waves (wave pipelining with on-device generation; caller pattern
    app.py:161-167 per clip).  A job of many clips is cut into waves that fit the GPU; wave k+1's input is PRODUCED on
    a side stream (two x buffers) while wave k runs, so the producer -- the synthetic generator, a loader front end,
    an H2D copy -- overlaps the spectra kernel instead of sitting between two waves.  (The fused SRC->EQ kernel owns
    every SM's registers, so the producer's CTAs get onto the SMs when that kernel drains, i.e. next to the FFT, which
    leaves a third of the HBM bandwidth unused.)

        ws = WaveScheduler(chain, clips_per_wave, n_in, device)
        ws.run(waves, produce, consume)      # waves: [(first_clip, count)], count <= clips_per_wave
            produce(x_view, first, count)    # fills x_view [count, n_in]; runs under the side stream
            consume(z_view, mag_view, first, count)   # optional; called after the wave's kernels are enqueued, with the
                                                      # scheduler's work stream current (enqueue work, do not block)

    z and the spectra of a wave live in one buffer each: consume them (or copy them out) before the next wave's
    kernels overwrite them -- the calls are stream-ordered, so a consume that only enqueues work is enough.  run()
    returns with the caller's stream ordered after the whole job."""

    def __init__(self, chain: "Chain", clips_per_wave: int, n_in: int, device):
        torch = _torch()
        self.chain = chain
        self.clips = int(clips_per_wave)
        self.n_in = int(n_in)
        self.n_out = chain.out_len(self.n_in)
        self.n_frames = self.n_out // chain.fft.n_fft
        dt = torch.float32 if chain.dtype_id == F32 else torch.float64
        self.device = torch.device(device)
        pitch = -(-self.n_in // 4) * 4 if chain.dtype_id == F32 else -(-self.n_in // 2) * 2
        self.x = [torch.empty((self.clips, pitch), dtype=dt, device=self.device)[:, :self.n_in] for _ in range(2)]
        self.z = torch.empty((self.clips, self.n_out), dtype=dt, device=self.device)
        self.mag = torch.empty((self.clips, self.n_frames, chain.fft.bins), dtype=dt, device=self.device)
        # the waves' kernels run on a HIGH-priority stream, the producer on a normal one: when the fused kernel drains,
        # the block scheduler places the FFT's CTAs first and the producer's CTAs take what is left of each SM (one
        # 128-thread CTA next to the FFT's five), instead of the producer's grid filling the SMs and the FFT waiting
        self.work = torch.cuda.Stream(device=self.device, priority=-1)
        self.side = torch.cuda.Stream(device=self.device)

    def run(self, waves, produce, consume=None):
        torch = _torch()
        waves = list(waves)
        if not waves:
            return
        if max(w[1] for w in waves) > self.clips:
            raise ValueError("a wave is larger than clips_per_wave")
        caller = torch.cuda.current_stream(self.device)
        main = self.work
        ready = [torch.cuda.Event(), torch.cuda.Event()]       # x[b] holds the wave produced into it
        taken = [torch.cuda.Event(), torch.cuda.Event()]       # the kernels that read x[b] have run
        start = torch.cuda.Event()
        start.record(caller)
        main.wait_event(start)

        def produce_into(b, first, count, after):
            with torch.cuda.stream(self.side):
                self.side.wait_event(after)
                produce(self.x[b][:count], first, count)
                ready[b].record(self.side)

        produce_into(0, waves[0][0], waves[0][1], start)
        with torch.cuda.stream(main):
            for i, (first, count) in enumerate(waves):
                b = i & 1
                main.wait_event(ready[b])
                xv, zv, mv = self.x[b][:count], self.z[:count], self.mag[:count]
                fused = (self.chain.dtype_id == F32 and self.chain.src is not None
                         and self.chain.kernel_kind(count, self.n_in) == "fused")
                if fused:
                    self.chain.run_fused(xv, out=zv)
                elif self.chain.src is not None:
                    self.chain.src.run(xv, out=zv)
                    self.chain.eq.run(zv, out=zv)
                else:
                    self.chain.eq.run(xv, out=zv)
                taken[b].record(main)                          # x[b] is free once the kernels above have run
                if i + 1 < len(waves):
                    # the next wave's producer: x[1 - b] was last read by wave i - 1 (already recorded)
                    produce_into(1 - b, waves[i + 1][0], waves[i + 1][1], taken[1 - b] if i >= 1 else start)
                if self.n_frames > 0:
                    self.chain.fft.magnitudes(zv, out=mv)
                if consume is not None:
                    consume(zv, mv, first, count)
        caller.wait_stream(main)
        caller.wait_stream(self.side)


def to_pcm16(z, out=None):
    """Playback export of app.py:349-354 on [rows, time] CUDA tensors: nan_to_num,
    divide by the row peak when it is > 0, * 32767, truncate to int16.
    Returns (int16 tensor, peaks)."""
    torch = _torch()
    dtype_id = _dtype_id(z.dtype)
    _check_tensor(z, dtype_id, "z")
    rows, n = z.shape
    if out is None:
        out = torch.empty((rows, n), dtype=torch.int16, device=z.device)
    peaks = torch.empty((rows,), dtype=z.dtype, device=z.device)
    fn = _lib.load().dspb200_pcm16_run_f32 if dtype_id == F32 else _lib.load().dspb200_pcm16_run_f64
    with torch.cuda.device(z.device):
        check(fn(z.data_ptr(), _row_stride(z), peaks.data_ptr(), out.data_ptr(), _row_stride(out), rows, n,
                 _stream_ptr(z)))
    return out, peaks


def mono_normalize(frames):
    """Loader front end of dsp_core.py:23-31 on a [clips, frames, channels_in]
    (or [clips, frames]) CUDA tensor of float64/float32 samples: mono mean in
    float64, float32 cast, division by the clip peak when it exceeds 1e-6.
    Returns (float32 [clips, frames], peaks [clips])."""
    torch = _torch()
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.is_contiguous()):
        raise TypeError("frames must be a contiguous CUDA tensor")
    if frames.dim() == 2:
        frames = frames.unsqueeze(-1)
    if frames.dim() != 3:
        raise ValueError("frames must be [clips, frames, channels_in]")
    clips, n, cin = frames.shape
    pitch = -(-n // 4) * 4
    mono = torch.empty((clips, pitch), dtype=torch.float32, device=frames.device)[:, :n]
    peaks = torch.empty((clips,), dtype=torch.float32, device=frames.device)
    if frames.dtype == torch.float64:
        fn = _lib.load().dspb200_mono_normalize_run_f64
    elif frames.dtype == torch.float32:
        fn = _lib.load().dspb200_mono_normalize_run_f32
    else:
        raise TypeError("frames must be float32 or float64")
    with torch.cuda.device(frames.device):
        check(fn(frames.data_ptr(), clips, n, cin, mono.data_ptr(), _row_stride(mono), peaks.data_ptr(),
                 _stream_ptr(frames)))
    return mono, peaks


def generate_uniform(out, seed: int, lo: float = -0.5, hi: float = 0.5, first_channel: int = 0):
    """Fill a [channels, n] CUDA tensor with the library's counter-based synthetic clips (the throughput configurations
    generate their inputs on the device wave by wave); the tests hold a numpy twin of the generator
    (synthetic_clips) and check this kernel against it."""
    torch = _torch()
    dtype_id = _dtype_id(out.dtype)
    _check_tensor(out, dtype_id, "out")
    ch, n = out.shape
    fn = _lib.load().dspb200_generate_uniform_f32 if dtype_id == F32 else _lib.load().dspb200_generate_uniform_f64
    with torch.cuda.device(out.device):
        check(fn(out.data_ptr(), _row_stride(out), ch, n, int(first_channel), int(seed) & (2 ** 64 - 1), float(lo), float(hi),
                 _stream_ptr(out)))
    return out


def current_device() -> int:
    """The CUDA device current on the calling thread (what a plan created now belongs to)."""
    d = C.c_int(-1)
    check(_lib.load().dspb200_current_device(C.byref(d)))
    return int(d.value)


@functools.lru_cache(maxsize=32)
def _cached_src_plan(L: int, M: int, dtype_id: int, device: int) -> SrcPlan:
    return SrcPlan(L, M, _np_dtype(dtype_id))


@functools.lru_cache(maxsize=32)
def _cached_fft_plan(n_fft: int, dtype_id: int, hann: bool, device: int) -> FftPlan:
    return FftPlan(n_fft, _np_dtype(dtype_id), hann)


def cached_src_plan(L: int, M: int, dtype_id: int) -> SrcPlan:
    """Plans own device tables: the cache is keyed by the current device as well."""
    return _cached_src_plan(L, M, dtype_id, current_device())


def cached_fft_plan(n_fft: int, dtype_id: int, hann: bool) -> FftPlan:
    return _cached_fft_plan(n_fft, dtype_id, hann, current_device())
