"""Thin functional layer over the C ABI: argument marshalling only.

Device path: ``torch`` CUDA tensors (PyTorch is used for device memory and
streams only).  Host path: numpy arrays / CPU tensors go through the ``*_host``
entry points of the C ABI (chunked H2D / compute / D2H inside the library) where
one exists, otherwise they are staged through a device tensor.  Either way the
arithmetic runs in the CUDA kernels; there is no CPU implementation here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

FRAME_MAJOR = 0
LANE_MAJOR = 1

KINDS = {
    "i8": (np.int8, torch.int8, C.c_int8),
    "i16": (np.int16, torch.int16, C.c_int16),
    "i32": (np.int32, torch.int32, C.c_int32),
    "i64": (np.int64, torch.int64, C.c_int64),
    "f32": (np.float32, torch.float32, C.c_float),
    "f64": (np.float64, torch.float64, C.c_double),
}
_TORCH2KIND = {v[1]: k for k, v in KINDS.items()}
_NP2KIND = {np.dtype(v[0]): k for k, v in KINDS.items()}


def kind_of(a) -> str:
    if isinstance(a, torch.Tensor):
        return _TORCH2KIND[a.dtype]
    return _NP2KIND[np.asarray(a).dtype]


def _small(values, kind):
    """Small host coefficient array -> ctypes array (kept alive by the caller)."""
    ct = KINDS[kind][2]
    vals = list(np.asarray(values).ravel().tolist())
    return (ct * len(vals))(*vals)


class PeerView:
    """A range of device memory given by address (e.g. a lane block inside a buffer another process
    owns, mapped with ``idsp_b200_ipc_open``): accepted wherever an output device tensor is."""

    def __init__(self, ptr: int, numel: int, dtype: torch.dtype):
        self.ptr, self._numel, self.dtype = int(ptr), int(numel), dtype

    def numel(self) -> int:
        return self._numel


def _is_dev(a) -> bool:
    return isinstance(a, PeerView) or (isinstance(a, torch.Tensor) and a.is_cuda)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, PeerView):
        return C.c_void_p(a.ptr)
    if isinstance(a, torch.Tensor):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(a.data_ptr())
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return C.c_void_p(a.ctypes.data)


class Context:
    """One device + one stream (``idsp_ctx``).  Not thread-safe, like ``&mut`` state."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        self._L = _lib.lib()
        self.device = int(device)
        h = C.c_void_p()
        self.stream_handle = None
        if use_torch_stream:
            torch.cuda.init()
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.stream_handle = int(stream)
            _lib.check(self._L.idsp_b200_init_on_stream(self.device, C.c_void_p(stream), C.byref(h)))
        else:
            _lib.check(self._L.idsp_b200_init(self.device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.idsp_b200_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._L.idsp_b200_sync(self._h))

    @property
    def launches(self) -> int:
        return int(self._L.idsp_b200_launch_count(self._h))

    def set_kernel_policy(self, policy: int):
        _lib.check(self._L.idsp_b200_set_kernel_policy(self._h, int(policy)))

    @property
    def last_kernel(self) -> str:
        """kernel family of the most recent launch (``idsp_b200_last_kernel``): makes the silent
        fall-back from the TMA / tiled kernels to the generic ones observable"""
        return self._L.idsp_b200_last_kernel(self._h).decode()

    # ------------------------------------------------------------------ helpers
    def _stage(self, a):
        """host array -> device tensor (for ops without a *_host entry point)."""
        if _is_dev(a):
            return a
        t = torch.from_numpy(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a
        return t.to(f"cuda:{self.device}")

    @staticmethod
    def _unstage(dev, host):
        if host is None or _is_dev(host):
            return
        if isinstance(host, torch.Tensor):
            host.copy_(dev.cpu())
        else:
            host[...] = dev.cpu().numpy().reshape(host.shape)

    def _out_like(self, x, out, n=None, dtype=None):
        if out is not None:
            return out
        n = x.numel() if (n is None and isinstance(x, torch.Tensor)) else (x.size if n is None else n)
        if isinstance(x, torch.Tensor):
            return torch.empty(n, dtype=dtype or x.dtype, device=x.device)
        return np.empty(n, dtype=dtype or x.dtype)

    @staticmethod
    def _n(a) -> int:
        return a.numel() if isinstance(a, (torch.Tensor, PeerView)) else a.size

    @staticmethod
    def _check_state(name, state, words: int, lanes: int, kind: Optional[str] = None):
        """SoA state [words, lanes] of the expected word type: the kernels index it from the operator's
        parameters, so a short or mistyped state would be read / written out of bounds on the device"""
        if state is None:
            raise ValueError(f"{name}: state is None")
        shp = tuple(state.shape)
        if len(shp) != 2 or shp[1] != lanes or shp[0] < words:
            raise ValueError(f"{name}: state has shape {shp}, needs [{words}, {lanes}] (words, lanes)")
        if kind is not None and kind_of(state) != kind:
            raise TypeError(f"{name}: state words are {kind_of(state)}, the operator needs {kind}")
        if isinstance(state, torch.Tensor) and not state.is_contiguous():
            raise ValueError(f"{name}: state must be contiguous")

    def _check_io(self, name, x, y, lanes: int, wi: int, wo: int, same_kind: bool = True) -> int:
        """whole frames, matching output length and sample type; returns the frame count"""
        if lanes <= 0:
            raise ValueError(f"{name}: lanes must be positive")
        nx, ny = self._n(x), self._n(y)
        if nx % (lanes * wi):
            raise ValueError(f"{name}: {nx} input samples are not a whole number of frames of {lanes} lanes x {wi}")
        frames = nx // (lanes * wi)
        if ny != frames * lanes * wo:
            raise ValueError(f"{name}: output holds {ny} samples, needs {frames * lanes * wo}")
        if same_kind and not isinstance(y, PeerView) and kind_of(x) != kind_of(y):
            raise TypeError(f"{name}: x is {kind_of(x)}, y is {kind_of(y)}")
        return frames

    def _run(self, name, host_name, host_capable, arrays, call):
        """arrays: dict of name -> array (samples/state). call(ptrs) -> rc."""
        any_dev = any(_is_dev(a) for a in arrays.values() if a is not None)
        all_dev = all(_is_dev(a) for a in arrays.values() if a is not None)
        if any_dev and not all_dev:
            raise ValueError(f"{name}: mix of device tensors and host arrays")
        if all_dev:
            for a in arrays.values():
                if a is not None and not isinstance(a, PeerView) and a.device.index != self.device:
                    raise ValueError(f"{name}: tensor on {a.device}, ctx on cuda:{self.device}")
            _lib.check(call(getattr(self._L, name), {k: _ptr(v) for k, v in arrays.items()}))
            return
        if host_capable:
            _lib.check(call(getattr(self._L, host_name), {k: _ptr(v) for k, v in arrays.items()}))
            return
        dev = {k: (None if v is None else self._stage(v)) for k, v in arrays.items()}
        _lib.check(call(getattr(self._L, name), {k: _ptr(v) for k, v in dev.items()}))
        self.sync()
        for k, v in arrays.items():
            if v is not None and k not in ("x", "accu_step"):
                self._unstage(dev[k], v)

    # ------------------------------------------------------------------ biquads
    def biquad(self, form: str, ba, F: int, clamp, state, x, out=None, *, lanes: int,
               layout: int = FRAME_MAJOR, nsec: int = 1):
        """form: df1 | df2t | df1wide | df1dither | cascade.  state is [words, lanes] SoA."""
        kind = kind_of(x)
        y = self._out_like(x, out)
        frames = self._check_io(f"biquad {form}", x, y, lanes, 1, 1)
        words = {"df1": 4, "df2t": 2, "df1wide": 6, "df1dither": 5, "cascade": 2 + 2 * nsec}.get(form)
        if words is None:
            raise ValueError(form)
        self._check_state(f"biquad {form}", state, words, lanes, kind)
        cba = _small(ba, kind)
        ccl = None if clamp is None else _small(clamp, kind)
        tail = (C.c_size_t(frames), C.c_size_t(lanes), C.c_int(layout))
        h = self._h
        arrays = {"state": state, "x": x, "y": y}
        if form == "df1":
            self._run(f"idsp_biquad_df1_{kind}", f"idsp_biquad_df1_{kind}_host", True, arrays,
                      lambda fn, p: fn(h, cba, F, ccl, p["state"], p["x"], p["y"], *tail))
        elif form == "cascade":
            self._run(f"idsp_biquad_cascade_{kind}", None, False, arrays,
                      lambda fn, p: fn(h, cba, F, nsec, p["state"], p["x"], p["y"], *tail))
        elif form == "df2t":
            self._run(f"idsp_biquad_df2t_{kind}", None, False, arrays,
                      lambda fn, p: fn(h, cba, ccl, p["state"], p["x"], p["y"], *tail))
        elif form in ("df1wide", "df1dither"):
            if kind != "i32":
                raise ValueError(f"{form} is i32 only")
            self._run(f"idsp_biquad_{form}_i32", None, False, arrays,
                      lambda fn, p: fn(h, cba, F, ccl, p["state"], p["x"], p["y"], *tail))
        else:
            raise ValueError(form)
        return y

    # ------------------------------------------------------------------ hbf
    def hbf_dec(self, taps, state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        taps = np.asarray(taps, np.float32)
        n = self._n(x) // (2 * lanes)
        y = self._out_like(x, out, n * lanes)
        self._check_io("hbf_dec", x, y, lanes, 2, 1)
        self._check_state("hbf_dec", state, 3 * int(taps.size) - 2, lanes, "f32")
        ct = _small(taps, "f32")
        self._run("idsp_hbf_dec_f32", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, ct, int(taps.size), p["state"], p["x"], p["y"], n, lanes, layout))
        return y

    def hbf_int(self, taps, state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        taps = np.asarray(taps, np.float32)
        n = self._n(x) // lanes
        y = self._out_like(x, out, 2 * n * lanes)
        self._check_io("hbf_int", x, y, lanes, 1, 2)
        self._check_state("hbf_int", state, 2 * int(taps.size) - 1, lanes, "f32")
        ct = _small(taps, "f32")
        self._run("idsp_hbf_int_f32", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, ct, int(taps.size), p["state"], p["x"], p["y"], n, lanes, layout))
        return y

    def fir(self, taps, odd: bool, sym: bool, state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        taps = np.asarray(taps, np.float32)
        n = self._n(x) // lanes
        y = self._out_like(x, out)
        self._check_io("fir", x, y, lanes, 1, 1)
        self._check_state("fir", state, 2 * int(taps.size) - 1 + int(bool(odd)), lanes, "f32")
        ct = _small(taps, "f32")
        self._run("idsp_fir_f32", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, ct, int(taps.size), int(odd), int(sym), p["state"], p["x"], p["y"], n, lanes, layout))
        return y

    def hbf_dec_cascade(self, log2_rate: int, state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        R = 1 << log2_rate
        n = self._n(x) // (R * lanes)
        y = self._out_like(x, out, n * lanes)
        self._check_io("hbf_dec_cascade", x, y, lanes, R, 1)
        self._check_state("hbf_dec_cascade", state, int(self._L.idsp_hbf_dec_state_words(log2_rate)), lanes, "f32")
        self._run("idsp_hbf_dec_cascade_f32", "idsp_hbf_dec_cascade_f32_host", True,
                  {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, log2_rate, p["state"], p["x"], p["y"], n, lanes, layout))
        return y

    def hbf_int_cascade(self, log2_rate: int, state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        R = 1 << log2_rate
        n = self._n(x) // lanes
        y = self._out_like(x, out, n * lanes * R)
        self._check_io("hbf_int_cascade", x, y, lanes, 1, R)
        self._check_state("hbf_int_cascade", state, int(self._L.idsp_hbf_int_state_words(log2_rate)), lanes, "f32")
        self._run("idsp_hbf_int_cascade_f32", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, log2_rate, p["state"], p["x"], p["y"], n, lanes, layout))
        return y

    def hbf_cascade_taps(self, decimate: bool, taps, state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        """half-band cascade over caller-supplied tap sets, ``taps[i]`` = stage of index i in the order of the
        reference's tap tuples (0 = lowest rate): ``idsp_hbf_{dec,int}_cascade_taps_f32``"""
        tl = [np.ascontiguousarray(t, np.float32).reshape(-1) for t in taps]
        k = len(tl)
        R = 1 << k
        Ms = (C.c_int * k)(*[int(t.size) for t in tl])
        ptrs = (C.c_void_p * k)(*[t.ctypes.data for t in tl])
        words = int(self._L.idsp_hbf_cascade_state_words(int(bool(decimate)), k, Ms))
        if words == 0:
            raise ValueError("hbf_cascade_taps: 1..5 stages with 1 <= M <= 32 taps each")
        if decimate:
            n = self._n(x) // (R * lanes)
            y = self._out_like(x, out, n * lanes)
            self._check_io("hbf_dec_cascade_taps", x, y, lanes, R, 1)
        else:
            n = self._n(x) // lanes
            y = self._out_like(x, out, n * lanes * R)
            self._check_io("hbf_int_cascade_taps", x, y, lanes, 1, R)
        self._check_state("hbf_cascade_taps", state, words, lanes, "f32")
        name = "idsp_hbf_dec_cascade_taps_f32" if decimate else "idsp_hbf_int_cascade_taps_f32"
        self._run(name, None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, k, ptrs, Ms, p["state"], p["x"], p["y"], n, lanes, layout))
        return y

    def chain(self, log2_rate: int, ba, state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        R = 1 << log2_rate
        n = self._n(x) // (R * lanes)
        y = self._out_like(x, out)
        self._check_io("chain", x, y, lanes, R, R)
        self._check_state("chain", state, int(self._L.idsp_chain_state_words(log2_rate)), lanes, "f32")
        cba = _small(ba, "f32")
        self._run("idsp_chain_f32", "idsp_chain_f32_host", True, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, log2_rate, cba, p["state"], p["x"], p["y"], n, lanes, layout))
        return y

    # ------------------------------------------------------------------ trig
    def cossin(self, phase, out=None):
        n = phase.numel() if isinstance(phase, torch.Tensor) else phase.size
        cs = self._out_like(phase, out, 2 * n)
        self._run("idsp_cossin_i32", "idsp_cossin_i32_host", True, {"x": phase, "y": cs},
                  lambda fn, p: fn(self._h, p["x"], p["y"], n))
        return cs.reshape(n, 2) if out is None else cs

    def atan2(self, xy, out=None):
        n = (xy.numel() if isinstance(xy, torch.Tensor) else xy.size) // 2
        pp = self._out_like(xy, out, n)
        self._run("idsp_atan2_i32", "idsp_atan2_i32_host", True, {"x": xy, "y": pp},
                  lambda fn, p: fn(self._h, p["x"], p["y"], n))
        return pp

    # ------------------------------------------------------------------ fm discriminator
    def fm_disc(self, carrier: int, ba: Sequence[int], F: int, state, x, out=None, *, lanes: int,
                layout: int = FRAME_MAJOR):
        b = np.asarray(ba, np.int32).reshape(5)
        frames = self._n(x) // (2 * lanes)
        y = self._out_like(x, out, frames * lanes)
        self._check_io("fm_disc", x, y, lanes, 2, 1)
        self._check_state("fm_disc", state, 7, lanes, "i32")
        car = int(carrier) - (1 << 32) if int(carrier) >= (1 << 31) else int(carrier)
        self._run("idsp_fm_disc_i32", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, car, C.c_void_p(b.ctypes.data), F, p["state"], p["x"], p["y"], frames, lanes, layout))
        return y

    # ------------------------------------------------------------------ pll
    def pll(self, ba: Sequence[int], state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        b = np.asarray(ba, np.int32).reshape(3)
        y = self._out_like(x, out)
        frames = self._check_io("pll", x, y, lanes, 1, 1)
        self._check_state("pll", state, 9, lanes, "i32")
        self._run("idsp_pll_i32", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, C.c_void_p(b.ctypes.data), p["state"], p["x"], p["y"], frames, lanes, layout))
        return y

    # ------------------------------------------------------------------ cic
    def cic(self, decimate: bool, N: int, M: int, rate: int, state, x, out=None, *, lanes: int,
            layout: int = FRAME_MAJOR):
        R = rate + 1
        nx = x.numel() if isinstance(x, torch.Tensor) else x.size
        frames = nx // (lanes * (R if decimate else 1))
        y = self._out_like(x, out, frames * lanes * (1 if decimate else R))
        kind = {torch.int32: "i32", torch.int64: "i64"}.get(x.dtype) if isinstance(x, torch.Tensor) else \
            {"int32": "i32", "int64": "i64"}.get(str(x.dtype))
        if kind is None:
            raise TypeError("Cic lanes: samples must be int32 or int64")
        self._check_io("cic", x, y, lanes, R if decimate else 1, 1 if decimate else R)
        self._check_state("cic", state, int(self._L.idsp_cic_state_words(N, M)), lanes, kind)
        self._run(f"idsp_cic_{'dec' if decimate else 'int'}_{kind}", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, N, M, rate, p["state"], p["x"], p["y"], frames, lanes, layout))
        return y

    # ------------------------------------------------------------------ lowpass / lockin
    def lowpass(self, k: Sequence[int], state, x, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        order = len(k)
        ck = _small(k, "i32")
        y = self._out_like(x, out)
        frames = self._check_io("lowpass", x, y, lanes, 1, 1)
        self._check_state("lowpass", state, order, lanes, "i64")
        if kind_of(x) != "i32":
            raise TypeError("lowpass: samples must be int32")
        self._run("idsp_lowpass_i32", None, False, {"state": state, "x": x, "y": y},
                  lambda fn, p: fn(self._h, order, ck, p["state"], p["x"], p["y"], frames, lanes, layout))
        return y

    def lockin(self, k: Sequence[int], accu_state, accu_step, lp_state, x, out=None, *, lanes: int,
               layout: int = FRAME_MAJOR):
        order = len(k)
        ck = _small(k, "i32")
        n = self._n(x)
        iq = self._out_like(x, out, 2 * n)
        frames = self._check_io("lockin", x, iq, lanes, 1, 2)
        self._check_state("lockin", lp_state, 2 * order, lanes, "i64")
        if kind_of(x) != "i32":
            raise TypeError("lockin: samples must be int32")
        for nm, a in (("accu_state", accu_state), ("accu_step", accu_step)):
            if self._n(a) != lanes or kind_of(a) != "i32":
                raise ValueError(f"lockin: {nm} must hold one int32 per lane")
        self._run("idsp_lockin_i32", "idsp_lockin_i32_host", True,
                  {"accu_state": accu_state, "accu_step": accu_step, "lp_state": lp_state, "x": x, "y": iq},
                  lambda fn, p: fn(self._h, order, ck, p["accu_state"], p["accu_step"], p["lp_state"],
                                   p["x"], p["y"], frames, lanes, layout))
        return iq


    def lockin_phase(self, k: Sequence[int], lp_state, xp, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        """``Lockin`` on (sample, phase) tuples (src/lockin.rs:30-39): xp = (x, phase) int32 pairs"""
        order = len(k)
        ck = _small(k, "i32")
        iq = self._out_like(xp, out, self._n(xp))
        frames = self._check_io("lockin_phase", xp, iq, lanes, 2, 2)
        self._check_state("lockin_phase", lp_state, 2 * order, lanes, "i64")
        if kind_of(xp) != "i32":
            raise TypeError("lockin_phase: samples must be int32")
        self._run("idsp_lockin_phase_i32", None, False, {"lp_state": lp_state, "x": xp, "y": iq},
                  lambda fn, p: fn(self._h, order, ck, p["lp_state"], p["x"], p["y"], frames, lanes, layout))
        return iq

    def lockin_lo(self, k: Sequence[int], lp_state, xlo, out=None, *, lanes: int, layout: int = FRAME_MAJOR):
        """``Lockin`` on (sample, LO) tuples (src/lockin.rs:17-28): xlo = (x, lo.re, lo.im) int32 triples"""
        order = len(k)
        ck = _small(k, "i32")
        n3 = self._n(xlo)
        iq = self._out_like(xlo, out, n3 // 3 * 2)
        frames = self._check_io("lockin_lo", xlo, iq, lanes, 3, 2)
        self._check_state("lockin_lo", lp_state, 2 * order, lanes, "i64")
        if kind_of(xlo) != "i32":
            raise TypeError("lockin_lo: samples must be int32")
        self._run("idsp_lockin_lo_i32", None, False, {"lp_state": lp_state, "x": xlo, "y": iq},
                  lambda fn, p: fn(self._h, order, ck, p["lp_state"], p["x"], p["y"], frames, lanes, layout))
        return iq


_default_ctx: dict = {}


def default_context(device: Optional[int] = None) -> Context:
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    # one ctx per (device, current torch stream): work launched under `with torch.cuda.stream(s)` is
    # ordered on s like every torch op, not on the stream that happened to be current at first use
    stream = int(torch.cuda.current_stream(device).cuda_stream) if torch.cuda.is_available() else 0
    c = _default_ctx.get((device, stream))
    if c is None:
        c = _default_ctx[(device, stream)] = Context(device)
    return c
