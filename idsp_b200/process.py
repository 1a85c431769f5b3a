"""Host-side mirror of the ``dsp-process`` trait surface for the GPU lane engine.

Mirrors (names, argument meaning, error behaviour):
  * ``SplitProcess::{process, block}`` / ``SplitInplace::inplace``
    (dsp-process/src/process.rs:111-142)
  * ``Split{config,state}`` and ``Split::lanes`` (dsp-process/src/split.rs:29-44, 272-277)
  * ``Lanes<C>`` (dsp-process/src/compose.rs:448-513)
  * ``View`` / ``ViewMut`` with ``FrameMajor`` / ``LaneMajor`` (dsp-process/src/view.rs)

The lane count is a run-time property of the state object (the reference's
``[S; N]`` const generic cannot hold 2^16..2^24 lanes).  Sample buffers are flat
torch CUDA tensors (device path) or numpy arrays (host path, streamed through
the device by the C ABI).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from .engine import FRAME_MAJOR, LANE_MAJOR, Context, default_context

__all__ = ["FrameMajor", "LaneMajor", "View", "ViewMut", "Split", "Lanes", "LaneState"]

FrameMajor = FRAME_MAJOR
LaneMajor = LANE_MAJOR


def _numel(a) -> int:
    return a.numel() if isinstance(a, torch.Tensor) else a.size


class View:
    """Typed view of a flat buffer (dsp-process/src/view.rs:24-36).

    ``FrameMajor``: flat[t*L + l] (view.rs:106-131); ``LaneMajor``: flat[l*frames + t]
    (view.rs:176-225).  Like ``View::from_flat`` this raises if
    ``len(flat) != frames * lanes`` (view.rs:181-182 asserts).
    """

    def __init__(self, flat, frames: int, layout: int, lanes: int, width: int = 1):
        if _numel(flat) != frames * lanes * width:
            raise AssertionError(
                f"View::from_flat: flat.len()={_numel(flat)} != frames*L={frames * lanes * width}"
            )
        self.flat, self.frames, self.layout, self.lanes, self.width = flat, frames, layout, lanes, width

    @classmethod
    def from_flat(cls, flat, frames: int, lanes: int, width: int = 1):
        """Lane-major view (view.rs:176-188)."""
        return cls(flat, frames, LANE_MAJOR, lanes, width)

    @classmethod
    def from_frames(cls, flat, lanes: int, width: int = 1):
        """Frame-major view over ``[[T; L]]`` (view.rs:106-115)."""
        n = _numel(flat)
        if n % (lanes * width):
            raise AssertionError("View::from_frames: length is not a whole number of frames")
        return cls(flat, n // (lanes * width), FRAME_MAJOR, lanes, width)

    def lane(self, i: int):
        """One contiguous lane slice (lane-major only, view.rs:192-195)."""
        if self.layout != LANE_MAJOR:
            raise TypeError("lane() needs a LaneMajor view")
        n = self.frames * self.width
        return self.flat.reshape(-1)[i * n:(i + 1) * n]

    def frame(self, t: int):
        if self.layout != FRAME_MAJOR:
            raise TypeError("frame() needs a FrameMajor view")
        n = self.lanes * self.width
        return self.flat.reshape(-1)[t * n:(t + 1) * n]


ViewMut = View


class LaneState:
    """Base of all filter states: ``words`` is the SoA array [W, lanes] of the C ABI."""

    WORDS = 0
    DTYPE: Any = None

    def __init__(self, words):
        self.words = words

    @property
    def lanes(self) -> int:
        return int(self.words.shape[1])

    @property
    def on_device(self) -> bool:
        return isinstance(self.words, torch.Tensor) and self.words.is_cuda

    @staticmethod
    def _alloc(nwords: int, lanes: int, np_dtype, device):
        if device is None or device == "host":
            return np.zeros((nwords, lanes), np_dtype)
        return torch.zeros((nwords, lanes), dtype=getattr(torch, np.dtype(np_dtype).name), device=device)

    def clone(self):
        w = self.words.clone() if isinstance(self.words, torch.Tensor) else self.words.copy()
        c = object.__new__(type(self))
        c.__dict__.update(self.__dict__)
        c.words = w
        return c

    def numpy(self) -> np.ndarray:
        return self.words.cpu().numpy() if isinstance(self.words, torch.Tensor) else self.words


class _Proc:
    """Common SplitProcess surface for configs (process.rs:111-142)."""

    def _ctx(self, state) -> Context:
        dev = state.words.device.index if state.on_device else None
        return default_context(dev)

    # to be provided: _block(ctx, state, x, y, layout) ; rate()=(in_width, out_width)
    def widths(self):
        return 1, 1

    def block(self, state, x, y, layout: int = FRAME_MAJOR):
        """``SplitProcess::block(&self, &mut S, &[X], &mut [Y])``; lengths must agree
        (debug_assert in the reference, process.rs:121-123; here ValueError)."""
        wi, wo = self.widths()
        if _numel(x) * wo != _numel(y) * wi:
            raise ValueError("block: x and y lengths do not match")
        if _numel(x) % (state.lanes * wi):
            raise ValueError("block: length is not a whole number of frames")
        self._block(self._ctx(state), state, x, y, layout)

    def inplace(self, state, xy, layout: int = FRAME_MAJOR):
        """``SplitInplace::inplace`` (process.rs:135-142)."""
        wi, wo = self.widths()
        if wi != wo:
            raise TypeError("inplace needs X == Y")
        self.block(state, xy, xy, layout)

    def process(self, state, x):
        """``SplitProcess::process``: one frame (one sample per lane) -> one frame."""
        wi, wo = self.widths()
        if isinstance(x, torch.Tensor):
            xx = x.reshape(-1).contiguous()
            y = torch.empty(xx.numel() * wo // wi, dtype=self._out_dtype(xx), device=xx.device)
        else:
            xx = np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=state.DTYPE if state.DTYPE else None)))
            y = np.empty(xx.size * wo // wi, dtype=self._out_dtype(xx))
        self.block(state, xx, y, FRAME_MAJOR)
        if state.on_device:
            self._ctx(state).sync()
        return y

    def _out_dtype(self, x):
        return x.dtype


class Lanes(_Proc):
    """``Lanes<C>``: one configuration, N independent states (compose.rs:448-513).

    ``block`` takes frame-major ``[[X; N]]`` (compose.rs:468-476); ``process_view``
    takes ``View<LaneMajor>`` pairs (compose.rs:478-494); ``inplace_view``
    (compose.rs:503-513).
    """

    def __init__(self, inner):
        self.inner = inner

    def into_inner(self):
        return self.inner

    def widths(self):
        return self.inner.widths()

    def _block(self, ctx, state, x, y, layout):
        self.inner._block(ctx, state, x, y, layout)

    def _out_dtype(self, x):
        return self.inner._out_dtype(x)

    def process_view(self, state, x: View, y: View):
        if x.frames != y.frames:
            raise ValueError("process_view: x.frames() != y.frames()")
        if x.lanes != state.lanes or y.lanes != state.lanes:
            raise ValueError("process_view: view lane count != state lane count")
        if x.layout != y.layout:
            raise TypeError("process_view: mixed layouts")
        self.inner._block(self._ctx(state), state, x.flat, y.flat, x.layout)

    def inplace_view(self, state, xy: View):
        self.process_view(state, xy, xy)


@dataclass
class Split:
    """``Split<C, S>``: owns config and state (dsp-process/src/split.rs:29-44)."""

    config: Any
    state: Any

    @classmethod
    def new(cls, config, state):
        return cls(config, state)

    def lanes(self, n: int) -> "Split":
        """``Split::lanes::<N>()`` (split.rs:272-277): same config, N copies of the state."""
        if self.state.lanes != 1:
            raise ValueError("lanes(): state already has lanes")
        st = self.state.clone()
        w = st.words
        st.words = w.repeat(1, n).contiguous() if isinstance(w, torch.Tensor) else np.repeat(w, n, axis=1)
        return Split(Lanes(self.config), st)

    def process(self, x):
        return self.config.process(self.state, x)

    def block(self, x, y, layout: int = FRAME_MAJOR):
        self.config.block(self.state, x, y, layout)

    def inplace(self, xy, layout: int = FRAME_MAJOR):
        self.config.inplace(self.state, xy, layout)

    def process_view(self, x: View, y: View):
        self.config.process_view(self.state, x, y)

    def inplace_view(self, xy: View):
        self.config.inplace_view(self.state, xy)
