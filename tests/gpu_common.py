"""Helpers shared by the -m gpu parity tests (all calls go through the C ABI)."""
import numpy as np
import torch

DEV = "cuda:0"

NP = {"i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64, "f32": np.float32, "f64": np.float64}
BITS = {"i8": 8, "i16": 16, "i32": 32, "i64": 64}


def rand_samples(rng, kind, n, amp_bits=None):
    if kind in BITS:
        b = amp_bits if amp_bits is not None else BITS[kind] - 3
        return rng.integers(-(1 << b), 1 << b, n, dtype=np.int64).astype(NP[kind])
    return rng.standard_normal(n).astype(NP[kind])


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def to_np(t):
    return t.detach().cpu().numpy()


def layout_flat(x_tl, layout):
    """x_tl: [frames, lanes, ...] -> flat array in the given layout"""
    if layout == 0:
        return np.ascontiguousarray(x_tl).reshape(-1)
    return np.ascontiguousarray(np.swapaxes(x_tl, 0, 1)).reshape(-1)


def assert_bits_equal(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        ai = a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])
        bi = b.view({4: np.uint32, 8: np.uint64}[b.dtype.itemsize])
        bad = np.nonzero(ai != bi)[0]
    else:
        bad = np.nonzero(a != b)[0]
    assert bad.size == 0, f"{what}: {bad.size} mismatches, first at {bad[:5]}: {a.reshape(-1)[bad[:5]]} vs {b.reshape(-1)[bad[:5]]}"
