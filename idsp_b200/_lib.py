"""ctypes binding of the C ABI in include/idsp_b200.h (libidsp_b200.so).

There is no CPU fallback: if the shared library is missing this raises.  Build
it with ``python -m idsp_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IDSP_B200_LIB selects an experimental build of the same library (tile-shape sweeps)
LIB_PATH = os.environ.get("IDSP_B200_LIB") or os.path.join(_HERE, "libidsp_b200.so")

_c_p = C.c_void_p
_sz = C.c_size_t
_i = C.c_int

# (name, argtypes) for every symbol declared in include/idsp_b200.h
_KINDS = ("i8", "i16", "i32", "i64", "f32", "f64")


def _signatures():
    sig = {
        "idsp_b200_init": ([_i, C.POINTER(_c_p)], _i),
        "idsp_b200_init_on_stream": ([_i, _c_p, C.POINTER(_c_p)], _i),
        "idsp_b200_free": ([_c_p], None),
        "idsp_b200_sync": ([_c_p], _i),
        "idsp_b200_last_error": ([], C.c_char_p),
        "idsp_b200_version": ([], _i),
        "idsp_b200_launch_count": ([_c_p], C.c_uint64),
        "idsp_b200_set_kernel_policy": ([_c_p, _i], _i),
        "idsp_b200_host_alloc": ([C.POINTER(_c_p), _sz], _i),
        "idsp_b200_host_free": ([_c_p], None),
        "idsp_b200_malloc": ([_c_p, _sz, C.POINTER(_c_p)], _i),
        "idsp_b200_mfree": ([_c_p, _c_p], _i),
        "idsp_b200_ipc_export": ([_c_p, _c_p, _c_p], _i),
        "idsp_b200_ipc_open": ([_c_p, _c_p, C.POINTER(_c_p)], _i),
        "idsp_b200_ipc_close": ([_c_p, _c_p], _i),
        "idsp_b200_memcpy": ([_c_p, _c_p, _c_p, _sz, _i], _i),
        "idsp_b200_memset": ([_c_p, _c_p, _i, _sz], _i),
        "idsp_hbf_taps": ([_i, C.POINTER(_i)], C.POINTER(C.c_float)),
        "idsp_hbf_dec_state_words": ([_i], _sz),
        "idsp_hbf_int_state_words": ([_i], _sz),
        "idsp_chain_state_words": ([_i], _sz),
    }
    lanes_tail = [_sz, _sz, _i]
    for k in _KINDS:
        # ctx, ba, F, clamp, state, x, y, frames, lanes, layout
        sig[f"idsp_biquad_df1_{k}"] = ([_c_p, _c_p, _i, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
        sig[f"idsp_biquad_df1_{k}_host"] = ([_c_p, _c_p, _i, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
        # ctx, ba, F, nsec, state, x, y, frames, lanes, layout
        sig[f"idsp_biquad_cascade_{k}"] = ([_c_p, _c_p, _i, _i, _c_p, _c_p, _c_p] + lanes_tail, _i)
    for k in ("f32", "f64"):
        sig[f"idsp_biquad_df2t_{k}"] = ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    for n in ("idsp_biquad_df1wide_i32", "idsp_biquad_df1dither_i32"):
        sig[n] = ([_c_p, _c_p, _i, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    # ctx, taps, M, state, x, y, n, lanes, layout
    sig["idsp_hbf_dec_f32"] = ([_c_p, _c_p, _i, _c_p, _c_p, _c_p] + lanes_tail, _i)
    sig["idsp_hbf_int_f32"] = ([_c_p, _c_p, _i, _c_p, _c_p, _c_p] + lanes_tail, _i)
    sig["idsp_fir_f32"] = ([_c_p, _c_p, _i, _i, _i, _c_p, _c_p, _c_p] + lanes_tail, _i)
    # ctx, log2_rate, state, x, y, n, lanes, layout
    for n in ("idsp_hbf_dec_cascade_f32", "idsp_hbf_dec_cascade_f32_host", "idsp_hbf_int_cascade_f32"):
        sig[n] = ([_c_p, _i, _c_p, _c_p, _c_p] + lanes_tail, _i)
    sig["idsp_chain_f32"] = ([_c_p, _i, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    for n in ("idsp_cossin_i32", "idsp_cossin_i32_host", "idsp_atan2_i32", "idsp_atan2_i32_host"):
        sig[n] = ([_c_p, _c_p, _c_p, _sz], _i)
    sig["idsp_lowpass_i32"] = ([_c_p, _i, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    for n in ("idsp_lockin_i32", "idsp_lockin_i32_host"):
        sig[n] = ([_c_p, _i, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    sig["idsp_fm_disc_i32"] = ([_c_p, C.c_int32, _c_p, _i, _c_p, _c_p, _c_p] + lanes_tail, _i)
    sig["idsp_pll_i32"] = ([_c_p, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    # ctx, N, M, rate, state, x, y, frames, lanes, layout
    sig["idsp_cic_state_words"] = ([_i, _i], _sz)
    for n in ("idsp_cic_dec_i32", "idsp_cic_dec_i64", "idsp_cic_int_i32", "idsp_cic_int_i64"):
        sig[n] = ([_c_p, _i, _i, C.c_uint32, _c_p, _c_p, _c_p] + lanes_tail, _i)
    sig["idsp_b200_last_kernel"] = ([_c_p], C.c_char_p)
    # ctx, order, k, lp_state, xp | xlo, iq, frames, lanes, layout
    for n in ("idsp_lockin_phase_i32", "idsp_lockin_lo_i32"):
        sig[n] = ([_c_p, _i, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    # caller-supplied tap sets: ctx, nstages, taps**, M*, state, x, y, n, lanes, layout
    sig["idsp_hbf_cascade_state_words"] = ([_i, _i, _c_p], _sz)
    for n in ("idsp_hbf_dec_cascade_taps_f32", "idsp_hbf_int_cascade_taps_f32"):
        sig[n] = ([_c_p, _i, _c_p, _c_p, _c_p, _c_p, _c_p] + lanes_tail, _i)
    sig["idsp_hbf_taps_98"] = ([_i, C.POINTER(_i)], C.POINTER(C.c_float))
    sig["idsp_chain_f32_host"] = sig["idsp_chain_f32"]
    # multi-GPU edges
    sig["idsp_b200_comm_unique_id"] = ([_c_p], _i)
    sig["idsp_b200_comm_init"] = ([_c_p, _i, _i, _c_p, C.POINTER(_c_p)], _i)
    sig["idsp_b200_comm_free"] = ([_c_p], _i)
    sig["idsp_b200_comm_rank"] = ([_c_p], _i)
    sig["idsp_b200_comm_size"] = ([_c_p], _i)
    sig["idsp_b200_nccl_version"] = ([], _i)
    sig["idsp_b200_lane_block"] = ([_sz, _i, _i, _sz, C.POINTER(_sz), C.POINTER(_sz)], _i)
    for n in ("idsp_scatter_lanes", "idsp_gather_lanes"):
        sig[n] = ([_c_p, _c_p, _c_p, _sz, _sz, _sz, _i, _i], _i)
    sig["idsp_broadcast"] = ([_c_p, _c_p, _sz, _i], _i)
    sig["idsp_comm_group_begin"] = ([_c_p], _i)
    sig["idsp_comm_group_end"] = ([_c_p], _i)
    sig["idsp_comm_send"] = ([_c_p, _c_p, _sz, _i], _i)
    sig["idsp_comm_recv"] = ([_c_p, _c_p, _sz, _i], _i)
    sig["idsp_b200_stream_wait"] = ([_c_p, _c_p], _i)
    # coefficient builders (host side)
    for s_, ft in (("f64", C.c_double), ("f32", C.c_float)):
        sig[f"idsp_filter_default_{s_}"] = ([_c_p], None)
        sig[f"idsp_filter_validate_{s_}"] = ([_c_p], _i)
        sig[f"idsp_filter_build_{s_}"] = ([_c_p, _i, _c_p], _i)
        sig[f"idsp_biquad_from_ba6_{s_}"] = ([_c_p, _i, _i, _c_p], _i)
        sig[f"idsp_biquad_from_ba5_{s_}"] = ([_c_p, _i, _i, _c_p], _i)
        sig[f"idsp_filter_build_biquad_{s_}"] = ([_c_p, _i, _i, _i, _c_p], _i)
        sig[f"idsp_pid_default_{s_}"] = ([_c_p], None)
        sig[f"idsp_pid_validate_{s_}"] = ([_c_p, ft], _i)
        sig[f"idsp_pid_build_{s_}"] = ([_c_p, ft, _i, _i, _c_p], _i)
    sig["idsp_biquad_from_zpk_f64"] = ([_c_p, _i, _c_p, _i, C.c_double, _i, _i, _c_p], _i)
    return sig


SIGNATURES = _signatures()


class FilterF64(C.Structure):
    _fields_ = [("frequency", C.c_double), ("gain", C.c_double), ("shelf", C.c_double),
                ("shape_kind", C.c_int), ("shape", C.c_double)]


class FilterF32(C.Structure):
    _fields_ = [("frequency", C.c_float), ("gain", C.c_float), ("shelf", C.c_float),
                ("shape_kind", C.c_int), ("shape", C.c_float)]


class PidF64(C.Structure):
    _fields_ = [("order", C.c_int), ("gain", C.c_double * 5), ("limit", C.c_double * 5)]


class PidF32(C.Structure):
    _fields_ = [("order", C.c_int), ("gain", C.c_float * 5), ("limit", C.c_float * 5)]


KIND_CODE = {"i8": 0, "i16": 1, "i32": 2, "i64": 3, "f32": 4, "f64": 5}

_lib = None


class IdspError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the CUDA library; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IdspError(
                f"{LIB_PATH} not found: the CUDA extension is not built "
                "(run `python -m idsp_b200.build`). There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is missing
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().idsp_b200_last_error().decode(errors="replace")
        raise IdspError(f"idsp_b200 error {rc}: {msg}")
