//! UNVERIFIED SOURCE: never compiled (no Rust toolchain in the build image).  `ffi.rs` is generated from
//! include/idsp_b200.h (tools/gen_rust_ffi.py) and checked structurally by tests/test_rust_shim.py; the
//! trait impls below are the binding a maintainer of quartiq/idsp would add, written against the trait
//! definitions in dsp-process/src/{process,view,split,compose}.rs of the reference tree.
//!
//! The shim puts the CUDA lane engine behind the reference's own traits, so a
//! `Split<GpuLanes<Biquad<Q32<F>>>, GpuDf1<i32>>` stands where
//! `Split<Lanes<Biquad<Q32<F>>>, [DirectForm1<i32>; N]>` stood (dsp-process/src/compose.rs:448-513,
//! split.rs:272-277).  The lane count is a run-time property of the state (the reference's `[S; N]` const
//! generic cannot hold 2^16..2^24 lanes); sample slices are flat frame-major `[t*lanes + l]` or lane-major
//! `[l*frames + t]` exactly as `View<FrameMajor|LaneMajor>` defines them (dsp-process/src/view.rs).
//!
//! Two families of entry points:
//!  * host slices (`SplitProcess::block`, `SplitInplace::inplace`, `SplitViewProcess::process_view`): the
//!    library streams them through the GPU (`*_host` symbols), one PCIe round trip per call;
//!  * device-resident buffers (`DeviceBuffer<T>` + the `*_dev` methods and `DecIntBiquad`): a chained graph
//!    (dsp-process/src/compose.rs:13-113) keeps samples and state in HBM between stages.
#![allow(non_camel_case_types)]
pub mod ffi;
pub use ffi::*;

use core::ffi::{c_int, c_void};
use core::marker::PhantomData;
use dsp_fixedpoint::Q32;
use dsp_process::{LaneMajor, SplitInplace, SplitProcess, SplitViewInplace, SplitViewProcess, View, ViewMut};
use idsp::iir::{Biquad, BiquadClamp, DirectForm1};
use idsp::{Complex, Lockin, Lowpass};

fn last_error() -> String {
    unsafe { core::ffi::CStr::from_ptr(idsp_b200_last_error()) }.to_string_lossy().into_owned()
}
#[track_caller]
fn check(rc: c_int) {
    // the reference hot path is infallible (length mismatches are caller preconditions,
    // dsp-process/src/process.rs:121-123); an ABI error is a broken precondition or a CUDA failure
    assert_eq!(rc, 0, "idsp_b200: {}", last_error());
}

/// Owns a device context (one CUDA stream); `!Sync`, mirrors `&mut` exclusivity of the states.
pub struct Engine(*mut idsp_ctx);
impl Engine {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut p = core::ptr::null_mut();
        let rc = unsafe { idsp_b200_init(device, &mut p) };
        if rc != 0 { Err(last_error()) } else { Ok(Self(p)) }
    }
    pub fn sync(&self) { check(unsafe { idsp_b200_sync(self.0) }) }
    /// kernel family of the last launch ("tma frame-major wide", "generic lane-major", ...)
    pub fn last_kernel(&self) -> String {
        unsafe { core::ffi::CStr::from_ptr(idsp_b200_last_kernel(self.0)) }.to_string_lossy().into_owned()
    }
}
impl Drop for Engine {
    fn drop(&mut self) { unsafe { idsp_b200_free(self.0) } }
}

/// A processor whose lanes run on the GPU: `GpuLanes<C>` is to `Lanes<C>` what the kernels are to the
/// reference's lane loop (dsp-process/src/compose.rs:468-513).
pub struct GpuLanes<'e, C> { pub engine: &'e Engine, pub inner: C }

// ------------------------------------------------------------------------------------------------
// iir::Biquad on DirectForm1 (src/iir/biquad.rs:366-404)
// ------------------------------------------------------------------------------------------------
/// SoA state of N DF1 lanes: `words[w * lanes + lane]`, w = [x0, x1, y0, y1] (field order of
/// `DirectForm<T,1,2>`, src/iir/biquad.rs:260-269).  The reference struct is not `#[repr(C)]`, so fields
/// are copied, never transmuted.
pub struct GpuDf1<T> { pub lanes: usize, pub words: Vec<T> }
impl<T: Copy + Default> GpuDf1<T> {
    pub fn new(lanes: usize) -> Self { Self { lanes, words: vec![T::default(); 4 * lanes] } }
    pub fn from_states(s: &[DirectForm1<T>]) -> Self {
        let n = s.len();
        let mut words = vec![T::default(); 4 * n];
        for (l, st) in s.iter().enumerate() {
            words[l] = st.x[0]; words[n + l] = st.x[1];
            words[2 * n + l] = st.y[0][0]; words[3 * n + l] = st.y[0][1];
        }
        Self { lanes: n, words }
    }
    pub fn to_state(&self, l: usize) -> DirectForm1<T> {
        let n = self.lanes;
        DirectForm1 { x: [self.words[l], self.words[n + l]], y: [[self.words[2 * n + l], self.words[3 * n + l]]] }
    }
}

impl<'e, const F: i8> GpuLanes<'e, Biquad<Q32<F>>> {
    fn run(&self, clamp: Option<[i32; 3]>, state: &mut GpuDf1<i32>, x: *const i32, y: *mut i32, len: usize, layout: c_int) {
        debug_assert_eq!(len % state.lanes, 0);
        let ba: [i32; 5] = self.inner.ba.map(|c| c.into_bits());
        let cl = clamp.as_ref().map_or(core::ptr::null(), |c| c.as_ptr());
        check(unsafe {
            idsp_biquad_df1_i32_host(self.engine.0, ba.as_ptr(), F as c_int, cl, state.words.as_mut_ptr(), x, y,
                len / state.lanes, state.lanes, layout)
        });
    }
}
/// Frame-major block: x, y are `[[i32; N]]` flattened (compose.rs:468-476 via process.rs:122-127).
impl<'e, const F: i8> SplitProcess<i32, i32, GpuDf1<i32>> for GpuLanes<'e, Biquad<Q32<F>>> {
    /// one sample of a ONE-lane state (a frame of an N-lane state is `block(&[x; N])`); correct but a full
    /// PCIe round trip per call -- use `block`
    fn process(&self, state: &mut GpuDf1<i32>, x: i32) -> i32 {
        assert_eq!(state.lanes, 1, "process(): one sample = one lane; pass a frame to block()");
        let mut y = 0i32;
        self.run(None, state, &x, &mut y, 1, IDSP_FRAME_MAJOR);
        y
    }
    fn block(&self, state: &mut GpuDf1<i32>, x: &[i32], y: &mut [i32]) {
        debug_assert_eq!(x.len(), y.len());
        self.run(None, state, x.as_ptr(), y.as_mut_ptr(), x.len(), IDSP_FRAME_MAJOR);
    }
}
impl<'e, const F: i8> SplitInplace<i32, GpuDf1<i32>> for GpuLanes<'e, Biquad<Q32<F>>> {
    fn inplace(&self, state: &mut GpuDf1<i32>, xy: &mut [i32]) {
        let p = xy.as_mut_ptr();
        self.run(None, state, p as *const i32, p, xy.len(), IDSP_FRAME_MAJOR); // x may alias y exactly
    }
}
/// Lane-major views (compose.rs:478-494): `View<LaneMajor, L>` with L == state.lanes.
impl<'a, 'b, 'e, const F: i8, const L: usize>
    SplitViewProcess<View<'a, i32, LaneMajor, L>, ViewMut<'b, i32, LaneMajor, L>, GpuDf1<i32>>
    for GpuLanes<'e, Biquad<Q32<F>>>
{
    fn process_view(&self, state: &mut GpuDf1<i32>, x: View<'a, i32, LaneMajor, L>, mut y: ViewMut<'b, i32, LaneMajor, L>) {
        debug_assert_eq!(x.frames(), y.frames());
        assert_eq!(L, state.lanes);
        let (xf, yf) = (x.flat(), y.flat_mut());
        self.run(None, state, xf.as_ptr(), yf.as_mut_ptr(), xf.len(), IDSP_LANE_MAJOR);
    }
}
impl<'a, 'e, const F: i8, const L: usize> SplitViewInplace<ViewMut<'a, i32, LaneMajor, L>, GpuDf1<i32>>
    for GpuLanes<'e, Biquad<Q32<F>>>
{
    fn inplace_view(&self, state: &mut GpuDf1<i32>, mut xy: ViewMut<'a, i32, LaneMajor, L>) {
        assert_eq!(L, state.lanes);
        let f = xy.flat_mut();
        let p = f.as_mut_ptr();
        self.run(None, state, p as *const i32, p, f.len(), IDSP_LANE_MAJOR);
    }
}
/// `BiquadClamp<Q32<F>, i32>`: offset and limits at the summing junction (biquad.rs:394-404)
impl<'e, const F: i8> SplitProcess<i32, i32, GpuDf1<i32>> for GpuLanes<'e, BiquadClamp<Q32<F>, i32>> {
    fn process(&self, state: &mut GpuDf1<i32>, x: i32) -> i32 {
        let mut y = [0i32];
        self.block(state, &[x], &mut y);
        y[0]
    }
    fn block(&self, state: &mut GpuDf1<i32>, x: &[i32], y: &mut [i32]) {
        debug_assert_eq!(x.len(), y.len());
        let c = &self.inner;
        GpuLanes { engine: self.engine, inner: Biquad { ba: c.coeff.ba } }
            .run(Some([c.u, c.min, c.max]), state, x.as_ptr(), y.as_mut_ptr(), x.len(), IDSP_FRAME_MAJOR);
    }
}
/// `Biquad<f32>` on `DirectForm1<f32>` (biquad.rs:366-383 with C = T = f32): bit-identical to the
/// reference (no FMA contraction, denormals kept)
impl<'e> SplitProcess<f32, f32, GpuDf1<f32>> for GpuLanes<'e, Biquad<f32>> {
    fn process(&self, state: &mut GpuDf1<f32>, x: f32) -> f32 {
        let mut y = [0f32];
        self.block(state, &[x], &mut y);
        y[0]
    }
    fn block(&self, state: &mut GpuDf1<f32>, x: &[f32], y: &mut [f32]) {
        debug_assert_eq!(x.len(), y.len());
        check(unsafe {
            idsp_biquad_df1_f32_host(self.engine.0, self.inner.ba.as_ptr(), 0, core::ptr::null(), state.words.as_mut_ptr(),
                x.as_ptr(), y.as_mut_ptr(), x.len() / state.lanes, state.lanes, IDSP_FRAME_MAJOR)
        });
    }
}

// ------------------------------------------------------------------------------------------------
// hbf: HBF_DEC_CASCADE / HBF_INT_CASCADE (src/hbf.rs:385-421, 476-512)
// ------------------------------------------------------------------------------------------------
/// `HBF_DEC_CASCADE` truncated to depth K (`.inner().1` ... in the reference): X = `[f32; 2^K]`, Y = f32.
pub struct HbfDecCascade<const K: usize>;
/// State of N lanes of `HbfDec{2,4,8,16,32}` (hbf.rs:363-383) as the ABI's SoA words: the stage states
/// (even history M-1 | odd history 2M-1, oldest first) concatenated, highest-rate stage first.  The
/// reference's state fields are private and sized by its CPU block length; `Default` = all zero in both.
pub struct GpuHbfDec<const K: usize> { pub lanes: usize, pub words: Vec<f32> }
impl<const K: usize> GpuHbfDec<K> {
    pub fn new(lanes: usize) -> Self {
        Self { lanes, words: vec![0.0; unsafe { idsp_hbf_dec_state_words(K as c_int) } * lanes] }
    }
}
macro_rules! impl_hbf_dec {
    ($k:literal, $r:literal) => {
        /// `SplitProcess<[f32; R], f32, HbfDecR>` (hbf.rs:412-421) over lanes: x = `[[[f32; R]; N]]` flattened
        impl<'e> SplitProcess<[f32; $r], f32, GpuHbfDec<$k>> for GpuLanes<'e, HbfDecCascade<$k>> {
            fn process(&self, state: &mut GpuHbfDec<$k>, x: [f32; $r]) -> f32 {
                let mut y = [0f32];
                self.block(state, &[x], &mut y);
                y[0]
            }
            fn block(&self, state: &mut GpuHbfDec<$k>, x: &[[f32; $r]], y: &mut [f32]) {
                debug_assert_eq!(x.len(), y.len());
                check(unsafe {
                    idsp_hbf_dec_cascade_f32_host(self.engine.0, $k, state.words.as_mut_ptr(), x.as_ptr() as *const f32,
                        y.as_mut_ptr(), y.len() / state.lanes, state.lanes, IDSP_FRAME_MAJOR)
                });
            }
        }
    };
}
impl_hbf_dec!(1, 2);
impl_hbf_dec!(2, 4);
impl_hbf_dec!(3, 8);
impl_hbf_dec!(4, 16);
impl_hbf_dec!(5, 32);

// ------------------------------------------------------------------------------------------------
// cossin / atan2 over slices (src/cossin.rs:14-67, src/atan2.rs:66-82; the shapes of src/py.rs:11-46)
// ------------------------------------------------------------------------------------------------
pub fn cossin_slice(e: &Engine, phase: &[i32], cs: &mut [(i32, i32)]) {
    debug_assert_eq!(phase.len(), cs.len());
    check(unsafe { idsp_cossin_i32_host(e.0, phase.as_ptr(), cs.as_mut_ptr() as *mut i32, phase.len()) });
}
/// rows are (x, y): `p[i] = atan2(xy[i].1, xy[i].0)`
pub fn atan2_slice(e: &Engine, xy: &[(i32, i32)], p: &mut [i32]) {
    debug_assert_eq!(xy.len(), p.len());
    check(unsafe { idsp_atan2_i32_host(e.0, xy.as_ptr() as *const i32, p.as_mut_ptr(), p.len()) });
}

// ------------------------------------------------------------------------------------------------
// Lockin<Lowpass<N>> (src/lockin.rs:17-39), phase from a per-lane Accu (src/accu.rs:29-38)
// ------------------------------------------------------------------------------------------------
/// per lane: `Accu<Wrapping<i32>>` {state, step} and `[LowpassState<N>; 2]` as SoA words [I (N) | Q (N)]
pub struct GpuLockin<const N: usize> { pub lanes: usize, pub accu_state: Vec<i32>, pub accu_step: Vec<i32>, pub lp: Vec<i64> }
impl<const N: usize> GpuLockin<N> {
    pub fn new(step: Vec<i32>) -> Self {
        let lanes = step.len();
        Self { lanes, accu_state: vec![0; lanes], accu_step: step, lp: vec![0; 2 * N * lanes] }
    }
}
impl<'e, const N: usize> SplitProcess<i32, Complex<i32>, GpuLockin<N>> for GpuLanes<'e, Lockin<Lowpass<N>>> {
    fn process(&self, state: &mut GpuLockin<N>, x: i32) -> Complex<i32> {
        let mut y = [Complex::new(0, 0)];
        self.block(state, &[x], &mut y);
        y[0]
    }
    /// x: `[[i32; N]]` flattened; every lane's phase advances by its step before each sample
    fn block(&self, state: &mut GpuLockin<N>, x: &[i32], y: &mut [Complex<i32>]) {
        debug_assert_eq!(x.len(), y.len());
        check(unsafe {
            idsp_lockin_i32_host(self.engine.0, N as c_int, self.inner.0.0.as_ptr(), state.accu_state.as_mut_ptr(),
                state.accu_step.as_ptr(), state.lp.as_mut_ptr(), x.as_ptr(), y.as_mut_ptr() as *mut i32,
                x.len() / state.lanes, state.lanes, IDSP_FRAME_MAJOR)
        });
    }
}

// ------------------------------------------------------------------------------------------------
// Device-resident buffers: graphs that stay in HBM between stages
// ------------------------------------------------------------------------------------------------
/// RAII device allocation (`idsp_b200_malloc`); `zeroed` = the reference's `Default` state.
pub struct DeviceBuffer<'e, T> { e: &'e Engine, ptr: *mut T, len: usize, _t: PhantomData<T> }
impl<'e, T: Copy> DeviceBuffer<'e, T> {
    pub fn zeroed(e: &'e Engine, len: usize) -> Self {
        let mut p: *mut c_void = core::ptr::null_mut();
        check(unsafe { idsp_b200_malloc(e.0, len * core::mem::size_of::<T>(), &mut p) });
        check(unsafe { idsp_b200_memset(e.0, p, 0, len * core::mem::size_of::<T>()) });
        Self { e, ptr: p as *mut T, len, _t: PhantomData }
    }
    pub fn len(&self) -> usize { self.len }
    pub fn upload(&mut self, host: &[T]) {
        assert!(host.len() <= self.len);
        check(unsafe { idsp_b200_memcpy(self.e.0, self.ptr as *mut c_void, host.as_ptr() as *const c_void, core::mem::size_of_val(host), 0) });
        self.e.sync(); // `host` may be dropped by the caller right after
    }
    pub fn download(&self, host: &mut [T]) {
        assert!(host.len() <= self.len);
        check(unsafe { idsp_b200_memcpy(self.e.0, host.as_mut_ptr() as *mut c_void, self.ptr as *const c_void, core::mem::size_of_val(host), 1) });
        self.e.sync();
    }
}
impl<'e, T> Drop for DeviceBuffer<'e, T> {
    fn drop(&mut self) { unsafe { idsp_b200_mfree(self.e.0, self.ptr as *mut c_void); } }
}

impl<'e, const F: i8> GpuLanes<'e, Biquad<Q32<F>>> {
    /// `block` on device-resident samples and state (`state` = 4 * lanes zeroed words); asynchronous
    pub fn block_dev(&self, state: &mut DeviceBuffer<i32>, lanes: usize, x: &DeviceBuffer<i32>, y: &mut DeviceBuffer<i32>, layout: c_int) {
        assert!(state.len >= 4 * lanes && x.len == y.len && x.len % lanes == 0);
        let ba: [i32; 5] = self.inner.ba.map(|c| c.into_bits());
        check(unsafe { idsp_biquad_df1_i32(self.engine.0, ba.as_ptr(), F as c_int, core::ptr::null(), state.ptr, x.ptr, y.ptr, x.len / lanes, lanes, layout) });
    }
}
impl<'e, const K: usize> GpuLanes<'e, HbfDecCascade<K>> {
    pub fn block_dev(&self, state: &mut DeviceBuffer<f32>, lanes: usize, x: &DeviceBuffer<f32>, y: &mut DeviceBuffer<f32>, layout: c_int) {
        assert!(x.len == y.len << K && y.len % lanes == 0);
        check(unsafe { idsp_hbf_dec_cascade_f32(self.engine.0, K as c_int, state.ptr, x.ptr, y.ptr, y.len / lanes, lanes, layout) });
    }
}
/// `HBF_INT_CASCADE` truncated to depth K (hbf.rs:476-512): X = f32, Y = `[f32; 2^K]`
pub struct HbfIntCascade<const K: usize>;
impl<'e, const K: usize> GpuLanes<'e, HbfIntCascade<K>> {
    pub fn block_dev(&self, state: &mut DeviceBuffer<f32>, lanes: usize, x: &DeviceBuffer<f32>, y: &mut DeviceBuffer<f32>, layout: c_int) {
        assert!(y.len == x.len << K && x.len % lanes == 0);
        check(unsafe { idsp_hbf_int_cascade_f32(self.engine.0, K as c_int, state.ptr, x.ptr, y.ptr, x.len / lanes, lanes, layout) });
    }
}
impl<'e, const N: usize> GpuLanes<'e, Lockin<Lowpass<N>>> {
    /// device-resident lock-in; `iq` may be another rank's buffer mapped with `idsp_b200_ipc_open` (the
    /// result tiles then leave the kernel epilogue over NVLink)
    pub fn block_dev(&self, accu_state: &mut DeviceBuffer<i32>, accu_step: &DeviceBuffer<i32>, lp: &mut DeviceBuffer<i64>,
                     lanes: usize, x: &DeviceBuffer<i32>, iq: *mut i32, layout: c_int) {
        check(unsafe { idsp_lockin_i32(self.engine.0, N as c_int, self.inner.0.0.as_ptr(), accu_state.ptr, accu_step.ptr, lp.ptr, x.ptr, iq, x.len / lanes, lanes, layout) });
    }
}
/// The tuple `(HbfDecCascade<K>, HbfIntCascade<K>, Biquad<f32>)` as one processor (compose.rs:13-113 chains
/// `SplitProcess` tuples): one library call, the low-rate stream never leaves the device.
pub struct DecIntBiquad<const K: usize> { pub iir: Biquad<f32> }
impl<'e, const K: usize> GpuLanes<'e, DecIntBiquad<K>> {
    pub fn state_words() -> usize { unsafe { idsp_chain_state_words(K as c_int) } }
    pub fn block_dev(&self, state: &mut DeviceBuffer<f32>, lanes: usize, x: &DeviceBuffer<f32>, y: &mut DeviceBuffer<f32>, layout: c_int) {
        assert!(state.len >= Self::state_words() * lanes && x.len == y.len && x.len % (lanes << K) == 0);
        check(unsafe { idsp_chain_f32(self.engine.0, K as c_int, self.inner.iir.ba.as_ptr(), state.ptr, x.ptr, y.ptr, x.len / (lanes << K), lanes, layout) });
    }
    /// host slices: one PCIe round trip for the three operators
    pub fn block_host(&self, state: &mut [f32], lanes: usize, x: &[f32], y: &mut [f32], layout: c_int) {
        assert!(state.len() == Self::state_words() * lanes && x.len() == y.len());
        check(unsafe { idsp_chain_f32_host(self.engine.0, K as c_int, self.inner.iir.ba.as_ptr(), state.as_mut_ptr(), x.as_ptr(), y.as_mut_ptr(), x.len() / (lanes << K), lanes, layout) });
    }
}

// ------------------------------------------------------------------------------------------------
// More than one GPU: one process per GPU, contiguous lane blocks, NCCL only at the edges
// ------------------------------------------------------------------------------------------------
pub struct Comm<'e> { e: &'e Engine, c: *mut idsp_comm }
impl<'e> Comm<'e> {
    /// rank 0 calls `Comm::unique_id()` and hands the 128 bytes to the other processes out of band
    pub fn unique_id() -> [u8; IDSP_COMM_ID_BYTES] {
        let mut id = [0u8; IDSP_COMM_ID_BYTES];
        check(unsafe { idsp_b200_comm_unique_id(id.as_mut_ptr()) });
        id
    }
    pub fn new(e: &'e Engine, nranks: i32, rank: i32, id: &[u8; IDSP_COMM_ID_BYTES]) -> Self {
        let mut c = core::ptr::null_mut();
        check(unsafe { idsp_b200_comm_init(e.0, nranks, rank, id.as_ptr(), &mut c) });
        Self { e, c }
    }
    /// lanes [lo, hi) of this rank (whole warps, the same partition on every rank)
    pub fn lane_block(&self, lanes: usize) -> (usize, usize) {
        let (mut lo, mut hi) = (0usize, 0usize);
        check(unsafe { idsp_b200_lane_block(lanes, idsp_b200_comm_size(self.c), idsp_b200_comm_rank(self.c), 0, &mut lo, &mut hi) });
        (lo, hi)
    }
    /// root's `full` (all lanes) -> every rank's `part` (its lane block, same layout); `None` off the root
    pub fn scatter_lanes<T: Copy>(&self, full: Option<&DeviceBuffer<T>>, part: &mut DeviceBuffer<T>, frames: usize, lanes: usize, width: usize, layout: c_int, root: i32) {
        let src = full.map_or(core::ptr::null(), |b| b.ptr as *const c_void);
        check(unsafe { idsp_scatter_lanes(self.c, src, part.ptr as *mut c_void, frames, lanes, width * core::mem::size_of::<T>(), layout, root) });
    }
    pub fn gather_lanes<T: Copy>(&self, part: &DeviceBuffer<T>, full: Option<&mut DeviceBuffer<T>>, frames: usize, lanes: usize, width: usize, layout: c_int, root: i32) {
        let dst = full.map_or(core::ptr::null_mut(), |b| b.ptr as *mut c_void);
        check(unsafe { idsp_gather_lanes(self.c, part.ptr as *const c_void, dst, frames, lanes, width * core::mem::size_of::<T>(), layout, root) });
    }
}
impl<'e> Drop for Comm<'e> {
    fn drop(&mut self) { let _ = self.e; unsafe { idsp_b200_comm_free(self.c); } }
}

// ------------------------------------------------------------------------------------------------
// Coefficient builders across the ABI (the crate's own `coefficients::Filter` / `pid::Builder` remain the
// natural choice on the Rust side; these exist for callers that only have the C ABI and are pinned to the
// same doctests, tests/test_coeff_builders.py)
// ------------------------------------------------------------------------------------------------
pub fn filter_build_biquad_q32<const F: i8>(f: &idsp_filter_f64, typ: c_int) -> Result<Biquad<Q32<F>>, String> {
    let mut raw = [0i32; 5];
    let rc = unsafe { idsp_filter_build_biquad_f64(f, typ, IDSP_I32, F as c_int, raw.as_mut_ptr() as *mut c_void) };
    if rc != 0 { return Err(last_error()); } // "OutOfRange(frequency)" etc. = iir::Error
    Ok(Biquad { ba: raw.map(Q32::<F>::from_bits) })
}
