//! UNVERIFIED SOURCE: never compiled (no Rust toolchain in the build image).
//!
//! Thin shim that puts the CUDA lane engine behind the reference's own traits, so a
//! `Split<GpuLanes<Biquad<Q32<F>>>, GpuDf1<i32>>` can stand where
//! `Split<Lanes<Biquad<Q32<F>>>, [DirectForm1<i32>; N]>` stood
//! (dsp-process/src/compose.rs:448-513, split.rs:272-277).
//!
//! The lane count is a run-time property of the state (the reference's `[S; N]`
//! const generic cannot hold 2^16..2^24 lanes); the sample slices are flat
//! frame-major `[t*lanes + l]` or lane-major `[l*frames + t]` exactly as
//! `View<FrameMajor|LaneMajor>` defines them (dsp-process/src/view.rs).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};
use dsp_fixedpoint::Q32;
use dsp_process::SplitProcess;
use idsp::iir::{Biquad, DirectForm1};

#[repr(C)]
pub struct idsp_ctx {
    _private: [u8; 0],
}

pub const IDSP_FRAME_MAJOR: c_int = 0;
pub const IDSP_LANE_MAJOR: c_int = 1;

unsafe extern "C" {
    pub fn idsp_b200_init(device: c_int, out: *mut *mut idsp_ctx) -> c_int;
    pub fn idsp_b200_free(ctx: *mut idsp_ctx);
    pub fn idsp_b200_sync(ctx: *mut idsp_ctx) -> c_int;
    pub fn idsp_b200_last_error() -> *const c_char;
    // peer memory (multi-GPU edges): the owner exports a device buffer, the other processes map it and
    // pass their lane block of it as the `y` of a kernel (stores over NVLink from the kernel epilogue)
    pub fn idsp_b200_malloc(ctx: *mut idsp_ctx, bytes: usize, ptr: *mut *mut c_void) -> c_int;
    pub fn idsp_b200_mfree(ctx: *mut idsp_ctx, ptr: *mut c_void) -> c_int;
    pub fn idsp_b200_ipc_export(ctx: *mut idsp_ctx, ptr: *const c_void, handle: *mut [u8; 64]) -> c_int;
    pub fn idsp_b200_ipc_open(ctx: *mut idsp_ctx, handle: *const [u8; 64], ptr: *mut *mut c_void) -> c_int;
    pub fn idsp_b200_ipc_close(ctx: *mut idsp_ctx, ptr: *mut c_void) -> c_int;
    /// replaces `Biquad<Q32<F>>::process` looped by `Lanes` (src/iir/biquad.rs:366-383)
    pub fn idsp_biquad_df1_i32_host(
        ctx: *mut idsp_ctx, ba: *const i32, f: c_int, clamp: *const i32, state: *mut i32,
        x: *const i32, y: *mut i32, frames: usize, lanes: usize, layout: c_int,
    ) -> c_int;
    /// replaces `HBF_DEC_CASCADE` (src/hbf.rs:385-421)
    pub fn idsp_hbf_dec_cascade_f32_host(
        ctx: *mut idsp_ctx, log2_rate: c_int, state: *mut f32, x: *const f32, y: *mut f32,
        n_out: usize, lanes: usize, layout: c_int,
    ) -> c_int;
    pub fn idsp_cossin_i32_host(ctx: *mut idsp_ctx, phase: *const i32, cs: *mut i32, n: usize) -> c_int;
    pub fn idsp_atan2_i32_host(ctx: *mut idsp_ctx, xy: *const i32, p: *mut i32, n: usize) -> c_int;
    /// replaces `Lockin<Lowpass<N>>` fed by an `Accu` (src/lockin.rs:30-39); device pointers
    pub fn idsp_lockin_i32(
        ctx: *mut idsp_ctx, order: c_int, k: *const i32, accu_state: *mut i32, accu_step: *const i32,
        lp_state: *mut i64, x: *const i32, iq: *mut i32, frames: usize, lanes: usize, layout: c_int,
    ) -> c_int;
    /// replaces `Split::stateful(Cic::<i64, N, M>::new(rate)).decimate()` (src/cic.rs:176-200)
    pub fn idsp_cic_dec_i64(
        ctx: *mut idsp_ctx, n: c_int, m: c_int, rate: u32, state: *mut i64, x: *const i64, y: *mut i64,
        frames: usize, lanes: usize, layout: c_int,
    ) -> c_int;
    /// replaces `PLL::process` (src/pll.rs:88-108); ba = raw `Q32<32>` bits
    pub fn idsp_pll_i32(
        ctx: *mut idsp_ctx, ba: *const i32, state: *mut i32, x: *const i32, y: *mut i32,
        frames: usize, lanes: usize, layout: c_int,
    ) -> c_int;
    /// replaces the graph of examples/fm_disc.rs:26-48
    pub fn idsp_fm_disc_i32(
        ctx: *mut idsp_ctx, carrier: i32, ba: *const i32, f: c_int, state: *mut i32, x: *const i32,
        y: *mut i32, frames: usize, lanes: usize, layout: c_int,
    ) -> c_int;
    // ... one declaration per remaining symbol of include/idsp_b200.h (same pattern)
}

/// Owns a device context (one CUDA stream); `!Sync`, mirrors `&mut` exclusivity.
pub struct Engine(*mut idsp_ctx);
impl Engine {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut p = core::ptr::null_mut();
        let rc = unsafe { idsp_b200_init(device, &mut p) };
        if rc != 0 { Err(last_error()) } else { Ok(Self(p)) }
    }
}
impl Drop for Engine {
    fn drop(&mut self) { unsafe { idsp_b200_free(self.0) } }
}
fn last_error() -> String {
    unsafe { core::ffi::CStr::from_ptr(idsp_b200_last_error()) }.to_string_lossy().into_owned()
}

/// SoA state of N DF1 lanes: `words[w * lanes + lane]`, w = [x0, x1, y0, y1]
/// (field order of `DirectForm<T,1,2>`, src/iir/biquad.rs:260-269).  The reference
/// struct is not `#[repr(C)]`, so fields are copied, never transmuted.
pub struct GpuDf1 { pub lanes: usize, pub words: Vec<i32> }
impl GpuDf1 {
    pub fn from_states(s: &[DirectForm1<i32>]) -> Self {
        let n = s.len();
        let mut words = vec![0; 4 * n];
        for (l, st) in s.iter().enumerate() {
            words[l] = st.x[0]; words[n + l] = st.x[1];
            words[2 * n + l] = st.y[0][0]; words[3 * n + l] = st.y[0][1];
        }
        Self { lanes: n, words }
    }
    pub fn to_state(&self, l: usize) -> DirectForm1<i32> {
        let n = self.lanes;
        DirectForm1 { x: [self.words[l], self.words[n + l]], y: [[self.words[2 * n + l], self.words[3 * n + l]]] }
    }
}

/// `Lanes<Biquad<Q32<F>>>` executed on the GPU.
pub struct GpuLanes<'e, C> { pub engine: &'e Engine, pub inner: C }

/// Frame-major block: x, y are `[[i32; N]]` flattened (compose.rs:468-476 via process.rs:122-127).
impl<'e, const F: i8> SplitProcess<i32, i32, GpuDf1> for GpuLanes<'e, Biquad<Q32<F>>> {
    fn process(&self, _state: &mut GpuDf1, _x: i32) -> i32 {
        unimplemented!("single-sample calls make no sense across PCIe; use block()")
    }
    fn block(&self, state: &mut GpuDf1, x: &[i32], y: &mut [i32]) {
        debug_assert_eq!(x.len(), y.len());
        debug_assert_eq!(x.len() % state.lanes, 0);
        let ba: [i32; 5] = self.inner.ba.map(|c| c.into_bits());
        let rc = unsafe {
            idsp_biquad_df1_i32_host(self.engine.0, ba.as_ptr(), F as c_int, core::ptr::null(),
                state.words.as_mut_ptr(), x.as_ptr(), y.as_mut_ptr(),
                x.len() / state.lanes, state.lanes, IDSP_FRAME_MAJOR)
        };
        assert_eq!(rc, 0, "{}", last_error());
    }
}
// `SplitViewProcess<View<LaneMajor>, ViewMut<LaneMajor>, GpuDf1>` is identical with
// `IDSP_LANE_MAJOR`, `x.flat()` / `y.flat_mut()` and `frames = x.frames()` (compose.rs:478-494).
#[allow(unused)] type _Unused = c_void;
