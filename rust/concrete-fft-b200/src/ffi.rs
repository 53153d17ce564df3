//! Raw bindings of include/cfft_b200.h (one `extern "C"` item per declared symbol).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub type cfft_status = i32;
pub const CFFT_OK: cfft_status = 0;
pub const CFFT_ELENGTH: cfft_status = -5;
pub const CFFT_METHOD_USER: c_int = 0;
pub const CFFT_METHOD_MEASURE: c_int = 1;

#[repr(C)]
pub struct cfft_plan {
    _private: [u8; 0],
}

extern "C" {
    pub fn cfft_ordered_plan_create(out: *mut *mut cfft_plan, device: c_int, n: u64, method: c_int, algo: c_int, allow_large: c_int) -> cfft_status;
    pub fn cfft_unordered_plan_create(out: *mut *mut cfft_plan, device: c_int, n: u64, method: c_int, base_algo: c_int, base_n: u64) -> cfft_status;
    pub fn cfft_f128_plan_create(out: *mut *mut cfft_plan, device: c_int, n: u64) -> cfft_status;
    pub fn cfft_plan_destroy(plan: *mut cfft_plan);
    pub fn cfft_plan_clone(plan: *const cfft_plan, out: *mut *mut cfft_plan) -> cfft_status;
    pub fn cfft_plan_fft_size(plan: *const cfft_plan) -> u64;
    pub fn cfft_plan_algo(plan: *const cfft_plan, algo: *mut c_int, base_n: *mut u64) -> cfft_status;
    pub fn cfft_plan_scratch_req(plan: *const cfft_plan, bytes: *mut u64, align: *mut u64) -> cfft_status;
    pub fn cfft_plan_kind(plan: *const cfft_plan) -> c_int;
    pub fn cfft_plan_device(plan: *const cfft_plan) -> c_int;
    pub fn cfft_plan_kernel_name(plan: *const cfft_plan) -> *const c_char;
    pub fn cfft_plan_autotune(plan: *mut cfft_plan, batch_hint: u64) -> cfft_status;
    pub fn cfft_plan_tuning_report(plan: *const cfft_plan, buf: *mut c_char, buf_len: u64) -> u64;
    pub fn cfft_c64_fwd(plan: *const cfft_plan, dev_buf: *mut c_void, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_inv(plan: *const cfft_plan, dev_buf: *mut c_void, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_plan_clone_to_device(plan: *const cfft_plan, device: c_int, out: *mut *mut cfft_plan) -> cfft_status;
    pub fn cfft_c64_host_multi(plans: *const *const cfft_plan, nplans: c_int, op: c_int, host_buf: *mut c_void, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_f128_host_multi(plans: *const *const cfft_plan, nplans: c_int, op: c_int, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_c64_fwd_strided(plan: *const cfft_plan, dev_buf: *mut c_void, row_stride: u64, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_inv_strided(plan: *const cfft_plan, dev_buf: *mut c_void, row_stride: u64, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_f128_fwd_strided(plan: *const cfft_plan, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, row_stride: u64, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_f128_inv_strided(plan: *const cfft_plan, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, row_stride: u64, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_fwd_host(plan: *const cfft_plan, host_buf: *mut c_void, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_c64_inv_host(plan: *const cfft_plan, host_buf: *mut c_void, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_c64_fwd_inv_host(plan: *const cfft_plan, host_buf: *mut c_void, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_unordered_fwd_monomial(plan: *const cfft_plan, degree: u64, dev_buf: *mut c_void, stream: *mut c_void) -> cfft_status;
    pub fn cfft_unordered_fwd_monomial_host(plan: *const cfft_plan, degree: u64, host_buf: *mut c_void, len: u64) -> cfft_status;
    pub fn cfft_unordered_permutation(plan: *const cfft_plan, out: *mut u64) -> cfft_status;
    pub fn cfft_unordered_to_standard(plan: *const cfft_plan, dev_src: *const c_void, dev_dst: *mut c_void, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_unordered_from_standard(plan: *const cfft_plan, dev_src: *const c_void, dev_dst: *mut c_void, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_unordered_to_standard_host(plan: *const cfft_plan, src: *const c_void, dst: *mut c_void) -> cfft_status;
    pub fn cfft_unordered_from_standard_host(plan: *const cfft_plan, src: *const c_void, count: u64, dst: *mut c_void) -> cfft_status;
    pub fn cfft_f128_fwd(plan: *const cfft_plan, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_f128_inv(plan: *const cfft_plan, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_f128_fwd_host(plan: *const cfft_plan, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_f128_inv_host(plan: *const cfft_plan, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_f128_binary_op(device: c_int, op: c_int, a_hi: *const f64, a_lo: *const f64, b_hi: *const f64, b_lo: *const f64, out_hi: *mut f64, out_lo: *mut f64, len: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_mul_assign(device: c_int, lhs_dev: *mut c_void, rhs_dev: *const c_void, len: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_mul_add_assign(device: c_int, acc_dev: *mut c_void, a_dev: *const c_void, b_dev: *const c_void, len: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_f128_fwd_mul_inv(plan: *const cfft_plan, l_re0: *mut f64, l_re1: *mut f64, l_im0: *mut f64, l_im1: *mut f64, r_re0: *const f64, r_re1: *const f64, r_im0: *const f64, r_im1: *const f64, rhs_row_stride: u64, factor: f64, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_fwd_mul_inv(plan: *const cfft_plan, a_dev: *const c_void, k_terms: u64, b_dev: *const c_void, b_row_stride: u64, out_dev: *mut c_void, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_fwd_mul_inv_multi(plan: *const cfft_plan, a_dev: *const c_void, k_terms: u64, b_dev: *const c_void, b_row_stride: u64, n_out: u64, out_dev: *mut c_void, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_plan_has_fused_mul2_kernel(plan: *const cfft_plan) -> c_int;
    pub fn cfft_c64_fwd_mul_add(plan: *const cfft_plan, a_dev: *const c_void, a_row_stride: u64, b_dev: *const c_void, b_row_stride: u64, acc_dev: *mut c_void, accumulate: c_int, batch: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_plan_has_fused_mul_kernel(plan: *const cfft_plan) -> c_int;
    pub fn cfft_f128_cplx_mul_scale(device: c_int, l_re0: *mut f64, l_re1: *mut f64, l_im0: *mut f64, l_im1: *mut f64, r_re0: *const f64, r_re1: *const f64, r_im0: *const f64, r_im1: *const f64, factor: f64, len: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_f128_fwd_inv_host(plan: *const cfft_plan, re0: *mut f64, re1: *mut f64, im0: *mut f64, im1: *mut f64, len: u64, batch: u64) -> cfft_status;
    pub fn cfft_status_string(st: cfft_status) -> *const c_char;
    pub fn cfft_last_error() -> *const c_char;
    pub fn cfft_launch_count() -> u64;
    pub fn cfft_version() -> *const c_char;
    pub fn cfft_plan_copy_twiddles(plan: *const cfft_plan, which: c_int, host_out: *mut c_void, bytes: u64) -> cfft_status;
    pub fn cfft_f128_unary_op(device: c_int, op: c_int, a_hi: *const f64, a_lo: *const f64, out_hi: *mut f64, out_lo: *mut f64, out2_hi: *mut f64, out2_lo: *mut f64, len: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_f128_compare(device: c_int, a_hi: *const f64, a_lo: *const f64, b_hi: *const f64, b_lo: *const f64, out: *mut i8, len: u64, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_poly_fwd(plan: *const cfft_plan, poly_dev: *const i64, fourier_dev: *mut c_void, batch: u64, flags: u32, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_poly_inv(plan: *const cfft_plan, fourier_dev: *const c_void, poly_dev: *mut i64, batch: u64, flags: u32, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_poly_mul(plan: *const cfft_plan, a_dev: *const i64, k_terms: u64, b_dev: *const c_void, b_row_stride: u64, out_dev: *mut i64, batch: u64, flags: u32, stream: *mut c_void) -> cfft_status;
    pub fn cfft_c64_poly_mul_host(plan: *const cfft_plan, a_host: *const i64, k_terms: u64, b_dev: *const c_void, b_row_stride: u64, out_host: *mut i64, batch: u64, flags: u32) -> cfft_status;
    pub fn cfft_plan_has_fused_poly_kernel(plan: *const cfft_plan, k_terms: u64) -> c_int;
    pub fn cfft_plan_copy_twist(plan: *const cfft_plan, host_out: *mut c_void, bytes: u64) -> cfft_status;
    pub fn cfft_twopass_timeouts(device: c_int, out: *mut u32) -> cfft_status;
    pub fn cfft_probe_fp64_issue_rate(device: c_int, dfma_per_s: *mut f64, dadd_per_s: *mut f64, mix_per_s: *mut f64, sm_mhz: *mut f64, sm_count: *mut c_int) -> cfft_status;
}

/// flags of the polynomial entry points (`cfft_c64_poly_*`)
pub const CFFT_POLY_INTEGER: u32 = 0;
pub const CFFT_POLY_TORUS: u32 = 1;
pub const CFFT_POLY_ACCUMULATE: u32 = 2;

/// Shared by the three `Plan` types: kernel family name, on-device autotune and its report.
pub(crate) fn kernel_name(h: &Handle) -> String {
    unsafe { core::ffi::CStr::from_ptr(cfft_plan_kernel_name(h.0)) }.to_string_lossy().into_owned()
}
pub(crate) fn autotune(h: &mut Handle, batch_hint: u64) -> String {
    check(unsafe { cfft_plan_autotune(h.0, batch_hint) });
    let mut buf = vec![0u8; 4096];
    let n = unsafe { cfft_plan_tuning_report(h.0, buf.as_mut_ptr().cast(), buf.len() as u64) } as usize;
    String::from_utf8_lossy(&buf[..n]).into_owned()
}

/// Turns a non-zero status into the panic the reference would have raised at the same place.
#[track_caller]
pub fn check(st: cfft_status) {
    if st != CFFT_OK {
        let msg = unsafe { core::ffi::CStr::from_ptr(cfft_last_error()) };
        panic!("cfft_b200: {} (status {st})", msg.to_string_lossy());
    }
}

/// Owning handle: Drop -> cfft_plan_destroy, Clone -> cfft_plan_clone.  Plans are immutable
/// after creation, so sharing `&Handle` across threads is sound (the reference's `&self`).
pub struct Handle(pub *mut cfft_plan);
unsafe impl Send for Handle {}
unsafe impl Sync for Handle {}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { cfft_plan_destroy(self.0) }
    }
}
impl Clone for Handle {
    fn clone(&self) -> Self {
        let mut out = core::ptr::null_mut();
        check(unsafe { cfft_plan_clone(self.0, &mut out) });
        Handle(out)
    }
}

/// CUDA device new plans are created on (`CFFT_B200_DEVICE`, default 0).
pub fn default_device() -> c_int {
    std::env::var("CFFT_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0)
}
