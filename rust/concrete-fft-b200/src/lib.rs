//! concrete-fft-b200: the public API of concrete-fft 0.5.1 (`ordered`, `unordered`, `fft128`)
//! served by hand-written sm_100a CUDA kernels through the C ABI of `include/cfft_b200.h`.
//!
//! Every public item mirrors the item of the same name in concrete-fft (file:line citations
//! refer to that crate); the bodies are single FFI calls.  Semantics kept:
//! * in-place transforms on host slices, unnormalised, `&self` shareable across threads;
//! * the unordered plan's permuted Fourier order for a given `(base_algo, base_n)`, index for
//!   index, and results bit-identical to concrete-fft's for the same plan;
//! * panics where concrete-fft asserts (sizes, lengths);
//! * `fft_scratch()` still reports the reference's stack requirement so callers that size a
//!   `PodStack` keep working (the device path does not use it).
//!
//! Extra, not in concrete-fft: `*_batch` methods (many polynomials per call) and `device`
//! sub-modules taking raw device pointers + a CUDA stream.
#![allow(non_camel_case_types)]

pub mod ffi;

/// `concrete_fft::c64`, src/lib.rs:84
pub type c64 = num_complex::Complex64;

pub mod ordered {
    //! src/ordered.rs
    use crate::{c64, ffi};
    use aligned_vec::CACHELINE_ALIGN;
    use dyn_stack::{PodStack, SizeOverflow, StackReq};

    /// src/ordered.rs:28-45
    #[derive(Clone, Copy, Debug, PartialEq, Eq)]
    #[non_exhaustive]
    pub enum FftAlgo {
        Dif2,
        Dit2,
        Dif4,
        Dit4,
        Dif8,
        Dit8,
        Dif16,
        Dit16,
    }
    impl FftAlgo {
        pub(crate) fn from_raw(v: i32) -> Self {
            use FftAlgo::*;
            [Dif2, Dit2, Dif4, Dit4, Dif8, Dit8, Dif16, Dit16][v as usize]
        }
    }

    /// src/ordered.rs:50-59
    #[derive(Clone, Copy, Debug, PartialEq, Eq)]
    #[non_exhaustive]
    pub enum Method {
        UserProvided(FftAlgo),
        #[cfg(feature = "std")]
        Measure(core::time::Duration),
    }

    /// src/ordered.rs:187-193
    #[derive(Clone)]
    pub struct Plan {
        h: ffi::Handle,
    }

    impl core::fmt::Debug for Plan {
        fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
            f.debug_struct("Plan").field("algo", &self.algo()).field("fft_size", &self.fft_size()).finish()
        }
    }

    impl Plan {
        /// src/ordered.rs:242-278.  Panics if `n` is not a power of two or exceeds 2^10.
        #[track_caller]
        pub fn new(n: usize, method: Method) -> Self {
            let (m, algo) = match method {
                Method::UserProvided(a) => (ffi::CFFT_METHOD_USER, a as i32),
                #[cfg(feature = "std")]
                Method::Measure(_) => (ffi::CFFT_METHOD_MEASURE, 0),
            };
            let mut out = core::ptr::null_mut();
            ffi::check(unsafe { ffi::cfft_ordered_plan_create(&mut out, ffi::default_device(), n as u64, m, algo, 0) });
            Self { h: ffi::Handle(out) }
        }
        /// src/ordered.rs:291-293
        pub fn fft_size(&self) -> usize {
            unsafe { ffi::cfft_plan_fft_size(self.h.0) as usize }
        }
        /// src/ordered.rs:305-307
        pub fn algo(&self) -> FftAlgo {
            let (mut a, mut b) = (0, 0);
            ffi::check(unsafe { ffi::cfft_plan_algo(self.h.0, &mut a, &mut b) });
            FftAlgo::from_raw(a)
        }
        /// src/ordered.rs:320-322
        pub fn fft_scratch(&self) -> Result<StackReq, SizeOverflow> {
            StackReq::try_new_aligned::<c64>(self.fft_size(), CACHELINE_ALIGN)
        }
        /// src/ordered.rs:342-347
        #[track_caller]
        pub fn fwd(&self, buf: &mut [c64], stack: PodStack) {
            let _ = stack;
            ffi::check(unsafe { ffi::cfft_c64_fwd_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, 1) });
        }
        /// src/ordered.rs:368-373
        #[track_caller]
        pub fn inv(&self, buf: &mut [c64], stack: PodStack) {
            let _ = stack;
            ffi::check(unsafe { ffi::cfft_c64_inv_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, 1) });
        }
        /// Extension: `buf.len() / fft_size()` independent transforms in one call.
        #[track_caller]
        pub fn fwd_batch(&self, buf: &mut [c64]) {
            let b = (buf.len() / self.fft_size()) as u64;
            ffi::check(unsafe { ffi::cfft_c64_fwd_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, b) });
        }
        #[track_caller]
        pub fn inv_batch(&self, buf: &mut [c64]) {
            let b = (buf.len() / self.fft_size()) as u64;
            ffi::check(unsafe { ffi::cfft_c64_inv_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, b) });
        }
        /// Extension: standard-order plans above the reference's 2^10 cap (up to 2^20).
        #[track_caller]
        pub fn new_large(n: usize, method: Method) -> Self {
            let (m, algo) = match method {
                Method::UserProvided(a) => (ffi::CFFT_METHOD_USER, a as i32),
                #[cfg(feature = "std")]
                Method::Measure(_) => (ffi::CFFT_METHOD_MEASURE, 0),
            };
            let mut out = core::ptr::null_mut();
            ffi::check(unsafe { ffi::cfft_ordered_plan_create(&mut out, ffi::default_device(), n as u64, m, algo, 1) });
            Self { h: ffi::Handle(out) }
        }
        /// The raw plan handle for the [`crate::device`] entry points (valid as long as `self` lives).
        pub fn as_raw(&self) -> *const ffi::cfft_plan {
            self.h.0
        }
        /// Which kernel family serves this plan.
        pub fn kernel_name(&self) -> String {
            ffi::kernel_name(&self.h)
        }
        /// On-device kernel-variant autotune (what `Method::Measure` runs implicitly); returns the timing report.
        pub fn autotune(&mut self, batch_hint: u64) -> String {
            ffi::autotune(&mut self.h, batch_hint)
        }
    }
}

pub mod unordered {
    //! src/unordered.rs
    use crate::{c64, ffi, ordered::FftAlgo};
    use aligned_vec::CACHELINE_ALIGN;
    use dyn_stack::{PodStack, SizeOverflow, StackReq};

    /// src/unordered.rs:526-537
    #[derive(Clone, Copy, Debug)]
    pub enum Method {
        UserProvided { base_algo: FftAlgo, base_n: usize },
        #[cfg(feature = "std")]
        Measure(core::time::Duration),
    }

    /// src/unordered.rs:496-512
    #[derive(Clone)]
    pub struct Plan {
        h: ffi::Handle,
    }

    impl core::fmt::Debug for Plan {
        fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
            let (a, b) = self.algo();
            f.debug_struct("Plan").field("base_algo", &a).field("base_size", &b).field("fft_size", &self.fft_size()).finish()
        }
    }

    impl Plan {
        /// src/unordered.rs:659-747
        #[track_caller]
        pub fn new(n: usize, method: Method) -> Self {
            let (m, algo, base_n) = match method {
                Method::UserProvided { base_algo, base_n } => (ffi::CFFT_METHOD_USER, base_algo as i32, base_n as u64),
                #[cfg(feature = "std")]
                Method::Measure(_) => (ffi::CFFT_METHOD_MEASURE, 0, 0),
            };
            let mut out = core::ptr::null_mut();
            ffi::check(unsafe { ffi::cfft_unordered_plan_create(&mut out, ffi::default_device(), n as u64, m, algo, base_n) });
            Self { h: ffi::Handle(out) }
        }
        /// src/unordered.rs:760-762
        pub fn fft_size(&self) -> usize {
            unsafe { ffi::cfft_plan_fft_size(self.h.0) as usize }
        }
        /// src/unordered.rs:783-785
        pub fn algo(&self) -> (FftAlgo, usize) {
            let (mut a, mut b) = (0, 0);
            ffi::check(unsafe { ffi::cfft_plan_algo(self.h.0, &mut a, &mut b) });
            (FftAlgo::from_raw(a), b as usize)
        }
        /// src/unordered.rs:798-800
        pub fn fft_scratch(&self) -> Result<StackReq, SizeOverflow> {
            StackReq::try_new_aligned::<c64>(self.algo().1, CACHELINE_ALIGN)
        }
        /// src/unordered.rs:826-839.  Panics when `buf.len() != fft_size()`.
        #[track_caller]
        pub fn fwd(&self, buf: &mut [c64], stack: PodStack) {
            let _ = stack;
            ffi::check(unsafe { ffi::cfft_c64_fwd_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, 1) });
        }
        /// src/unordered.rs:927-940
        #[track_caller]
        pub fn inv(&self, buf: &mut [c64], stack: PodStack) {
            let _ = stack;
            ffi::check(unsafe { ffi::cfft_c64_inv_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, 1) });
        }
        /// src/unordered.rs:844-900
        #[track_caller]
        pub fn fwd_monomial(&self, degree: usize, buf: &mut [c64]) {
            ffi::check(unsafe { ffi::cfft_unordered_fwd_monomial_host(self.h.0, degree as u64, buf.as_mut_ptr().cast(), buf.len() as u64) });
        }
        /// Extension: `buf.len() / fft_size()` polynomials per call (one upload, one download).
        #[track_caller]
        pub fn fwd_batch(&self, buf: &mut [c64]) {
            let b = (buf.len() / self.fft_size()) as u64;
            ffi::check(unsafe { ffi::cfft_c64_fwd_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, b) });
        }
        #[track_caller]
        pub fn inv_batch(&self, buf: &mut [c64]) {
            let b = (buf.len() / self.fft_size()) as u64;
            ffi::check(unsafe { ffi::cfft_c64_inv_host(self.h.0, buf.as_mut_ptr().cast(), buf.len() as u64, b) });
        }

        /// The raw plan handle for the [`crate::device`] entry points (valid as long as `self` lives).
        pub fn as_raw(&self) -> *const ffi::cfft_plan {
            self.h.0
        }
        /// Which kernel family serves this plan.
        pub fn kernel_name(&self) -> String {
            ffi::kernel_name(&self.h)
        }
        /// On-device kernel-variant autotune (what `Method::Measure` runs implicitly); returns the timing report.
        /// Never changes `(base_algo, base_n)`, the Fourier-domain order or any output bit.
        pub fn autotune(&mut self, batch_hint: u64) -> String {
            ffi::autotune(&mut self.h, batch_hint)
        }
        /// `perm[i]` = index in this plan's buffers of Fourier coefficient `i` (`bit_rev_twice`, src/unordered.rs:1046-1051).
        pub fn permutation(&self) -> Vec<u64> {
            let mut out = vec![0u64; self.fft_size()];
            ffi::check(unsafe { ffi::cfft_unordered_permutation(self.h.0, out.as_mut_ptr()) });
            out
        }
        /// Extension: a whole negacyclic product step on integer polynomials in host memory --
        /// `out[r] (+)= round(untwist(inv(sum_k fwd(twist(fold(a[r][k]))) * b[r][k])))` with the Fourier-domain operand `b_dev`
        /// resident on the GPU (`[k_terms][n]` c64 shared by every row when `b_row_stride == 0`).  `a`: `batch * k_terms`
        /// polynomials of `2 * fft_size()` coefficients, `out`: `batch` polynomials.  `flags`: `ffi::CFFT_POLY_*`.
        /// # Safety
        /// `b_dev` must be a device pointer on this plan's GPU holding the operand described above.
        #[track_caller]
        pub unsafe fn poly_mul_host(&self, a: &[i64], k_terms: usize, b_dev: *const core::ffi::c_void, b_row_stride: u64, out: &mut [i64], flags: u32) {
            let npoly = 2 * self.fft_size();
            assert!(k_terms >= 1 && out.len() % npoly == 0 && a.len() == out.len() * k_terms);
            ffi::check(ffi::cfft_c64_poly_mul_host(self.h.0, a.as_ptr(), k_terms as u64, b_dev, b_row_stride, out.as_mut_ptr(), (out.len() / npoly) as u64, flags));
        }

        /// src/unordered.rs:951-972
        #[cfg(feature = "serde")]
        pub fn serialize_fourier_buffer<S: serde::Serializer>(&self, serializer: S, buf: &[c64]) -> Result<S::Ok, S::Error> {
            use serde::ser::SerializeSeq;
            let n = self.fft_size();
            assert_eq!(n, buf.len());
            let mut std_order = vec![c64::default(); n];
            ffi::check(unsafe { ffi::cfft_unordered_to_standard_host(self.h.0, buf.as_ptr().cast(), std_order.as_mut_ptr().cast()) });
            let mut seq = serializer.serialize_seq(Some(n))?;
            for z in &std_order {
                seq.serialize_element(z)?;
            }
            seq.end()
        }

        /// src/unordered.rs:982-1036
        #[cfg(feature = "serde")]
        pub fn deserialize_fourier_buffer<'de, D: serde::Deserializer<'de>>(&self, deserializer: D, buf: &mut [c64]) -> Result<(), D::Error> {
            use serde::de::{SeqAccess, Visitor};
            let n = self.fft_size();
            assert_eq!(n, buf.len());
            struct SeqVisitor<'a> {
                perm: Vec<u64>,
                buf: &'a mut [c64],
            }
            impl<'de, 'a> Visitor<'de> for SeqVisitor<'a> {
                type Value = ();
                fn expecting(&self, f: &mut core::fmt::Formatter) -> core::fmt::Result {
                    write!(f, "a sequence of {} 64-bit complex numbers", self.buf.len())
                }
                fn visit_seq<S: SeqAccess<'de>>(self, mut seq: S) -> Result<(), S::Error> {
                    // element i of the standard-order sequence lands at perm[i]; like the reference, the first n elements are
                    // written even when the sequence turns out too short or too long, and the length is judged at the end
                    let n = self.buf.len();
                    let mut count = 0usize;
                    while let Some(v) = seq.next_element::<c64>()? {
                        if count < n {
                            self.buf[self.perm[count] as usize] = v;
                        }
                        count += 1;
                    }
                    if count == n {
                        Ok(())
                    } else {
                        Err(serde::de::Error::invalid_length(count, &self))
                    }
                }
            }
            deserializer.deserialize_seq(SeqVisitor { perm: self.permutation(), buf })
        }
    }
}

#[cfg(feature = "fft128")]
pub mod fft128 {
    //! src/fft128/mod.rs
    use crate::ffi;

    /// src/fft128/mod.rs:3-7
    #[derive(Copy, Clone, Debug)]
    #[repr(C)]
    pub struct f128(pub f64, pub f64);

    /// the scalar operator surface of `f128` (src/fft128/f128_ops.rs:48-618): `+ - * /` with `f128` and `f64` operands,
    /// `Neg`, `PartialEq`, `PartialOrd`, the named `*_f128_f64`-style functions, `sqr`, `abs`, `to_f64`, `is_nan`, `sincospi`
    mod f128_ops; // src/fft128/f128_ops.rs

    /// src/fft128/mod.rs:1832-1838
    #[derive(Clone)]
    pub struct Plan {
        h: ffi::Handle,
    }

    impl core::fmt::Debug for Plan {
        fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
            f.debug_struct("Plan").field("fft_size", &self.fft_size()).finish()
        }
    }

    impl Plan {
        /// src/fft128/mod.rs:1864-1881.  Panics unless `n` is a power of two >= 32.
        #[track_caller]
        pub fn new(n: usize) -> Self {
            let mut out = core::ptr::null_mut();
            ffi::check(unsafe { ffi::cfft_f128_plan_create(&mut out, ffi::default_device(), n as u64) });
            Self { h: ffi::Handle(out) }
        }
        /// src/fft128/mod.rs:1891-1893
        pub fn fft_size(&self) -> usize {
            unsafe { ffi::cfft_plan_fft_size(self.h.0) as usize }
        }
        /// src/fft128/mod.rs:1905-1928
        #[track_caller]
        pub fn fwd(&self, buf_re0: &mut [f64], buf_re1: &mut [f64], buf_im0: &mut [f64], buf_im1: &mut [f64]) {
            let n = self.fft_size();
            assert_eq!(buf_re0.len(), n);
            assert_eq!(buf_re1.len(), n);
            assert_eq!(buf_im0.len(), n);
            assert_eq!(buf_im1.len(), n);
            ffi::check(unsafe {
                ffi::cfft_f128_fwd_host(self.h.0, buf_re0.as_mut_ptr(), buf_re1.as_mut_ptr(), buf_im0.as_mut_ptr(), buf_im1.as_mut_ptr(), n as u64, 1)
            });
        }
        /// src/fft128/mod.rs:1938-1960
        #[track_caller]
        pub fn inv(&self, buf_re0: &mut [f64], buf_re1: &mut [f64], buf_im0: &mut [f64], buf_im1: &mut [f64]) {
            let n = self.fft_size();
            assert_eq!(buf_re0.len(), n);
            assert_eq!(buf_re1.len(), n);
            assert_eq!(buf_im0.len(), n);
            assert_eq!(buf_im1.len(), n);
            ffi::check(unsafe {
                ffi::cfft_f128_inv_host(self.h.0, buf_re0.as_mut_ptr(), buf_re1.as_mut_ptr(), buf_im0.as_mut_ptr(), buf_im1.as_mut_ptr(), n as u64, 1)
            });
        }
        /// The raw plan handle for the [`crate::device`] entry points (valid as long as `self` lives).
        pub fn as_raw(&self) -> *const ffi::cfft_plan {
            self.h.0
        }
        /// Which kernel family serves this plan.
        pub fn kernel_name(&self) -> String {
            ffi::kernel_name(&self.h)
        }
        /// On-device autotune of the tile size / group shape; returns the timing report.
        pub fn autotune(&mut self, batch_hint: u64) -> String {
            ffi::autotune(&mut self.h, batch_hint)
        }
        /// Extension: `len / fft_size()` transforms per call on planar arrays.
        #[track_caller]
        pub fn fwd_batch(&self, re0: &mut [f64], re1: &mut [f64], im0: &mut [f64], im1: &mut [f64]) {
            let (len, n) = (re0.len(), self.fft_size());
            assert!(re1.len() == len && im0.len() == len && im1.len() == len && len % n == 0);
            ffi::check(unsafe {
                ffi::cfft_f128_fwd_host(self.h.0, re0.as_mut_ptr(), re1.as_mut_ptr(), im0.as_mut_ptr(), im1.as_mut_ptr(), len as u64, (len / n) as u64)
            });
        }
        #[track_caller]
        pub fn inv_batch(&self, re0: &mut [f64], re1: &mut [f64], im0: &mut [f64], im1: &mut [f64]) {
            let (len, n) = (re0.len(), self.fft_size());
            assert!(re1.len() == len && im0.len() == len && im1.len() == len && len % n == 0);
            ffi::check(unsafe {
                ffi::cfft_f128_inv_host(self.h.0, re0.as_mut_ptr(), re1.as_mut_ptr(), im0.as_mut_ptr(), im1.as_mut_ptr(), len as u64, (len / n) as u64)
            });
        }
    }
}

/// One host call, several GPUs: replicas of a plan on the devices that share the work; `fwd_batch` / `inv_batch` cut the
/// batch into contiguous row ranges, one per replica, each on its own host thread and copy pipeline inside the library
/// (`cfft_c64_host_multi` / `cfft_f128_host_multi`).  No collective: polynomials are independent.
pub mod multi_gpu {
    use crate::{c64, ffi};

    pub struct Replicas {
        h: Vec<*mut ffi::cfft_plan>,
        n: usize,
    }
    unsafe impl Send for Replicas {}
    unsafe impl Sync for Replicas {}

    impl Replicas {
        /// `plan`: `Plan::as_raw()` of any plan type; `devices`: CUDA device indices (an index may repeat).
        /// # Safety
        /// `plan` must be a live plan handle.
        pub unsafe fn new(plan: *const ffi::cfft_plan, devices: &[i32]) -> Self {
            assert!(!devices.is_empty());
            let mut h = Vec::with_capacity(devices.len());
            for &d in devices {
                let mut out = core::ptr::null_mut();
                ffi::check(ffi::cfft_plan_clone_to_device(plan, d, &mut out));
                h.push(out);
            }
            Self { h, n: ffi::cfft_plan_fft_size(plan) as usize }
        }
        fn raw(&self) -> Vec<*const ffi::cfft_plan> {
            self.h.iter().map(|p| *p as *const ffi::cfft_plan).collect()
        }
        /// `Plan::fwd` on every `fft_size` chunk of `buf` (op 0), `inv` (1) or `fwd` then `inv` (2).
        pub fn c64(&self, op: i32, buf: &mut [c64]) {
            assert_eq!(buf.len() % self.n, 0);
            let r = self.raw();
            ffi::check(unsafe { ffi::cfft_c64_host_multi(r.as_ptr(), r.len() as i32, op, buf.as_mut_ptr().cast(), buf.len() as u64, (buf.len() / self.n) as u64) });
        }
        pub fn f128(&self, op: i32, re0: &mut [f64], re1: &mut [f64], im0: &mut [f64], im1: &mut [f64]) {
            assert_eq!(re0.len() % self.n, 0);
            assert_eq!(re0.len(), re1.len());
            assert_eq!(re0.len(), im0.len());
            assert_eq!(re0.len(), im1.len());
            let r = self.raw();
            ffi::check(unsafe {
                ffi::cfft_f128_host_multi(r.as_ptr(), r.len() as i32, op, re0.as_mut_ptr(), re1.as_mut_ptr(), im0.as_mut_ptr(), im1.as_mut_ptr(), re0.len() as u64, (re0.len() / self.n) as u64)
            });
        }
    }
    impl Drop for Replicas {
        fn drop(&mut self) {
            for p in self.h.drain(..) {
                unsafe { ffi::cfft_plan_destroy(p) };
            }
        }
    }
}

/// Device-pointer entry points (stream ordered, no host copies) for callers that keep their
/// polynomials on the GPU.
pub mod device {
    use crate::ffi;
    use core::ffi::c_void;

    /// # Safety
    /// `dev_buf` must point to `batch * fft_size` c64 on the plan's device; `stream` is a `cudaStream_t`.
    /// `plan` comes from `Plan::as_raw()` of any of the three plan types.
    pub unsafe fn c64_fwd(plan: *const ffi::cfft_plan, dev_buf: *mut c_void, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_fwd(plan, dev_buf, batch, stream));
    }
    /// # Safety
    /// See [`c64_fwd`].
    pub unsafe fn c64_inv(plan: *const ffi::cfft_plan, dev_buf: *mut c_void, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_inv(plan, dev_buf, batch, stream));
    }
    /// `fwd` on `batch` rows that start `row_stride >= n` c64 apart (polynomials inside larger records, transformed where
    /// they are; elements between the rows are not touched).
    /// # Safety
    /// `dev_buf` addresses `(batch - 1) * row_stride + n` c64 on the plan's device.
    pub unsafe fn c64_fwd_strided(plan: *const ffi::cfft_plan, dev_buf: *mut c_void, row_stride: u64, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_fwd_strided(plan, dev_buf, row_stride, batch, stream));
    }
    /// # Safety
    /// See [`c64_fwd_strided`].
    pub unsafe fn c64_inv_strided(plan: *const ffi::cfft_plan, dev_buf: *mut c_void, row_stride: u64, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_inv_strided(plan, dev_buf, row_stride, batch, stream));
    }
    /// fft128 `fwd` / `inv` on rows `row_stride >= n` doubles apart in each of the four planes.
    /// # Safety
    /// Every plane addresses `(batch - 1) * row_stride + n` doubles on the plan's device.
    pub unsafe fn f128_fwd_strided(plan: *const ffi::cfft_plan, p: [*mut f64; 4], row_stride: u64, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_f128_fwd_strided(plan, p[0], p[1], p[2], p[3], row_stride, batch, stream));
    }
    /// # Safety
    /// See [`f128_fwd_strided`].
    pub unsafe fn f128_inv_strided(plan: *const ffi::cfft_plan, p: [*mut f64; 4], row_stride: u64, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_f128_inv_strided(plan, p[0], p[1], p[2], p[3], row_stride, batch, stream));
    }
    /// `out[r][o] = inv(sum_k fwd(a[r][k]) * b[r][k][o])`, `o < n_out`: the GLWE external product, every forward transform feeding
    /// all outputs (one kernel for `n_out == 2`, `n` = 512 / 1024 / 2048); bit-identical to [`c64_fwd_mul_inv`] once per output.
    /// # Safety
    /// `a`: `batch * k_terms * n` c64, `b`: `k_terms * n_out * n` (shared, `b_row_stride == 0`) or `batch` rows `b_row_stride` apart,
    /// `out`: `batch * n_out * n` c64 overlapping neither, all on the plan's device.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn c64_fwd_mul_inv_multi(plan: *const ffi::cfft_plan, a: *const c_void, k_terms: u64, b: *const c_void, b_row_stride: u64, n_out: u64, out: *mut c_void, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_fwd_mul_inv_multi(plan, a, k_terms, b, b_row_stride, n_out, out, batch, stream));
    }
    /// `lhs[i] *= rhs[i]` on `len` device c64 (the Fourier-domain step between `fwd` and `inv`; the bits of
    /// `num_complex`'s `*`).
    /// # Safety
    /// Both pointers must address `len` c64 on `device`; `stream` is a `cudaStream_t`.
    pub unsafe fn c64_mul_assign(device: i32, lhs: *mut c_void, rhs: *const c_void, len: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_mul_assign(device, lhs, rhs, len, stream));
    }
    /// `acc[i] += a[i] * b[i]`.
    /// # Safety
    /// See [`c64_mul_assign`].
    pub unsafe fn c64_mul_add_assign(device: i32, acc: *mut c_void, a: *const c_void, b: *const c_void, len: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_mul_add_assign(device, acc, a, b, len, stream));
    }
    /// `acc[r] <- [acc[r] +] fwd(a[r]) * b[r]` in the Fourier domain (no inverse): forward transform and multiply[-accumulate]
    /// in one call, for loops that produce terms one at a time or feed several accumulators from one input; finish with
    /// [`c64_inv`].  Strides in c64 elements; `a_row_stride` a positive multiple of `n`, `b_row_stride == 0` shares `b`.
    /// # Safety
    /// `a`, `b`, `acc` address `batch` rows of `n` c64 at those strides on the plan's device; `acc` aliases neither input.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn c64_fwd_mul_add(plan: *const ffi::cfft_plan, a: *const c_void, a_row_stride: u64, b: *const c_void, b_row_stride: u64, acc: *mut c_void, accumulate: bool, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_fwd_mul_add(plan, a, a_row_stride, b, b_row_stride, acc, accumulate as i32, batch, stream));
    }
    /// fft128: `lhs <- inv((fwd(lhs) * rhs) * factor)` on `batch` transforms in one call (one kernel for
    /// `n <= 4096`): the negacyclic product of the reference's tests, bit-identical to `fwd`, the scalar
    /// `cplx_mul` loop and `inv`.  `rhs_row_stride` is 0 (one Fourier-domain `rhs` row shared by the batch) or `n`.
    /// # Safety
    /// The `l_*` planes hold `batch * n` doubles, the `r_*` planes `n` or `batch * n`, all on the plan's device.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn f128_fwd_mul_inv(plan: *const ffi::cfft_plan, l: [*mut f64; 4], r: [*const f64; 4], rhs_row_stride: u64, factor: f64, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_f128_fwd_mul_inv(plan, l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], rhs_row_stride, factor, batch, stream));
    }
    /// Integer polynomials (2 n signed 64-bit coefficients per row) to the Fourier domain with the fold, the conversion and
    /// the negacyclic twist fused into the forward transform (`cfft_c64_poly_fwd`); `flags`: `ffi::CFFT_POLY_TORUS` or 0.
    /// # Safety
    /// `poly`: `batch * 2 n` i64, `fourier`: `batch * n` c64 (16-byte aligned), both on the plan's device, not overlapping.
    pub unsafe fn c64_poly_fwd(plan: *const ffi::cfft_plan, poly: *const i64, fourier: *mut c_void, batch: u64, flags: u32, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_poly_fwd(plan, poly, fourier, batch, flags, stream));
    }
    /// The way back: inverse transform, untwist, 1 / n, rounding (`f64::round`; torus: fractional part x 2^64), optionally added
    /// to `poly` modulo 2^64 (`cfft_c64_poly_inv`); `fourier` is not modified.
    /// # Safety
    /// See [`c64_poly_fwd`].
    pub unsafe fn c64_poly_inv(plan: *const ffi::cfft_plan, fourier: *const c_void, poly: *mut i64, batch: u64, flags: u32, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_poly_inv(plan, fourier, poly, batch, flags, stream));
    }
    /// `out[r] (+)= round(untwist(inv(sum_k fwd(twist(fold(a[r][k]))) * b[r][k])))`: a whole negacyclic product / external
    /// product step, integers in, integers out, one kernel for `n <= 4096` (`cfft_c64_poly_mul`).
    /// # Safety
    /// `a`: `batch * k_terms * 2 n` i64, `b`: Fourier-domain c64 as in [`c64_fwd_mul_inv`], `out`: `batch * 2 n` i64, all on the
    /// plan's device; `out` may equal `a` only when `k_terms == 1` and not accumulating.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn c64_poly_mul(plan: *const ffi::cfft_plan, a: *const i64, k_terms: u64, b: *const c_void, b_row_stride: u64, out: *mut i64, batch: u64, flags: u32, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_poly_mul(plan, a, k_terms, b, b_row_stride, out, batch, flags, stream));
    }
    /// Element-wise `f128` operator on device planes, bit-identical to the host scalars of `fft128::f128`
    /// (`op`: `CFFT_F128_*` of include/cfft_b200.h; the lo plane of an f64 operand may be null).
    /// # Safety
    /// Every non-null plane addresses `len` doubles on `device`.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn f128_binary_op(device: i32, op: i32, a: [*const f64; 2], b: [*const f64; 2], out: [*mut f64; 2], len: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_f128_binary_op(device, op, a[0], a[1], b[0], b[1], out[0], out[1], len, stream));
    }
    /// `sqr`, `abs`, `neg`, `sincospi` (second output = cos), `is_nan` on device planes.
    /// # Safety
    /// See [`f128_binary_op`]; `out2` is only written by `sincospi`.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn f128_unary_op(device: i32, op: i32, a: [*const f64; 2], out: [*mut f64; 2], out2: [*mut f64; 2], len: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_f128_unary_op(device, op, a[0], a[1], out[0], out[1], out2[0], out2[1], len, stream));
    }
    /// `PartialOrd` of `f128` element-wise: -1 / 0 / 1 / 2 (unordered) per element; null `b[1]` compares with f64 values.
    /// # Safety
    /// See [`f128_binary_op`]; `out` addresses `len` bytes.
    pub unsafe fn f128_compare(device: i32, a: [*const f64; 2], b: [*const f64; 2], out: *mut i8, len: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_f128_compare(device, a[0], a[1], b[0], b[1], out, len, stream));
    }
    /// `out[r] = inv(sum_k fwd(a[r][k]) * b[r][k])` for `batch` rows of `k_terms` polynomials: forward transforms,
    /// element-wise multiply-accumulate and the inverse transform in one call (one kernel for plans of the
    /// `(Dif16, 256)` family with `n <= 4096`), bit-identical to the separate calls.  `b_row_stride == 0` shares
    /// `b` (`[k_terms][n]`) between all rows.
    /// # Safety
    /// `a`: `batch * k_terms * n` c64, `b`: `k_terms * n` (shared) or `batch` rows `b_row_stride` apart, `out`:
    /// `batch * n` c64, all on the plan's device; `out` may equal `a` only when `k_terms == 1`.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn c64_fwd_mul_inv(plan: *const ffi::cfft_plan, a: *const c_void, k_terms: u64, b: *const c_void, b_row_stride: u64, out: *mut c_void, batch: u64, stream: *mut c_void) {
        ffi::check(ffi::cfft_c64_fwd_mul_inv(plan, a, k_terms, b, b_row_stride, out, batch, stream));
    }
}
