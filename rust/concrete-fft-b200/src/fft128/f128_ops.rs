//! Host-side scalar arithmetic of `f128` (a double-double: value = hi + lo), the operator surface of concrete-fft's
//! `src/fft128/f128_ops.rs:48-618`.  These are plain host scalars -- callers use them around the transform -- so they are pure
//! Rust here; the array forms of the same operators on GPU buffers are `device::f128_binary_op` / `f128_unary_op` /
//! `f128_compare` (C ABI `cfft_f128_*`), which return the same bits.
//!
//! Building blocks are the classical error-free transformations (Knuth two-sum, Dekker fast-two-sum, FMA two-product).
use super::f128;
use core::cmp::Ordering;
use core::ops::{Add, AddAssign, Div, DivAssign, Mul, MulAssign, Neg, Sub, SubAssign};

/// s = fl(a + b) and the rounding error of that sum, valid when |a| >= |b|.
#[inline(always)]
fn fast_two_sum(a: f64, b: f64) -> (f64, f64) {
    let s = a + b;
    (s, b - (s - a))
}
/// s = fl(a + b) and its rounding error, no ordering assumption.
#[inline(always)]
fn two_sum(a: f64, b: f64) -> (f64, f64) {
    let s = a + b;
    let bv = s - a;
    (s, (a - (s - bv)) + (b - bv))
}
/// d = fl(a - b) and its rounding error.
#[inline(always)]
fn two_diff(a: f64, b: f64) -> (f64, f64) {
    let d = a - b;
    let bv = d - a;
    (d, (a - (d - bv)) - (b + bv))
}
/// p = fl(a * b) and its rounding error (one fused multiply-add).
#[inline(always)]
fn two_prod(a: f64, b: f64) -> (f64, f64) {
    let p = a * b;
    (p, a.mul_add(b, -p))
}
#[inline(always)]
fn renorm(hi: f64, lo: f64) -> f128 {
    let (h, l) = fast_two_sum(hi, lo);
    f128(h, l)
}

impl From<f64> for f128 {
    #[inline(always)]
    fn from(v: f64) -> Self {
        f128(v, 0.0)
    }
}

impl f128 {
    /// 128-bit approximation of pi (f128_ops.rs:579)
    pub const PI: Self = f128(3.141592653589793, 1.2246467991473532e-16);

    // ---- sums --------------------------------------------------------------------------------------
    /// f128_ops.rs:279-283
    #[inline(always)]
    pub fn add_f64_f64(a: f64, b: f64) -> Self {
        let (s, e) = two_sum(a, b);
        f128(s, e)
    }
    /// f128_ops.rs:286-291
    #[inline(always)]
    pub fn add_f128_f64(a: f128, b: f64) -> Self {
        let (s, e) = two_sum(a.0, b);
        renorm(s, e + a.1)
    }
    /// f128_ops.rs:294-298
    #[inline(always)]
    pub fn add_f64_f128(a: f64, b: f128) -> Self {
        Self::add_f128_f64(b, a)
    }
    /// Cheaper sum with a slightly larger error bound (what the transform's butterflies use), f128_ops.rs:302-307
    #[inline(always)]
    pub fn add_estimate_f128_f128(a: f128, b: f128) -> Self {
        let (s, e) = two_sum(a.0, b.0);
        renorm(s, e + (a.1 + b.1))
    }
    /// f128_ops.rs:311-321
    #[inline(always)]
    pub fn add_f128_f128(a: f128, b: f128) -> Self {
        let (s, e) = two_sum(a.0, b.0);
        let (t, f) = two_sum(a.1, b.1);
        let f128(s, e) = renorm(s, e + t);
        renorm(s, e + f)
    }

    // ---- differences -------------------------------------------------------------------------------
    /// f128_ops.rs:324-328
    #[inline(always)]
    pub fn sub_f64_f64(a: f64, b: f64) -> Self {
        let (d, e) = two_diff(a, b);
        f128(d, e)
    }
    /// f128_ops.rs:331-336
    #[inline(always)]
    pub fn sub_f128_f64(a: f128, b: f64) -> Self {
        let (d, e) = two_diff(a.0, b);
        renorm(d, e + a.1)
    }
    /// f128_ops.rs:339-345
    #[inline(always)]
    pub fn sub_f64_f128(a: f64, b: f128) -> Self {
        let (d, e) = two_diff(a, b.0);
        renorm(d, e - b.1)
    }
    /// f128_ops.rs:350-356
    #[inline(always)]
    pub fn sub_estimate_f128_f128(a: f128, b: f128) -> Self {
        let (d, e) = two_diff(a.0, b.0);
        let e = e + a.1;
        renorm(d, e - b.1)
    }
    /// f128_ops.rs:360-370
    #[inline(always)]
    pub fn sub_f128_f128(a: f128, b: f128) -> Self {
        let (d, e) = two_diff(a.0, b.0);
        let (t, f) = two_diff(a.1, b.1);
        let f128(d, e) = renorm(d, e + t);
        renorm(d, e + f)
    }

    // ---- products ----------------------------------------------------------------------------------
    /// f128_ops.rs:373-377
    #[inline(always)]
    pub fn mul_f64_f64(a: f64, b: f64) -> Self {
        let (p, e) = two_prod(a, b);
        f128(p, e)
    }
    /// f128_ops.rs:380-385
    #[inline(always)]
    pub fn mul_f128_f64(a: f128, b: f64) -> Self {
        let (p, e) = two_prod(a.0, b);
        renorm(p, e + (a.1 * b))
    }
    /// f128_ops.rs:388-391
    #[inline(always)]
    pub fn mul_f64_f128(a: f64, b: f128) -> Self {
        Self::mul_f128_f64(b, a)
    }
    /// f128_ops.rs:395-400
    #[inline(always)]
    pub fn mul_f128_f128(a: f128, b: f128) -> Self {
        let (p, e) = two_prod(a.0, b.0);
        renorm(p, e + (a.0 * b.1 + a.1 * b.0))
    }
    /// f128_ops.rs:404-409
    #[inline(always)]
    pub fn sqr(self) -> Self {
        let (p, e) = two_prod(self.0, self.0);
        renorm(p, e + 2.0 * (self.0 * self.1))
    }

    // ---- quotients ---------------------------------------------------------------------------------
    /// f128_ops.rs:413-428
    #[inline(always)]
    pub fn div_f64_f64(a: f64, b: f64) -> Self {
        let q1 = a / b;
        let (p, pe) = two_prod(q1, b);
        let (s, e) = two_diff(a, p);
        let q2 = (s + (e - pe)) / b;
        renorm(q1, q2)
    }
    /// f128_ops.rs:431-448
    #[inline(always)]
    pub fn div_f128_f64(a: f128, b: f64) -> Self {
        let q1 = a.0 / b;
        let (p, pe) = two_prod(q1, b);
        let (s, e) = two_diff(a.0, p);
        let e = e + a.1;
        let q2 = (s + (e - pe)) / b;
        renorm(q1, q2)
    }
    /// f128_ops.rs:451-454
    #[inline(always)]
    pub fn div_f64_f128(a: f64, b: f128) -> Self {
        Self::div_f128_f128(a.into(), b)
    }
    /// f128_ops.rs:457-474
    #[inline(always)]
    pub fn div_estimate_f128_f128(a: f128, b: f128) -> Self {
        let q1 = a.0 / b.0;
        let r = b * q1;
        let (s, e) = two_diff(a.0, r.0);
        let e = e - r.1;
        let e = e + a.1;
        let q2 = (s + e) / b.0;
        renorm(q1, q2)
    }
    /// Three quotient digits, f128_ops.rs:477-491
    #[inline(always)]
    pub fn div_f128_f128(a: f128, b: f128) -> Self {
        let q1 = a.0 / b.0;
        let r = a - b * q1;
        let q2 = r.0 / b.0;
        let r = r - q2 * b;
        let q3 = r.0 / b.0;
        renorm(q1, q2) + q3
    }

    // ---- helpers -----------------------------------------------------------------------------------
    /// f128_ops.rs:494-496
    #[inline(always)]
    pub fn to_f64(self) -> f64 {
        self.0
    }
    /// f128_ops.rs:499-501
    #[inline(always)]
    pub fn is_nan(self) -> bool {
        self.0.is_nan() || self.1.is_nan()
    }
    /// f128_ops.rs:506-511
    #[inline(always)]
    pub fn abs(self) -> Self {
        if self.0 < 0.0 {
            -self
        } else {
            self
        }
    }

    /// Taylor part of `sincospi` on the reduced argument (|x| <= 1/32): returns (sin(pi x), cos(pi x)), f128_ops.rs:514-532
    fn sincospi_taylor(self) -> (Self, Self) {
        let x2 = self.sqr();
        let (mut sin_over_x, mut cos, mut power) = (Self::PI, f128(1.0, 0.0), f128(1.0, 0.0));
        for (s, c) in SINPI_TAYLOR.iter().zip(COSPI_TAYLOR.iter()) {
            power *= x2;
            sin_over_x += *s * power;
            cos += *c * power;
        }
        (sin_over_x * self, cos)
    }

    /// (sin(pi x), cos(pi x)) for x in [-1, 1]; panics outside, like the reference (f128_ops.rs:534-575).
    pub fn sincospi(self) -> (Self, Self) {
        if self > 1.0 || self < -1.0 {
            panic!("only inputs in [-1, 1] are currently supported, received: {self:?}");
        }
        // reduce by the nearest multiple of 1/2, then of 1/16
        let half_turns = (self.0 * 2.0).round();
        let r = self - half_turns * 0.5;
        let sixteenths = (r.0 * 16.0).round();
        let r = r - sixteenths * (1.0 / 16.0);
        let (p, q) = (half_turns as isize, sixteenths as isize);
        let (sr, cr) = r.sincospi_taylor();
        let (s, c) = if q == 0 {
            (sr, cr)
        } else {
            let k = q.unsigned_abs() - 1;
            let (u, v) = (COS_K_PI_OVER_16[k], SIN_K_PI_OVER_16[k]);
            if q > 0 {
                (u * sr + v * cr, u * cr - v * sr)
            } else {
                (u * sr - v * cr, u * cr + v * sr)
            }
        };
        match p {
            0 => (s, c),
            1 => (c, -s),
            -1 => (-c, s),
            _ => (-s, -c),
        }
    }
}

// Taylor coefficients of sin(pi x) / x - pi and cos(pi x) - 1 in x^2 and the values at k pi / 16, each as an f128
// (the reference's tables, f128_ops.rs:581-617; regenerated there from a 1024-bit pi, :1217-1272).
const SINPI_TAYLOR: [f128; 9] = [
    f128(-5.16771278004997, 2.2665622825789447e-16),
    f128(2.5501640398773455, -7.931006345326556e-17),
    f128(-0.5992645293207921, 2.845026112698218e-17),
    f128(0.08214588661112823, -3.847292805297656e-18),
    f128(-0.0073704309457143504, -3.328281165603432e-19),
    f128(0.00046630280576761255, 1.0704561733683463e-20),
    f128(-2.1915353447830217e-5, 1.4648526682685598e-21),
    f128(7.952054001475513e-7, 1.736540361519021e-23),
    f128(-2.2948428997269873e-8, -7.376346207041088e-26),
];
const COSPI_TAYLOR: [f128; 9] = [
    f128(-4.934802200544679, -3.1326477543698557e-16),
    f128(4.0587121264167685, -2.6602000824298645e-16),
    f128(-1.3352627688545895, 3.1815237892149862e-18),
    f128(0.2353306303588932, -1.2583065576724427e-18),
    f128(-0.02580689139001406, 1.170191067939226e-18),
    f128(0.0019295743094039231, -9.669517939986956e-20),
    f128(-0.0001046381049248457, -2.421206183964864e-21),
    f128(4.303069587032947e-6, -2.864010082936791e-22),
    f128(-1.3878952462213771e-7, -7.479362090417238e-24),
];
const SIN_K_PI_OVER_16: [f128; 4] = [
    f128(0.19509032201612828, -7.991079068461731e-18),
    f128(0.3826834323650898, -1.0050772696461588e-17),
    f128(0.5555702330196022, 4.709410940561677e-17),
    f128(0.7071067811865476, -4.833646656726457e-17),
];
const COS_K_PI_OVER_16: [f128; 4] = [
    f128(0.9807852804032304, 1.8546939997825006e-17),
    f128(0.9238795325112867, 1.7645047084336677e-17),
    f128(0.8314696123025452, 1.4073856984728024e-18),
    f128(0.7071067811865476, -4.833646656726457e-17),
];

// ---- operator impls (f128_ops.rs:48-274): every combination of f128 and f64 operands --------------------------------
macro_rules! binop {
    ($tr:ident, $f:ident, $atr:ident, $af:ident, $ff:ident, $fd:ident, $df:ident) => {
        impl $tr<f128> for f128 {
            type Output = f128;
            #[inline(always)]
            fn $f(self, rhs: f128) -> f128 {
                f128::$ff(self, rhs)
            }
        }
        impl $tr<f64> for f128 {
            type Output = f128;
            #[inline(always)]
            fn $f(self, rhs: f64) -> f128 {
                f128::$fd(self, rhs)
            }
        }
        impl $tr<f128> for f64 {
            type Output = f128;
            #[inline(always)]
            fn $f(self, rhs: f128) -> f128 {
                f128::$df(self, rhs)
            }
        }
        impl $atr<f128> for f128 {
            #[inline(always)]
            fn $af(&mut self, rhs: f128) {
                *self = $tr::$f(*self, rhs)
            }
        }
        impl $atr<f64> for f128 {
            #[inline(always)]
            fn $af(&mut self, rhs: f64) {
                *self = $tr::$f(*self, rhs)
            }
        }
    };
}
binop!(Add, add, AddAssign, add_assign, add_f128_f128, add_f128_f64, add_f64_f128);
binop!(Sub, sub, SubAssign, sub_assign, sub_f128_f128, sub_f128_f64, sub_f64_f128);
binop!(Mul, mul, MulAssign, mul_assign, mul_f128_f128, mul_f128_f64, mul_f64_f128);
binop!(Div, div, DivAssign, div_assign, div_f128_f128, div_f128_f64, div_f64_f128);

impl Neg for f128 {
    type Output = f128;
    #[inline(always)]
    fn neg(self) -> f128 {
        f128(-self.0, -self.1)
    }
}

impl PartialEq<f128> for f128 {
    #[inline(always)]
    fn eq(&self, other: &f128) -> bool {
        self.0 == other.0 && self.1 == other.1
    }
}
impl PartialEq<f64> for f128 {
    #[inline(always)]
    fn eq(&self, other: &f64) -> bool {
        *self == f128(*other, 0.0)
    }
}
impl PartialEq<f128> for f64 {
    #[inline(always)]
    fn eq(&self, other: &f128) -> bool {
        *other == *self
    }
}
/// Lexicographic on (hi, lo): the high words decide unless they compare equal (f128_ops.rs:261-274).
impl PartialOrd<f128> for f128 {
    #[inline(always)]
    fn partial_cmp(&self, other: &f128) -> Option<Ordering> {
        match self.0.partial_cmp(&other.0) {
            Some(Ordering::Equal) => self.1.partial_cmp(&other.1),
            decided => decided,
        }
    }
}
impl PartialOrd<f64> for f128 {
    #[inline(always)]
    fn partial_cmp(&self, other: &f64) -> Option<Ordering> {
        self.partial_cmp(&f128(*other, 0.0))
    }
}
impl PartialOrd<f128> for f64 {
    #[inline(always)]
    fn partial_cmp(&self, other: &f128) -> Option<Ordering> {
        f128(*self, 0.0).partial_cmp(other)
    }
}
