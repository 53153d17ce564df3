// tables.cc -- see tables.h.  Built with -ffp-contract=off: every fused multiply-add below is
// an explicit std::fma, nothing else may be contracted.
#include "tables.h"

#include <cmath>
#include <limits>

namespace cfft {

void sincospi64(double a, double &s_out, double &c_out)
{
    const double az = a * 0.0;
    if (!(std::fabs(a) < 9007199254740992.0)) a = az;

    double r = std::round(a + a);
    const long long quad = static_cast<long long>(r);
    const double t = std::fma(-0.5, r, a);
    const double t2 = t * t;

    static const double C[8] = {-1.0369917389758117e-4, 1.9294935641298806e-3, -2.5806887942825395e-2,
                                2.3533063028328211e-1,  -1.3352627688538006e+0, 4.0587121264167623e+0,
                                -4.9348022005446790e+0, 1.0000000000000000e+0};
    static const double S[6] = {4.6151442520157035e-4,  -7.3700183130883555e-3, 8.2145868949323936e-2,
                                -5.9926452893214921e-1, 2.5501640398732688e+0,  -5.1677127800499516e+0};
    double c = C[0];
    for (int i = 1; i < 8; i++) c = std::fma(c, t2, C[i]);
    r = S[0];
    for (int i = 1; i < 6; i++) r = std::fma(r, t2, S[i]);
    const double t3 = t2 * t;
    r = r * t3;
    double s = std::fma(t, 3.1415926535897931e+0, r);

    if (quad & 2) { s = 0.0 - s; c = 0.0 - c; }
    if (quad & 1) { const double ns = 0.0 - s; s = c; c = ns; }
    if (a == std::floor(a)) s = az;
    s_out = s;
    c_out = c;
}

void init_wt(size_t r, size_t n, cplx *w, cplx *w_inv)
{
    if (n < r) return;
    const size_t nr = n / r;
    const double theta = -2.0 / static_cast<double>(n);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (size_t i = 0; i < 2 * n; i++) w[i] = cplx{nan, nan};
    for (size_t p = 0; p < nr; p++)
        for (size_t k = 1; k < r; k++) {
            double s, c;
            sincospi64(theta * static_cast<double>(k * p), s, c);
            w[p + k * nr] = w[n + r * p + k] = cplx{c, s};
            w_inv[p + k * nr] = w_inv[n + r * p + k] = cplx{c, -s};
        }
}

int algo_radix(int algo) { return 2 << (algo >> 1); }
bool algo_is_dit(int algo) { return (algo & 1) != 0; }
unsigned ilog2(uint64_t n) { unsigned b = 0; while ((n >> b) > 1) b++; return b; }
bool is_pow2(uint64_t n) { return n != 0 && (n & (n - 1)) == 0; }
int top_radix(uint64_t n, uint64_t base_n) { return n == 2 * base_n ? 2 : (n == 4 * base_n ? 4 : 8); }

void init_unordered_twiddles(size_t n, size_t base_n, size_t base_r, std::vector<cplx> &w,
                             std::vector<cplx> &w_inv)
{
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const size_t total = n + base_n;
    w.assign(total, cplx{nan, nan});
    w_inv.assign(total, cplx{nan, nan});
    size_t head = 0, tail = total; // fwd levels grow from the front, inv levels from the back
    size_t cur = n;
    while (cur > base_n) {
        const size_t r = static_cast<size_t>(top_radix(cur, base_n));
        const size_t m = cur / r, lvl = (r - 1) * m;
        const double theta = 2.0 / static_cast<double>(cur);
        cplx *wf = w.data() + head;
        cplx *wi = w_inv.data() + (tail - lvl);
        for (size_t p = 0; p < m; p++)
            for (size_t k = 1; k < r; k++) {
                double s, c;
                sincospi64(theta * static_cast<double>(k * p), s, c);
                wf[(r - 1) * p + (k - 1)] = cplx{c, -s};
                wi[(r - 1) * p + (k - 1)] = cplx{c, s};
            }
        head += lvl;
        tail -= lvl;
        cur = m;
    }
    // base table: 2 * base_n entries at w[head..], w_inv[0..2 base_n) (tail == 2 base_n here)
    init_wt(base_r, cur, w.data() + head, w_inv.data());
}

uint64_t bit_rev(unsigned nbits, uint64_t i)
{
    uint64_t r = 0;
    for (unsigned b = 0; b < nbits; b++) r |= ((i >> b) & 1) << (nbits - 1 - b);
    return r;
}
uint64_t bit_rev_twice(unsigned nbits, unsigned base_nbits, uint64_t i)
{
    const uint64_t i_rev = bit_rev(nbits, i);
    const uint64_t mask = (uint64_t{1} << base_nbits) - 1;
    return (i_rev & ~mask) | bit_rev(base_nbits, i_rev & mask);
}
uint64_t bit_rev_twice_inv(unsigned nbits, unsigned base_nbits, uint64_t i)
{
    const uint64_t mask = (uint64_t{1} << base_nbits) - 1;
    return bit_rev(nbits, (i & ~mask) | bit_rev(base_nbits, i & mask));
}

// ---------------------------------------------------------------------------------------------
// double-double arithmetic used only to build the fft128 twiddles (host, cold path)
// src/fft128/f128_ops.rs:6-40, 311-321, 331-336, 360-370, 395-409
// ---------------------------------------------------------------------------------------------
namespace {
struct dd { double hi, lo; };

inline dd quick_two_sum(double a, double b) { double s = a + b; return {s, b - (s - a)}; }
inline dd two_sum(double a, double b) { double s = a + b, bb = s - a; return {s, (a - (s - bb)) + (b - bb)}; }
inline dd two_diff(double a, double b) { double s = a - b, bb = s - a; return {s, (a - (s - bb)) - (b + bb)}; }
inline dd two_prod(double a, double b) { double p = a * b; return {p, std::fma(a, b, -p)}; }

inline dd add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s = quick_two_sum(s.hi, s.lo + t.hi);
    return quick_two_sum(s.hi, s.lo + t.lo);
}
inline dd sub(dd a, dd b)
{
    dd s = two_diff(a.hi, b.hi), t = two_diff(a.lo, b.lo);
    s = quick_two_sum(s.hi, s.lo + t.hi);
    return quick_two_sum(s.hi, s.lo + t.lo);
}
inline dd sub_d(dd a, double b)
{
    dd s = two_diff(a.hi, b);
    return quick_two_sum(s.hi, s.lo + a.lo);
}
inline dd mul(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    return quick_two_sum(p.hi, p.lo + (a.hi * b.lo + a.lo * b.hi));
}
inline dd sqr(dd a)
{
    dd p = two_prod(a.hi, a.hi);
    return quick_two_sum(p.hi, p.lo + 2.0 * (a.hi * a.lo));
}
inline dd neg(dd a) { return {-a.hi, -a.lo}; }

// src/fft128/f128_ops.rs:578-618
const dd kPi = {3.141592653589793, 1.2246467991473532e-16};
const dd kSinTaylor[9] = {
    {-5.16771278004997, 2.2665622825789447e-16},      {2.5501640398773455, -7.931006345326556e-17},
    {-0.5992645293207921, 2.845026112698218e-17},     {0.08214588661112823, -3.847292805297656e-18},
    {-0.0073704309457143504, -3.328281165603432e-19}, {0.00046630280576761255, 1.0704561733683463e-20},
    {-2.1915353447830217e-5, 1.4648526682685598e-21}, {7.952054001475513e-7, 1.736540361519021e-23},
    {-2.2948428997269873e-8, -7.376346207041088e-26}};
const dd kCosTaylor[9] = {
    {-4.934802200544679, -3.1326477543698557e-16},   {4.0587121264167685, -2.6602000824298645e-16},
    {-1.3352627688545895, 3.1815237892149862e-18},   {0.2353306303588932, -1.2583065576724427e-18},
    {-0.02580689139001406, 1.170191067939226e-18},   {0.0019295743094039231, -9.669517939986956e-20},
    {-0.0001046381049248457, -2.421206183964864e-21}, {4.303069587032947e-6, -2.864010082936791e-22},
    {-1.3878952462213771e-7, -7.479362090417238e-24}};
const dd kSin16[4] = {{0.19509032201612828, -7.991079068461731e-18}, {0.3826834323650898, -1.0050772696461588e-17},
                      {0.5555702330196022, 4.709410940561677e-17},   {0.7071067811865476, -4.833646656726457e-17}};
const dd kCos16[4] = {{0.9807852804032304, 1.8546939997825006e-17}, {0.9238795325112867, 1.7645047084336677e-17},
                      {0.8314696123025452, 1.4073856984728024e-18}, {0.7071067811865476, -4.833646656726457e-17}};

// src/fft128/f128_ops.rs:514-575
void dd_sincospi(dd x, dd &s_out, dd &c_out)
{
    const double p = std::round(x.hi * 2.0);
    dd r = sub_d(x, p * 0.5);
    const double q = std::round(r.hi * 16.0);
    r = sub_d(r, q * (1.0 / 16.0));

    dd sinc = kPi, cosv = {1.0, 0.0}, pw = {1.0, 0.0};
    const dd r2 = sqr(r);
    for (int i = 0; i < 9; i++) {
        pw = mul(pw, r2);
        sinc = add(sinc, mul(kSinTaylor[i], pw));
        cosv = add(cosv, mul(kCosTaylor[i], pw));
    }
    const dd sin_r = mul(sinc, r), cos_r = cosv;

    dd s = sin_r, c = cos_r;
    const long qi = static_cast<long>(q);
    if (qi != 0) {
        const dd u = kCos16[(qi < 0 ? -qi : qi) - 1], v = kSin16[(qi < 0 ? -qi : qi) - 1];
        if (qi > 0) {
            s = add(mul(u, sin_r), mul(v, cos_r));
            c = sub(mul(u, cos_r), mul(v, sin_r));
        } else {
            s = sub(mul(u, sin_r), mul(v, cos_r));
            c = add(mul(u, cos_r), mul(v, sin_r));
        }
    }
    switch (static_cast<long>(p)) {
    case 0: s_out = s; c_out = c; break;
    case 1: s_out = c; c_out = neg(s); break;
    case -1: s_out = neg(c); c_out = s; break;
    default: s_out = neg(s); c_out = neg(c); break;
    }
}
} // namespace

void init_negacyclic_twiddles(size_t n, double *re0, double *re1, double *im0, double *im1)
{
    const unsigned bits2n = ilog2(2 * n);
    for (size_t m = 1; m < n; m *= 2)
        for (size_t i = 0; i < m; i++) {
            const size_t k = 2 * m + i, pos = m + i;
            dd s, c;
            dd_sincospi(dd{static_cast<double>(bit_rev(bits2n, k)) / static_cast<double>(2 * n), 0.0}, s, c);
            re0[pos] = c.hi; re1[pos] = c.lo;
            im0[pos] = s.hi; im1[pos] = s.lo;
        }
}

} // namespace cfft
