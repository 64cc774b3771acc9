// CPU emulation of the lean fp64 metric epilogue (same operation sequence as the device code,
// explicit fma), compared with the straightforward libm formulas of the oracle.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double seed_err = 0.0;   // relative perturbation of the reciprocal seed (models MUFU.RCP64H ~ 2^-20)
static double rcp_seed(double b) { float f = (float)(1.0 / b); uint32_t u; memcpy(&u, &f, 4); u &= 0xFFFFF000u; memcpy(&f, &u, 4); return (double)f * (1.0 + seed_err); }
static double rsq_seed(double b) { float f = (float)(1.0 / sqrt(b)); uint32_t u; memcpy(&u, &f, 4); u &= 0xFFFFF000u; memcpy(&f, &u, 4); return (double)f * (1.0 + seed_err); }

static double rcp_full(double b)
{
    double r = rcp_seed(b);
    double e = fma(-b, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    return r;
}
static double div_fast(double a, double b, double r)   // r = rcp_full(b)
{
    double q = a * r;
    double rem = fma(-b, q, a);
    return fma(rem, r, q);
}
static double sqrt_fast(double a)   // a > 0 normal
{
    double y = rsq_seed(a);                 // ~ 1/sqrt(a)
    double h = 0.5 * y;
    double g = a * y;                       // ~ sqrt(a)
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    double d = fma(-g, g, a);
    return fma(d, h, g);
}
static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
    Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
    Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
static double log_fast(double x)   // x > 0 normal
{
    uint64_t u; memcpy(&u, &x, 8);
    int32_t hx = (int32_t)(u >> 32);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int i = (hx + 0x95f64) & 0x100000;      // mantissa above sqrt(2): halve it
    uint64_t v = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (u & 0xffffffffu);
    k += i >> 20;
    double m; memcpy(&m, &v, 8);
    double f = m - 1.0;
    double t = 2.0 + f;
    double s = div_fast(f, t, rcp_full(t));
    double dk = (double)k;
    double z = s * s, w = z * z;
    double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    return fma(dk, ln2_hi, -((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f));
}
static void metrics_fast(int same, int ts, int tv, int gap, double out[4])
{
    const double n = (double)(same + ts + tv);
    if (!(n > 0.0)) { out[0] = out[1] = out[2] = out[3] = NAN; return; }
    const double d = (double)(ts + tv), g = (double)gap;
    const double rn = rcp_full(n);
    const double p = div_fast(d, n, rn);
    out[0] = p;
    out[1] = div_fast(d + g, n + g, rcp_full(n + g));
    const double P = div_fast((double)ts, n, rn), Q = div_fast((double)tv, n, rn);
    const double u = 1.0 - div_fast(4.0 * p, 3.0, 1.0 / 3.0);
    const double b = 1.0 - 2.0 * Q, a = (1.0 - 2.0 * P - Q);
    out[2] = u > 0.0 ? -0.75 * log_fast(u) + 0.0 : NAN;
    double v = (b > 0.0 && a > 0.0) ? a * sqrt_fast(b) : -1.0;
    out[3] = v > 0.0 ? -0.5 * log_fast(v) + 0.0 : NAN;
}
static void metrics_ref(int same, int ts, int tv, int gap, double out[4])
{
    const double s = same, t = ts, v = tv, g = gap, n = s + t + v;
    if (!(n > 0)) { out[0] = out[1] = out[2] = out[3] = NAN; return; }
    const double p = (t + v) / n;
    out[0] = p; out[1] = (t + v + g) / (n + g);
    const double P = t / n, Q = v / n;
    double jc = -0.75 * log(1.0 - 4.0 * p / 3.0);
    double k2p = -0.5 * log((1.0 - 2.0 * P - Q) * sqrt(1.0 - 2.0 * Q));
    out[2] = isfinite(jc) ? jc + 0.0 : NAN; out[3] = isfinite(k2p) ? k2p + 0.0 : NAN;
}
int main(int argc, char** argv)
{
    double worst[4] = {0, 0, 0, 0}; long bad_nan = 0, bitdiff01 = 0, total = 0;
    const double errs[] = {0.0, 4e-7, -4e-7, 9e-7, -9e-7};
    for (int e = 0; e < 5; ++e) {
        seed_err = errs[e];
        srand(7);
        for (int it = 0; it < 3000000; ++it) {
            int n = 1 + rand() % (it % 3 == 0 ? 700 : (it % 3 == 1 ? 60 : 40000));
            int ts = rand() % (n + 1); if (it % 2) ts = ts % (1 + n / 10);
            int tv = rand() % (n - ts + 1); if (it % 2) tv = tv % (1 + n / 10);
            int same = n - ts - tv; int gap = (it % 5 == 0) ? rand() % 50 : 0;
            double a[4], b[4];
            metrics_fast(same, ts, tv, gap, a); metrics_ref(same, ts, tv, gap, b);
            ++total;
            for (int k = 0; k < 4; ++k) {
                if (isnan(a[k]) != isnan(b[k])) { ++bad_nan; if (bad_nan < 10) printf("nan mismatch k=%d same=%d ts=%d tv=%d gap=%d fast=%g ref=%g\n", k, same, ts, tv, gap, a[k], b[k]); continue; }
                if (isnan(a[k])) continue;
                if (k < 2 && memcmp(&a[k], &b[k], 8)) { ++bitdiff01; if (bitdiff01 < 10) printf("bit diff k=%d %d %d %d %d %.17g %.17g\n", k, same, ts, tv, gap, a[k], b[k]); }
                double rel = fabs(a[k] - b[k]) / fmax(fabs(b[k]), 1e-300);
                if (b[k] == 0.0) rel = fabs(a[k]);
                if (rel > worst[k]) worst[k] = rel;
            }
        }
    }
    printf("total %ld nan mismatches %ld bit diffs in p/p-gaps %ld worst rel p %.3g pg %.3g jc %.3g k2p %.3g\n", total, bad_nan, bitdiff01, worst[0], worst[1], worst[2], worst[3]);
    return 0;
}
