// CPU check of the table form of the JC / K2P epilogue used by the alignment-free kernels for rows of
// at most 2048 columns: ln k as a 64-bit fixed-point table (58 fractional bits), the logarithm of a
// ratio of counts as an exact integer difference.  Compared with the oracle's libm formulas
// (oracle/taxi_oracle.c:282-285) over every n <= NMAX and a dense sample of (ts, tv): largest
// relative deviation (must stay well under the 1e-12 of north_star) and equality of the NaN pattern.
// Usage: gcc -O2 -o /tmp/mtc tools/metrics_table_check.c -lm && /tmp/mtc
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define NMAX 2048
static int64_t T[3 * NMAX + 1];

int main(void)
{
    for (int k = 1; k <= 3 * NMAX; ++k) T[k] = (int64_t)llroundl(logl((long double)k) * 0x1p58L);
    const double scale = 0x1p-58;
    double worst_jc = 0, worst_k2p = 0;
    long long cases = 0, nan_mismatch = 0;
    uint64_t rng = 88172645463325252ULL;
    for (int n = 1; n <= NMAX; ++n) {
        const int exhaustive = n <= 160;                             // every (ts, tv) of short rows, a dense sample of longer ones
        const int reps = exhaustive ? (n + 1) * (n + 1) : 4000;
        for (int rep = 0; rep < reps; ++rep) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            int ts, tv;
            if (exhaustive) { ts = rep % (n + 1); tv = rep / (n + 1); }
            else if (rep < 64) { ts = rep % 8; tv = rep / 8; }      // the smallest distances: the worst relative errors
            else if (rep < 1000) { tv = (int)(rng % (uint64_t)(n / 2 + 1)); ts = (n - tv) / 2 - (int)((rng >> 40) % 3); if (ts < 0) ts = 0; }   // around 1 - 2P - Q = 0
            else { ts = (int)(rng % (uint64_t)(n + 1)); tv = (int)((rng >> 20) % (uint64_t)(n + 1)); }
            if (ts + tv > n) continue;
            const int d = ts + tv;
            const double p = (double)d / n, P = (double)ts / n, Q = (double)tv / n;
            double jc = -0.75 * log(1.0 - 4.0 * p / 3.0);
            double k2p = -0.5 * log((1.0 - 2.0 * P - Q) * sqrt(1.0 - 2.0 * Q));
            if (!isfinite(jc)) jc = NAN;
            if (!isfinite(k2p)) k2p = NAN;
            const int A = 3 * n - 4 * d, a = n - 2 * ts - tv, b = n - 2 * tv;
            // an argument that is exactly zero goes through the floating-point formula (the reference's rounding
            // decides between -inf and the logarithm of a tiny residue there), as in the kernels
            // 3n = 4d makes d / n = 0.75 and 4p/3 = 1 exactly (jc = -inf = None), n = 2 tv makes Q = 0.5 and the
            // square root 0 exactly (None); only n = 2 ts + tv leaves a rounding residue of 1 - 2P - Q that the
            // floating-point formula has to decide, as in the kernels
            double jt = A > 0 ? -0.75 * ((double)(T[A] - T[3 * n]) * scale) + 0.0 : NAN;
            double kt = (a > 0 && b > 0) ? -0.25 * ((double)(2 * T[a] + T[b] - 3 * T[n]) * scale) + 0.0 : (a == 0 && b > 0) ? k2p : NAN;
            ++cases;
            if (isnan(jc) != isnan(jt) || isnan(k2p) != isnan(kt)) { ++nan_mismatch; continue; }
            if (!isnan(jc) && jc != 0.0) { double e = fabs(jt - jc) / fabs(jc); if (e > worst_jc) worst_jc = e; }
            if (!isnan(jc) && jc == 0.0 && jt != 0.0) ++nan_mismatch;
            if (!isnan(k2p) && k2p != 0.0) { double e = fabs(kt - k2p) / fabs(k2p); if (e > worst_k2p) worst_k2p = e; }
            if (!isnan(k2p) && k2p == 0.0 && kt != 0.0) ++nan_mismatch;
        }
    }
    printf("{\"cases\": %lld, \"nmax\": %d, \"worst_rel_jc\": %.3g, \"worst_rel_k2p\": %.3g, \"nan_or_zero_mismatches\": %lld}\n",
           cases, NMAX, worst_jc, worst_k2p, nan_mismatch);
    return nan_mismatch != 0 || worst_jc > 5e-13 || worst_k2p > 5e-13;
}
