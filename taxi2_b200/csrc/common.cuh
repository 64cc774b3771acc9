// Shared declarations of the taxi2_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TAXI_FULL_MASK 0xffffffffu

// Debug build (make bounds / -DTAXI_BOUNDS_CHECK): every access to the traceback arena, the
// stripe-boundary buffer and the gapped-string slots is checked against its extent and a
// violation sets the sticky device status (reported by the next taxi_sync as a TAXI_E_CUDA-class
// error -90 - k).  compute-sanitizer is closed on the pool this was developed on; the GPU tests run
// against this build instead (tools/bounds_check.sh).
#ifdef TAXI_BOUNDS_CHECK
#define TAXI_CHECK(a, cond, k) do { if (!(cond)) atomicMin((a).status, -90 - (k)); } while (0)
#else
#define TAXI_CHECK(a, cond, k) do { } while (0)
#endif

namespace taxi {

// Values in the DP are score*64 + 6 tag bits.  The tag is the priority of the predecessor
// state, replicated three times (bits 0-1, 2-3, 4-5) so that one VIMNMX3 both maximises the
// score and resolves ties in Biopython's order, and the winner's identity survives in the low
// bits of the result.  See DESIGN.md "tagged maxima".
constexpr int TAG_BITS = 6;
constexpr int TAG_MASK = 63;
constexpr int TAG_REP = 21;  // 0b010101: replicate a 2-bit tag into the three fields

struct ScoreSet {
    int32_t match, mismatch;      // scaled by 64
    int32_t io, ie, eo, ee;       // internal/end open/extend, scaled by 64
    int32_t tagM, tagX, tagY;     // replicated priority tags (1..3)*21
    int32_t pM, pX, pY;           // plain priorities (1..3)
};

// constants of the packed 16-bit fast path (gotoh_pair16.cuh), transformed score space, x16
struct Fast16 {
    int32_t D16;                       // (match - mismatch) * 16: penalty of a mismatching column (a match costs 0)
    uint32_t negD;                     // 2^32 - D16: (1 per differing half) * negD + H is H - D16 in those halves, one IMAD
    int32_t PoX, PeX, PeoX, PeeX;      // vertical   (Ix) penalties 16*(match - g): internal open/extend, end open/extend
    int32_t PoY, PeY, PeoY, PeeY;      // horizontal (Iy) penalties -16*g
    int32_t beta;                      // row potential of the transformed space (= match), unscaled
    int32_t neg;                       // "minus infinity" of the dead slots (multiple of 16, above every penalty)
    int32_t bias;                      // value representing transformed score 0 (multiple of 16), placed by the host so that
                                       // every reachable value of the padded DP fits the unsigned 16-bit window
    uint32_t class_lut[2];             // nibble k = base_class of the symbol with code k (traceback counts work on codes)
    uint32_t ascii[4];                 // byte k = the symbol with code k (gapped strings are written from codes)
    int32_t has_gap_symbol;            // '-' occurs inside the loaded sequences (input that was not normalized)
    int32_t dead_extra;                // largest difference of the x lengths of the two pairs of a unit that `neg` is budgeted for
};

// Symbol codes are 0..14 (the fifteen IUPAC nucleotide symbols fit; more distinct symbols send the
// job to the general kernel), 15 = pad.
constexpr int CODE_SYMBOLS = 15;
constexpr uint32_t CODE_PADSYM = 15u;

__device__ __forceinline__ int code_class(const Fast16& f, int code)      // base_class of a symbol code
{
    return (int)((f.class_lut[code >> 3] >> (4 * (code & 7))) & 7u);
}

__device__ __forceinline__ uint8_t code_ascii(const Fast16& f, int code)  // the symbol itself
{
    return (uint8_t)(f.ascii[code >> 2] >> (8 * (code & 3)));
}

// Symbol codes are stored per sequence with CODE_LEAD pad codes (15 = "matches no symbol") in front
// and one behind, so the DP's one-step-ahead fetch of the next column symbol needs no lower bound
// check at all (the wavefront skew is at most 31 columns) and only a clamp at the upper end.
constexpr int CODE_LEAD = 32;
constexpr int CODE_PAD = CODE_LEAD + 1;
__host__ __device__ __forceinline__ long long code_offset(long long byte_offset, long long index)
{
    return byte_offset + index * CODE_PAD + CODE_LEAD;
}

struct PairIndex { int xi, yi; long long out; };

struct AlignArgs {
    const uint8_t* xb; const int64_t* xoff;   // row set (x): bytes + offsets
    const uint8_t* yb; const int64_t* yoff;   // column set (y)
    const uint8_t* xc; const uint8_t* yc;     // 3-bit symbol codes at code_offset(offset, index), fast path only
    Fast16 f16;
    const int32_t* px; const int32_t* py;     // explicit pair list, or nullptr for rect mode
    int32_t x0, y0, ny;                       // rect mode: pair p = (x0 + p / ny, y0 + p % ny)
    const int32_t* xrows;                     // rect mode, optional: row r of the launch is sequence xrows[r] (a subset of the
                                              // rectangle's rows); results still land at ((x - x0) * ny + y)
    const int32_t* ycols;                     // rect mode, optional: column c of the launch is sequence ycols[c] (a permutation
                                              // of the rectangle's columns, longest first, so that the two pairs of a warp and
                                              // consecutive work units have similar lengths)
    long long npairs;
    // packed kernel (two pairs per warp unit).  The two pairs of a unit must have (nearly) the same
    // x length: the bottom-aligned variants keep the shorter pair's extra slots dead, and the dead
    // band is budgeted for at most Fast16::dead_extra of them.  Rect mode: a unit is two consecutive
    // columns of ONE row (the last unit of an odd-width row holds a single pair), so the x lengths
    // are equal by construction.  Pair-list mode: the host pairs the list by length (unit_pairs).
    long long nunits;
    const int32_t* unit_pairs;                // pair-list mode: unit u aligns pairs unit_pairs[2u], unit_pairs[2u+1] (equal = one pair)
    ScoreSet sc;
    int32_t* score;                           // [npairs] or nullptr
    int32_t* counts;                          // [npairs][4] or nullptr
    double* metrics;                          // [npairs][4] or nullptr
    uint8_t* aln_x; uint8_t* aln_y;           // gapped strings (right-aligned in slots) or nullptr
    const int64_t* aln_off; int64_t* aln_start;
    // "both orientations" launches (SYM kernels): results of (y, x) are the results of (x, y) unless a
    // tie between Ix and Iy was decided on the traced path; then the pair is appended to `redo`
    int32_t nx;                               // rows of the rectangle (transposed index = col * nx + row)
    int32_t* t_score; int32_t* t_counts; double* t_metrics;   // [ny][nx] outputs of the mirrored pairs, or nullptr
    long long* redo; unsigned long long* redo_count;          // out indices of orientation-sensitive pairs
    uint8_t* trace; long long trace_per_warp; // traceback arena
    int32_t* bnd; long long bnd_per_warp;     // stripe-boundary rows (2 ints per column)
    int32_t coop_stripes;                     // intra-task kernel: boundary buffers per CTA (the most stripes any pair of the launch has)
    unsigned long long* counter;              // dynamic work counter
    int* status;                              // sticky error flag
};

__device__ __forceinline__ PairIndex pair_index(const AlignArgs& a, long long p)
{
    PairIndex r;
    if (a.px) { r.xi = a.px[p]; r.yi = a.py[p]; r.out = p; return r; }
    const int row = (int)(p / a.ny), col = (int)(p % a.ny);
    r.xi = a.xrows ? a.xrows[row] : a.x0 + row;
    r.yi = a.ycols ? a.ycols[col] : a.y0 + col;
    r.out = (long long)(r.xi - a.x0) * a.ny + (r.yi - a.y0);
    return r;
}

// 0..3 = A,G,C,T (bit1 = pyrimidine: a transition flips only bit0); 4 = '-'; 5 = missing
__host__ __device__ __forceinline__ int base_class(int c)
{
    const int u = c & 0xDF;  // fold ASCII case
    int k = 5;
    k = (u == 'A') ? 0 : k;
    k = (u == 'G') ? 1 : k;
    k = (u == 'C') ? 2 : k;
    k = (u == 'T') ? 3 : k;
    k = (c == '-') ? 4 : k;
    return k;
}

// ---- lean fp64 arithmetic for the metric epilogue ----------------------------------------------
// The metrics cost more instructions than the counting itself when written with the compiler's
// generic `/`, log() and sqrt() (special-case paths, 64-bit immediates materialised per use).  The
// operands here are tame -- ratios of small non-negative integers, logarithm arguments in (0, 1] --
// so the epilogue uses straight-line versions: a Newton reciprocal from MUFU.RCP64H shared by all
// divisions with the same divisor, each quotient finished with the remainder correction
// (q + (a - b*q) * r), which makes it the correctly rounded IEEE quotient (so p, p-gaps and every
// intermediate ratio are bit-identical to `a / b`), and the fdlibm log kernel on the reduced
// mantissa.  tools/metrics_emul.c replays the same operation sequence on the CPU against libm:
// p / p-gaps bit-exact, jc / k2p within 3e-16 relative over 15 M count tuples.
__constant__ double kLogCoef[9] = {
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
    6.93147180369123816490e-01 /* ln2_hi */, 1.90821492927058770002e-10 /* ln2_lo */};

__device__ __forceinline__ double rcp_full(double b)   // 1 / b to full precision (b > 0, normal)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    return fma(r, e, r);
}

__device__ __forceinline__ double div_by(double a, double b, double r)   // a / b correctly rounded, r = rcp_full(b)
{
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

__device__ __forceinline__ double sqrt_pos(double a)   // sqrt(a), a > 0 normal
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double h = 0.5 * y, g = a * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    return fma(fma(-g, g, a), h, g);
}

__device__ __forceinline__ double log_pos(double x)   // log(x), x > 0 normal (fdlibm e_log.c kernel)
{
    int hx = __double2hiint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;            // mantissa above sqrt(2): use half of it
    k += i >> 20;
    const double f = __hiloint2double(hx | (i ^ 0x3ff00000), __double2loint(x)) - 1.0;
    const double t = 2.0 + f;
    const double s = div_by(f, t, rcp_full(t));
    const double dk = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;   // (double)k without I2F: 2^52 + 2^31 + k, minus both
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, kLogCoef[5], kLogCoef[3]), kLogCoef[1]);
    const double t2 = z * fma(w, fma(w, fma(w, kLogCoef[6], kLogCoef[4]), kLogCoef[2]), kLogCoef[0]);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return fma(dk, kLogCoef[7], -((hfsq - fma(s, hfsq + R, dk * kLogCoef[8])) - f));
}

// exact int -> double for 0 <= k < 2^31 without the XU-pipe I2F (which shares a quarter-rate pipe
// with POPC and MUFU): 2^52 + k assembled as bits, minus 2^52
__device__ __forceinline__ double count_to_double(int k)
{
    return __hiloint2double(0x43300000, k) - 4503599627370496.0;
}

// distances.py:319-348 formulas in fp64; NaN where the reference yields None.
__device__ __forceinline__ void metrics_from_counts(int same, int ts, int tv, int gap, double out[4])
{
    const double n = count_to_double(same + ts + tv);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (!(n > 0.0)) { out[0] = out[1] = out[2] = out[3] = nan; return; }
    const double d = count_to_double(ts + tv), g = count_to_double(gap);
    const double rn = rcp_full(n);
    const double p = div_by(d, n, rn);
    out[0] = p;
    out[1] = div_by(d + g, n + g, rcp_full(n + g));
    const double P = div_by(count_to_double(ts), n, rn), Q = div_by(count_to_double(tv), n, rn);
    const double u = 1.0 - div_by(4.0 * p, 3.0, 1.0 / 3.0);           // jc = -3/4 ln(1 - 4p/3)
    const double b = 1.0 - 2.0 * Q, a = 1.0 - 2.0 * P - Q;             // k2p = -1/2 ln((1 - 2P - Q) sqrt(1 - 2Q))
    const double v = (a > 0.0 && b > 0.0) ? a * sqrt_pos(b) : -1.0;
    // a non-positive argument is -inf / NaN in the reference (None); "+ 0.0" folds -0.0 into +0.0
    out[2] = u > 0.0 ? -0.75 * log_pos(u) + 0.0 : nan;
    out[3] = v > 0.0 ? -0.5 * log_pos(v) + 0.0 : nan;
}

// Table form of the same four metrics for the alignment-free kernels (rows of at most LN_TABLE_COLS
// columns).  The arguments of both logarithms are ratios of small integers,
//     1 - 4p/3 = (3n - 4d) / 3n,   (1 - 2P - Q) sqrt(1 - 2Q) = ((n - 2ts - tv) / n) ((n - 2tv) / n)^(1/2),
// so with ln k tabulated as a 64-bit fixed-point number (58 fractional bits, rounded from long double
// on the host) each logarithm is an exact integer difference of table entries, converted once.  That
// is ~60 instructions per pair instead of ~190 and more accurate than the floating-point formula;
// against the reference's own formula (whose 1 - x loses about 1e-16 / x) it stays within 1.4e-13
// relative for n <= 2048 (tools/metrics_table_check.c, 4.2 M count tuples), inside the 1e-12 of
// north_star.  Longer rows keep the operation-for-operation form above, whose deviation does not
// grow with n.  The one case the integers cannot decide (n = 2 ts + tv, where 1 - 2P - Q is a rounding
// residue in the reference) goes through the floating-point form.
// p and p-gaps are the same correctly rounded quotients in both forms.
constexpr int LN_TABLE_COLS = 2048;
constexpr int LN_TABLE_SIZE = 3 * LN_TABLE_COLS + 1;     // entries 0 .. 3 n_max (entry 0 unused)

template <class Lookup>
__device__ __forceinline__ void metrics_from_counts_table(int same, int ts, int tv, int gap, double out[4], Lookup ln58)
{
    const int n = same + ts + tv, d = ts + tv;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (n <= 0) { out[0] = out[1] = out[2] = out[3] = nan; return; }
    const int A = 3 * n - 4 * d, a = n - 2 * ts - tv, b = n - 2 * tv;
    const double nd = count_to_double(n), dd = count_to_double(d), g = count_to_double(gap);
    const double rn = rcp_full(nd);
    out[0] = div_by(dd, nd, rn);
    out[1] = div_by(dd + g, nd + g, rcp_full(nd + g));
    const long long tn = ln58(n);
    // -3/4 and -1/4 times 2^-58 are exact constants: one rounding per metric after the conversion.
    // 3n = 4d makes d / n = 0.75 and 4p/3 = 1 exactly, n = 2 tv makes Q = 0.5 and the square root 0 exactly:
    // the reference's logarithm is -inf (None) there, like for negative arguments.
    out[2] = A > 0 ? __ll2double_rn(ln58(A) - ln58(3 * n)) * (-0.75 * 0x1p-58) + 0.0 : nan;
    if (a > 0 && b > 0) out[3] = __ll2double_rn(2 * ln58(a) + ln58(b) - 3 * tn) * (-0.25 * 0x1p-58) + 0.0;
    else if (a == 0 && b > 0) {
        // n = 2 ts + tv: 1 - 2P - Q is a rounding residue in the reference (zero, negative or ~1e-17), and its
        // logarithm a finite number or None accordingly: the floating-point form decides, operation for operation
        const double P = div_by(count_to_double(ts), nd, rn), Q = div_by(count_to_double(tv), nd, rn);
        const double bf = 1.0 - 2.0 * Q, af = 1.0 - 2.0 * P - Q;
        const double v = (af > 0.0 && bf > 0.0) ? af * sqrt_pos(bf) : -1.0;
        out[3] = v > 0.0 ? -0.5 * log_pos(v) + 0.0 : nan;
    }
    else out[3] = nan;
}

}  // namespace taxi
