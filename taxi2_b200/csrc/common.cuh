// Shared declarations of the taxi2_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TAXI_FULL_MASK 0xffffffffu

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
    int32_t D16;                       // (match - mismatch) * 16 <= 127: penalty of a mismatching column (a match costs 0)
    uint32_t tlo, thi;                 // PRMT table indexed by the XOR of two symbol codes: byte 0 = 0, bytes 1..7 = D16
    int32_t PoX, PeX, PeoX, PeeX;      // vertical   (Ix) penalties 16*(match - g): internal open/extend, end open/extend
    int32_t PoY, PeY, PeoY, PeeY;      // horizontal (Iy) penalties -16*g
    int32_t beta;                      // row potential of the transformed space (= match), unscaled
    int32_t neg;                       // "minus infinity" of the dead slots (multiple of 16, above every penalty)
    int32_t bias;                      // value representing transformed score 0 (multiple of 16), placed by the host so that
                                       // every reachable value of the padded DP fits the unsigned 16-bit window
    uint32_t class_lut;                // nibble k = base_class of the symbol with code k (traceback counts work on codes)
    uint32_t ascii_lo, ascii_hi;       // byte k = the symbol with code k (gapped strings are written from codes)
    int32_t has_gap_symbol;            // '-' occurs inside the loaded sequences (input that was not normalized)
    int32_t dead_extra;                // largest difference of the x lengths of the two pairs of a unit that `neg` is budgeted for
};

// Symbol codes are stored per sequence with CODE_LEAD pad codes (7 = "matches nothing") in front
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
    uint8_t* trace; long long trace_per_warp; // traceback arena
    int32_t* bnd; long long bnd_per_warp;     // stripe-boundary rows (2 ints per column)
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

// distances.py:319-348 formulas in fp64; NaN where the reference yields None.
__device__ __forceinline__ void metrics_from_counts(int same, int ts, int tv, int gap, double out[4])
{
    const double n = (double)same + (double)ts + (double)tv;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (!(n > 0.0)) { out[0] = out[1] = out[2] = out[3] = nan; return; }
    const double d = (double)ts + (double)tv;
    const double p = d / n;
    out[0] = p;
    out[1] = (d + (double)gap) / (n + (double)gap);
    const double P = (double)ts / n, Q = (double)tv / n;
    const double jc = -0.75 * log(1.0 - 4.0 * p / 3.0);
    const double k2 = -0.5 * log((1.0 - 2.0 * P - Q) * sqrt(1.0 - 2.0 * Q));
    out[2] = isfinite(jc) ? jc + 0.0 : nan;
    out[3] = isfinite(k2) ? k2 + 0.0 : nan;
}

}  // namespace taxi
