// Alignment-free mode (params.pairs.align = False): per-pair same / transition / transversion /
// gap-column counts straight from pre-aligned sequences, bit-sliced.
//
// Replaces calc.seq_distances_{p,p_gaps,jukes_cantor,kimura2p} applied to the raw strings
// (reference: src/itaxotools/taxi2/distances.py:319-348, called from versus_all.py:546-552 when
// versus_all.py:522-530 skipped normalisation and alignment).
//
// The packer turns every sequence into four bit planes of 32 columns per word:
//   b0, b1 : 2-bit nucleotide (A=00 G=01 C=10 T=11; a transition flips only b0)
//   R      : "real" mask (A/C/G/T)
//   G      : gap mask ('-')
// Columns past a sequence's end have R = G = 0, which is exactly "truncate to the shorter".
#pragma once
#include "common.cuh"

namespace taxi {

struct Planes {
    const uint32_t* w;   // [nseq][4][W] words: b0, b1, R, G
    int32_t W;           // words per plane
    int32_t nseq;
};

// one thread per sequence: bytes -> planes.  Each thread walks its own sequence; the byte loads
// of a warp are scattered, but this kernel runs once per load (O(N*L)) and is not on the O(N^2) path.
__global__ void pack_planes_kernel(const uint8_t* __restrict__ bytes, const int64_t* __restrict__ off,
                                   int32_t nseq, int32_t W, uint32_t* __restrict__ out)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)nseq * W;
    if (gid >= total) return;
    const int seq = (int)(gid / W), w = (int)(gid % W);
    const int64_t o = off[seq];
    const int len = (int)(off[seq + 1] - o);
    uint32_t b0 = 0, b1 = 0, R = 0, G = 0;
    const int base = w * 32;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const int pos = base + k;
        if (pos < len) {
            const int c = base_class((int)bytes[o + pos]);
            if (c < 4) { R |= 1u << k; b0 |= (uint32_t)(c & 1) << k; b1 |= (uint32_t)(c >> 1) << k; }
            else if (c == 4) G |= 1u << k;
        }
    }
    uint32_t* dst = out + (size_t)seq * 4 * W;
    dst[w] = b0; dst[W + w] = b1; dst[2 * W + w] = R; dst[3 * W + w] = G;
}

struct CountArgs {
    Planes x, y;
    const int32_t* px; const int32_t* py;   // explicit list or nullptr (rect)
    int32_t x0, y0, ny;
    long long npairs;
    int32_t* counts;   // [npairs][4] or nullptr
    double* metrics;   // [npairs][4] or nullptr
};

// One thread per pair.  In rect mode consecutive threads take consecutive y for the same x, so
// the x planes are a warp-broadcast load and the results are written fully coalesced (16 B and
// 32 B per pair); the y planes of a tile stay L1/L2 resident (the whole plane set is a few MB).
__global__ void __launch_bounds__(256) count_planes_kernel(const CountArgs a)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npairs) return;
    int xi, yi;
    if (a.px) { xi = a.px[p]; yi = a.py[p]; }
    else { xi = a.x0 + (int)(p / a.ny); yi = a.y0 + (int)(p % a.ny); }
    const int W = min(a.x.W, a.y.W);
    const uint32_t* __restrict__ X = a.x.w + (size_t)xi * 4 * a.x.W;
    const uint32_t* __restrict__ Y = a.y.w + (size_t)yi * 4 * a.y.W;
    const int WX = a.x.W, WY = a.y.W;
    int same = 0, ts = 0, tv = 0, gapc = 0, pend = 0;
    bool seen = false;
    for (int w = 0; w < W; ++w) {
        const uint32_t xr = __ldg(X + 2 * WX + w), yr = __ldg(Y + 2 * WY + w);
        const uint32_t both = xr & yr;
        const uint32_t gapw = (__ldg(X + 3 * WX + w) & yr) | (xr & __ldg(Y + 3 * WY + w));
        if (both) {
            const uint32_t d0 = __ldg(X + w) ^ __ldg(Y + w);
            const uint32_t d1 = __ldg(X + WX + w) ^ __ldg(Y + WY + w);
            tv += __popc(both & d1);
            ts += __popc(both & d0 & ~d1);
            same += __popc(both & ~(d0 | d1));
            const int lo = __ffs(both) - 1, hi = 31 - __clz(both);
            const uint32_t upto_hi = (hi == 31) ? 0xffffffffu : ((2u << hi) - 1u);
            const uint32_t from_lo = seen ? 0xffffffffu : (0xffffffffu << lo);
            if (seen) gapc += pend;
            gapc += __popc(gapw & upto_hi & from_lo);
            pend = __popc(gapw & ~upto_hi);
            seen = true;
        } else {
            pend += __popc(gapw);
        }
    }
    if (a.counts) *reinterpret_cast<int4*>(a.counts + 4 * p) = make_int4(same, ts, tv, gapc);
    if (a.metrics) {
        double m[4];
        metrics_from_counts(same, ts, tv, gapc, m);
        double2* dst = reinterpret_cast<double2*>(a.metrics + 4 * p);
        dst[0] = make_double2(m[0], m[1]);
        dst[1] = make_double2(m[2], m[3]);
    }
}

// versus_reference.py:184-188 / decontaminate.py:258-264: first minimum per query row, NaN skipped.
// One warp per row; ties resolve to the smallest column index.
__global__ void argmin_rows_kernel(const double* __restrict__ metrics, int32_t nx, int32_t ny, int32_t metric,
                                   int32_t* __restrict__ out_idx, double* __restrict__ out_val)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= nx) return;
    double best = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    int bidx = 0x7fffffff;
    for (int c = lane; c < ny; c += 32) {
        const double v = metrics[((size_t)row * ny + c) * 4 + metric];
        if (v == v && (v < best || (v == best && c < bidx))) { best = v; bidx = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(TAXI_FULL_MASK, best, o);
        const int oi = __shfl_down_sync(TAXI_FULL_MASK, bidx, o);
        if (oi != 0x7fffffff && (bidx == 0x7fffffff || ov < best || (ov == best && oi < bidx))) { best = ov; bidx = oi; }
    }
    if (lane == 0) {
        out_idx[row] = (bidx == 0x7fffffff) ? -1 : bidx;
        out_val[row] = (bidx == 0x7fffffff) ? __longlong_as_double(0x7ff8000000000000LL) : best;
    }
}

}  // namespace taxi
