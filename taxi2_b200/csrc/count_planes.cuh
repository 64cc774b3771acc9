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
// The planes are stored plane-major with the sequence index fastest ([4][W][n]), so a warp whose
// threads hold consecutive sequences reads every plane word with one coalesced transaction.
#pragma once
#include "common.cuh"

namespace taxi {

struct Planes {
    const uint32_t* w;   // [4][W][nseq] words (plane-major, sequence fastest): b0, b1, R, G
    int32_t W;           // words per plane
    int32_t nseq;
    __device__ __forceinline__ uint32_t at(int plane, int word, int seq) const
    {
        return __ldg(w + ((size_t)plane * W + word) * nseq + seq);
    }
};

// one thread per (sequence, word): bytes -> planes.  Consecutive threads take consecutive
// sequences so the plane words are written coalesced; the byte reads of a warp are scattered, but
// this kernel runs once per load (O(N*L)) and is not on the O(N^2) path.
__global__ void pack_planes_kernel(const uint8_t* __restrict__ bytes, const int64_t* __restrict__ off,
                                   int32_t nseq, int32_t W, uint32_t* __restrict__ out)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)nseq * W;
    if (gid >= total) return;
    const int w = (int)(gid / nseq), seq = (int)(gid % nseq);
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
    const size_t plane = (size_t)W * nseq;
    uint32_t* dst = out + (size_t)w * nseq + seq;
    dst[0] = b0; dst[plane] = b1; dst[2 * plane] = R; dst[3 * plane] = G;
}

struct CountArgs {
    Planes x, y;
    const int32_t* px; const int32_t* py;   // explicit list or nullptr (rect)
    int32_t x0, y0, nx, ny;
    int32_t slab;                           // rect mode: x rows staged in shared memory per block (<= COUNT_SLAB)
    long long npairs;
    int32_t* counts;   // [npairs][4] or nullptr
    double* metrics;   // [npairs][4] or nullptr
};

// running state of one pair's scan (columns in order): the trim-to-first/last-both-real rule
struct CountState {
    int same, ts, tv, gapc, pend;
    bool seen;
    __device__ __forceinline__ void init() { same = ts = tv = gapc = pend = 0; seen = false; }
    __device__ __forceinline__ void word(uint32_t x0, uint32_t x1, uint32_t xr, uint32_t xg,
                                         uint32_t y0, uint32_t y1, uint32_t yr, uint32_t yg)
    {
        const uint32_t both = xr & yr;
        const uint32_t gapw = (xg & yr) | (xr & yg);
        if (both) {
            const uint32_t d0 = x0 ^ y0, d1 = x1 ^ y1;
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
    __device__ __forceinline__ void store(const CountArgs& a, long long p) const
    {
        if (a.counts) *reinterpret_cast<int4*>(a.counts + 4 * p) = make_int4(same, ts, tv, gapc);
        if (a.metrics) {
            double m[4];
            metrics_from_counts(same, ts, tv, gapc, m);
            double2* dst = reinterpret_cast<double2*>(a.metrics + 4 * p);
            dst[0] = make_double2(m[0], m[1]);
            dst[1] = make_double2(m[2], m[3]);
        }
    }
};

// Rectangle mode.  A block owns COUNT_TY consecutive y columns (one per thread) and walks the x
// rows of its slab in groups of COUNT_RX: the y words of a thread are loaded once per group
// (coalesced across the warp, plane-major layout) and reused for COUNT_RX pairs whose x words come
// from shared memory as warp broadcasts.  Results are written coalesced, 16 B + 32 B per pair.
constexpr int COUNT_TY = 256;
constexpr int COUNT_RX = 8;
constexpr int COUNT_SLAB = 64;   // x rows staged in shared memory per block (fewer when the sequences are long)

__global__ void __launch_bounds__(COUNT_TY, 2) count_rect_kernel(const CountArgs a)
{
    extern __shared__ uint32_t xs[];   // [slab][4][W]
    const int W = min(a.x.W, a.y.W);
    const int yj = blockIdx.x * COUNT_TY + threadIdx.x;          // column inside the rectangle
    const int xbase = blockIdx.y * a.slab;                       // first row of the slab
    const int rows = min(a.slab, a.nx - xbase);
    for (int k = threadIdx.x; k < rows * 4 * W; k += COUNT_TY) {
        const int r = k / (4 * W), pw = k % (4 * W);
        xs[k] = a.x.at(pw / W, pw % W, a.x0 + xbase + r);
    }
    __syncthreads();
    if (yj >= a.ny) return;
    const int ys = a.y0 + yj;
    for (int g = 0; g < rows; g += COUNT_RX) {
        CountState st[COUNT_RX];
#pragma unroll
        for (int k = 0; k < COUNT_RX; ++k) st[k].init();
        for (int w = 0; w < W; ++w) {
            const uint32_t y0 = a.y.at(0, w, ys), y1 = a.y.at(1, w, ys), yr = a.y.at(2, w, ys), yg = a.y.at(3, w, ys);
#pragma unroll
            for (int k = 0; k < COUNT_RX; ++k) {
                if (g + k < rows) {
                    const uint32_t* xw = xs + (size_t)(g + k) * 4 * W + w;
                    st[k].word(xw[0], xw[W], xw[2 * W], xw[3 * W], y0, y1, yr, yg);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < COUNT_RX; ++k)
            if (g + k < rows) st[k].store(a, (long long)(xbase + g + k) * a.ny + yj);
    }
}

// Explicit pair list: one thread per pair.
__global__ void __launch_bounds__(256) count_pairs_kernel(const CountArgs a)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.npairs) return;
    const int xi = a.px[p], yi = a.py[p];
    const int W = min(a.x.W, a.y.W);
    CountState st;
    st.init();
    for (int w = 0; w < W; ++w)
        st.word(a.x.at(0, w, xi), a.x.at(1, w, xi), a.x.at(2, w, xi), a.x.at(3, w, xi),
                a.y.at(0, w, yi), a.y.at(1, w, yi), a.y.at(2, w, yi), a.y.at(3, w, yi));
    st.store(a, p);
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

// Same reduction, returning the whole winning pair: index (offset by col0, -1 = no defined
// distance in the row), its four metrics and its four counts -- what VersusReference needs per
// query (versus_reference.py:119-129,184-188), so the nx x ny matrix never leaves the device.
__global__ void best_rows_kernel(const double* __restrict__ metrics, const int32_t* __restrict__ counts, int32_t nx, int32_t ny,
                                 int32_t metric, int32_t col0, int32_t* __restrict__ out_idx, double* __restrict__ out_metrics,
                                 int32_t* __restrict__ out_counts)
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
        const double ov = __shfl_xor_sync(TAXI_FULL_MASK, best, o);
        const int oi = __shfl_xor_sync(TAXI_FULL_MASK, bidx, o);
        if (oi != 0x7fffffff && (bidx == 0x7fffffff || ov < best || (ov == best && oi < bidx))) { best = ov; bidx = oi; }
    }
    const bool none = bidx == 0x7fffffff;
    if (lane == 0) out_idx[row] = none ? -1 : col0 + bidx;
    if (lane < 4) {
        const size_t at = ((size_t)row * ny + (none ? 0 : bidx)) * 4 + lane;
        out_metrics[(size_t)row * 4 + lane] = none ? __longlong_as_double(0x7ff8000000000000LL) : metrics[at];
        if (counts && out_counts) out_counts[(size_t)row * 4 + lane] = none ? 0 : counts[at];
    }
}

}  // namespace taxi
