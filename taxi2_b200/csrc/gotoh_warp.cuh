// Global affine-gap alignment, one pair per warp (inter-task), exact Biopython first-path
// semantics, with the distance counts fused into the traceback epilogue.
//
// Replaces Bio.Align.PairwiseAligner.align(x, y)[0] + _format_pretty + the four
// calc.seq_distances_* scans (reference: src/itaxotools/taxi2/align.py:151-157,
// distances.py:319-348).
//
// Layout of one pair's DP:  rows = x (1..nA), columns = y (1..nB).  The rows are cut into
// stripes of 32*H rows; inside a stripe lane l owns rows [l*H, l*H+H) in registers and sweeps
// the columns one per step, one column behind lane l-1 (a wavefront skewed by lane).  The
// bottom row of each lane moves to the next lane with two warp shuffles per step; the bottom
// row of a stripe goes through a small per-warp boundary buffer.
//
// Tagged maxima: every value is score*64 + tag, tag = priority of the predecessor state
// (replicated in three 2-bit fields).  max3 over tagged candidates therefore returns the best
// score AND, on ties, the predecessor Biopython's path generator visits first (M, Ix, Iy for
// Gotoh; Iy, Ix, M for Needleman-Wunsch).  The three pointers of a cell are the low bits of
// the three results, so the 6-bit trace code costs two LOP3s per cell.
#pragma once
#include "common.cuh"

namespace taxi {

__device__ __forceinline__ int bitselect(int a, int b, int mask)
{
    // (a & mask) | (b & ~mask) in one LOP3
    int r;
    // LUT over inputs (mask, a, b): (0xF0 & 0xCC) | (~0xF0 & 0xAA) = 0xCA
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(mask), "r"(a), "r"(b));
    return r;
}

__device__ __forceinline__ int and_or(int a, int mask, int c)
{
    // (a & mask) | c ; LUT over inputs (a, mask, c): (0xF0 & 0xCC) | 0xAA = 0xEA
    int r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(c));
    return r;
}

template <int H> struct TraceGeom {
    static constexpr int HB = (H + 7) / 8 * 8;   // trace bytes per (lane, step), 8-byte granules
    static constexpr int SL = 32 * H;            // rows per stripe
};

// priority tag (1..3) -> state id: 0 = M (diagonal), 1 = X (vertical, gap in y), 2 = Y (horizontal)
__device__ __forceinline__ int state_of_tag(int tag, const ScoreSet& sc)
{
    return tag == sc.pM ? 0 : (tag == sc.pX ? 1 : 2);
}

template <int H>
__device__ __forceinline__ void align_one(const AlignArgs& a, long long p, int lane, uint8_t* trace, int32_t* bnd)
{
    using G = TraceGeom<H>;
    constexpr int HB = G::HB;
    constexpr int SL = G::SL;
    const ScoreSet& sc = a.sc;

    int xi, yi;
    if (a.px) { xi = a.px[p]; yi = a.py[p]; }
    else { xi = a.x0 + (int)(p / a.ny); yi = a.y0 + (int)(p % a.ny); }
    const int64_t xo = a.xoff[xi], yo = a.yoff[yi];
    const uint8_t* __restrict__ x = a.xb + xo;
    const uint8_t* __restrict__ y = a.yb + yo;
    const int nA = (int)(a.xoff[xi + 1] - xo);
    const int nB = (int)(a.yoff[yi + 1] - yo);

    if (nA <= 0 || nB <= 0) {  // Biopython raises ValueError; flag it and emit "undefined"
        if (lane == 0) {
            atomicExch(a.status, -2);
            if (a.score) a.score[p] = 0;
            if (a.counts) { a.counts[4 * p] = a.counts[4 * p + 1] = a.counts[4 * p + 2] = a.counts[4 * p + 3] = 0; }
            if (a.metrics) { double m[4]; metrics_from_counts(0, 0, 0, 0, m); for (int k = 0; k < 4; ++k) a.metrics[4 * p + k] = m[k]; }
            if (a.aln_start) a.aln_start[p] = a.aln_off[p + 1];
        }
        return;
    }

    const int nstripes = (nA + SL - 1) / SL;
    const int step_stride = nB + 31;
    int Hl[H];   // H(i, j-1) of my rows: best of the three states, tagged by the winning state

    for (int s = 0; s < nstripes; ++s) {
        const int itop = s * SL + lane * H + 1;  // DP row held in register slot 0
        const int rows_here = min(SL, nA - s * SL);
        const int nlive = (rows_here + H - 1) / H;
        const bool live = lane < nlive;
        const int nsteps = nB + nlive - 1;

        int ach[H], Yn[H], cYo[H], cYe[H];
#pragma unroll
        for (int r = 0; r < H; ++r) {
            const int i = itop + r;
            ach[r] = (i <= nA) ? (int)__ldg(x + (i <= nA ? i - 1 : 0)) : 0x100;
            const int xc0 = sc.eo + (i - 1) * sc.ee;  // Ix(i, 0): leading end gap of i rows
            Hl[r] = xc0 | sc.tagX;
            cYo[r] = (i == nA) ? sc.eo : sc.io;       // horizontal gaps on the last row are end gaps
            cYe[r] = (i == nA) ? sc.ee : sc.ie;
            Yn[r] = (xc0 + cYo[r]) | sc.tagX;         // Iy(i, 1) opens from Ix(i, 0)
        }
        // H(itop-1, 0): diagonal input of my first row at column 1
        int Hd_saved = (itop == 1) ? sc.tagM : ((sc.eo + (itop - 2) * sc.ee) | sc.tagX);
        int outX = 0, outH = 0;
        uint8_t* tbase = trace + ((size_t)s * step_stride * 32 + lane) * HB;

        for (int t = 0; t < nsteps; ++t) {
            const int j = t - lane + 1;
            int rX = __shfl_up_sync(TAXI_FULL_MASK, outX, 1);
            int rH = __shfl_up_sync(TAXI_FULL_MASK, outH, 1);
            const bool active = live && j >= 1 && j <= nB;
            if (lane == 0 && active) {
                if (s == 0) {
                    // row 0 of the matrix: only Iy is alive there (leading end gap of j columns)
                    const int y0j = sc.eo + (j - 1) * sc.ee;
                    rH = y0j | sc.tagY;
                    rX = (y0j + (j == nB ? sc.eo : sc.io)) | sc.tagY;
                } else {
                    rX = __ldcg(bnd + 2 * j);
                    rH = __ldcg(bnd + 2 * j + 1);
                }
            }
            if (active) {
                const int b = (int)__ldg(y + j - 1);
                const int cXo = (j == nB) ? sc.eo : sc.io;  // vertical gaps in the last column are end gaps
                const int cXe = (j == nB) ? sc.ee : sc.ie;
                int Hd = Hd_saved;
                int Xin = rX;
                uint32_t tw[HB / 4];
#pragma unroll
                for (int k = 0; k < HB / 4; ++k) tw[k] = 0;
                int tq[4];
#pragma unroll
                for (int r = 0; r < H; ++r) {
                    const int sub = (ach[r] == b) ? sc.match : sc.mismatch;
                    const int Mr = Hd + sub;
                    const int Yin = Yn[r];
                    // trace code: bits 0-1 pred of M, 2-3 pred of Ix, 4-5 pred of Iy (+ garbage above)
                    tq[r & 3] = bitselect(bitselect(Mr, Xin, 3), Yin, 15);
                    const int Mt = and_or(Mr, ~TAG_MASK, sc.tagM);
                    const int Xt = and_or(Xin, ~TAG_MASK, sc.tagX);
                    const int Yt = and_or(Yin, ~TAG_MASK, sc.tagY);
                    Hd = Hl[r];
                    Hl[r] = __vimax3_s32(Mt, Xt, Yt);
                    Xin = __viaddmax_s32(max(Mt, Yt), cXo, Xt + cXe);          // Ix(i+1, j)
                    Yn[r] = __viaddmax_s32(max(Mt, Xt), cYo[r], Yt + cYe[r]);  // Iy(i, j+1)
                    if ((r & 3) == 3 || r == H - 1) {
                        const int n = (r & 3) + 1;
                        uint32_t lo = __byte_perm(tq[0], n > 1 ? tq[1] : 0, 0x0040);
                        uint32_t hi = __byte_perm(n > 2 ? tq[2] : 0, n > 3 ? tq[3] : 0, 0x0040);
                        tw[r >> 2] = __byte_perm(lo, hi, 0x5410);
                    }
                }
                outX = Xin;
                outH = Hl[H - 1];
                Hd_saved = rH;
                uint8_t* dst = tbase + (size_t)t * 32 * HB;
                if constexpr (HB % 16 == 0) {
#pragma unroll
                    for (int k = 0; k < HB / 16; ++k)
                        __stcg(reinterpret_cast<uint4*>(dst) + k, make_uint4(tw[4 * k], tw[4 * k + 1], tw[4 * k + 2], tw[4 * k + 3]));
                } else {
#pragma unroll
                    for (int k = 0; k < HB / 8; ++k)
                        __stcg(reinterpret_cast<uint2*>(dst) + k, make_uint2(tw[2 * k], tw[2 * k + 1]));
                }
                if (lane == 31 && s + 1 < nstripes) {
                    __stcg(bnd + 2 * j, outX);
                    __stcg(bnd + 2 * j + 1, outH);
                }
            }
        }
        __syncwarp();
    }

    // ---- end state and score: H(nA, nB) sits in the register slot of row nA ------------------
    const int l_last = ((nA - 1) % SL) / H;
    const int r_last = (nA - 1) % H;
    int fin = 0;
#pragma unroll
    for (int r = 0; r < H; ++r) fin = (r == r_last) ? Hl[r] : fin;
    fin = __shfl_sync(TAXI_FULL_MASK, fin, l_last);
    __syncwarp();

    if (lane != 0) return;

    // ---- first-path traceback + fused distance counts (lane 0) -------------------------------
    int state = state_of_tag(fin & 3, sc);
    int i = nA, j = nB;
    int same = 0, ts = 0, tv = 0, gapc = 0, pend = 0;
    bool seen = false;
    const bool strings = a.aln_x != nullptr;
    int64_t wpos = strings ? a.aln_off[p + 1] : 0;
    uint8_t* ox = a.aln_x;
    uint8_t* oy = a.aln_y;
    while (i > 0 && j > 0) {
        const int q = (i - 1) % SL;
        const int l = q / H, r = q % H, s = (i - 1) / SL;
        const int tb = (int)__ldcg(trace + ((size_t)(s * step_stride + (j - 1 + l)) * 32 + l) * HB + r);
        const int ca = (int)__ldg(x + i - 1), cb = (int)__ldg(y + j - 1);
        const int ka = base_class(ca), kb = base_class(cb);
        int tag;
        if (state == 0) {
            tag = tb & 3;
            if (ka < 4 && kb < 4) {
                if (seen) gapc += pend;
                pend = 0; seen = true;
                const int d = ka ^ kb;
                same += (d == 0); ts += (d == 1); tv += (d > 1);
            } else if ((ka == 4 && kb < 4) || (kb == 4 && ka < 4)) {
                ++pend;
            }
            if (strings) { --wpos; ox[wpos] = (uint8_t)ca; oy[wpos] = (uint8_t)cb; }
            --i; --j;
        } else if (state == 1) {
            tag = (tb >> 2) & 3;
            pend += (ka < 4);
            if (strings) { --wpos; ox[wpos] = (uint8_t)ca; oy[wpos] = '-'; }
            --i;
        } else {
            tag = (tb >> 4) & 3;
            pend += (kb < 4);
            if (strings) { --wpos; ox[wpos] = '-'; oy[wpos] = (uint8_t)cb; }
            --j;
        }
        state = state_of_tag(tag, sc);
    }
    if (strings) {
        // leading end gap: whatever is left of x (vertical) or y (horizontal)
        while (i > 0) { --wpos; ox[wpos] = __ldg(x + i - 1); oy[wpos] = '-'; --i; }
        while (j > 0) { --wpos; ox[wpos] = '-'; oy[wpos] = __ldg(y + j - 1); --j; }
        a.aln_start[p] = wpos;
    }
    if (a.score) a.score[p] = fin >> TAG_BITS;
    if (a.counts) {
        *reinterpret_cast<int4*>(a.counts + 4 * p) = make_int4(same, ts, tv, gapc);
    }
    if (a.metrics) {
        double m[4];
        metrics_from_counts(same, ts, tv, gapc, m);
        double2* dst = reinterpret_cast<double2*>(a.metrics + 4 * p);
        dst[0] = make_double2(m[0], m[1]);
        dst[1] = make_double2(m[2], m[3]);
    }
}

constexpr int GOTOH_WARPS_PER_BLOCK = 4;

template <int H>
__global__ void __launch_bounds__(GOTOH_WARPS_PER_BLOCK * 32)
gotoh_warp_kernel(const AlignArgs a)
{
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * GOTOH_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    uint8_t* trace = a.trace + gw * a.trace_per_warp;
    int32_t* bnd = a.bnd + gw * a.bnd_per_warp;
    for (;;) {
        unsigned long long p = 0;
        if (lane == 0) p = atomicAdd(a.counter, 1ULL);
        p = __shfl_sync(TAXI_FULL_MASK, p, 0);
        if (p >= (unsigned long long)a.npairs) break;
        align_one<H>(a, (long long)p, lane, trace, bnd);
        __syncwarp();
    }
}

}  // namespace taxi
