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

// COOP = false: the warp runs the stripes of its pair one after the other (inter-task: one pair per
// warp).  COOP = true (intra-task, long pairs): the `nw` warps of a CTA share ONE pair -- warp `wid`
// takes stripes wid, wid + nw, ... and the stripes run as a pipeline: stripe s + 1 follows stripe s
// a few dozen columns behind, reading the bottom row of stripe s from that stripe's own boundary
// buffer as soon as `progress[s]` (shared memory) says the column has been published.  The warp
// that owns the last stripe then walks the path.  Same arithmetic, same trace layout.
template <int H, bool COOP = false>
__device__ __forceinline__ void align_one(const AlignArgs& a, long long p, int lane, uint8_t* trace, int32_t* bnd,
                                          int wid = 0, int nw = 1, volatile int* progress = nullptr)
{
    using G = TraceGeom<H>;
    constexpr int HB = G::HB;
    constexpr int SL = G::SL;
    const ScoreSet& sc = a.sc;

    const PairIndex pi = pair_index(a, p);
    const int xi = pi.xi, yi = pi.yi;
    p = pi.out;   // from here on p only addresses the outputs
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

    for (int s = COOP ? wid : 0; s < nstripes; s += COOP ? nw : 1) {
        // boundary rows: one buffer per warp (sequential stripes reuse it: the writer trails the
        // reader by 31 columns), one per stripe when the stripes of a pair run concurrently
        int32_t* bnd_in = COOP ? bnd + (size_t)max(s - 1, 0) * a.bnd_per_warp : bnd;
        int32_t* bnd_out = COOP ? bnd + (size_t)s * a.bnd_per_warp : bnd;
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

        // the y symbol of a column and, for the top rows of a later stripe, the values the previous
        // stripe left in the boundary buffer are fetched one step ahead: their latency hides behind a
        // whole step instead of stalling the warp at the top of it
        int nb = 0, nbX = 0, nbH = 0;
        int seen = 0;                                  // COOP: columns of the stripe above known to be published
        auto fetch_next = [&](int jj) {
            if (live && jj >= 1 && jj <= nB) {
                nb = (int)__ldg(y + jj - 1);
                if (lane == 0 && s > 0) {
                    TAXI_CHECK(a, 2LL * jj + 1 < a.bnd_per_warp, 11);
                    if (COOP && jj > seen) {                   // poll only when the last look did not already cover this column
                        int got;
                        while ((got = progress[s - 1]) < jj) { }   // the stripe above has not published this column yet
                        seen = got;
                        __threadfence_block();
                    }
                    nbX = __ldcg(bnd_in + 2 * jj); nbH = __ldcg(bnd_in + 2 * jj + 1);
                }
            }
        };
        fetch_next(1 - lane);
        for (int t = 0; t < nsteps; ++t) {
            const int j = t - lane + 1;
            int rX = __shfl_up_sync(TAXI_FULL_MASK, outX, 1);
            int rH = __shfl_up_sync(TAXI_FULL_MASK, outH, 1);
            const bool active = live && j >= 1 && j <= nB;
            const int b = nb;
            if (lane == 0 && active) {
                if (s == 0) {
                    // row 0 of the matrix: only Iy is alive there (leading end gap of j columns)
                    const int y0j = sc.eo + (j - 1) * sc.ee;
                    rH = y0j | sc.tagY;
                    rX = (y0j + (j == nB ? sc.eo : sc.io)) | sc.tagY;
                } else {
                    rX = nbX;
                    rH = nbH;
                }
            }
            fetch_next(j + 1);
            if (active) {
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
                TAXI_CHECK(a, dst >= trace && dst + HB <= trace + a.trace_per_warp, 12);
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
                    TAXI_CHECK(a, 2LL * j + 1 < a.bnd_per_warp, 13);
                    __stcg(bnd_out + 2 * j, outX);
                    __stcg(bnd_out + 2 * j + 1, outH);
                    if (COOP && ((j & 7) == 0 || j == nB)) {   // publish every 8th column (and the last): one fence per 8 steps
                        __threadfence_block();
                        progress[s] = j;                       // columns 1..j of this stripe's bottom row are visible
                    }
                }
            }
        }
        __syncwarp();
    }

    if (COOP) {
        __syncthreads();                               // every stripe of the pair is complete, its trace codes written
        if ((nstripes - 1) % nw != wid) return;        // only the owner of the last stripe holds H(nA, nB) and walks the path
    }

    // ---- end state and score: H(nA, nB) sits in the register slot of row nA ------------------
    const int l_last = ((nA - 1) % SL) / H;
    const int r_last = (nA - 1) % H;
    int fin = 0;
#pragma unroll
    for (int r = 0; r < H; ++r) fin = (r == r_last) ? Hl[r] : fin;
    fin = __shfl_sync(TAXI_FULL_MASK, fin, l_last);
    __syncwarp();

    // ---- first-path traceback + fused distance counts, warp-parallel --------------------------
    // The walk is sequential by nature, but it moves in long straight runs (diagonal runs between
    // indels).  Each iteration the 32 lanes fetch the trace codes of the next 32 cells straight
    // ahead in the current direction, a ballot finds how far the path really goes that way, and
    // the columns of that run are classified and counted with ballots.  ~70 iterations per
    // 650 bp pair instead of ~1300 dependent loads.
    int state = state_of_tag(fin & 3, sc);
    int i = nA, j = nB;
    int same = 0, ts = 0, tv = 0, gapc = 0, pend = 0;
    bool seen = false;
    const bool strings = a.aln_x != nullptr;
    int64_t wpos = strings ? a.aln_off[p + 1] : 0;
    while (i > 0 && j > 0) {
        const int di = (state != 2), dj = (state != 1);
        const int ii = i - lane * di, jj = j - lane * dj;
        const bool valid = ii >= 1 && jj >= 1;
        int tb = 0, ca = 0, cb = 0;
        if (valid) {
            const int q = (ii - 1) % SL;
            const int l = q / H, r = q % H, s = (ii - 1) / SL;
            TAXI_CHECK(a, (long long)(((size_t)(s * step_stride + (jj - 1 + l)) * 32 + l) * HB + r) < a.trace_per_warp, 14);
            tb = (int)__ldcg(trace + ((size_t)(s * step_stride + (jj - 1 + l)) * 32 + l) * HB + r);
            ca = (int)__ldg(x + ii - 1);
            cb = (int)__ldg(y + jj - 1);
        }
        const int tag = (tb >> (2 * state)) & 3;
        const int own = state == 0 ? sc.pM : (state == 1 ? sc.pX : sc.pY);
        const unsigned cont = __ballot_sync(TAXI_FULL_MASK, valid && tag == own);
        const unsigned vmask = __ballot_sync(TAXI_FULL_MASK, valid);
        // lanes 0..f are visited in this state, f = first lane whose pointer leaves the run
        const int f = __ffs(~cont) - 1;                     // 0..31, or -1 when all 32 continue
        int V = (f < 0) ? 32 : f + 1;
        V = min(V, __popc(vmask));
        const unsigned visited = (V == 32) ? 0xffffffffu : ((1u << V) - 1u);
        const int last_tag = __shfl_sync(TAXI_FULL_MASK, tag, V - 1);
        // classify my column
        const int ka = (state == 2) ? 4 : base_class(ca);
        const int kb = (state == 1) ? 4 : base_class(cb);
        const bool both = ka < 4 && kb < 4;
        const int d = ka ^ kb;
        const unsigned bm = __ballot_sync(TAXI_FULL_MASK, both) & visited;
        const unsigned gm = __ballot_sync(TAXI_FULL_MASK, (ka == 4) != (kb == 4) && (ka < 4 || kb < 4)) & visited;
        const unsigned tsm = __ballot_sync(TAXI_FULL_MASK, both && d == 1) & visited;
        const unsigned tvm = __ballot_sync(TAXI_FULL_MASK, both && d > 1) & visited;
        if (bm) {
            ts += __popc(tsm); tv += __popc(tvm); same += __popc(bm & ~(tsm | tvm));
            const int fb = __ffs(bm) - 1, lb = 31 - __clz(bm);
            const unsigned below = (1u << fb) - 1u;                           // visited before the first both-real column
            const unsigned upto = (lb == 31) ? 0xffffffffu : ((2u << lb) - 1u);
            if (seen) gapc += pend + __popc(gm & below);
            gapc += __popc(gm & upto & ~below);
            pend = __popc(gm & ~upto);
            seen = true;
        } else {
            pend += __popc(gm);
        }
        if (strings && lane < V) {
            TAXI_CHECK(a, wpos - 1 - lane >= a.aln_off[p] && wpos - 1 - lane < a.aln_off[p + 1], 15);
            a.aln_x[wpos - 1 - lane] = (state == 2) ? (uint8_t)'-' : (uint8_t)ca;
            a.aln_y[wpos - 1 - lane] = (state == 1) ? (uint8_t)'-' : (uint8_t)cb;
        }
        wpos -= V;
        i -= V * di; j -= V * dj;
        state = state_of_tag(last_tag, sc);
    }
    if (strings) {
        // leading end gap: whatever is left of x (vertical) or y (horizontal)
        for (int k = lane; k < i; k += 32) { a.aln_x[wpos - 1 - k] = __ldg(x + i - 1 - k); a.aln_y[wpos - 1 - k] = '-'; }
        wpos -= i;
        for (int k = lane; k < j; k += 32) { a.aln_x[wpos - 1 - k] = '-'; a.aln_y[wpos - 1 - k] = __ldg(y + j - 1 - k); }
        wpos -= j;
        if (lane == 0) a.aln_start[p] = wpos;
    }
    if (lane != 0) return;
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

// Intra-task kernel for long pairs: one pair per CTA, its stripes pipelined over the CTA's warps.
// Taken when a launch has fewer pairs than the GPU has resident warps and the pairs span several
// stripes (a single 12 kbp x 9 kbp pair: 18 stripes over 8 warps instead of one warp doing all).
constexpr int GOTOH_COOP_WARPS = 12;
constexpr int GOTOH_COOP_MAX_STRIPES = 1024;

template <int H>
__global__ void __launch_bounds__(GOTOH_COOP_WARPS * 32, 1)
gotoh_coop_kernel(const AlignArgs a)
{
    __shared__ volatile int progress[GOTOH_COOP_MAX_STRIPES];
    __shared__ unsigned long long next_pair;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* trace = a.trace + (long long)blockIdx.x * a.trace_per_warp;                       // per CTA here
    int32_t* bnd = a.bnd + (long long)blockIdx.x * a.bnd_per_warp * a.coop_stripes;             // one buffer per stripe
    for (;;) {
        __syncthreads();                                   // everyone is done with the previous pair (and its progress flags)
        for (int k = threadIdx.x; k < GOTOH_COOP_MAX_STRIPES; k += blockDim.x) progress[k] = 0;
        if (threadIdx.x == 0) next_pair = atomicAdd(a.counter, 1ULL);
        __syncthreads();
        const unsigned long long p = next_pair;
        if (p >= (unsigned long long)a.npairs) break;
        align_one<H, true>(a, (long long)p, lane, trace, bnd, wid, GOTOH_COOP_WARPS, progress);
    }
}

}  // namespace taxi
