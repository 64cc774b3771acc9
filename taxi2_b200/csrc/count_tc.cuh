// Alignment-free mode on the tensor cores: the per-pair column counts of count_planes.cuh as ONE
// int8 contraction per 128 x 128 tile of the pair matrix, hand-written for sm_100a
// (tcgen05.mma kind::i8 issued by one elected thread, operands staged by TMA with the 128-byte
// swizzle, four int32 accumulators in TMEM, read back with tcgen05.ld by the epilogue warps).
//
// Replaces, like count_rect_kernel, calc.seq_distances_{p,p_gaps,jukes_cantor,kimura2p} applied to
// pre-aligned strings (reference: src/itaxotools/taxi2/distances.py:319-348 from
// versus_all.py:546-552 with params.pairs.align = False).  north_star: "an int8 one-hot
// tensor-core contraction is kept only if ncu shows it beats popcount" -- it does: the popcount
// kernel is bound by POPC (16 lanes/clk/SM) at 1.35 ms for 9000 x 9000 x 618 columns, the
// contraction below takes 0.3-0.4 ms (tools/tc_onehot.cu, profiles/tc_onehot_r02.json).
//
// Encoding.  With the 2-bit nucleotide code of the planes (A=00 G=01 C=10 T=11) write u0, u1 = +1 /
// -1 for bit 0 / bit 1 of a real base and 0 for anything else, R = 1 for a real base.  Summed over
// the columns of a pair,
//     n    = R.R                          both real
//     S1   = u1.u1                        purine/pyrimidine class agrees minus disagrees
//     S0+S01 = u0.u0 + (u0 u1).(u0 u1)
//     same = (n + S1 + (S0 + S01)) / 4,   transitions = (n + S1 - (S0 + S01)) / 4,
//     transversions = (n - S1) / 2
// because [same] = (1 + u0x u0y)(1 + u1x u1y) / 4 on a both-real column.  Gap columns are
// G'x.Ry + Rx.G'y with the clipped gap plane G' of count_planes.cuh.  Every sequence therefore
// becomes a row of 8 * Lp signed bytes (Lp = columns padded to 128):
//     [0, Lp)     R                       -> accumulator 0 (x . y)
//     [Lp, 2Lp)   u1                      -> accumulator 1
//     [2Lp, 4Lp)  u0, u0 u1 interleaved   -> accumulator 2
//     [4Lp, 6Lp)  G', R interleaved       -> accumulator 3, the row's role as x
//     [6Lp, 8Lp)  R, G' interleaved       ->                its role as y
// 6 Lp bytes of K per pair.  The few gap columns outside the first / last both-real column are
// subtracted in the epilogue exactly as in the popcount kernel (gaps_outside_trim on the planes),
// and the fp64 metrics are computed there too, so the kernel writes the same 16 B + 32 B per pair.
//
// One CTA per tile of 128 y columns x TX x rows (TcGeom: 32 warps and all 512 TMEM columns for TX = 128,
// 16 warps and half of them for TX = 64, two CTAs per SM).
// The y tile is the A operand, so a TMEM lane -- an epilogue thread -- is a y column and the 32
// threads of a warp write 32 consecutive pairs of one x row: the same coalesced 16 B + 32 B per pair
// as the popcount kernel.  Lane 0 of warp 0 feeds TMA, lane 0 of warp 1 issues the MMAs, then all
// warps (4 TMEM lane quarters x TX / 16 groups of 16 x rows) do trim, metrics and stores -- the
// epilogue, not the contraction, is what bounds the kernel.
#pragma once
#include <cuda.h>

#include "count_planes.cuh"

namespace taxi {

constexpr int TC_TILE = 128;              // pairs per tile side; also bytes of K per pipeline stage
constexpr int TC_BAND = 12;                // x tiles per band of the CTA order (operand reuse in L2)
constexpr int TC_UMMA_K = 32;             // bytes of K per tcgen05.mma (8-bit operands)
constexpr int TC_ROW_SEGMENTS = 8;        // row length in units of Lp
// Tile geometry: TC_TILE y columns (TMEM lanes) x TX x rows (TMEM columns per accumulator).  TX = 128: one
// CTA of 32 warps per SM, the four accumulators fill the 512 TMEM columns.  TX = 64: CTAs of 16 warps, two per
// SM (256 TMEM columns and 97 KB of shared memory each), so one CTA's contraction runs under the other's epilogue.
template <int TX> struct TcGeom {
    static constexpr int THREADS = TX * 8;                          // 16 x rows per warp and TMEM lane quarter
    static constexpr int STAGES = TX == 128 ? 5 : 4;
    static constexpr int STAGE_X = TX * TC_TILE, STAGE_Y = TC_TILE * TC_TILE;
    static constexpr size_t SMEM = (size_t)STAGES * (STAGE_X + STAGE_Y) + 1024;
    static constexpr int TMEM_COLS = 4 * TX;
    static constexpr int CTAS_PER_SM = TX == 128 ? 1 : 2;
};

// one thread per (sequence, 4 columns): planes -> operand row
__global__ void tc_operands_kernel(const uint4* __restrict__ planes, int32_t nseq, int32_t W, int32_t Lp, int8_t* __restrict__ out)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int groups = Lp / 4;
    if (gid >= (long long)nseq * groups) return;
    const int seq = (int)(gid % nseq), c0 = (int)(gid / nseq) * 4;   // sequence fastest: coalesced plane reads
    const int w = c0 / 32;
    uint4 pw = make_uint4(0, 0, 0, 0);
    if (w < W) pw = __ldg(planes + (size_t)w * nseq + seq);
    uint32_t r4 = 0, u14 = 0;
    uint32_t m2[2] = {0, 0}, gx[2] = {0, 0}, gy[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int bit = (c0 + k) & 31;
        const int R = (pw.z >> bit) & 1, G = (pw.w >> bit) & 1;
        const int u0 = R ? (((pw.x >> bit) & 1) ? -1 : 1) : 0, u1 = R ? (((pw.y >> bit) & 1) ? -1 : 1) : 0;
        r4 |= (uint32_t)R << (8 * k);
        u14 |= (uint32_t)(uint8_t)u1 << (8 * k);
        m2[k >> 1] |= ((uint32_t)(uint8_t)u0 | ((uint32_t)(uint8_t)(u0 * u1) << 8)) << (16 * (k & 1));
        gx[k >> 1] |= ((uint32_t)G | ((uint32_t)R << 8)) << (16 * (k & 1));
        gy[k >> 1] |= ((uint32_t)R | ((uint32_t)G << 8)) << (16 * (k & 1));
    }
    int8_t* row = out + (size_t)seq * TC_ROW_SEGMENTS * Lp;
    *reinterpret_cast<uint32_t*>(row + c0) = r4;
    *reinterpret_cast<uint32_t*>(row + Lp + c0) = u14;
    *reinterpret_cast<uint2*>(row + 2 * Lp + 2 * c0) = make_uint2(m2[0], m2[1]);
    *reinterpret_cast<uint2*>(row + 4 * Lp + 2 * c0) = make_uint2(gx[0], gx[1]);
    *reinterpret_cast<uint2*>(row + 6 * Lp + 2 * c0) = make_uint2(gy[0], gy[1]);
}

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same with a pause between polls: for the thousand epilogue threads that wait out the whole contraction --
// polling flat out they take issue slots from the two threads that drive it (and from a co-resident CTA)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile as TMA writes it with the 128-byte swizzle: rows of 128 bytes, groups of 8 rows 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(const void* tile, int k_byte_offset)
{
    const uint32_t addr = smem_u32(tile) + (uint32_t)k_byte_offset;
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, int (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
}

}  // namespace tc

struct CountTcArgs {
    CountArgs c;          // planes (trim correction), rectangle, outputs
    int32_t Lp;           // padded columns of the operand rows (the smaller of the two sets')
    int32_t LpX, LpY;     // padded columns of each set's operand rows (segment offsets)
};

template <int TX>
__global__ void __launch_bounds__(TcGeom<TX>::THREADS, TcGeom<TX>::CTAS_PER_SM)
count_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, const CountTcArgs args)
{
    using namespace tc;
    using G = TcGeom<TX>;
    constexpr int TC_STAGES = G::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sY = smem;                                              // operand tiles: rows of 128 bytes (TC_TILE y rows, TX x rows)
    uint8_t* sX = smem + TC_STAGES * G::STAGE_Y;
    uint64_t* full = reinterpret_cast<uint64_t*>(sX + TC_STAGES * G::STAGE_X);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* accum_full = empty + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
    const CountArgs& a = args.c;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Tile of this CTA.  The grid is one-dimensional and walks the tile matrix in bands of TC_BAND x tiles,
    // inside a band column by column (the band's x tiles of one y tile, then the next y tile): a wave of ~148
    // CTAs then touches about 12 x tiles and 12 y tiles (16 MB of operand rows) instead of 2 and all 71 (48 MB),
    // and the operands stay in L2 under the stream of results.
    const int tiles_y = (a.ny + TC_TILE - 1) / TC_TILE, tiles_x = (a.nx + TX - 1) / TX;
    const int band = (int)blockIdx.x / (TC_BAND * tiles_y), within = (int)blockIdx.x % (TC_BAND * tiles_y);
    const int band_rows = min(TC_BAND, tiles_x - band * TC_BAND);
    const int xt = (band * TC_BAND + within % band_rows) * TX, yt = (within / band_rows) * TC_TILE;   // tile origin inside the rectangle
    const int Lp = args.Lp;
    const int blocks_per_L = Lp / TC_TILE;
    const int kblocks = 6 * blocks_per_L;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&map_y) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    // k-block kb of the pair's 6 Lp bytes: which accumulator it feeds and where it sits in the x / y rows
    auto segment = [&](int kb, int& acc, int& kx, int& ky) {
        const int unit = kb / blocks_per_L, within = (kb % blocks_per_L) * TC_TILE;   // unit: 0 R, 1 u1, 2-3 u0/u0u1, 4-5 gap
        acc = unit == 0 ? 0 : unit == 1 ? 1 : unit < 4 ? 2 : 3;
        // the two-unit segments are interleaved byte pairs: 2 Lp consecutive bytes of the pair's common columns,
        // starting at the segment's base in each row (bases depend on the row's own Lp)
        const int first = acc == 2 ? 2 : acc == 3 ? 4 : unit;               // first unit of this accumulator
        const int into = (unit - first) * Lp + within;                      // byte offset inside the segment
        kx = first * args.LpX + into;
        ky = (acc == 3 ? 6 : first) * args.LpY + into;                      // the gap segment of the y role
    };

    if (warp == 0 && lane == 0) {
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % TC_STAGES;
            int acc, kx, ky;
            segment(kb, acc, kx, ky);
            mbar_wait(empty + s, ((kb / TC_STAGES) & 1) ^ 1);
            mbar_expect_tx(full + s, G::STAGE_X + G::STAGE_Y);
            tma_load_2d(sX + s * G::STAGE_X, &map_x, full + s, kx, a.x0 + xt);
            tma_load_2d(sY + s * G::STAGE_Y, &map_y, full + s, ky, a.y0 + yt);
        }
    } else if (warp == 1 && lane == 0) {
        // S32 accumulators, signed 8-bit A and B, both K-major, N = TX, M = 128; A = the y tile (TMEM lanes), B = the x tile
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TX >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
        int prev_acc = -1;
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % TC_STAGES;
            int acc, kx, ky;
            segment(kb, acc, kx, ky);
            mbar_wait(full + s, (kb / TC_STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < TC_TILE / TC_UMMA_K; ++k)
                umma_i8(tmem + (uint32_t)(acc * TX), umma_desc(sY + s * G::STAGE_Y, k * TC_UMMA_K),
                        umma_desc(sX + s * G::STAGE_X, k * TC_UMMA_K), idesc, (acc == prev_acc || k > 0) ? 1u : 0u);
            prev_acc = acc;
            umma_commit(empty + s);
        }
        umma_commit(accum_full);
    }
    __syncwarp();

    // ---- epilogue: every warp ------------------------------------------------------------------
    mbar_wait_relaxed(accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // every MMA has completed, so the operand stages are free: the logarithm table of the metric epilogue
    // (common.cuh: metrics_from_counts_table) takes their place, entries 0 .. 3 * (common columns)
    const long long* shared_tab = nullptr;
    if (a.lntab != nullptr && a.metrics != nullptr) {
        long long* tab = reinterpret_cast<long long*>(smem);
        const int entries = min(3 * min(a.x.W, a.y.W) * 32 + 1, LN_TABLE_SIZE);
        for (int k = threadIdx.x; k < entries; k += G::THREADS) tab[k] = __ldg(a.lntab + k);
        __syncthreads();
        shared_tab = tab;
    }
    {
        const int q = warp & 3;                          // this warp reaches TMEM lanes [32 q, 32 q + 32): its 32 y columns
        const int cg = warp >> 2;                        // its 16 x rows of the tile
        const int yc = yt + q * 32 + lane;               // column inside the rectangle
        const bool col_ok = yc < a.ny;
        const int ys = a.y0 + min(yc, a.ny - 1);
        const int2 sy = __ldg(a.y.span + ys);
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
            const int r0 = cg * 16 + ch * 8;
            int A0[8], A1[8], A2[8], A3[8];
            tmem_ld8(lane_addr + (uint32_t)(0 * TX + r0), A0);
            tmem_ld8(lane_addr + (uint32_t)(1 * TX + r0), A1);
            tmem_ld8(lane_addr + (uint32_t)(2 * TX + r0), A2);
            tmem_ld8(lane_addr + (uint32_t)(3 * TX + r0), A3);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int xr = xt + r0 + k;              // row inside the rectangle (the same for the whole warp)
                if (xr < a.nx && col_ok) {
                    const int n = A0[k];
                    const int ts = (n + A1[k] - A2[k]) >> 2, tv = (n - A1[k]) >> 1;
                    int gap = A3[k];
                    const int xs = a.x0 + xr;
                    if (n > 0 && gap > 0)
                        gap -= gaps_outside_trim([&](int w) { return a.x.at(w, xs); }, [&](int w) { return a.y.at(w, ys); },
                                                 __ldg(a.x.span + xs), sy);
                    store_pair<true>(a, (long long)xr * a.ny + yc, n, tv, ts, n > 0 ? gap : 0, shared_tab);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(G::TMEM_COLS));
}

// ---- persistent form: the contraction of the next tile runs under the epilogue of this one -----------------
// count_tc_kernel's CTAs all start together, so a whole wave is in its contraction phase (no result traffic)
// and then in its epilogue (DRAM-bound on the 48 B per pair) at the same time: the two phases add up
// (0.37 + 0.6 ms on C2) on an SM and across the GPU, with one CTA per SM or with two.  Here one CTA per SM
// walks a static list of 128 x 64 tiles; the TMA and MMA threads run up to two tiles ahead into the second
// of two accumulator sets in TMEM (2 x 4 x 64 columns) while 16 epilogue warps drain the first.
constexpr int TCP_TX = 64;                 // x rows per tile
constexpr int TCP_STAGES = 5;
constexpr int TCP_EPI_WARPS = 16;          // warps 4 .. 19: TMEM lane quarter = warp % 4, 16 x rows each
constexpr int TCP_THREADS = (4 + TCP_EPI_WARPS) * 32;
constexpr int TCP_STAGE_X = TCP_TX * TC_TILE, TCP_STAGE_Y = TC_TILE * TC_TILE;
constexpr size_t TCP_SMEM = (size_t)TCP_STAGES * (TCP_STAGE_X + TCP_STAGE_Y) + (size_t)LN_TABLE_SIZE * 8 + 1024;

namespace tc {
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
}  // namespace tc

__global__ void __launch_bounds__(TCP_THREADS, 1)
count_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, const CountTcArgs args)
{
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sY = smem;
    uint8_t* sX = smem + TCP_STAGES * TCP_STAGE_Y;
    long long* tab = reinterpret_cast<long long*>(sX + TCP_STAGES * TCP_STAGE_X);
    uint64_t* full = reinterpret_cast<uint64_t*>(tab + LN_TABLE_SIZE);
    uint64_t* empty = full + TCP_STAGES;
    uint64_t* acc_full = empty + TCP_STAGES;      // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const CountArgs& a = args.c;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lp = args.Lp;
    const int blocks_per_L = Lp / TC_TILE;
    const int kblocks = 6 * blocks_per_L;
    const int tiles_y = (a.ny + TC_TILE - 1) / TC_TILE, tiles_x = (a.nx + TCP_TX - 1) / TCP_TX;
    const int ntiles = tiles_x * tiles_y;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&map_y) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TCP_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, TCP_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const bool use_tab = a.lntab != nullptr && a.metrics != nullptr;
    if (use_tab) {
        const int entries = min(3 * min(a.x.W, a.y.W) * 32 + 1, LN_TABLE_SIZE);
        for (int k = threadIdx.x; k < entries; k += TCP_THREADS) tab[k] = __ldg(a.lntab + k);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    auto segment = [&](int kb, int& acc, int& kx, int& ky) {   // as in count_tc_kernel
        const int unit = kb / blocks_per_L, within = (kb % blocks_per_L) * TC_TILE;
        acc = unit == 0 ? 0 : unit == 1 ? 1 : unit < 4 ? 2 : 3;
        const int first = acc == 2 ? 2 : acc == 3 ? 4 : unit;
        const int into = (unit - first) * Lp + within;
        kx = first * args.LpX + into;
        ky = (acc == 3 ? 6 : first) * args.LpY + into;
    };
    auto tile_origin = [&](int tile, int& xt, int& yt) {       // bands of TC_BAND x tiles, column by column (operand reuse in L2)
        const int band = tile / (TC_BAND * tiles_y), within = tile % (TC_BAND * tiles_y);
        const int band_rows = min(TC_BAND, tiles_x - band * TC_BAND);
        xt = (band * TC_BAND + within % band_rows) * TCP_TX;
        yt = (within / band_rows) * TC_TILE;
    };

    if (warp == 0) {
        if (lane == 0) {
            long long kbg = 0;   // k-blocks issued so far, over all tiles: stage and phase of the smem ring
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int xt, yt;
                tile_origin(tile, xt, yt);
                for (int kb = 0; kb < kblocks; ++kb, ++kbg) {
                    const int s = (int)(kbg % TCP_STAGES);
                    int acc, kx, ky;
                    segment(kb, acc, kx, ky);
                    mbar_wait(empty + s, (uint32_t)((kbg / TCP_STAGES) & 1) ^ 1u);
                    mbar_expect_tx(full + s, TCP_STAGE_X + TCP_STAGE_Y);
                    tma_load_2d(sX + s * TCP_STAGE_X, &map_x, full + s, kx, a.x0 + xt);
                    tma_load_2d(sY + s * TCP_STAGE_Y, &map_y, full + s, ky, a.y0 + yt);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TCP_TX >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
            long long kbg = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                mbar_wait(acc_empty + as, (uint32_t)((it >> 1) & 1) ^ 1u);     // the epilogue has drained this accumulator set
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                int prev_acc = -1;
                for (int kb = 0; kb < kblocks; ++kb, ++kbg) {
                    const int s = (int)(kbg % TCP_STAGES);
                    int acc, kx, ky;
                    segment(kb, acc, kx, ky);
                    mbar_wait(full + s, (uint32_t)((kbg / TCP_STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < TC_TILE / TC_UMMA_K; ++k)
                        umma_i8(tmem + (uint32_t)(as * 4 * TCP_TX + acc * TCP_TX), umma_desc(sY + s * TCP_STAGE_Y, k * TC_UMMA_K),
                                umma_desc(sX + s * TCP_STAGE_X, k * TC_UMMA_K), idesc, (acc == prev_acc || k > 0) ? 1u : 0u);
                    prev_acc = acc;
                    umma_commit(empty + s);
                }
                umma_commit(acc_full + as);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                          // TMEM lanes [32 q, 32 q + 32): 32 y columns of the tile
        const int cg = (warp - 4) >> 2;                  // 16 x rows of the tile
        const long long* shared_tab = use_tab ? tab : nullptr;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            int xt, yt;
            tile_origin(tile, xt, yt);
            const int yc = yt + q * 32 + lane;
            const bool col_ok = yc < a.ny;
            const int ys = a.y0 + min(yc, a.ny - 1);
            const int2 sy = __ldg(a.y.span + ys);
            mbar_wait_relaxed(acc_full + as, (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 4 * TCP_TX);
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                const int r0 = cg * 16 + ch * 8;
                int A0[8], A1[8], A2[8], A3[8];
                tmem_ld8(lane_addr + (uint32_t)(0 * TCP_TX + r0), A0);
                tmem_ld8(lane_addr + (uint32_t)(1 * TCP_TX + r0), A1);
                tmem_ld8(lane_addr + (uint32_t)(2 * TCP_TX + r0), A2);
                tmem_ld8(lane_addr + (uint32_t)(3 * TCP_TX + r0), A3);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch == 1) {   // this warp holds everything it needs of the accumulator set: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + as);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int xr = xt + r0 + k;
                    if (xr < a.nx && col_ok) {
                        const int n = A0[k];
                        const int ts = (n + A1[k] - A2[k]) >> 2, tv = (n - A1[k]) >> 1;
                        int gap = A3[k];
                        const int xs = a.x0 + xr;
                        if (n > 0 && gap > 0)
                            gap -= gaps_outside_trim([&](int w) { return a.x.at(w, xs); }, [&](int w) { return a.y.at(w, ys); },
                                                     __ldg(a.x.span + xs), sy);
                        store_pair<true>(a, (long long)xr * a.ny + yc, n, tv, ts, n > 0 ? gap : 0, shared_tab);
                    }
                }
            }
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

}  // namespace taxi
