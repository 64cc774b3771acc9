// Tensor-core alternative for the alignment-free counts (north star: "an int8 one-hot tensor-core
// contraction is kept only if ncu shows it beats popcount"): a hand-written tcgen05 kernel that
// computes  same[i][j] = sum_c [x_i[c] == y_j[c], both A/C/G/T]  as the int8 GEMM  X . X^T  over
// one-hot rows (4 bytes per column: A, C, G, T), TMA-fed (128-byte swizzle), accumulators in TMEM,
// one 128 x 256 output tile per CTA, warp-specialised (TMA producer / MMA issuer / 4 epilogue
// warps).  It is a measurement tool, not part of the library: it times the contraction at the
// BASELINE C2 size (9000 x 9000 x 618 columns) for K = 2560 bytes (the `same` product alone) and
// K = 5760 bytes (as many k-blocks as all the products the distances need: same 4L, class match 2L,
// both-real L, gap columns 2L), and checks the K = 2560 result against a direct count.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/tc_onehot tools/tc_onehot.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int BM = 128, BN = 256, BK = 128;       // tile: rows x rows x bytes of K per stage
constexpr int UMMA_K = 32;                          // bytes of K per tcgen05.mma (8-bit operands)
constexpr int STAGES = 4;
constexpr int TMEM_COLS = 256;
constexpr int THREADS = 192;                        // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile written by TMA with the 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(const void* tile, int k_byte_offset)
{
    const uint32_t addr = smem_u32(tile) + (uint32_t)k_byte_offset;
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);     // start address, 16-byte units
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: next group of 8 rows
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
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

__global__ void __launch_bounds__(THREADS, 1)
onehot_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   int32_t* __restrict__ out, int nrows, int ncols, int ldo, int kblocks)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * BM * BK;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * BN * BK);
    uint64_t* empty = full + STAGES;
    uint64_t* accum_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(empty + s, ((kb / STAGES) & 1) ^ 1);
                mbar_expect_tx(full + s, (BM + BN) * BK);
                tma_load_2d(sA + s * BM * BK, &map_a, full + s, kb * BK, m0);
                tma_load_2d(sB + s * BN * BK, &map_b, full + s, kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: S32 accumulate, unsigned 8-bit A and B, both K-major, N = 256, M = 128
            const uint32_t idesc = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(full + s, (kb / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_i8(tmem, umma_desc(sA + s * BM * BK, k * UMMA_K), umma_desc(sB + s * BN * BK, k * UMMA_K), idesc, (kb | k) != 0);
                umma_commit(empty + s);        // frees the stage once the MMAs that read it are done
            }
            umma_commit(accum_full);
        }
    } else {
        mbar_wait(accum_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                // a warp reaches TMEM lanes [32 * (warp % 4), +32)
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int col = n0 + c * 32;
            if (row < nrows && col < ncols) {
                int4* dst = reinterpret_cast<int4*>(out + (size_t)row * ldo + col);
#pragma unroll
                for (int k = 0; k < 8; ++k) dst[k] = make_int4((int)v[4 * k], (int)v[4 * k + 1], (int)v[4 * k + 2], (int)v[4 * k + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS));
}

// bytes -> one-hot rows: 4 bytes per column (A, C, G, T), zero for anything else; rows padded with zeros
__global__ void onehot_kernel(const uint8_t* __restrict__ seq, int n, int L, uint8_t* __restrict__ out, int K)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * L) return;
    const int i = (int)(gid / L), c = (int)(gid % L);
    const int ch = seq[(size_t)i * L + c] & 0xDF;
    const int k = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
    uint32_t w = k < 0 ? 0u : (1u << (8 * k));
    *reinterpret_cast<uint32_t*>(out + (size_t)i * K + 4 * c) = w;
}

__global__ void direct_same_kernel(const uint8_t* __restrict__ seq, int L, const int* __restrict__ pi, const int* __restrict__ pj, int np, int* __restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    int s = 0;
    for (int c = 0; c < L; ++c) {
        const int a = seq[(size_t)pi[p] * L + c] & 0xDF, b = seq[(size_t)pj[p] * L + c] & 0xDF;
        s += (a == b) && (a == 'A' || a == 'C' || a == 'G' || a == 'T');
    }
    out[p] = s;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiled encode, void* base, uint64_t K, uint64_t rows, uint32_t box_rows)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {K, rows};
    cuuint64_t strides[1] = {K};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 9000;
    const int L = 640;                                   // 618 columns padded to 20 words
    const int rows_pad = (n + BN - 1) / BN * BN;
    const int Kmax = 5760;
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    std::vector<uint8_t> h((size_t)n * L);
    srand(9000);
    std::vector<uint8_t> root(L);
    for (int c = 0; c < L; ++c) root[c] = "ACGT"[rand() & 3];
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < L; ++c) {
            const int r = rand() % 100;
            h[(size_t)i * L + c] = c >= 618 ? '-' : (r < 85 ? root[c] : r < 97 ? "ACGT"[rand() & 3] : r < 99 ? '-' : 'N');
        }
    uint8_t *d_seq, *d_hot;
    int32_t* d_out;
    CHECK(cudaMalloc(&d_seq, h.size()));
    CHECK(cudaMemcpy(d_seq, h.data(), h.size(), cudaMemcpyHostToDevice));
    CHECK(cudaMalloc(&d_hot, (size_t)rows_pad * Kmax));
    CHECK(cudaMemset(d_hot, 0, (size_t)rows_pad * Kmax));
    const int ldo = rows_pad;
    CHECK(cudaMalloc(&d_out, (size_t)rows_pad * ldo * sizeof(int32_t)));

    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { fprintf(stderr, "cuTensorMapEncodeTiled unavailable\n"); return 1; }
    EncodeTiled encode = (EncodeTiled)fn;

    const size_t smem = (size_t)STAGES * (BM + BN) * BK + 1024;
    CHECK(cudaFuncSetAttribute(onehot_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n + BN - 1) / BN, (n + BM - 1) / BM);
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));

    for (int K : {2560, 5760}) {
        // K = 2560: the real one-hot rows.  K = 5760: the same rows followed by repeats of them, only to time
        // as many k-blocks as the full set of products needs (the result is not meaningful)
        CHECK(cudaMemset(d_hot, 0, (size_t)rows_pad * Kmax));
        onehot_kernel<<<(unsigned)(((long long)n * L + 255) / 256), 256>>>(d_seq, n, L, d_hot, K);
        if (K > 2560) {
            for (int off = 2560; off < K; off += 640)
                CHECK(cudaMemcpy2D(d_hot + off, K, d_hot, K, std::min(640, K - off), n, cudaMemcpyDeviceToDevice));
        }
        CHECK(cudaDeviceSynchronize());
        CUtensorMap map_a = make_map(encode, d_hot, K, rows_pad, BM);
        CUtensorMap map_b = make_map(encode, d_hot, K, rows_pad, BN);
        const int kblocks = K / BK;
        float best = 1e30f;
        for (int it = 0; it < 6; ++it) {
            CHECK(cudaEventRecord(e0));
            onehot_gemm_kernel<<<grid, THREADS, smem>>>(map_a, map_b, d_out, n, n, ldo, kblocks);
            CHECK(cudaEventRecord(e1));
            CHECK(cudaDeviceSynchronize());
            CHECK(cudaGetLastError());
            float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 && ms < best) best = ms;
        }
        int bad = -1, np = 0;
        if (K == 2560) {
            np = 4096;
            std::vector<int> pi(np), pj(np), want(np), got(np);
            for (int p = 0; p < np; ++p) { pi[p] = rand() % n; pj[p] = rand() % n; }
            pi[0] = 0; pj[0] = 0; pi[1] = n - 1; pj[1] = n - 1; pi[2] = 0; pj[2] = n - 1;
            int *d_pi, *d_pj, *d_w;
            CHECK(cudaMalloc(&d_pi, np * 4)); CHECK(cudaMalloc(&d_pj, np * 4)); CHECK(cudaMalloc(&d_w, np * 4));
            CHECK(cudaMemcpy(d_pi, pi.data(), np * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(d_pj, pj.data(), np * 4, cudaMemcpyHostToDevice));
            direct_same_kernel<<<(np + 127) / 128, 128>>>(d_seq, L, d_pi, d_pj, np, d_w);
            CHECK(cudaMemcpy(want.data(), d_w, np * 4, cudaMemcpyDeviceToHost));
            bad = 0;
            for (int p = 0; p < np; ++p) {
                CHECK(cudaMemcpy(&got[p], d_out + (size_t)pi[p] * ldo + pj[p], 4, cudaMemcpyDeviceToHost));
                if (got[p] != want[p]) { if (bad < 5) fprintf(stderr, "mismatch (%d,%d): got %d want %d\n", pi[p], pj[p], got[p], want[p]); ++bad; }
            }
        }
        const double ops = 2.0 * (double)n * n * K;
        printf("{\"kernel\": \"onehot_gemm_kernel (tcgen05.mma kind::i8, TMA, TMEM)\", \"n\": %d, \"K_bytes\": %d, \"ms\": %.4f, \"pairs_per_s\": %.4g, "
               "\"tera_ops_per_s\": %.1f, \"output_bytes\": %.3g, \"checked_pairs\": %d, \"mismatches\": %d, \"sms\": %d}\n",
               n, K, best, (double)n * n / best * 1e3, ops / best / 1e9, (double)n * n * 4, np, bad, prop.multiProcessorCount);
    }
    return 0;
}
