// Integer-pipe microbenchmark for sm_100a: measures issue throughput (lane-ops / clk / SM) of the
// instructions the Gotoh kernels are built from, so that DP roofline fractions are quoted against
// a MEASURED INT32 peak (SURVEY.md 8d / BASELINE.md 2 leave it "to be measured by the build").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/int_peak int_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

constexpr int ITERS = 2048;
constexpr int ILP = 8;

enum Op { IADD3, LOP3, VIMNMX, VIMNMX3, VIADDMNMX, IMAD, SELP, PRMT, MAX16, ADDMAX16, MAX316, ADD16, MIX_ALU_FMA, SHFL, DPCELL, POPC, DFMA, COUNTWORD, NOPS };
const char* kNames[] = {"IADD3", "LOP3", "VIMNMX", "VIMNMX3", "VIADDMNMX", "IMAD", "ISETP+SEL", "PRMT",
                        "VIMNMX.S16x2", "VIADDMNMX.S16x2", "VIMNMX3.S16x2", "VIADD.16x2", "VIMNMX+IMAD (2 ops)", "SHFL.UP", "DP-cell mix (16 ops)",
                        "POPC", "DFMA", "count-word mix (7 LOP3 + 4 POPC + 4 IADD)"};
const int kOpsPerIter[] = {1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 2, 1, 16, 1, 1, 15};

template <int OP>
__global__ void __launch_bounds__(1024) bench(int* out, int a0, int b0, long long* cycles)
{
    int v[ILP], w[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { v[k] = threadIdx.x * 7 + k * a0; w[k] = threadIdx.x ^ (k + b0); }
    const int c1 = a0 | 1, c2 = b0 | 3;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == IADD3) { asm volatile("add.s32 %0, %0, %1;" : "+r"(v[k]) : "r"(w[k])); }
            else if (OP == LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[k]) : "r"(w[k]), "r"(c1)); }
            else if (OP == VIMNMX) { asm volatile("max.s32 %0, %0, %1;" : "+r"(v[k]) : "r"(w[k])); w[k] ^= 0; }
            else if (OP == VIMNMX3) { v[k] = __vimax3_s32(v[k], w[k], v[(k + 1) % ILP]); }
            else if (OP == VIADDMNMX) { v[k] = __viaddmax_s32(v[k], c1, w[k]); }
            else if (OP == IMAD) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(c1), "r"(w[k])); }
            else if (OP == SELP) { v[k] = (v[k] == w[k]) ? c1 : (v[k] + c2); }
            else if (OP == PRMT) { asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(w[k]), "r"(c2)); }
            else if (OP == MAX16) { v[k] = __vmaxs2(v[k], w[k]); w[k] += 0; }
            else if (OP == ADDMAX16) { v[k] = __viaddmax_s16x2(v[k], c1, w[k]); }
            else if (OP == MAX316) { v[k] = __vimax3_s16x2(v[k], w[k], v[(k + 1) % ILP]); }
            else if (OP == ADD16) { v[k] = __vadd2(v[k], w[k]); }
            else if (OP == MIX_ALU_FMA) {
                asm volatile("max.s32 %0, %0, %1;" : "+r"(v[k]) : "r"(w[k]));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w[k]) : "r"(c1), "r"(c2));
            }
            else if (OP == SHFL) { v[k] = __shfl_up_sync(0xffffffffu, v[k], 1); }
            else if (OP == POPC) { asm volatile("popc.b32 %0, %0;" : "+r"(v[k])); v[k] ^= w[k]; }
            else if (OP == DFMA) {
                double d = __hiloint2double(v[k], w[k]);
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5));
                v[k] = __double2hiint(d); w[k] = __double2loint(d);
            }
            else if (OP == COUNTWORD) {
                // one 32-column word of the alignment-free kernel: both / transversion / transition / gap masks + 4 popcounts
                const unsigned x0 = v[k], x1 = w[k], xr = v[(k + 1) % ILP], xg = w[(k + 1) % ILP], y0 = c1 ^ it, y1 = c2 + it, yr = ~it, yg = it * 3;
                const unsigned both = xr & yr, d1 = x1 ^ y1, tvm = both & d1, tsm = (x0 ^ y0) & both & ~d1, gm = (xg & yr) | (xr & yg);
                v[k] += __popc(both) + __popc(tvm);
                w[k] += __popc(tsm) + __popc(gm);
            }
            else if (OP == DPCELL) {
                // the op mix of one tagged Gotoh cell: 5 LOP3, 1 max3, 2 max, 2 addmax, 3 add, cmp+sel, prmt
                int M = v[k], X = w[k], Y = v[(k + 1) % ILP];
                int sub = (M == c2) ? c1 : c2;
                int Mr = X + sub;
                int t; asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(t) : "r"(3), "r"(Mr), "r"(X));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(t) : "r"(15), "r"(t), "r"(Y));
                int Mt, Xt, Yt;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(Mt) : "r"(Mr), "r"(~63), "r"(c1));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(Xt) : "r"(X), "r"(~63), "r"(c2));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(Yt) : "r"(Y), "r"(~63), "r"(c1 + 1));
                int Hc = __vimax3_s32(Mt, Xt, Yt);
                int Xn = __viaddmax_s32(max(Mt, Yt), c1, Xt + c2);
                int Yn = __viaddmax_s32(max(Mt, Xt), c2, Yt + c1);
                v[k] = Hc ^ __byte_perm(t, Yn, 0x0040);
                w[k] = Xn;
            }
        }
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc ^= v[k] ^ w[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Every op is timed over back-to-back launches lasting >= 60 ms, after the whole device has been
// busy for >= 200 ms (main), so that the SM clock has ramped: `sm_mhz` = median per-block cycle
// count / per-launch time is the clock the measurement actually ran at, and gops_per_s is only
// meaningful together with it.  The hardware constant is lane_ops_per_clk_per_sm.
template <int OP> void run(int sms, int* d_out, long long* d_cyc, int threads, int blocks_per_sm)
{
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<OP><<<blocks, threads>>>(d_out, 3, 5, d_cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<OP><<<blocks, threads>>>(d_out, 3, 5, d_cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms1 = 0; cudaEventElapsedTime(&ms1, e0, e1);
    const int reps = std::max(1, (int)(60.0f / std::max(ms1, 1e-3f)) + 1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) bench<OP><<<blocks, threads>>>(d_out, 3, 5, d_cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    std::vector<long long> cyc(blocks);
    cudaMemcpy(cyc.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(cyc.begin(), cyc.end());
    const double med = (double)cyc[blocks / 2];
    const double ops_per_block = (double)threads * ITERS * ILP * kOpsPerIter[OP];
    const double per_clk_sm = ops_per_block * blocks_per_sm / med;
    const double total = ops_per_block * blocks;
    printf("{\"op\": \"%s\", \"threads_per_sm\": %d, \"lane_ops_per_clk_per_sm\": %.2f, \"gops_per_s\": %.1f, \"ms_per_launch\": %.4f, \"launches\": %d, \"timed_ms\": %.1f, \"sm_mhz\": %.0f}\n",
           kNames[OP], threads * blocks_per_sm, per_clk_sm, total / (ms * 1e6), ms, reps, ms * reps, med / (ms * 1e3));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    int* d_out; long long* d_cyc;
    cudaMalloc(&d_out, sizeof(int) * sms * 2 * 1024);
    cudaMalloc(&d_cyc, sizeof(long long) * sms * 2);
    {   // ramp the clocks: keep every SM busy for >= 200 ms before anything is timed
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms = 0;
        cudaEventRecord(e0);
        do {
            for (int r = 0; r < 64; ++r) bench<IADD3><<<sms * 2, 1024>>>(d_out, 3, 5, d_cyc);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        } while (ms < 200.f);
    }
    for (int bps = 1; bps <= 2; ++bps) {
        run<IADD3>(sms, d_out, d_cyc, 1024, bps);
        run<LOP3>(sms, d_out, d_cyc, 1024, bps);
        run<VIMNMX>(sms, d_out, d_cyc, 1024, bps);
        run<VIMNMX3>(sms, d_out, d_cyc, 1024, bps);
        run<VIADDMNMX>(sms, d_out, d_cyc, 1024, bps);
        run<IMAD>(sms, d_out, d_cyc, 1024, bps);
        run<SELP>(sms, d_out, d_cyc, 1024, bps);
        run<PRMT>(sms, d_out, d_cyc, 1024, bps);
        run<MAX16>(sms, d_out, d_cyc, 1024, bps);
        run<ADDMAX16>(sms, d_out, d_cyc, 1024, bps);
        run<MAX316>(sms, d_out, d_cyc, 1024, bps);
        run<ADD16>(sms, d_out, d_cyc, 1024, bps);
        run<MIX_ALU_FMA>(sms, d_out, d_cyc, 1024, bps);
        run<SHFL>(sms, d_out, d_cyc, 1024, bps);
        run<DPCELL>(sms, d_out, d_cyc, 1024, bps);
        run<POPC>(sms, d_out, d_cyc, 1024, bps);
        run<DFMA>(sms, d_out, d_cyc, 1024, bps);
        run<COUNTWORD>(sms, d_out, d_cyc, 1024, bps);
    }
    return 0;
}
