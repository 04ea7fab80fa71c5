// microbench.cu -- per-SM issue rates of the integer instructions the rollout kernels use (sm_100a).
// Each test runs ITER iterations of an unrolled body with 8 independent dependency chains per
// thread, on 148*k CTAs of 256 threads that are all resident at once; the rate is reported as
// warp-instructions per clock per SM from clock64() deltas.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define ITER 4096
#define CHAINS 8

enum Op { LOP3 = 0, SHF_C, SHF_V, IADD, IMAD, IMADHI, IMADWIDE, POPC, SEL, PRMT, LDS8, FLO, MIX_LOP_IMAD, MIX_LOP_IMADHI, MIX_SHF_IMAD, MIX_LOP_POPC, MIX_LOP_LDS, FUNNEL, MIX3, NOPS };
static const char* NAMES[] = {"lop3", "shf.r const", "shf.r var", "iadd3", "imad", "imad.hi", "imad.wide", "popc", "sel(setp+selp)", "prmt", "lds.u8 random", "flo(clz)", "lop3+imad 1:1", "lop3+imad.hi 1:1", "shf+imad 1:1", "lop3+popc 4:1", "lop3+lds 4:1", "shf.r.u64 funnel", "lop3+imad+lds 4:4:1"};

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, long long* cycles, uint32_t seed) {
    __shared__ uint8_t lut[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) lut[i] = (uint8_t)(i * 37 + 11);
    __syncthreads();
    uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = seed * (threadIdx.x + 1) + c * 0x9E3779B9u; y[c] = x[c] ^ 0x5555u; }
    const uint32_t k1 = seed | 1u, k2 = (seed >> 3) | 5u;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
            if (OP == SHF_C) asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(k1));
            if (OP == SHF_V) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2));
            if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(k1));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2));
            if (OP == IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(k1));
            if (OP == IMADWIDE) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[c]), "r"(k1)); x[c] = (uint32_t)w ^ (uint32_t)(w >> 32); }
            if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
            if (OP == SEL) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(x[c]) : "r"(k1), "r"(k2));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(x[c]) : "r"(k1));
            if (OP == LDS8) { x[c] = lut[x[c] & 1023] + (x[c] >> 3); }
            if (OP == FLO) asm volatile("clz.b32 %0, %0;" : "+r"(x[c]));
            if (OP == MIX_LOP_IMAD) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(k1), "r"(k2)); }
            if (OP == MIX_LOP_IMADHI) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2)); asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(y[c]) : "r"(k1)); }
            if (OP == MIX_SHF_IMAD) { asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(k1)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(k1), "r"(k2)); }
            if (OP == MIX_LOP_POPC) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                if ((c & 3) == 0) asm volatile("popc.b32 %0, %0;" : "+r"(y[c]));
            }
            if (OP == MIX_LOP_LDS) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                if ((c & 3) == 0) y[c] = lut[y[c] & 1023] + (y[c] >> 3);
            }
            if (OP == FUNNEL) { uint64_t w = ((uint64_t)y[c] << 32) | x[c]; w >>= 7; x[c] = (uint32_t)w; y[c] = (uint32_t)(w >> 32) | k1; }
            if (OP == MIX3) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[c]) : "r"(k1), "r"(k2));
                if ((c & 3) == 0) y[c] = lut[y[c] & 1023] + y[c];
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cycles[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

template <int OP>
void run(int sms, double instr_per_chain_iter) {
    const int ctas_per_sm = 4, threads = 256;
    const int grid = sms * ctas_per_sm;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, grid * threads * 4);
    cudaMalloc(&cyc, grid * threads / 32 * 8);
    bench<OP><<<grid, threads>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    bench<OP><<<grid, threads>>>(out, cyc, 12345u);
    cudaError_t e = cudaDeviceSynchronize();
    const int nw = grid * threads / 32;
    long long* h = (long long*)malloc(nw * 8);
    cudaMemcpy(h, cyc, nw * 8, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int i = 0; i < nw; ++i) sum += (double)h[i];
    const double avg = sum / nw;
    const double warps_per_sm = ctas_per_sm * threads / 32.0;
    const double winst = (double)ITER * CHAINS * instr_per_chain_iter * warps_per_sm;
    printf("%-22s %8.3f warp-instr/clk/SM  (%6.1f lanes/clk/SM)  avg %.0f cycles  %s\n", NAMES[OP], winst / avg,
           32.0 * winst / avg, avg, e == cudaSuccess ? "" : cudaGetErrorString(e));
    free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, instruction counts are PTX-level (see SASS for the exact mix)\n", p.name, p.multiProcessorCount);
    const int s = p.multiProcessorCount;
    run<LOP3>(s, 1); run<SHF_C>(s, 1); run<SHF_V>(s, 1); run<IADD>(s, 1); run<IMAD>(s, 1); run<IMADHI>(s, 1);
    run<IMADWIDE>(s, 1); run<POPC>(s, 1); run<SEL>(s, 2); run<PRMT>(s, 1); run<LDS8>(s, 1); run<FLO>(s, 1);
    run<MIX_LOP_IMAD>(s, 2); run<MIX_LOP_IMADHI>(s, 2); run<MIX_SHF_IMAD>(s, 2); run<MIX_LOP_POPC>(s, 1.25);
    run<MIX_LOP_LDS>(s, 1.25); run<FUNNEL>(s, 1); run<MIX3>(s, 2.25);
    return 0;
}
