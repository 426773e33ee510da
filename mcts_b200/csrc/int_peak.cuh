// int_peak.cuh -- integer-issue micro-benchmarks: the roofline denominator for the playout kernels.
//
// The playout path is bound by integer instruction issue (LOP3 / SHF / IADD3 on the ALU pipe, IMAD on
// the FMA pipe, POPC on its own slower pipe), not by HBM or tensor cores.  Each kernel below runs 8
// independent dependency chains per thread so that the measured rate is the PIPE's throughput, not its
// latency; inline PTX keeps the compiler from folding the chains.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mr {

constexpr int INT_PEAK_KINDS = 10;
constexpr int INT_PEAK_CHAINS = 8;
constexpr int INT_PEAK_UNROLL = 8;
constexpr int INT_PEAK_OPS_PER_ITER = INT_PEAK_CHAINS * INT_PEAK_UNROLL;

template <int KIND>
__device__ __forceinline__ void int_op(uint32_t &x, uint32_t a, uint32_t b, int u) {
    if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    if (KIND == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (KIND == 2) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (KIND == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (KIND == 4) asm volatile("popc.b32 %0, %0;" : "+r"(x));
    if (KIND == 6) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (KIND == 7) { // 64-bit product of two 32-bit values (IMAD.WIDE), folded back to 32 bits with one xor
        unsigned long long w;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(a));
        x = (uint32_t)w ^ (uint32_t)(w >> 32);
    }
    if (KIND == 8) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(a)); // funnel shift by an immediate
    if (KIND == 9) { // compare + select pair (ISETP + SEL)
        asm volatile("{ .reg .pred q; setp.lt.u32 q, %0, %1; selp.u32 %0, %2, %0, q; }" : "+r"(x) : "r"(a), "r"(b));
    }
    if (KIND == 5) { // alternate ALU-pipe and FMA-pipe instructions
        if (u & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
        else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
    }
}

template <int KIND>
__global__ void __launch_bounds__(1024) int_peak_kernel(int iters, uint32_t *sink) {
    uint32_t x[INT_PEAK_CHAINS];
    const uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 7u;
#pragma unroll
    for (int c = 0; c < INT_PEAK_CHAINS; ++c) x[c] = a + c * 977u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < INT_PEAK_UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < INT_PEAK_CHAINS; ++c) int_op<KIND>(x[c], a | 3u, b | 5u, u);
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < INT_PEAK_CHAINS; ++c) acc ^= x[c];
    if (acc == 0x9e3779b9u) sink[threadIdx.x & 1023] = acc; // practically never; defeats dead-code elimination
}

inline void launch_int_peak(int kind, int blocks, int threads, int iters, uint32_t *sink, cudaStream_t s) {
    switch (kind) {
    case 0: int_peak_kernel<0><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 1: int_peak_kernel<1><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 2: int_peak_kernel<2><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 3: int_peak_kernel<3><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 4: int_peak_kernel<4><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 5: int_peak_kernel<5><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 6: int_peak_kernel<6><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 7: int_peak_kernel<7><<<blocks, threads, 0, s>>>(iters, sink); break;
    case 8: int_peak_kernel<8><<<blocks, threads, 0, s>>>(iters, sink); break;
    default: int_peak_kernel<9><<<blocks, threads, 0, s>>>(iters, sink); break;
    }
}

} // namespace mr
