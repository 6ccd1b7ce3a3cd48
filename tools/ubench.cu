// Instruction-throughput microbenchmark for the integer pipes that bound the field multiply.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE> __global__ void k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a[CHAINS], b[CHAINS];
    uint64_t w[CHAINS];
    double d[CHAINS];
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i; w[i] = a[i]; d[i] = 1.0 + i; }
    uint32_t x = seed | 1, y = seed * 5 + 3;
    double dx = 1.0000001, dy = 0.5;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y));
            if (MODE == 3) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(x), "r"(y));
            if (MODE == 4) asm volatile("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(x), "r"(y));
            if (MODE == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));
            if (MODE == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x)); }
            if (MODE == 7) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(x), "r"(y)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x)); }
            if (MODE == 8) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dx), "d"(dy));
            if (MODE == 9) { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dx), "d"(dy)); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(x), "r"(y)); }
            if (MODE == 10) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(y)); }
            if (MODE == 11) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(x));
            if (MODE == 12) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 13) { asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dx), "d"(dy)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y)); }
            if (MODE == 14) { asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dx), "d"(dy)); asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(x), "r"(y)); }
            if (MODE == 15) { asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dx), "d"(dy)); asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dy), "d"(dx)); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y)); asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(x), "r"(y)); }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < CHAINS; i++) acc += a[i] + b[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int ops_per_iter) {
    int blocks = 148 * 4, threads = 256;
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 12345, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 4]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double warp_instr_per_sm = 4.0 * (threads / 32) * (double)ITERS * CHAINS * ops_per_iter;   // 4 blocks per SM
    double total_ops = (double)blocks * threads * ITERS * CHAINS * ops_per_iter;
    printf("%-34s %8.3f ms  cycles/block %10.0f  warp-instr/clk/SM %6.3f  lane-ops/s %8.2f T  eff.clk %5.0f MHz\n", name, ms, avg,
           warp_instr_per_sm / avg, total_ops / ms / 1e9, avg / ms / 1e3);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IMAD (mad.lo)", 1);
    run<1>("IMAD.HI (mad.hi)", 1);
    run<2>("IMAD.WIDE (mad.wide)", 1);
    run<3>("mad.lo.cc+madc.hi.cc (count 1)", 1);
    run<4>("add.cc+addc.cc (count 2)", 2);
    run<5>("IADD (add)", 1);
    run<6>("IMAD.WIDE + IADD (count 2)", 2);
    run<7>("IMAD.lo + IADD (count 2)", 2);
    run<8>("DFMA", 1);
    run<9>("DFMA + IMAD.lo (count 2)", 2);
    run<10>("IMAD.WIDE + 2 IADD (count 3)", 3);
    run<11>("SHF", 1);
    run<12>("LOP3", 1);
    run<13>("DFMA + IMAD.WIDE (count 2)", 2);
    run<14>("DFMA + add.cc/addc (count 3)", 3);
    run<15>("2 DFMA + IMAD.WIDE + 2 IADD (5)", 5);
    return 0;
}
