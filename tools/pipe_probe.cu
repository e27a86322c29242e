// Micro-probe: latency (1 dependent chain) and throughput (8 chains x many warps) of the pipes the CTC
// lattice can run on: fp64 add/mul/fma, fp32 fma, MUFU ex2/lg2.  Build: nvcc -arch=sm_100a -O3 -o pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int CH>
__global__ void probe(double* out, long long* cyc, int iters, double seed) {
    double a[CH];
    float f[CH];
    unsigned u[CH];
    for (int c = 0; c < CH; ++c) { a[c] = seed + c * 1e-3 + threadIdx.x * 1e-6; f[c] = (float)a[c]; u[c] = (unsigned)(c * 977 + threadIdx.x); }
    double m = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == 0) a[c] = a[c] + b;
            if (OP == 1) a[c] = a[c] * m;
            if (OP == 2) a[c] = fma(a[c], m, b);
            if (OP == 3) f[c] = fmaf(f[c], 1.0000001f, 1e-9f);
            if (OP == 4) f[c] = exp2f(f[c]) * 0.25f;      // MUFU.EX2 + FMUL
            if (OP == 5) f[c] = __log2f(f[c]) + 3.0f;     // MUFU.LG2 + FADD
            if (OP == 6) f[c] = __fadd_rn(__fmul_rn(f[c], 1.0000001f), 1e-9f);   // FMUL + FADD, never fused (the sampler's arithmetic)
            if (OP == 7) { u[c] = (u[c] ^ 0x9E3779B9u) + (u[c] >> 7); }          // LOP3 + SHF + IADD3 (the edit distance's arithmetic)
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += a[c] + f[c] + (double)u[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int CH>
void run(const char* name, int threads, int blocks) {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaMalloc(&cyc, 8);
    int iters = 4096;
    probe<OP, CH><<<blocks, threads>>>(out, cyc, iters, 1.5);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP, CH><<<blocks, threads>>>(out, cyc, iters, 1.5);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)iters * CH * threads * blocks;
    printf("%-8s chains=%d threads=%4d blocks=%4d : %8.2f cyc/iter(block0,warp0)  %8.3f cyc per op-per-chain  %10.2f Gop/s  (%.3f ms)\n",
           name, CH, threads, blocks, (double)h / iters, (double)h / iters / CH, ops / ms / 1e6, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s sm_%d%d, %d SMs, clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    // latency: 1 warp, 1 chain
    run<0, 1>("dadd", 32, 1); run<1, 1>("dmul", 32, 1); run<2, 1>("dfma", 32, 1);
    run<3, 1>("ffma", 32, 1); run<4, 1>("ex2", 32, 1); run<5, 1>("lg2", 32, 1);
    // one warp, 8 independent chains (what one CTC warp sees)
    run<0, 8>("dadd", 32, 1); run<2, 8>("dfma", 32, 1); run<3, 8>("ffma", 32, 1); run<4, 8>("ex2", 32, 1);
    // 4 warps per SM (one per scheduler)
    run<2, 8>("dfma", 128, 148); run<4, 8>("ex2", 128, 148);
    // saturate: 32 warps per SM
    run<0, 8>("dadd", 1024, 148); run<2, 8>("dfma", 1024, 148); run<3, 8>("ffma", 1024, 148);
    run<4, 8>("ex2", 1024, 148); run<5, 8>("lg2", 1024, 148);
    // what one scheduler partition issues: 1, 4 and 8 warps per partition, 8 independent chains per thread
    run<3, 8>("ffma", 128, 148); run<3, 8>("ffma", 512, 148);
    run<6, 8>("fmul+fadd", 128, 148); run<6, 8>("fmul+fadd", 512, 148); run<6, 8>("fmul+fadd", 1024, 148);
    run<7, 8>("lop+shf+iadd", 128, 148); run<7, 8>("lop+shf+iadd", 512, 148); run<7, 8>("lop+shf+iadd", 1024, 148);
    run<2, 8>("dfma", 512, 148);
    return 0;
}
