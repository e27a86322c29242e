// Micro-probe: the arithmetic of one CTC walker frame (8 states per lane: 8 DADD, 4 DFMA, 8 DMUL with the real
// dependency pattern) in a single warp, with and without the halo shuffle, to find the latency floor of the chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/walk_probe tools/walk_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double lds_v(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ double lds_v2(unsigned addr) {      // volatile, but no memory clobber
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}

// 0: arithmetic only, 1: + halo shuffle, 2: + 5 shared loads of the next frame's probabilities (plain C++),
// 3: loads two frames ahead (plain C++), 4: volatile-asm loads one frame ahead, 5: volatile-asm loads two frames ahead,
// 6: as 2 without the shuffle
template <int MODE>
__global__ void probe(double* out, long long* cyc, int frames, const double* ptab, double2* lat, int* ex) {
    __shared__ double tile[64 * 32];
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) tile[i] = ptab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double a[8], sk[4], p[5], pn[5], pn2[5];
    const unsigned tbase = (unsigned)__cvta_generic_to_shared(tile);
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-3 * (lane * 8 + j);
    for (int i = 0; i < 4; ++i) sk[i] = (lane + i) & 1 ? 1.0 : 0.0;
    for (int i = 0; i < 5; ++i) pn[i] = pn2[i] = 0.2501 + 1e-4 * i;
    double h0 = 0.5;
    int row = 0, mxp = 0;
    double2* latp = lat;
    double bcp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int f = 0; f < frames; ++f) {
#pragma unroll
        for (int i = 0; i < 5; ++i) p[i] = pn[i];
        if (MODE == 2 || MODE == 6) {
            row = (row + 1) & 63;
#pragma unroll
            for (int i = 0; i < 5; ++i) pn[i] = tile[row * 32 + ((lane * 5 + i * 7) & 31)];
        }
        if (MODE == 17 || MODE == 18) {
            row = (row + 1) & 63;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                pn[i] = pn2[i];
                pn2[i] = lds_v2(tbase + (unsigned)((row * 32 + ((lane * 5 + i * 7) & 31)) * 8));
            }
        } else if (MODE == 3 || MODE == 5 || MODE >= 7) {   // (modes 7+ build on mode 5)
            row = (row + 1) & 63;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                pn[i] = pn2[i];
                pn2[i] = MODE == 3 ? tile[row * 32 + ((lane * 5 + i * 7) & 31)]
                                   : lds_v(tbase + (unsigned)((row * 32 + ((lane * 5 + i * 7) & 31)) * 8));
            }
        }
        if (MODE == 4) {
            row = (row + 1) & 63;
#pragma unroll
            for (int i = 0; i < 5; ++i) pn[i] = lds_v(tbase + (unsigned)((row * 32 + ((lane * 5 + i * 7) & 31)) * 8));
        }
        // presum, in place, descending (alpha)
#pragma unroll
        for (int j = 7; j >= 2; --j) {
            if (j & 1) a[j] = fma(sk[j >> 1], a[j - 2], a[j] + a[j - 1]);
            else a[j] = a[j] + a[j - 1];
        }
        a[1] = fma(sk[0], h0, a[1] + a[0]);
        a[0] = a[0] + h0;
        if (MODE == 16 || MODE == 18) {                   // store the PREVIOUS frame's sums from a copy, keep a copy of this frame's
            double2* q = latp + lane;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) q[jj * 32] = make_double2(bcp[2 * jj], bcp[2 * jj + 1]);
            latp += 128;
            if ((f & 255) == 255) latp = lat;
#pragma unroll
            for (int j = 0; j < 8; ++j) bcp[j] = a[j];
        }
        if (MODE >= 12 && MODE <= 15) {                   // store variants on top of mode 5
            double2* lp = lat + (size_t)(f & 255) * 128 + lane;
            if (MODE == 12) {                             // 4 STG.128, no exponent
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) lp[jj * 32] = make_double2(a[2 * jj], a[2 * jj + 1]);
            } else if (MODE == 13) {                      // one STG.128 only
                lp[0] = make_double2(a[0], a[1]);
            } else if (MODE == 14) {                      // 4 STS.128 to shared memory instead
                double2* sp = reinterpret_cast<double2*>(tile) + ((f & 3) * 128) + lane;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) sp[jj * 32] = make_double2(a[2 * jj], a[2 * jj + 1]);
            } else {                                      // running pointer instead of index arithmetic
                double2* q = latp + lane;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) q[jj * 32] = make_double2(a[2 * jj], a[2 * jj + 1]);
                latp += 128;
                if ((f & 255) == 255) latp = lat;
            }
        }
        if (MODE == 7 || MODE == 9) {                     // lattice row: 4 x 16-byte stores per lane, one exponent per frame
            double2* lp = lat + (size_t)(f & 255) * 128 + lane;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) lp[jj * 32] = make_double2(a[2 * jj], a[2 * jj + 1]);
            if (lane == 0) ex[f & 255] = f;
        }
        if (MODE == 8 || MODE == 9) {                     // the rescale test of every 8th frame
            if ((f & 7) == 7) {
                int mx = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) mx = max(mx, __double2hiint(a[j]));
                mx = __reduce_max_sync(0xffffffffu, mx);
                if (mx >= 0x00100000) {
                    const int e = (mx >> 20) - 1023;
                    const double sc = __hiloint2double((1023 - e) << 20, 0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] *= sc;
                    h0 *= sc;
                }
            }
        }
        if (MODE == 10 || MODE == 11) {                   // pipelined rescale: measure at f % 8 == 3, apply at f % 8 == 7
            if ((f & 7) == 3) {
                int mx = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) mx = max(mx, __double2hiint(a[j]));
                mxp = __reduce_max_sync(0xffffffffu, mx);
            }
            if ((f & 7) == 7 && mxp >= 0x00100000) {
                const int e = (mxp >> 20) - 1023;
                const double sc = __hiloint2double((1023 - e) << 20, 0);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] *= sc;
                h0 *= sc;
            }
        }
        if (MODE == 11) {                                 // stores without the per-frame exponent
            double2* lp = lat + (size_t)(f & 255) * 128 + lane;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) lp[jj * 32] = make_double2(a[2 * jj], a[2 * jj + 1]);
            if ((f & 7) == 7 && lane == 0) ex[(f >> 3) & 255] = f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] *= (j & 1) ? p[1 + (j >> 1)] : p[0];
        if (MODE >= 1 && MODE != 6) {   // (modes 7-9 keep the shuffle)
            h0 = __shfl_up_sync(0xffffffffu, a[7], 1);
            h0 = lane == 0 ? 0.0 : h0;
        } else {
            h0 = a[7] * 0.5;
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += a[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
    double* out; long long* cyc; double* ptab; long long h; double2* lat; int* ex;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8); cudaMalloc(&ptab, 64 * 32 * 8);
    cudaMalloc(&lat, 256 * 128 * 16 * 8); cudaMalloc(&ex, 4096);
    double hp[64 * 32];
    for (int i = 0; i < 64 * 32; ++i) hp[i] = 0.25 + 1e-5 * (i % 97);
    cudaMemcpy(ptab, hp, sizeof(hp), cudaMemcpyHostToDevice);
    const int frames = 4000;
    probe<MODE><<<1, 32 * warps>>>(out, cyc, frames, ptab, lat, ex);
    cudaDeviceSynchronize();
    probe<MODE><<<1, 32 * warps>>>(out, cyc, frames, ptab, lat, ex);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps=%d : %7.1f cycles per frame\n", name, warps, (double)h / frames);
    cudaFree(out); cudaFree(cyc); cudaFree(ptab);
}

int main() {
    run<0>("arithmetic only (8 DADD, 4 DFMA, 8 DMUL)", 1);
    run<1>("+ halo shuffle", 1);
    run<2>("+ halo shuffle + 5 LDS (prefetched one frame)", 1);
    run<6>("5 LDS one frame ahead, no shuffle", 1);
    run<3>("shuffle + 5 LDS two frames ahead (C++)", 1);
    run<4>("shuffle + 5 volatile LDS one frame ahead", 1);
    run<5>("shuffle + 5 volatile LDS two frames ahead", 1);
    run<7>("mode 5 + lattice stores (4 STG.128 + exponent)", 1);
    run<8>("mode 5 + rescale test every 8 frames", 1);
    run<9>("mode 5 + stores + rescale", 1);
    run<12>("mode 5 + 4 STG.128 (no exponent store)", 1);
    run<13>("mode 5 + 1 STG.128", 1);
    run<14>("mode 5 + 4 STS.128 (shared memory)", 1);
    run<15>("mode 5 + 4 STG.128 through a running pointer", 1);
    run<16>("mode 5 + 4 STG.128 of the previous frame from copies", 1);
    run<17>("loads volatile without memory clobber, no stores", 1);
    run<18>("same + delayed stores from copies", 1);
    return 0;
}
