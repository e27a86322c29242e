// K5: CTC alpha-beta loss and gradient (no upstream code; spec = DESIGN.md "CTC spec").
//
// Design (B200): one CTA per utterance, two warps.  Warp 0 runs the alpha recurrence forward in time, warp 1
// runs beta backward, at the same time, and they meet in the middle: each stores the half of its lattice the
// other one needs (pre-emission sums, fp64, [frame][j][lane] so every store/load is a coalesced 256 B line),
// one __syncthreads, then each warp finishes its pass and emits the gradient rows of the frames it now owns.
// The serial depth is T frame steps instead of 2T, and nothing but the 2 halo values crosses lanes per frame
// (warp shuffles; every lane owns SPL consecutive label states in registers).
//
// Arithmetic: the lattice is kept in LINEAR space in fp64 with exact power-of-two rescaling (the exponent is
// read off the warp maximum with one redux.sync every 4 frames).  On B200 the fp64 pipe issues an add or a
// multiply at half the fp32 rate, whereas a log-space recurrence needs 3 ex2 + 1 lg2 per state per frame on
// the quarter-rate MUFU pipe and loses ~1e-4 relative in fp32 at |log alpha| ~ 1e3.  The occupancies
// gamma_t(s) = alpha_t(s) beta'_t(s) / Z_t are normalised per frame by Z_t = sum_s alpha beta', so no scale
// bookkeeping crosses the two directions; nll = -(log(alpha_T(S-1)+alpha_T(S-2)) + E ln 2).
#include "pgasr_common.cuh"

namespace pgasr {

constexpr int kCtcChunk = 32;   // frames staged per cp.async batch

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ int hi32(double x) { return __double2hiint(x); }

// In-kernel softmax for callers that do not hand in probabilities: thread per frame.
__device__ void softmax_rows(const float* __restrict__ logits, float* __restrict__ probs, int Tb, int V) {
    for (int t = threadIdx.x; t < Tb; t += blockDim.x) {
        const float* z = logits + (size_t)t * V;
        float m = -INFINITY;
        for (int v = 0; v < V; ++v) m = fmaxf(m, z[v]);
        float s = 0.0f;
        for (int v = 0; v < V; ++v) s += __expf(z[v] - m);
        const float inv = 1.0f / s;
        for (int v = 0; v < V; ++v) probs[(size_t)t * V + v] = __expf(z[v] - m) * inv;
    }
}

template <int SPL>
struct CtcLane {
    double a[SPL];          // alpha-hat or beta-hat of states lane*SPL + j (including the emission)
    int lab[SPL / 2];       // class of the odd (label) states, -1 beyond the transcript
    unsigned skip;          // bit j (odd j): the two-state transition into (alpha) / out of (beta) state j is legal
    unsigned evalid;        // bit j (even j): blank state j exists (s < S)
};

// One direction of the lattice.  kAlpha: t runs 0..Tb-1, else Tb-1..0.
template <int SPL, bool kAlpha>
__device__ void ctc_direction(const float* __restrict__ probs_u,   // [T,V] of this utterance
                              const int32_t* __restrict__ lab_u, int Tb, int L, int V, int blank,
                              float grad_scale, int accumulate, float* __restrict__ nll_out,
                              float* __restrict__ dlog_u, double* __restrict__ lat_u,
                              float* stage, float* racc) {
    const int lane = threadIdx.x & 31;
    const int S = 2 * L + 1;
    const int tm = Tb / 2;
    CtcLane<SPL> st;
    st.skip = 0u;
    st.evalid = 0u;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int s = lane * SPL + j;
        st.a[j] = 0.0;
        if (j & 1) {
            const int li = (s - 1) >> 1;
            const int c = li < L ? lab_u[li] : -1;
            st.lab[j >> 1] = c;
            if (kAlpha) {
                if (c >= 0 && li >= 1 && lab_u[li - 1] != c) st.skip |= 1u << j;
            } else {
                if (c >= 0 && li + 1 < L && lab_u[li + 1] != c) st.skip |= 1u << j;
            }
        } else if (s < S) {
            st.evalid |= 1u << j;
        }
    }
    int E = 0;                                  // true value = hat value * 2^E

    // frames are visited in "steps" 0..Tb-1; frame index t = kAlpha ? step : Tb-1-step.
    // first half: steps with a frame on this direction's side of tm (store pre-emission sums);
    // second half: the rest (load the other direction's sums, emit gradient rows).
    const int n_first = kAlpha ? tm : Tb - tm;
    for (int half = 0; half < 2; ++half) {
        const int step_lo = half == 0 ? 0 : n_first;
        const int step_hi = half == 0 ? n_first : Tb;
        if (half == 1) __syncthreads();         // the other warp's half-lattice is now visible
        int staged_lo = 0, staged_hi = 0;       // steps currently in the stage buffer
        for (int step = step_lo; step < step_hi; ++step) {
            if (step >= staged_hi) {
                // stage the probability rows of the next <= kCtcChunk steps (contiguous frames)
                staged_lo = step;
                staged_hi = min(step + kCtcChunk, step_hi);
                const int nfr = staged_hi - staged_lo;
                const int f0 = kAlpha ? staged_lo : Tb - staged_hi;   // lowest frame of the chunk
                const float* src = probs_u + (size_t)f0 * V;
                __syncwarp();
                for (int i = lane; i < nfr * V; i += 32) cp_async4(stage + i, src + i);
                cp_async_commit();
                cp_async_wait<0>();
                __syncwarp();
            }
            const int t = kAlpha ? step : Tb - 1 - step;
            const int f0 = kAlpha ? staged_lo : Tb - staged_hi;
            const float* row = stage + (size_t)(t - f0) * V;

            // ---- pre-emission sums into st.a (in place) -------------------------------------
            if (step == 0) {
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    const int s = lane * SPL + j;
                    const bool on = kAlpha ? (s <= 1 && s < S) : (s < S && s >= S - 2);
                    st.a[j] = on ? 1.0 : 0.0;
                }
            } else if (kAlpha) {
                double h = __shfl_up_sync(kFull, st.a[SPL - 1], 1);
                if (lane == 0) h = 0.0;
#pragma unroll
                for (int j = SPL - 1; j >= 2; --j) {
                    double x = st.a[j] + st.a[j - 1];
                    if ((j & 1) && (st.skip >> j & 1u)) x += st.a[j - 2];
                    st.a[j] = x;
                }
                {
                    double x = st.a[1] + st.a[0];
                    if (st.skip >> 1 & 1u) x += h;
                    st.a[1] = x;
                    st.a[0] = st.a[0] + h;
                }
            } else {
                double h0 = __shfl_down_sync(kFull, st.a[0], 1);
                double h1 = __shfl_down_sync(kFull, st.a[1], 1);
                if (lane == 31) { h0 = 0.0; h1 = 0.0; }
#pragma unroll
                for (int j = 0; j < SPL - 2; ++j) {
                    double x = st.a[j] + st.a[j + 1];
                    if ((j & 1) && (st.skip >> j & 1u)) x += st.a[j + 2];
                    st.a[j] = x;
                }
                st.a[SPL - 2] = st.a[SPL - 2] + st.a[SPL - 1];
                {
                    double x = st.a[SPL - 1] + h0;
                    if (st.skip >> (SPL - 1) & 1u) x += h1;
                    st.a[SPL - 1] = x;
                }
            }

            double* lat_t = lat_u + (size_t)t * (SPL * 32) + lane;
            double other[SPL];
            if (half == 0) {
#pragma unroll
                for (int j = 0; j < SPL; ++j) lat_t[j * 32] = st.a[j];
            } else {
#pragma unroll
                for (int j = 0; j < SPL; ++j) other[j] = lat_t[j * 32];
            }

            // ---- emission ---------------------------------------------------------------------
            const double pb = (double)row[blank];
#pragma unroll
            for (int j = 0; j < SPL; ++j) {
                double p;
                if (j & 1) {
                    const int c = st.lab[j >> 1];
                    p = c >= 0 ? (double)row[c] : 0.0;
                } else {
                    p = (st.evalid >> j & 1u) ? pb : 0.0;
                }
                st.a[j] *= p;
            }

            // ---- gradient row of frame t --------------------------------------------------------
            if (half == 1) {
                double zl = 0.0, zb = 0.0;
                double w[SPL];
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    w[j] = st.a[j] * other[j];
                    zl += w[j];
                    if (!(j & 1)) zb += w[j];
                }
                const double Z = warp_sum(zl);
                float* out = dlog_u + (size_t)t * V;
                if (Z > 0.0) {
                    const double inv = 1.0 / Z;
                    const float gb = warp_sum((float)(zb * inv));
                    for (int v = lane; v < V; v += 32) racc[v] = 0.0f;
                    __syncwarp();
#pragma unroll
                    for (int j = 1; j < SPL; j += 2) {
                        const int c = st.lab[j >> 1];
                        if (c >= 0) atomicAdd(&racc[c], (float)(w[j] * inv));
                    }
                    __syncwarp();
                    for (int v = lane; v < V; v += 32) {
                        const float occ = v == blank ? gb : racc[v];
                        const float g = grad_scale * (row[v] - occ);
                        out[v] = accumulate ? out[v] + g : g;
                    }
                    __syncwarp();
                } else if (!accumulate) {
                    for (int v = lane; v < V; v += 32) out[v] = 0.0f;
                }
            }

            // ---- exact power-of-two rescale every 4 steps ---------------------------------------
            if ((step & 3) == 3) {
                int mx = 0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) mx = max(mx, hi32(st.a[j]));
                mx = __reduce_max_sync(kFull, mx);
                if (mx >= 0x00100000) {
                    const int e = (mx >> 20) - 1023;
                    const double sc = __hiloint2double((1023 - e) << 20, 0);
                    E += e;
#pragma unroll
                    for (int j = 0; j < SPL; ++j) st.a[j] *= sc;
                }
            }
        }
    }

    if (kAlpha) {
        // nll = -(log(alpha_hat(S-1) + alpha_hat(S-2)) + E ln2)
        double fin = 0.0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const int s = lane * SPL + j;
            if (s < S && s >= S - 2) fin += st.a[j];
        }
        fin = warp_sum(fin);
        if (lane == 0)
            *nll_out = fin > 0.0 ? (float)(-(log(fin) + (double)E * 0.69314718055994530942)) : INFINITY;
    }
}

template <int SPL>
__global__ void __launch_bounds__(64)
ctc_kernel(const float* __restrict__ logits, const float* __restrict__ probs_in,
           float* __restrict__ probs_ws, const int32_t* __restrict__ targets,
           const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len, int T, int V,
           int Lmax, int blank, float grad_scale, int accumulate, float* __restrict__ nll,
           float* __restrict__ dlogits, double* __restrict__ lattice) {
    extern __shared__ float smem[];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int Tb = in_len ? in_len[b] : T;
    Tb = min(max(Tb, 0), T);
    int L = tgt_len ? tgt_len[b] : Lmax;
    L = min(max(L, 0), Lmax);
    float* dlog_u = dlogits + (size_t)b * T * V;
    // rows beyond the utterance carry no gradient
    if (!accumulate)
        for (int i = Tb * V + threadIdx.x; i < T * V; i += blockDim.x) dlog_u[i] = 0.0f;
    if (Tb == 0) {
        if (threadIdx.x == 0) nll[b] = L == 0 ? 0.0f : INFINITY;
        return;
    }
    const float* probs_u;
    if (probs_in) {
        probs_u = probs_in + (size_t)b * T * V;
    } else {
        float* pw = probs_ws + (size_t)b * T * V;
        softmax_rows(logits + (size_t)b * T * V, pw, Tb, V);
        __syncthreads();
        probs_u = pw;
    }
    float* stage = smem + (size_t)warp * (kCtcChunk * V + V);
    float* racc = stage + kCtcChunk * V;
    double* lat_u = lattice + (size_t)b * T * (SPL * 32);
    const int32_t* lab_u = targets + (size_t)b * Lmax;
    (void)lane;
    if (warp == 0)
        ctc_direction<SPL, true>(probs_u, lab_u, Tb, L, V, blank, grad_scale, accumulate, nll + b, dlog_u,
                                 lat_u, stage, racc);
    else
        ctc_direction<SPL, false>(probs_u, lab_u, Tb, L, V, blank, grad_scale, accumulate, nll + b, dlog_u,
                                  lat_u, stage, racc);
}

static int ctc_spl(int Lmax) {
    const int S = 2 * Lmax + 1;
    if (S <= 4 * 32) return 4;
    if (S <= 8 * 32) return 8;
    if (S <= 16 * 32) return 16;
    if (S <= 32 * 32) return 32;
    return 0;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace pgasr

extern "C" size_t pgasr_ctc_workspace_bytes(int B, int T, int V, int Lmax) {
    using namespace pgasr;
    const int spl = ctc_spl(Lmax);
    if (spl == 0 || B < 0 || T <= 0 || V <= 0) return 0;
    size_t lat = align_up((size_t)B * T * spl * 32 * sizeof(double), 256);
    size_t pr = align_up((size_t)B * T * V * sizeof(float), 256);
    return lat + pr;
}

extern "C" int pgasr_ctc_loss_grad(const float* logits, const float* probs, const int32_t* targets,
                                   const int32_t* in_len, const int32_t* tgt_len, int B, int T, int V,
                                   int Lmax, int blank, float grad_scale, int accumulate, float* nll,
                                   float* dlogits, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace pgasr;
    if ((!logits && !probs) || !targets || !nll || !dlogits || !workspace || B < 0 || T <= 0 || V <= 0 ||
        Lmax <= 0 || blank < 0 || blank >= V)
        return PGASR_ERR_INVALID_ARG;
    const int spl = ctc_spl(Lmax);
    if (spl == 0) return PGASR_ERR_UNSUPPORTED;
    if (workspace_bytes < pgasr_ctc_workspace_bytes(B, T, V, Lmax)) return PGASR_ERR_WORKSPACE;
    if (B == 0) return PGASR_OK;
    double* lattice = reinterpret_cast<double*>(workspace);
    float* probs_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) +
                                               align_up((size_t)B * T * spl * 32 * sizeof(double), 256));
    const size_t smem = 2 * ((size_t)kCtcChunk * V + V) * sizeof(float);
    cudaStream_t st = as_stream(stream);
#define PGASR_CTC(SPLv)                                                                                   \
    do {                                                                                                  \
        if (smem > 48 * 1024)                                                                             \
            PGASR_CUDA_TRY(cudaFuncSetAttribute(ctc_kernel<SPLv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)smem));                                              \
        ctc_kernel<SPLv><<<B, 64, smem, st>>>(logits, probs, probs_ws, targets, in_len, tgt_len, T, V, Lmax, \
                                              blank, grad_scale, accumulate, nll, dlogits, lattice);      \
    } while (0)
    switch (spl) {
        case 4: PGASR_CTC(4); break;
        case 8: PGASR_CTC(8); break;
        case 16: PGASR_CTC(16); break;
        default: PGASR_CTC(32); break;
    }
#undef PGASR_CTC
    PGASR_LAUNCH_CHECK();
    return PGASR_OK;
}
