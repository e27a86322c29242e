// K5: CTC alpha-beta loss and gradient (no upstream code; spec = DESIGN.md "CTC spec").
//
// Design (B200): one CTA per utterance.  Warp 0 runs the alpha recurrence forward in time, warp 1 runs beta
// backward, at the same time, and they meet in the middle: each stores the half of its lattice the other one
// needs (pre-emission sums, fp64, [frame][j][lane] so every store/load is a coalesced 256 B line), one
// __syncthreads, then each warp finishes its pass and emits the gradient rows of the frames it now owns.
// The serial depth is T frame steps instead of 2T, and only the 1-2 halo values cross lanes per frame
// (warp shuffles; every lane owns SPL consecutive label states in registers).
//
// Arithmetic: the lattice is kept in LINEAR space in fp64 with exact power-of-two rescaling (the exponent is
// read off the warp maximum with one redux.sync every 4 frames).  Measured on B200 (tools/pipe_probe.cu): the
// fp64 pipe sustains 17.3 T fma/s against 4.6 T ex2/s on the MUFU pipe, latency 8.5 vs ~40 cycles, and a
// log-space recurrence needs 3 ex2 + 1 lg2 per state per frame and loses ~1e-4 relative in fp32 at
// |log alpha| ~ 1e3.  The frame loop is branch free: illegal two-state transitions are a multiply by 0.0,
// label states beyond the transcript read a zero slot of the probability row.
//
// Gradient: occupancies gamma_t(s) = alpha_t(s) beta'_t(s) / P.  P is measured once, at the first frame of a
// warp's second half (Z0 = sum_s alpha beta'); later frames reuse it through the recorded power-of-two exponents,
// so no per-frame reduction over states is needed.  gamma is converted to 2^-30 fixed point by one fma with the
// 2^52+2^51 magic constant, the blank column is summed with one integer redux.sync, label columns with integer
// shared-memory atomics (exact, order independent => deterministic).
#include "pgasr_common.cuh"

namespace pgasr {

constexpr int kCtcChunk = 32;     // frames staged per cp.async batch
constexpr int kCtcThreads = 128;  // warp 0: alpha, warp 1: beta, all four: softmax rows

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__host__ __device__ inline int ctc_row_stride(int V) { return (V + 2) & ~1; }   // doubles per probability row

// softmax rows of one utterance into the fp64 workspace [T][RS]; slots V..RS-1 are zero (slot V is what
// label states beyond the transcript read).  One thread per frame.
__device__ void softmax_rows_f64(const float* __restrict__ logits_u, const float* __restrict__ probs_in_u,
                                 double* __restrict__ pw, int Tb, int V, int RS) {
    for (int t = threadIdx.x; t < Tb; t += blockDim.x) {
        double* o = pw + (size_t)t * RS;
        if (probs_in_u) {
            const float* p = probs_in_u + (size_t)t * V;
            for (int v = 0; v < V; ++v) o[v] = (double)p[v];
        } else {
            const float* z = logits_u + (size_t)t * V;
            float m = -INFINITY;
            for (int v = 0; v < V; ++v) m = fmaxf(m, z[v]);
            float s = 0.0f;
            for (int v = 0; v < V; ++v) s += __expf(z[v] - m);
            const float inv = 1.0f / s;
            for (int v = 0; v < V; ++v) o[v] = (double)(__expf(z[v] - m) * inv);
        }
        for (int v = V; v < RS; ++v) o[v] = 0.0;
    }
}

template <int SPL>
struct CtcLane {
    double a[SPL];            // alpha-hat / beta-hat of states lane*SPL + j (after the emission)
    double skipm[SPL / 2];    // odd state 2i+1: 1.0 if its two-state transition is legal, else 0.0
    int loff[SPL / 2];        // odd state 2i+1: slot of its class in a probability row (V = the zero slot)
    int E;                    // true value = hat value * 2^E
};

struct GradNorm {
    double invZ0;             // 1 / Z0, Z0 = sum_s alpha beta' at the first gradient frame
    int E0;                   // exponent sum (own + other) at that frame
    bool have;                // Z0 measured
    bool dead;                // Z0 == 0: no valid alignment
};

// 2^e as a double, e clamped to the normal range
__device__ __forceinline__ double pow2i(int e) {
    e = max(-1022, min(1023, e));
    return __hiloint2double((1023 + e) << 20, 0);
}

template <int SPL, bool kAlpha, bool kGrad>
__device__ __forceinline__ void ctc_frames(CtcLane<SPL>& st, GradNorm& gn, int step_lo, int step_hi, int Tb,
                                           int S, int V, int RS, int blank, const double* __restrict__ probs_u,
                                           double* __restrict__ lat_u, int* __restrict__ exp_u,
                                           float grad_scale, int accumulate, float* __restrict__ dlog_u,
                                           double* stage, int* racc) {
    const int lane = threadIdx.x & 31;
    const bool edge = kAlpha ? lane == 0 : lane == 31;
    constexpr double kMagic = 6755399441055744.0;       // 2^52 + 2^51
    constexpr double kFix = 1073741824.0;               // 2^30
    if (step_lo >= step_hi) return;

    auto issue_chunk = [&](int lo, int buf) {           // steps [lo, hi) -> contiguous frames
        const int hi = min(lo + kCtcChunk, step_hi);
        const int f0 = kAlpha ? lo : Tb - hi;
        const char* src = reinterpret_cast<const char*>(probs_u + (size_t)f0 * RS);
        char* dst = reinterpret_cast<char*>(stage + (size_t)buf * kCtcChunk * RS);
        const int n16 = (hi - lo) * RS / 2;
        for (int i = lane; i < n16; i += 32) cp_async16(dst + (size_t)i * 16, src + (size_t)i * 16);
        cp_async_commit();
    };

    double o[SPL];                                       // other direction's pre-emission sums, frame of `step`
    int eo = 0;
    if (kGrad) {
        const int t = kAlpha ? step_lo : Tb - 1 - step_lo;
        const double* lp = lat_u + (size_t)t * (SPL * 32) + lane;
#pragma unroll
        for (int j = 0; j < SPL; ++j) o[j] = lp[j * 32];
        eo = exp_u[t];
    }

    int buf = 0;
    issue_chunk(step_lo, 0);
    for (int lo = step_lo; lo < step_hi; lo += kCtcChunk, buf ^= 1) {
        const int hi = min(lo + kCtcChunk, step_hi);
        if (hi < step_hi) {
            issue_chunk(hi, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const double* chunk = stage + (size_t)buf * kCtcChunk * RS;
        for (int step = lo; step < hi; ++step) {
            const int t = kAlpha ? step : Tb - 1 - step;
            const double* row = chunk + (size_t)(kAlpha ? step - lo : hi - 1 - step) * RS;

            // ---- prefetch the other direction's values of the next frame -----------------------
            double on[SPL];
            int eon = 0;
            if (kGrad && step + 1 < step_hi) {
                const int tn = kAlpha ? t + 1 : t - 1;
                const double* lp = lat_u + (size_t)tn * (SPL * 32) + lane;
#pragma unroll
                for (int j = 0; j < SPL; ++j) on[j] = lp[j * 32];
                eon = exp_u[tn];
            }

            // ---- pre-emission sums, in place ---------------------------------------------------
            if (step == 0) {
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    const int s = lane * SPL + j;
                    const bool on_ = kAlpha ? (s <= 1 && s < S) : (s < S && s >= S - 2);
                    st.a[j] = on_ ? 1.0 : 0.0;
                }
            } else if (kAlpha) {
                double h = __shfl_up_sync(kFull, st.a[SPL - 1], 1);
                h = edge ? 0.0 : h;
#pragma unroll
                for (int j = SPL - 1; j >= 2; --j) {
                    if (j & 1) st.a[j] = fma(st.skipm[j >> 1], st.a[j - 2], st.a[j] + st.a[j - 1]);
                    else st.a[j] = st.a[j] + st.a[j - 1];
                }
                st.a[1] = fma(st.skipm[0], h, st.a[1] + st.a[0]);
                st.a[0] = st.a[0] + h;
            } else {
                double h0 = __shfl_down_sync(kFull, st.a[0], 1);
                double h1 = __shfl_down_sync(kFull, st.a[1], 1);
                h0 = edge ? 0.0 : h0;
                h1 = edge ? 0.0 : h1;
#pragma unroll
                for (int j = 0; j < SPL - 2; ++j) {
                    if (j & 1) st.a[j] = fma(st.skipm[j >> 1], st.a[j + 2], st.a[j] + st.a[j + 1]);
                    else st.a[j] = st.a[j] + st.a[j + 1];
                }
                st.a[SPL - 2] = st.a[SPL - 2] + st.a[SPL - 1];
                st.a[SPL - 1] = fma(st.skipm[SPL / 2 - 1], h1, st.a[SPL - 1] + h0);
            }

            if (!kGrad) {
                double* lp = lat_u + (size_t)t * (SPL * 32) + lane;
#pragma unroll
                for (int j = 0; j < SPL; ++j) lp[j * 32] = st.a[j];
                if (lane == 0) exp_u[t] = st.E;
            }

            // ---- emission ----------------------------------------------------------------------
            const double pb = row[blank];
#pragma unroll
            for (int j = 0; j < SPL; ++j) st.a[j] *= (j & 1) ? row[st.loff[j >> 1]] : pb;

            // ---- gradient row of frame t ---------------------------------------------------------
            if (kGrad) {
                double w[SPL];
                double zb = 0.0, zl = 0.0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) {
                    w[j] = st.a[j] * o[j];
                    if (j & 1) zl += w[j]; else zb += w[j];
                }
                if (!gn.have) {                           // first gradient frame: measure Z0 once
                    const double Z0 = warp_sum(zb + zl);
                    gn.have = true;
                    gn.dead = !(Z0 > 0.0);
                    gn.invZ0 = gn.dead ? 0.0 : 1.0 / Z0;
                    gn.E0 = st.E + eo;
                }
                const double c = gn.invZ0 * pow2i(st.E + eo - gn.E0) * kFix;
                const int ib = __double2loint(fma(zb, c, kMagic));
                const int gb = __reduce_add_sync(kFull, ib);
#pragma unroll
                for (int j = 1; j < SPL; j += 2)
                    atomicAdd(&racc[st.loff[j >> 1]], __double2loint(fma(w[j], c, kMagic)));
                __syncwarp();
                float* out = dlog_u + (size_t)t * V;
                for (int v = lane; v < V; v += 32) {
                    const int occ = v == blank ? gb : racc[v];
                    racc[v] = 0;
                    float g = gn.dead ? 0.0f : grad_scale * ((float)row[v] - (float)occ * 9.31322574615478515625e-10f);
                    out[v] = accumulate ? out[v] + g : g;
                }
                if (lane == 0) racc[V] = 0;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < SPL; ++j) o[j] = on[j];
                eo = eon;
            }

            // ---- exact power-of-two rescale every 4 steps ---------------------------------------
            if ((step & 3) == 3) {
                int mx = 0;
#pragma unroll
                for (int j = 0; j < SPL; ++j) mx = max(mx, __double2hiint(st.a[j]));
                mx = __reduce_max_sync(kFull, mx);
                if (mx >= 0x00100000) {
                    const int e = (mx >> 20) - 1023;
                    const double sc = __hiloint2double((1023 - e) << 20, 0);
                    st.E += e;
#pragma unroll
                    for (int j = 0; j < SPL; ++j) st.a[j] *= sc;
                }
            }
        }
        __syncwarp();
    }
}

template <int SPL, bool kAlpha>
__device__ void ctc_direction(const double* __restrict__ probs_u, const int32_t* __restrict__ lab_u, int Tb,
                              int L, int V, int RS, int blank, float grad_scale, int accumulate,
                              float* __restrict__ nll_out, float* __restrict__ dlog_u,
                              double* __restrict__ lat_u, int* __restrict__ exp_u, double* stage, int* racc) {
    const int lane = threadIdx.x & 31;
    const int S = 2 * L + 1;
    const int tm = Tb / 2;
    CtcLane<SPL> st;
    st.E = 0;
#pragma unroll
    for (int i = 0; i < SPL / 2; ++i) {
        const int li = (lane * SPL) / 2 + i;              // label index of odd state lane*SPL + 2i + 1
        const int c = li < L ? lab_u[li] : -1;
        st.loff[i] = c >= 0 ? c : V;
        bool legal;
        if (kAlpha) legal = c >= 0 && li >= 1 && lab_u[li - 1] != c;
        else legal = c >= 0 && li + 1 < L && lab_u[li + 1] != c;
        st.skipm[i] = legal ? 1.0 : 0.0;
    }
#pragma unroll
    for (int j = 0; j < SPL; ++j) st.a[j] = 0.0;
    for (int v = lane; v <= V; v += 32) racc[v] = 0;
    GradNorm gn;
    gn.have = false; gn.dead = false; gn.invZ0 = 0.0; gn.E0 = 0;

    // steps 0..Tb-1 visit frames 0..Tb-1 (alpha) or Tb-1..0 (beta).  First half: frames on this direction's
    // side of tm (store); second half: load the other direction's sums and emit gradient rows.
    const int n_first = kAlpha ? tm : Tb - tm;
    ctc_frames<SPL, kAlpha, false>(st, gn, 0, n_first, Tb, S, V, RS, blank, probs_u, lat_u, exp_u, grad_scale,
                                   accumulate, dlog_u, stage, racc);
    __syncthreads();                                      // the other warp's half-lattice is now visible
    ctc_frames<SPL, kAlpha, true>(st, gn, n_first, Tb, Tb, S, V, RS, blank, probs_u, lat_u, exp_u, grad_scale,
                                  accumulate, dlog_u, stage, racc);

    if (kAlpha) {
        double fin = 0.0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const int s = lane * SPL + j;
            if (s < S && s >= S - 2) fin += st.a[j];
        }
        fin = warp_sum(fin);
        if (lane == 0)
            *nll_out = fin > 0.0 ? (float)(-(log(fin) + (double)st.E * 0.69314718055994530942)) : INFINITY;
    }
}

template <int SPL>
__global__ void __launch_bounds__(kCtcThreads)
ctc_kernel(const float* __restrict__ logits, const float* __restrict__ probs_in,
           double* __restrict__ probs_ws, const int32_t* __restrict__ targets,
           const int32_t* __restrict__ in_len, const int32_t* __restrict__ tgt_len, int T, int V,
           int Lmax, int blank, float grad_scale, int accumulate, float* __restrict__ nll,
           float* __restrict__ dlogits, double* __restrict__ lattice, int* __restrict__ lat_exp) {
    extern __shared__ double smem_d[];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5;
    const int RS = ctc_row_stride(V);
    int Tb = in_len ? in_len[b] : T;
    Tb = min(max(Tb, 0), T);
    int L = tgt_len ? tgt_len[b] : Lmax;
    L = min(max(L, 0), Lmax);
    float* dlog_u = dlogits + (size_t)b * T * V;
    if (!accumulate)                                      // rows beyond the utterance carry no gradient
        for (int i = Tb * V + threadIdx.x; i < T * V; i += blockDim.x) dlog_u[i] = 0.0f;
    if (Tb == 0) {
        if (threadIdx.x == 0) nll[b] = L == 0 ? 0.0f : INFINITY;
        return;
    }
    double* pw = probs_ws + (size_t)b * T * RS;
    softmax_rows_f64(logits ? logits + (size_t)b * T * V : nullptr,
                     probs_in ? probs_in + (size_t)b * T * V : nullptr, pw, Tb, V, RS);
    int* exp_u = lat_exp + (size_t)b * (T + 1);
    __syncthreads();
    if (warp >= 2) {                                      // only two warps walk the lattice
        __syncthreads();                                  // (matches the mid-point barrier)
        return;
    }
    double* stage = smem_d + (size_t)warp * (2 * kCtcChunk * RS);
    int* racc = reinterpret_cast<int*>(smem_d + (size_t)2 * (2 * kCtcChunk * RS)) + warp * (V + 2);
    double* lat_u = lattice + (size_t)b * T * (SPL * 32);
    const int32_t* lab_u = targets + (size_t)b * Lmax;
    if (warp == 0)
        ctc_direction<SPL, true>(pw, lab_u, Tb, L, V, RS, blank, grad_scale, accumulate, nll + b, dlog_u,
                                 lat_u, exp_u, stage, racc);
    else
        ctc_direction<SPL, false>(pw, lab_u, Tb, L, V, RS, blank, grad_scale, accumulate, nll + b, dlog_u,
                                  lat_u, exp_u, stage, racc);
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

struct CtcWs { size_t lat, probs, exps, total; };
static CtcWs ctc_ws(int B, int T, int V, int spl) {
    CtcWs w;
    w.lat = align_up((size_t)B * T * spl * 32 * sizeof(double), 256);
    w.probs = align_up((size_t)B * T * ctc_row_stride(V) * sizeof(double), 256);
    w.exps = align_up((size_t)B * (T + 1) * sizeof(int), 256);
    w.total = w.lat + w.probs + w.exps;
    return w;
}

}  // namespace pgasr

extern "C" size_t pgasr_ctc_workspace_bytes(int B, int T, int V, int Lmax) {
    using namespace pgasr;
    const int spl = ctc_spl(Lmax);
    if (spl == 0 || B < 0 || T <= 0 || V <= 0) return 0;
    return ctc_ws(B, T, V, spl).total;
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
    const CtcWs w = ctc_ws(B, T, V, spl);
    if (workspace_bytes < w.total) return PGASR_ERR_WORKSPACE;
    if (B == 0) return PGASR_OK;
    char* base = reinterpret_cast<char*>(workspace);
    double* lattice = reinterpret_cast<double*>(base);
    double* probs_ws = reinterpret_cast<double*>(base + w.lat);
    int* lat_exp = reinterpret_cast<int*>(base + w.lat + w.probs);
    const int RS = ctc_row_stride(V);
    const size_t smem = (size_t)2 * (2 * kCtcChunk * RS) * sizeof(double) + (size_t)2 * (V + 2) * sizeof(int);
    if (smem > 200 * 1024) return PGASR_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
#define PGASR_CTC(SPLv)                                                                                   \
    do {                                                                                                  \
        if (smem > 48 * 1024)                                                                             \
            PGASR_CUDA_TRY(cudaFuncSetAttribute(ctc_kernel<SPLv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)smem));                                              \
        ctc_kernel<SPLv><<<B, kCtcThreads, smem, st>>>(logits, probs, probs_ws, targets, in_len, tgt_len, T, V, \
                                                       Lmax, blank, grad_scale, accumulate, nll, dlogits, \
                                                       lattice, lat_exp);                                 \
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
