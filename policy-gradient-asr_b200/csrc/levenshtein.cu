// K3: batched Levenshtein distance (upstream metrics.py:4-21 edit_dist:
// unit costs, hypothesis on rows, reference on columns; policy_grad.py:10-15 reads the last column).
//
// Two kernels, both exact integer arithmetic (bit-exact against the DP table):
//  * myers_u8: Myers' bit-vector algorithm in Hyyro's edit-distance form.  The reference string is the
//    bit pattern (<= 512 symbols = W 32-bit words held in registers), the hypothesis is consumed one symbol
//    per step, one THREAD per hypothesis; the K hypotheses of an utterance share the match table Peq[c][W]
//    in shared memory.  After step i the running score is dp[i, len(ref)] = ED(ref, hyp[:i]), i.e. the
//    column reward() needs, for free.  Per symbol it costs ~11 W integer instructions instead of the
//    len(ref) cell updates of the DP.
//  * wavefront_i32: the anti-diagonal DP over the table held in shared memory, one CTA per pair, for
//    arbitrary int32 tokens (word ids for WER, code points) and as an independent cross-check.
#include "myers_core.cuh"

namespace pgasr {

template <int W, bool kLastCol>
__global__ void myers_u8_kernel(const uint8_t* __restrict__ hyps, const int32_t* __restrict__ hyp_len,
                                int N, int hyp_stride, const int32_t* __restrict__ refs,
                                const int32_t* __restrict__ ref_len, int rows_per_ref, int ref_stride,
                                int vocab, int32_t* __restrict__ dist, int32_t* __restrict__ last_col) {
    extern __shared__ __align__(16) uint32_t peq[];   // [vocab + 1][W]
    const int g = blockIdx.x;
    int m = ref_len ? ref_len[g] : ref_stride;
    m = min(max(m, 0), ref_stride);
    myers_build_peq<W>(peq, vocab, refs + (size_t)g * ref_stride, m);
    if ((int)threadIdx.x >= rows_per_ref) return;
    const int row = g * rows_per_ref + threadIdx.x;
    if (row >= N) return;
    int n = hyp_len[row];
    n = min(max(n, 0), hyp_stride);
    int32_t* col = kLastCol ? last_col + (size_t)row * (hyp_stride + 1) : nullptr;
    dist[row] = myers_row<W, kLastCol>(hyps + (size_t)row * hyp_stride, n, peq, vocab, m, col);
}

// The same distance with the work of a row spread out (what the single-launch step does, fused_impl.cuh P3): the W
// words of a row on 4 adjacent lanes as a block-skewed pipeline, and the hypothesis cut in two -- forward half on the
// first warps, reversed second half on the others (myers_core.cuh, "Meeting in the middle").  One CTA per reference
// group; the rows are staged in shared memory (8-byte aligned, second halves reversed).  Used when no last column is
// asked for, W >= 4 and the group fits (rows_per_ref <= 64).
template <int W>
__global__ void myers_u8_split_kernel(const uint8_t* __restrict__ hyps, const int32_t* __restrict__ hyp_len,
                                      int N, int hyp_stride, const int32_t* __restrict__ refs,
                                      const int32_t* __restrict__ ref_len, int rows_per_ref, int ref_stride,
                                      int vocab, int32_t* __restrict__ dist) {
    constexpr int P = 4;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int K = rows_per_ref;
    const int Tp = (hyp_stride + 15) & ~15, Tp2 = (hyp_stride / 2 + 16) & ~15;
    uint32_t* peq = reinterpret_cast<uint32_t*>(sm_raw);                       // [vocab + 1][W]
    uint32_t* peq_r = peq + (size_t)(vocab + 1) * W;                           // reversed reference
    int16_t* fg = reinterpret_cast<int16_t*>(peq_r + (size_t)(vocab + 1) * W); // [(K + 7) & ~7][W * 32 + 2]
    uint8_t* hf = reinterpret_cast<uint8_t*>(fg + (size_t)((K + 7) & ~7) * (W * 32 + 2));
    hf += (16 - (reinterpret_cast<uintptr_t>(hf) & 15)) & 15;                  // [K][Tp]
    uint8_t* hr = hf + (size_t)K * Tp;                                         // [K][Tp2]
    const int g = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int m = ref_len ? ref_len[g] : ref_stride;
    m = min(max(m, 0), ref_stride);
    for (int i = threadIdx.x; i < 2 * (vocab + 1) * W; i += blockDim.x) peq[i] = 0u;
    __syncthreads();
    const int32_t* ref = refs + (size_t)g * ref_stride;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const uint32_t c = (uint32_t)ref[j];
        if (c < (uint32_t)vocab) {
            atomicOr(&peq[c * W + (j >> 5)], 1u << (j & 31));
            const int jr = m - 1 - j;
            atomicOr(&peq_r[c * W + (jr >> 5)], 1u << (jr & 31));
        }
    }
    for (int k = warp; k < K; k += blockDim.x / 32) {      // stage the rows: first part forwards, second half reversed
        const int row = g * K + k;
        int n = row < N ? hyp_len[row] : 0;
        n = min(max(n, 0), hyp_stride);
        const uint8_t* src = hyps + (size_t)min(row, N - 1) * hyp_stride;
        const int n1 = myers_split_point(n);
        for (int i = lane; i < n1; i += 32) hf[(size_t)k * Tp + i] = src[i];
        for (int i = lane; i < n - n1; i += 32) hr[(size_t)k * Tp2 + i] = src[n - 1 - i];
    }
    __syncthreads();
    const int nw = (K * P + 31) / 32;
    uint32_t VPh[W / P], VNh[W / P];
    const bool fwd = warp < nw;
    const int tid = (int)threadIdx.x - (fwd ? 0 : nw * 32);
    const int k = tid / P, p = tid % P;
    const int kc = min(k, K - 1);
    const int row = g * K + kc;
    int n = (k < K && row < N) ? hyp_len[row] : 0;
    n = min(max(n, 0), hyp_stride);
    const int n1 = myers_split_point(n);
    const int nsym = fwd ? n1 : n - n1;
    const int nmax = __reduce_max_sync(kFull, nsym);
    myers_half<W, P>(fwd ? hf + (size_t)kc * Tp : hr + (size_t)kc * Tp2, nsym, fwd ? peq : peq_r, vocab, p, nmax, VPh, VNh);
    if (!fwd) myers_store_column<W, P>(VPh, VNh, n - n1, p, fg + (size_t)k * (W * 32 + 2));
    __syncthreads();
    if (fwd) {
        const int d = myers_meet<W, P>(VPh, VNh, n1, m, p, fg + (size_t)k * (W * 32 + 2));
        if (k < K && p == 0 && g * K + k < N) dist[g * K + k] = d;
    }
}

static size_t myers_split_smem(int W, int K, int hyp_stride, int vocab) {
    const int Tp = (hyp_stride + 15) & ~15, Tp2 = (hyp_stride / 2 + 16) & ~15;
    return (size_t)2 * (vocab + 1) * W * 4 + (size_t)((K + 7) & ~7) * (W * 32 + 2) * 2 + 16 + (size_t)K * (Tp + Tp2);
}

// Anti-diagonal wavefront.  Cell (i, j), i over the hypothesis, j over the reference, sits on diagonal
// i + j.  Three rotating diagonals of len(ref)+1 entries live in shared memory, indexed by j.
__global__ void wavefront_i32_kernel(const int32_t* __restrict__ hyps, const int32_t* __restrict__ hyp_len,
                                     int hyp_stride, const int32_t* __restrict__ refs,
                                     const int32_t* __restrict__ ref_len, int rows_per_ref, int ref_stride,
                                     int32_t* __restrict__ dist) {
    extern __shared__ int32_t sm[];
    const int row = blockIdx.x;
    const int g = row / rows_per_ref;
    int m = ref_len ? ref_len[g] : ref_stride;
    m = min(max(m, 0), ref_stride);
    int n = hyp_len ? hyp_len[row] : hyp_stride;
    n = min(max(n, 0), hyp_stride);
    int32_t* rs = sm;                                  // ref symbols [m]
    int32_t* d0 = sm + ref_stride;                     // diagonal d-2
    int32_t* d1 = d0 + (ref_stride + 1);               // diagonal d-1
    int32_t* d2 = d1 + (ref_stride + 1);               // diagonal d
    const int32_t* h = hyps + (size_t)row * hyp_stride;
    for (int j = threadIdx.x; j < m; j += blockDim.x) rs[j] = refs[(size_t)g * ref_stride + j];
    if (threadIdx.x == 0) {
        d0[0] = 0;                                     // diagonal 0: cell (0,0)
        d1[0] = 1;                                     // diagonal 1: (1,0) and (0,1)
        if (m >= 1) d1[1] = 1;
    }
    __syncthreads();
    if (n == 0 || m == 0) {
        if (threadIdx.x == 0) dist[row] = n + m;
        return;
    }
    for (int d = 2; d <= n + m; ++d) {
        const int jlo = max(0, d - n), jhi = min(m, d);
        for (int j = jlo + threadIdx.x; j <= jhi; j += blockDim.x) {
            const int i = d - j;
            int v;
            if (j == 0) v = i;
            else if (i == 0) v = j;
            else if (h[i - 1] == rs[j - 1]) v = d0[j - 1];
            else v = 1 + min(min(d1[j - 1], d0[j - 1]), d1[j]);
            d2[j] = v;
        }
        __syncthreads();
        int32_t* tmp = d0; d0 = d1; d1 = d2; d2 = tmp;
    }
    if (threadIdx.x == 0) dist[row] = d1[m];
}

template <int W>
static int launch_myers(const uint8_t* hyps, const int32_t* hyp_len, int N, int hyp_stride,
                        const int32_t* refs, const int32_t* ref_len, int rows_per_ref, int ref_stride,
                        int vocab, int32_t* dist, int32_t* last_col, cudaStream_t st) {
    const int groups = (N + rows_per_ref - 1) / rows_per_ref;
    if constexpr (W >= 4) {
        const size_t need = myers_split_smem(W, rows_per_ref, hyp_stride, vocab);
        if (!last_col && rows_per_ref <= 64 && need <= 200 * 1024) {
            if (need > 48 * 1024)
                PGASR_CUDA_TRY(cudaFuncSetAttribute(myers_u8_split_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            const int nw = (rows_per_ref * 4 + 31) / 32;
            myers_u8_split_kernel<W><<<groups, 2 * nw * 32, need, st>>>(hyps, hyp_len, N, hyp_stride, refs, ref_len, rows_per_ref,
                                                                      ref_stride, vocab, dist);
            PGASR_LAUNCH_CHECK();
            return PGASR_OK;
        }
    }
    const int threads = ((rows_per_ref + 31) / 32) * 32;
    const size_t smem = (size_t)(vocab + 1) * W * sizeof(uint32_t);
    if (last_col)
        myers_u8_kernel<W, true><<<groups, threads, smem, st>>>(hyps, hyp_len, N, hyp_stride, refs, ref_len,
                                                               rows_per_ref, ref_stride, vocab, dist, last_col);
    else
        myers_u8_kernel<W, false><<<groups, threads, smem, st>>>(hyps, hyp_len, N, hyp_stride, refs, ref_len,
                                                                rows_per_ref, ref_stride, vocab, dist, last_col);
    PGASR_LAUNCH_CHECK();
    return PGASR_OK;
}

}  // namespace pgasr

extern "C" int pgasr_edit_distance_u8(const uint8_t* hyps, const int32_t* hyp_len, int N, int hyp_stride,
                                      const int32_t* refs, const int32_t* ref_len, int rows_per_ref,
                                      int ref_stride, int vocab, int32_t* dist, int32_t* last_col,
                                      void* stream) {
    using namespace pgasr;
    if (!hyps || !hyp_len || !refs || !dist || N < 0 || hyp_stride <= 0 || ref_stride <= 0 ||
        rows_per_ref <= 0 || rows_per_ref > 1024 || vocab <= 0 || vocab > 256)
        return PGASR_ERR_INVALID_ARG;
    if (ref_stride > 512) return PGASR_ERR_UNSUPPORTED;
    if (N == 0) return PGASR_OK;
    cudaStream_t st = as_stream(stream);
#define PGASR_MYERS(Wv) \
    return launch_myers<Wv>(hyps, hyp_len, N, hyp_stride, refs, ref_len, rows_per_ref, ref_stride, vocab, dist, last_col, st)
    if (ref_stride <= 32) PGASR_MYERS(1);
    if (ref_stride <= 64) PGASR_MYERS(2);
    if (ref_stride <= 128) PGASR_MYERS(4);
    if (ref_stride <= 256) PGASR_MYERS(8);
    PGASR_MYERS(16);
#undef PGASR_MYERS
}

extern "C" int pgasr_edit_distance_i32(const int32_t* hyps, const int32_t* hyp_len, int N, int hyp_stride,
                                       const int32_t* refs, const int32_t* ref_len, int rows_per_ref,
                                       int ref_stride, int32_t* dist, void* stream) {
    using namespace pgasr;
    if (!hyps || !refs || !dist || N < 0 || hyp_stride <= 0 || ref_stride <= 0 || rows_per_ref <= 0)
        return PGASR_ERR_INVALID_ARG;
    if (ref_stride > 4096) return PGASR_ERR_UNSUPPORTED;
    if (N == 0) return PGASR_OK;
    int threads = ((ref_stride + 1 + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = ((size_t)ref_stride + 3 * ((size_t)ref_stride + 1)) * sizeof(int32_t);
    if (smem > 48 * 1024)
        PGASR_CUDA_TRY(cudaFuncSetAttribute(wavefront_i32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
    wavefront_i32_kernel<<<N, threads, smem, as_stream(stream)>>>(hyps, hyp_len, hyp_stride, refs, ref_len,
                                                                  rows_per_ref, ref_stride, dist);
    PGASR_LAUNCH_CHECK();
    return PGASR_OK;
}
