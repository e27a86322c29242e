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
