// K2: repeat merge + blank drop (upstream CTCdecoder.py:119-131 collapse_fn keeps a
// symbol iff it differs from its predecessor; blanks are dropped after the merge as decode(blank=0),
// CTCdecoder.py:41, does).  One warp per row: ballot of the keep flags, popc prefix -> compaction.
#include "pgasr_common.cuh"

namespace pgasr {

constexpr int kCollapseWarps = 4;

__global__ void __launch_bounds__(kCollapseWarps* kWarp)
collapse_u8_kernel(const uint8_t* __restrict__ seqs, const int32_t* __restrict__ seq_len,
                   int rows_per_len, int N, int T, int blank, uint8_t* __restrict__ out,
                   int32_t* __restrict__ out_len) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * kCollapseWarps + (threadIdx.x >> 5);
    if (r >= N) return;
    int len = seq_len ? seq_len[r / rows_per_len] : T;
    len = min(max(len, 0), T);
    const uint8_t* in = seqs + (size_t)r * T;
    uint8_t* o = out + (size_t)r * T;
    int base = 0;
    int carry = -1;                                   // symbol before the chunk; -1 = none
    for (int t0 = 0; t0 < len; t0 += 32) {
        const int t = t0 + lane;
        const int x = t < len ? (int)in[t] : -2;
        int p = __shfl_up_sync(kFull, x, 1);
        if (lane == 0) p = carry;
        const bool keep = t < len && x != p && x != blank;   // blank < 0 never matches a symbol
        const unsigned mask = __ballot_sync(kFull, keep);
        if (keep) o[base + __popc(mask & ((1u << lane) - 1u))] = (uint8_t)x;
        base += __popc(mask);
        carry = __shfl_sync(kFull, x, 31);
    }
    for (int t = base + lane; t < T; t += 32) o[t] = 0;
    if (lane == 0) out_len[r] = base;
}

}  // namespace pgasr

extern "C" int pgasr_collapse_u8(const uint8_t* seqs, const int32_t* seq_len, int rows_per_len, int N,
                                 int T, int blank, uint8_t* out, int32_t* out_len, void* stream) {
    using namespace pgasr;
    if (!seqs || !out || !out_len || N < 0 || T <= 0 || rows_per_len <= 0 || blank > 255)
        return PGASR_ERR_INVALID_ARG;
    if (N == 0) return PGASR_OK;
    const int grid = (N + kCollapseWarps - 1) / kCollapseWarps;
    collapse_u8_kernel<<<grid, kCollapseWarps * kWarp, 0, as_stream(stream)>>>(
        seqs, seq_len, rows_per_len, N, T, blank, out, out_len);
    PGASR_LAUNCH_CHECK();
    return PGASR_OK;
}
