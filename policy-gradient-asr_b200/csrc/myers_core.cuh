// Bit-parallel Levenshtein (Myers 1999 in Hyyro's 2003 edit-distance form), one thread per hypothesis.
// Shared by levenshtein.cu and fused.cu.  Upstream semantics: metrics.py:4-21 (unit costs; the hypothesis
// indexes the rows of the DP table, the reference the columns); the running score after i hypothesis symbols is
// dp[i, len(ref)], the column policy_grad.py:10-15 reads.
#pragma once
#include "pgasr_common.cuh"

namespace pgasr {

// Match table: peq[c*W + w] bit j%32 of word j/32 is set iff ref[j] == c; row `vocab` is all zero and is what
// out-of-vocabulary hypothesis symbols read.  Call from all threads of the CTA; ends with __syncthreads().
template <int W>
__device__ __forceinline__ void myers_build_peq(uint32_t* peq, int vocab, const int32_t* __restrict__ ref, int m) {
    for (int i = threadIdx.x; i < (vocab + 1) * W; i += blockDim.x) peq[i] = 0u;
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const uint32_t c = (uint32_t)ref[j];
        if (c < (uint32_t)vocab) atomicOr(&peq[c * W + (j >> 5)], 1u << (j & 31));
    }
    __syncthreads();
}

template <int W>
struct MyersState {
    uint32_t VP[W], VN[W];
};

template <int W>
__device__ __forceinline__ void myers_load_eq(const uint32_t* __restrict__ peq, uint32_t c, uint32_t (&eq)[W]) {
    if (W == 1) {
        eq[0] = peq[c];
    } else if (W == 2) {
        const uint2 e = reinterpret_cast<const uint2*>(peq)[c];
        eq[0] = e.x; eq[1] = e.y;
    } else {
#pragma unroll
        for (int q = 0; q < W / 4; ++q) {
            const uint4 e = reinterpret_cast<const uint4*>(peq)[c * (W / 4) + q];
            eq[4 * q] = e.x; eq[4 * q + 1] = e.y; eq[4 * q + 2] = e.z; eq[4 * q + 3] = e.w;
        }
    }
}

// One hypothesis symbol.  Returns the change of dp[i, m] (+1, 0, -1) when kDelta, else 0.
template <int W, bool kDelta>
__device__ __forceinline__ int myers_step(MyersState<W>& s, const uint32_t (&eq)[W], const uint32_t (&sel)[W]) {
    uint32_t D0[W], HP[W], HN[W];
    uint32_t carry = 0u;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint64_t sum = (uint64_t)(eq[w] & s.VP[w]) + s.VP[w] + carry;
        carry = (uint32_t)(sum >> 32);
        D0[w] = (((uint32_t)sum ^ s.VP[w]) | eq[w]) | s.VN[w];
        HP[w] = s.VN[w] | ~(D0[w] | s.VP[w]);
        HN[w] = D0[w] & s.VP[w];
    }
    int delta = 0;
    if (kDelta) {
        uint32_t hp = 0u, hn = 0u;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            hp |= HP[w] & sel[w];
            hn |= HN[w] & sel[w];
        }
        delta = (hp != 0u) - (hn != 0u);
    }
#pragma unroll
    for (int w = W - 1; w >= 0; --w) {
        const uint32_t hps = (HP[w] << 1) | (w ? HP[w - 1] >> 31 : 1u);
        const uint32_t hns = (HN[w] << 1) | (w ? HN[w - 1] >> 31 : 0u);
        s.VP[w] = hns | ~(D0[w] | hps);
        s.VN[w] = hps & D0[w];
    }
    return delta;
}

// ED(ref[:m], h[:n]).  col (kLastCol) receives dp[i, m] for i = 0..n.  h may live in shared or global memory.
template <int W, bool kLastCol>
__device__ __forceinline__ int myers_row(const uint8_t* __restrict__ h, int n, const uint32_t* __restrict__ peq,
                                         int vocab, int m, int32_t* __restrict__ col) {
    MyersState<W> s;
    uint32_t sel[W];
    const int wm = m > 0 ? (m - 1) >> 5 : 0;
    const uint32_t bm = m > 0 ? 1u << ((m - 1) & 31) : 0u;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        s.VP[w] = 0xffffffffu;
        s.VN[w] = 0u;
        sel[w] = (w == wm) ? bm : 0u;
    }
    int score = m;
    if (kLastCol) col[0] = m;
    const uint32_t vmax = (uint32_t)vocab;
    int i = 0;
    if ((reinterpret_cast<uintptr_t>(h) & 3u) == 0u) {    // four symbols per load, next word prefetched
        const uint32_t* h4 = reinterpret_cast<const uint32_t*>(h);
        uint32_t pack = n >= 4 ? h4[0] : 0u;
        for (; i + 4 <= n; i += 4) {
            const uint32_t nxt = i + 8 <= n ? h4[(i >> 2) + 1] : 0u;
            uint32_t eq[4][W];
#pragma unroll
            for (int q = 0; q < 4; ++q) myers_load_eq<W>(peq, min((pack >> (8 * q)) & 0xffu, vmax), eq[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int d = myers_step<W, kLastCol>(s, eq[q], sel);
                if (kLastCol) {
                    score += d;
                    col[i + q + 1] = m > 0 ? score : i + q + 1;
                }
            }
            pack = nxt;
        }
    }
    for (; i < n; ++i) {
        uint32_t eq[W];
        myers_load_eq<W>(peq, min((uint32_t)h[i], vmax), eq);
        const int d = myers_step<W, kLastCol>(s, eq, sel);
        if (kLastCol) {
            score += d;
            col[i + 1] = m > 0 ? score : i + 1;
        }
    }
    // dp[n, m] = dp[n, 0] + sum_{j<m} (VP_j - VN_j)
    int d = n;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const int lo = w * 32;
        const uint32_t msk = m >= lo + 32 ? 0xffffffffu : (m > lo ? (1u << (m - lo)) - 1u : 0u);
        d += __popc(s.VP[w] & msk) - __popc(s.VN[w] & msk);
    }
    return d;
}

}  // namespace pgasr
