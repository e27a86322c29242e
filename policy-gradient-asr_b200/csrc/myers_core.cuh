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
template <int W, bool kLastCol, typename ColT = int32_t>
__device__ __forceinline__ int myers_row(const uint8_t* __restrict__ h, int n, const uint32_t* __restrict__ peq,
                                         int vocab, int m, ColT* __restrict__ col) {
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
    if (kLastCol) col[0] = (ColT)m;
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
                    col[i + q + 1] = (ColT)(m > 0 ? score : i + q + 1);
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
            col[i + 1] = (ColT)(m > 0 ? score : i + 1);
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

// The W words of one hypothesis can be spread over P adjacent lanes (WL = W / P words each) as a block-skewed pipeline:
// the hypothesis is cut into blocks of 8 symbols and in outer iteration `it` lane p works on block it - p, one block
// behind lane p - 1 (myers_half below).  (A skew of one SYMBOL was measured slower than one thread per sample, 190
// against 115 cycles per symbol: the shuffle and the loads then sit on the dependent chain of every symbol.)

// One block of up to 8 hypothesis symbols (packed little endian in sy) on this lane's WL words of a P-lane group.
// Every horizontal dependency of the recurrence runs from low words to high words: the carry of the addition and the
// top bits of HP and HN that are shifted into the next word.  They travel as three words per block -- bit (n-1-q) of
// cin.c / .p / .n is what symbol q of an n-symbol block receives (a funnel shift per symbol collects them most
// significant first: one instruction each, where packing three bits per symbol into one word cost eight).
struct MyersCarry { uint32_t c, p, n; };
__device__ __forceinline__ MyersCarry myers_carry_low(int nsym) {      // what the lowest lane of a group sees: (0, 1, 0)
    MyersCarry k;
    k.c = 0u; k.p = nsym >= 32 ? 0xffffffffu : (1u << nsym) - 1u; k.n = 0u;
    return k;
}
template <int W, int WL, bool kWhole, bool kClamp>
__device__ __forceinline__ MyersCarry myers_block8(uint32_t (&VP)[WL], uint32_t (&VN)[WL], MyersCarry cin, uint2 sy,
                                                   const uint32_t* __restrict__ pq, uint32_t vmax, int nvalid) {
    constexpr int BS = 8;
    uint32_t eq[BS][WL];
#pragma unroll
    for (int q = 0; q < BS; ++q) {
        uint32_t c = __byte_perm(q < 4 ? sy.x : sy.y, 0u, 0x4440u + (q & 3));   // byte q, zero extended
        if (kClamp || !kWhole) c = min(c, vmax);          // (a partial block's tail bytes are whatever the buffer holds)
        if constexpr (WL == 1) {
            eq[q][0] = pq[c * W];
        } else if constexpr (WL == 2) {
            const uint2 e = *reinterpret_cast<const uint2*>(pq + c * W);
            eq[q][0] = e.x; eq[q][1] = e.y;
        } else {
            const uint4 e = *reinterpret_cast<const uint4*>(pq + c * W);
            eq[q][0] = e.x; eq[q][1] = e.y; eq[q][2] = e.z; eq[q][3] = e.w;
        }
    }
    MyersCarry out;
    out.c = out.p = out.n = 0u;
    const int nv = kWhole ? BS : nvalid;
#pragma unroll
    for (int q = 0; q < BS; ++q) {
        if (kWhole || q < nvalid) {
            const int sh = nv - 1 - q;                    // (a compile-time constant for whole blocks)
            uint32_t D0[WL], HP[WL], HN[WL];
            uint32_t carry = (cin.c >> sh) & 1u;
#pragma unroll
            for (int w = 0; w < WL; ++w) {
                const uint64_t sum = (uint64_t)(eq[q][w] & VP[w]) + VP[w] + carry;
                carry = (uint32_t)(sum >> 32);
                D0[w] = (((uint32_t)sum ^ VP[w]) | eq[q][w]) | VN[w];
                HP[w] = VN[w] | ~(D0[w] | VP[w]);
                HN[w] = D0[w] & VP[w];
            }
            out.c = (out.c << 1) | carry;
            out.p = __funnelshift_l(HP[WL - 1], out.p, 1);    // (out.p << 1) | (HP >> 31)
            out.n = __funnelshift_l(HN[WL - 1], out.n, 1);
#pragma unroll
            for (int w = WL - 1; w >= 0; --w) {
                const uint32_t hps = (HP[w] << 1) | (w ? HP[w - 1] >> 31 : (cin.p >> sh) & 1u);
                const uint32_t hns = (HN[w] << 1) | (w ? HN[w - 1] >> 31 : (cin.n >> sh) & 1u);
                VP[w] = hns | ~(D0[w] | hps);
                VN[w] = hps & D0[w];
            }
        }
    }
    return out;
}

// Meeting in the middle: ED(ref[:m], h[:n]) = min_j ED(ref[:j], h[:n1]) + ED(ref[j:], h[n1:]).  The forward half
// walks h[:n1] against the match table of the reference, the backward half walks the REVERSED second half
// (hrev[i] = h[n-1-i], n - n1 symbols) against the table of the reversed reference; the two halves run in DIFFERENT
// warps (the phase is bound by the integer pipe of the scheduler partition a warp sits on -- two interleaved chains in
// one warp took exactly as long as one chain of twice the length, measured -- so the halves must sit on different
// partitions), each as the block-skewed pipeline of myers_half.  The column of a half is its VP / VN vectors
// (vertical differences): the backward warps write theirs out as prefix sums (G), and after a CTA barrier the forward
// warps form theirs on the fly and take the minimum over the m + 1 meeting points.
__host__ __device__ __forceinline__ int myers_split_point(int n) { return min(n, ((n + 1) / 2 + 7) & ~7); }

// one half: nsym symbols of h (8-byte aligned, readable to the next multiple of 8) on this lane's WL words; the state
// is left in VP / VN.  Call with all 32 lanes; nmax = the largest nsym of the warp.
template <int W, int P, bool kClamp = true>
__device__ __forceinline__ void myers_half(const uint8_t* __restrict__ h, int nsym, const uint32_t* __restrict__ peq_any,
                                           int vocab, int p, int nmax, uint32_t (&VP)[W / P], uint32_t (&VN)[W / P]) {
    constexpr int WL = W / P;
    constexpr int BS = 8;
#pragma unroll
    for (int w = 0; w < WL; ++w) { VP[w] = 0xffffffffu; VN[w] = 0u; }
    const uint32_t vmax = (uint32_t)vocab;
    const uint32_t* pq = peq_any + p * WL;
    MyersCarry cin = myers_carry_low(BS);
    const int iters = (nmax + BS - 1) / BS + P - 1;
    for (int it = 0; it < iters; ++it) {
        const int i0 = (it - p) * BS;
        MyersCarry out;
        out.c = out.p = out.n = 0u;
        if (i0 >= 0 && i0 < nsym) {
            const uint2 sy = *reinterpret_cast<const uint2*>(h + i0);
            if (i0 + BS <= nsym) {
                out = myers_block8<W, WL, true, kClamp>(VP, VN, cin, sy, pq, vmax, BS);
            } else {
                if (p == 0) cin = myers_carry_low(nsym - i0);
                out = myers_block8<W, WL, false, kClamp>(VP, VN, cin, sy, pq, vmax, nsym - i0);
            }
        }
        cin.c = __shfl_up_sync(kFull, out.c, 1);
        cin.p = __shfl_up_sync(kFull, out.p, 1);
        cin.n = __shfl_up_sync(kFull, out.n, 1);
        if (p == 0) cin = myers_carry_low(BS);
    }
}

// backward warps: G[j] = ED(rev ref[:j], rev h second half) = n2 + sum_{i<j} (VP_i - VN_i), j = 0 .. W*32, as int16
template <int W, int P>
__device__ __forceinline__ void myers_store_column(const uint32_t (&VP)[W / P], const uint32_t (&VN)[W / P], int n2, int p,
                                                   int16_t* __restrict__ G) {
    constexpr int WL = W / P;
    int tot = 0;
#pragma unroll
    for (int w = 0; w < WL; ++w) tot += __popc(VP[w]) - __popc(VN[w]);
    int incl = tot;
#pragma unroll
    for (int o = 1; o < P; o <<= 1) {
        const int x = __shfl_up_sync(kFull, incl, o);
        if (p >= o) incl += x;
    }
    int base = n2 + incl - tot;
    if (p == 0) G[0] = (int16_t)n2;
#pragma unroll
    for (int w = 0; w < WL; ++w) {
        const int j0 = (p * WL + w) * 32;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            base += (int)((VP[w] >> i) & 1u) - (int)((VN[w] >> i) & 1u);
            G[j0 + i + 1] = (int16_t)base;
        }
    }
}

// forward warps, after the barrier: min_j F[j] + G[m - j]; every lane of a group returns the distance
template <int W, int P>
__device__ __forceinline__ int myers_meet(const uint32_t (&VP)[W / P], const uint32_t (&VN)[W / P], int n1, int m, int p,
                                          const int16_t* __restrict__ G) {
    constexpr int WL = W / P;
    int tot = 0;
#pragma unroll
    for (int w = 0; w < WL; ++w) tot += __popc(VP[w]) - __popc(VN[w]);
    int incl = tot;
#pragma unroll
    for (int o = 1; o < P; o <<= 1) {
        const int x = __shfl_up_sync(kFull, incl, o);
        if (p >= o) incl += x;
    }
    int base = n1 + incl - tot;
    int best = p == 0 ? n1 + (int)G[m] : 1 << 20;          // j = 0
#pragma unroll
    for (int w = 0; w < WL; ++w) {
        const int j0 = (p * WL + w) * 32;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            base += (int)((VP[w] >> i) & 1u) - (int)((VN[w] >> i) & 1u);      // F[j0 + i + 1]
            const int j = j0 + i + 1;
            if (j <= m) best = min(best, base + (int)G[m - j]);
        }
    }
#pragma unroll
    for (int o = 1; o < P; o <<= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
    return best;
}

}  // namespace pgasr
