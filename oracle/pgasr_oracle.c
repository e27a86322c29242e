/*
 * pgasr_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the Policy-Gradient-ASR sequence-level training hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library; the product path (policy-gradient-asr_b200/) never does.
 *
 * Where the reference has code, each function follows it and cites file:line
 * (paths are relative to the upstream repository root):
 *   orc_edit_distance       metrics.py:4-21       PINNED  (tests/golden/reference_vectors.json)
 *   orc_collapse            CTCdecoder.py:119-131 PINNED  (same fixture), + blank drop as
 *                                                 CTCdecoder.py:41 (blank=0) does in decode
 *   orc_reward_positions    policy_grad.py:10-15  PINNED  against the intent restatement
 *                                                 (the upstream function raises TypeError)
 *   orc_nll_sum             loss.py:13-17         PINNED
 * Where the reference has NO code the oracle implements the written spec in DESIGN.md
 * ("parity unpinned" -- no upstream code, tests or vectors exist for these):
 *   orc_pg_togo_loss_grad   DESIGN.md "reward-to-go spec": r_i is the PINNED last column of orc_edit_distance
 *                           (policy_grad.py:10-15), the credit assignment over frames is this repo's spec (fp64)
 *   orc_softmax_sample      DESIGN.md "sampler spec"  (bit-exact contract with the CUDA kernel)
 *   orc_pg_loss_grad        DESIGN.md "policy gradient spec" (fp64)
 *   orc_ctc_loss_grad       DESIGN.md "CTC spec" (fp64 log-space alpha-beta; cross-checked
 *                           against torch.nn.functional.ctc_loss on CPU in tests/)
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off matters: the sampler contract is "every fp32 operation rounds once".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_version(void) { return 1; }

/* bench.py's CPU arm uses every host core even when a launcher (torchrun) exported OMP_NUM_THREADS=1 */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------
 * a1  edit distance -- metrics.py:4-21
 *   dp has len(s2)+1 rows (hypothesis) and len(s1)+1 columns (reference)   metrics.py:12
 *   row 0 = 0..len(s1), column 0 = 0..len(s2)                              metrics.py:13-14
 *   match copies the diagonal, otherwise 1 + min(left, diag, up)          metrics.py:17-20
 *   returns dp[-1,-1]  (the caller pairs it with len(s1))                  metrics.py:21
 * The reference keeps the table in float64; the values are small integers, so int32 is exact.
 * last_col (optional, n+1 entries) receives dp[i, len(s1)] = ED(ref, hyp[:i]), which is what
 * policy_grad.py:11-15 evaluates one slice at a time.
 * ------------------------------------------------------------------------------------ */
ORC_API int orc_edit_distance(const int32_t* ref, int m, const int32_t* hyp, int n,
                              int32_t* last_col) {
    int W = m + 1;
    int32_t* dp = (int32_t*)malloc((size_t)(n + 1) * (size_t)W * sizeof(int32_t));
    if (!dp) return -1;
    for (int j = 0; j <= m; ++j) dp[j] = j;
    for (int i = 0; i <= n; ++i) dp[(size_t)i * W] = i;
    for (int i = 1; i <= n; ++i) {
        int32_t* row = dp + (size_t)i * W;
        const int32_t* up = row - W;
        for (int j = 1; j <= m; ++j) {
            if (hyp[i - 1] == ref[j - 1]) {
                row[j] = up[j - 1];
            } else {
                int32_t a = row[j - 1], b = up[j - 1], c = up[j];
                int32_t mn = a < b ? a : b;
                if (c < mn) mn = c;
                row[j] = 1 + mn;
            }
        }
    }
    int d = dp[(size_t)n * W + m];
    if (last_col)
        for (int i = 0; i <= n; ++i) last_col[i] = dp[(size_t)i * W + m];
    free(dp);
    return d;
}

/* ------------------------------------------------------------------------------------
 * a3  collapse -- CTCdecoder.py:119-131 keeps a symbol iff it differs from its predecessor
 * (first symbol always kept, :123-127).  The reference does not drop blanks there; the CTC
 * label map B() the sampler needs drops them AFTER merging repeats, as decode(blank=0) does
 * (CTCdecoder.py:41,78-106).  blank < 0 reproduces collapse_fn exactly.
 * Returns the output length.
 * ------------------------------------------------------------------------------------ */
ORC_API int orc_collapse(const int32_t* in, int n, int blank, int32_t* out) {
    int len = 0;
    int have_prev = 0;
    int32_t prev = 0;
    for (int t = 0; t < n; ++t) {
        int32_t c = in[t];
        if (have_prev && c == prev) continue;
        prev = c;
        have_prev = 1;
        if (blank >= 0 && c == blank) continue;
        out[len++] = c;
    }
    return len;
}

/* ------------------------------------------------------------------------------------
 * a4  per-position reward -- policy_grad.py:10-15, with edit_dist(...)[0] (the upstream code
 * subtracts the tuples and raises).  c[i] = ED(y*, yhat[:i]); python slices saturate.
 *   t > 1 : r_t = -(c[t+1] - c[t])              policy_grad.py:10-13
 *   t == 1: r_1 = -(c[2]   - len(y*))           policy_grad.py:14-15
 * r[0] is unused (the reference leaves r_t unbound for t < 1).  r has tmax+1 entries.
 * ------------------------------------------------------------------------------------ */
ORC_API void orc_reward_positions(const int32_t* ref, int m, const int32_t* hyp, int n,
                                  int tmax, int32_t* r) {
    int32_t* col = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t));
    orc_edit_distance(ref, m, hyp, n, col);
    r[0] = 0;
    for (int t = 1; t <= tmax; ++t) {
        int hi = t + 1 < n ? t + 1 : n;
        int lo = t < n ? t : n;
        if (t > 1) r[t] = -(col[hi] - col[lo]);
        else       r[t] = -(col[hi] - m);
    }
    free(col);
}

/* ------------------------------------------------------------------------------------
 * a5  customNLLLoss.forward -- loss.py:13-17: sum over steps i of NLLLoss(inp[i], target[:,i]),
 * each NLLLoss mean-reduced over the batch (loss.py:12; ignore_index=0 is falsy, loss.py:9, so
 * nothing is ever ignored unless ignore_index is a non-zero class).
 * inp [L,B,V] log-probs, target [B,L].  ignore_index < 0 : none.
 * grad (optional) [L,B,V] receives d loss / d inp.
 * ------------------------------------------------------------------------------------ */
ORC_API double orc_nll_sum(const float* inp, const int64_t* target, int L, int B, int V,
                           int ignore_index, double* grad) {
    double total = 0.0;
    if (grad) memset(grad, 0, sizeof(double) * (size_t)L * B * V);
    for (int i = 0; i < L; ++i) {
        double s = 0.0;
        int cnt = 0;
        for (int b = 0; b < B; ++b) {
            int64_t c = target[(size_t)b * L + i];
            if (ignore_index >= 0 && c == ignore_index) continue;
            s += -(double)inp[((size_t)i * B + b) * V + c];
            cnt++;
        }
        if (cnt > 0) {
            total += s / cnt;
            if (grad)
                for (int b = 0; b < B; ++b) {
                    int64_t c = target[(size_t)b * L + i];
                    if (ignore_index >= 0 && c == ignore_index) continue;
                    grad[((size_t)i * B + b) * V + c] = -1.0 / cnt;
                }
        } else {
            total += NAN; /* torch: mean over an empty set */
        }
    }
    return total;
}

/* ------------------------------------------------------------------------------------
 * a6  sampler spec.  Every operation below is ONE IEEE fp32 operation, rounded to
 * nearest even, never fused.  The CUDA kernel issues the same operations with __fmul_rn /
 * __fadd_rn / __fsub_rn, so the comparison below picks the same class on both sides.
 * ------------------------------------------------------------------------------------ */
static inline float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

ORC_API float orc_exp_spec(float x) {
    /* x <= 0 in the sampler (x = z - max).  Below -87 the result is defined as +0. */
    if (!(x >= -87.0f)) return 0.0f;
    float t = x * 1.44269504088896341f;
    float n = rintf(t);                       /* ties to even (default rounding mode) */
    float r = x - n * 0.693359375f;           /* hi part of ln2: product is exact */
    r = r - n * -2.12194440e-4f;              /* lo part of ln2 */
    float p = 1.9875691500e-4f;
    p = p * r + 1.3981999507e-3f;
    p = p * r + 8.3334519073e-3f;
    p = p * r + 4.1665795894e-2f;
    p = p * r + 1.6666665459e-1f;
    p = p * r + 5.0000001201e-1f;
    float r2 = r * r;
    float y = p * r2;
    y = y + r;
    y = y + 1.0f;
    int32_t ni = (int32_t)n;                  /* -126 .. 0 */
    float scale = f_from_bits((uint32_t)(ni + 127) << 23);
    return y * scale;
}

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the sampler uses when no
 * uniforms are injected. */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k[0];
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c[3] ^ k[1];
    uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* uniform for (utterance b, frame t, sample k): counter (t, b, k/4, 'PGAS'), key = seed,
 * lane k%4 of the block, top 24 bits -> [0,1). */
ORC_API float orc_philox_uniform(uint64_t seed, int b, int t, int k) {
    uint32_t ctr[4] = {(uint32_t)t, (uint32_t)b, (uint32_t)(k >> 2), 0x50474153u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return (float)(out[k & 3] >> 8) * 5.9604644775390625e-08f; /* 2^-24, exact */
}

/* One frame: cdf[v] = running fp32 sum of exp_spec(z[v] - max), v ascending.  Returns S. */
static float frame_cdf(const float* z, int V, float* cdf, float* zmax) {
    float m = z[0];
    for (int v = 1; v < V; ++v) if (z[v] > m) m = z[v];
    float c = 0.0f;
    for (int v = 0; v < V; ++v) {
        float e = orc_exp_spec(z[v] - m);
        c = c + e;
        cdf[v] = c;
    }
    *zmax = m;
    return c;
}

/* pick: number of classes whose cdf is <= u*S, clamped to V-1 (first v with u*S < cdf[v]). */
static inline int frame_pick(const float* cdf, int V, float S, float u) {
    float tau = u * S;
    int cnt = 0;
    for (int v = 0; v < V; ++v) cnt += (cdf[v] <= tau);
    return cnt < V - 1 ? cnt : V - 1;
}

/*
 * logits   [B,T,V] fp32, in_len [B], uniforms [B,K,T] fp32 or NULL (then Philox with seed)
 * samples  [B,K,T] u8 (frames t >= in_len[b] are written as 0)
 * logp     [B,K]   double: sum_t log_softmax(z)[t, pi_t] evaluated in fp64 (tolerance-checked)
 */
ORC_API void orc_softmax_sample(const float* logits, const int32_t* in_len, const float* uniforms,
                                uint64_t seed, int B, int T, int V, int K,
                                uint8_t* samples, double* logp) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* cdf = (float*)malloc(sizeof(float) * (size_t)V);
        int Tb = in_len ? in_len[b] : T;
        for (int k = 0; k < K; ++k) logp[(size_t)b * K + k] = 0.0;
        for (int t = 0; t < T; ++t) {
            const float* z = logits + ((size_t)b * T + t) * V;
            if (t >= Tb) {
                for (int k = 0; k < K; ++k) samples[((size_t)b * K + k) * T + t] = 0;
                continue;
            }
            float m;
            float S = frame_cdf(z, V, cdf, &m);
            /* fp64 log-softmax denominator for the tolerance-checked log-prob */
            double lse = 0.0;
            for (int v = 0; v < V; ++v) lse += exp((double)z[v] - (double)m);
            lse = (double)m + log(lse);
            for (int k = 0; k < K; ++k) {
                float u = uniforms ? uniforms[((size_t)b * K + k) * T + t]
                                   : orc_philox_uniform(seed, b, t, k);
                int pi = frame_pick(cdf, V, S, u);
                samples[((size_t)b * K + k) * T + t] = (uint8_t)pi;
                logp[(size_t)b * K + k] += (double)z[pi] - lse;
            }
        }
        free(cdf);
    }
}

/* collapse + edit distance for every (b,k): hyps [B,K,T] u8 (prefix valid), hyp_len, dist */
ORC_API void orc_collapse_score(const uint8_t* samples, const int32_t* in_len,
                                const int32_t* targets, const int32_t* tgt_len,
                                int B, int T, int K, int Lmax, int blank,
                                uint8_t* hyps, int32_t* hyp_len, int32_t* dist) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int Tb = in_len ? in_len[b] : T;
        int m = tgt_len ? tgt_len[b] : Lmax;
        int32_t* path = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
        int32_t* hyp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
        for (int k = 0; k < K; ++k) {
            const uint8_t* s = samples + ((size_t)b * K + k) * T;
            for (int t = 0; t < Tb; ++t) path[t] = s[t];
            int n = orc_collapse(path, Tb, blank, hyp);
            uint8_t* h = hyps + ((size_t)b * K + k) * T;
            for (int t = 0; t < T; ++t) h[t] = t < n ? (uint8_t)hyp[t] : 0;
            hyp_len[(size_t)b * K + k] = n;
            dist[(size_t)b * K + k] =
                orc_edit_distance(targets + (size_t)b * Lmax, m, hyp, n, NULL);
        }
        free(path);
        free(hyp);
    }
}

/* ------------------------------------------------------------------------------------
 * a7  rewards, baseline, policy-gradient loss and gradient (DESIGN.md "policy gradient spec").
 *   reward_mode 0: R = -ED            1: R = -ED / len(ref)   (metrics.py:24-25, the CER ratio)
 *   baseline    0: none  1: mean over the K samples of the utterance
 *               2: leave-one-out mean  3: external scalar `baseline_value`
 *   A = R - b;  L_pg = -(1/(B K)) sum_{b,k} A_bk logp_bk
 *   g[b,t,v] = (1/(B K)) ( p_tv sum_k A_bk - sum_k A_bk [pi_bkt = v] ),  t < in_len[b]
 * R in fp32 is part of the bit-exact contract (one fp32 division); everything else is fp64.
 * ------------------------------------------------------------------------------------ */
ORC_API double orc_pg_loss_grad(const float* logits, const int32_t* in_len,
                                const uint8_t* samples, const double* logp,
                                const int32_t* dist, const int32_t* tgt_len,
                                int B, int T, int V, int K, int Lmax,
                                int reward_mode, int baseline_mode, double baseline_value,
                                float* rewards, double* adv, double* grad) {
    double loss = 0.0;
    for (int b = 0; b < B; ++b) {
        int m = tgt_len ? tgt_len[b] : Lmax;
        double sumR = 0.0;
        for (int k = 0; k < K; ++k) {
            float R = -(float)dist[(size_t)b * K + k];
            if (reward_mode == 1) R = R / (float)m;
            rewards[(size_t)b * K + k] = R;
            sumR += (double)R;
        }
        for (int k = 0; k < K; ++k) {
            double R = (double)rewards[(size_t)b * K + k], base = 0.0;
            if (baseline_mode == 1) base = sumR / K;
            else if (baseline_mode == 2) base = K > 1 ? (sumR - R) / (K - 1) : 0.0;
            else if (baseline_mode == 3) base = baseline_value;
            adv[(size_t)b * K + k] = R - base;
            loss += -(R - base) * logp[(size_t)b * K + k];
        }
    }
    loss /= (double)B * K;
    if (!grad) return loss;
    double inv = 1.0 / ((double)B * K);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int Tb = in_len ? in_len[b] : T;
        double sumA = 0.0;
        for (int k = 0; k < K; ++k) sumA += adv[(size_t)b * K + k];
        for (int t = 0; t < T; ++t) {
            double* g = grad + ((size_t)b * T + t) * V;
            for (int v = 0; v < V; ++v) g[v] = 0.0;
            if (t >= Tb) continue;
            const float* z = logits + ((size_t)b * T + t) * V;
            double mx = z[0];
            for (int v = 1; v < V; ++v) if (z[v] > mx) mx = z[v];
            double den = 0.0;
            for (int v = 0; v < V; ++v) den += exp((double)z[v] - mx);
            for (int v = 0; v < V; ++v) g[v] = inv * sumA * exp((double)z[v] - mx) / den;
            for (int k = 0; k < K; ++k)
                g[samples[((size_t)b * K + k) * T + t]] -= inv * adv[(size_t)b * K + k];
        }
    }
    return loss;
}

/* ------------------------------------------------------------------------------------
 * 8f.1  per-position rewards and reward-to-go (DESIGN.md "reward-to-go spec").
 * Upstream's policy_grad.reward (policy_grad.py:10-15) scores position i of the COLLAPSED hypothesis with
 *   r_i = -(c[i+1] - c[i]),  c[i] = ED(y*, yhat[:i])  (c[0] = len(y*); the t == 1 branch is r_0 + r_1),
 * i.e. with the increments of the last column of the edit-distance table (orc_edit_distance's last_col,
 * pinned by the reward_positions golden vectors).  A frame t of the sampled path can only influence the symbols
 * emitted at frames >= t, so the return credited to the action at frame t is the reward-to-go
 *   G_t = sum_{i >= pos(t)} r_i = c[pos(t)] - c[n],   pos(t) = number of symbols emitted at frames < t,
 * where frame t emits iff (t == 0 or pi_t != pi_{t-1}) and pi_t != blank (orc_collapse), n = len(yhat).
 *   baseline (per frame, over the K samples of the utterance): 0 none | 1 mean_k G_kt | 2 leave-one-out | 3 value
 *   A_kt = G_kt - b_kt;  L_pg = -(1/(B K)) sum_{b,k,t<T_b} A_kt log p_t(pi_kt)
 *   g[b,t,v] = (1/(B K)) ( p_tv sum_k A_kt - sum_k A_kt [pi_kt = v] )
 * rewards[b,k] = G_0 = len(y*) - ED (fp32, exact);  to_go [B,K,T] int16 (0 for t >= T_b);  r_pos [B,K,T] int8
 * (0 for i >= n).  Integers are part of the bit-exact contract, the rest is fp64.
 * ------------------------------------------------------------------------------------ */
ORC_API double orc_pg_togo_loss_grad(const float* logits, const int32_t* in_len, const uint8_t* samples,
                                     const int32_t* targets, const int32_t* tgt_len,
                                     int B, int T, int V, int K, int Lmax, int blank,
                                     int baseline_mode, double baseline_value,
                                     float* rewards, int16_t* to_go, int8_t* r_pos, double* grad) {
    double loss = 0.0;
    double inv = 1.0 / ((double)B * K);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : loss)
    for (int b = 0; b < B; ++b) {
        int Tb = in_len ? in_len[b] : T;
        int m = tgt_len ? tgt_len[b] : Lmax;
        int32_t* path = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
        int32_t* hyp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
        int32_t* col = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
        int32_t* G = (int32_t*)malloc(sizeof(int32_t) * (size_t)K * (size_t)(T + 1));
        for (int k = 0; k < K; ++k) {
            const uint8_t* s = samples + ((size_t)b * K + k) * T;
            for (int t = 0; t < Tb; ++t) path[t] = s[t];
            int n = orc_collapse(path, Tb, blank, hyp);
            orc_edit_distance(targets + (size_t)b * Lmax, m, hyp, n, col);
            int8_t* rp = r_pos ? r_pos + ((size_t)b * K + k) * T : NULL;
            if (rp) for (int i = 0; i < T; ++i) rp[i] = i < n ? (int8_t)(-(col[i + 1] - col[i])) : 0;
            int pos = 0;
            for (int t = 0; t < T; ++t) {
                int g = t < Tb ? col[pos] - col[n] : 0;
                G[(size_t)k * T + t] = g;
                if (to_go) to_go[((size_t)b * K + k) * T + t] = (int16_t)g;
                if (t < Tb) {
                    int keep = (t == 0 || path[t] != path[t - 1]) && path[t] != blank;
                    pos += keep;
                }
            }
            rewards[(size_t)b * K + k] = (float)(col[0] - col[n]);
        }
        for (int t = 0; t < T; ++t) {
            double* g = grad ? grad + ((size_t)b * T + t) * V : NULL;
            if (g) for (int v = 0; v < V; ++v) g[v] = 0.0;
            if (t >= Tb) continue;
            const float* z = logits + ((size_t)b * T + t) * V;
            double mx = z[0];
            for (int v = 1; v < V; ++v) if (z[v] > mx) mx = z[v];
            double den = 0.0;
            for (int v = 0; v < V; ++v) den += exp((double)z[v] - mx);
            double lz = mx + log(den);
            double sumG = 0.0, sumA = 0.0;
            for (int k = 0; k < K; ++k) sumG += (double)G[(size_t)k * T + t];
            for (int k = 0; k < K; ++k) {
                double Gk = (double)G[(size_t)k * T + t], base = 0.0;
                if (baseline_mode == 1) base = sumG / K;
                else if (baseline_mode == 2) base = K > 1 ? (sumG - Gk) / (K - 1) : 0.0;
                else if (baseline_mode == 3) base = baseline_value;
                double A = Gk - base;
                int pi = samples[((size_t)b * K + k) * T + t];
                loss += -A * ((double)z[pi] - lz);
                sumA += A;
                if (g) g[pi] -= inv * A;
            }
            if (g) for (int v = 0; v < V; ++v) g[v] += inv * sumA * exp((double)z[v] - lz);
        }
        free(path); free(hyp); free(col); free(G);
    }
    return loss * inv;
}

/* ------------------------------------------------------------------------------------
 * a8  CTC alpha-beta in fp64 log space (DESIGN.md "CTC spec"; Graves et al. 2006).
 *   extended labels l' = (blank, l1, blank, ..., lL, blank), S = 2L+1, blank id given
 *   alpha_t(s) = lse(alpha_{t-1}(s), alpha_{t-1}(s-1), [alpha_{t-1}(s-2) if l'_s != blank and
 *                l'_s != l'_{s-2}]) + y_t(l'_s)
 *   nll = -lse(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
 *   d nll / d z_t(v) = p_t(v) - sum_{s: l'_s = v} exp(alpha_t(s) + beta_t(s) - y_t(v) + nll)
 * An utterance with no valid alignment gets nll = +inf and a zero gradient.
 * logits [B,T,V] fp32 (log-softmax is taken here, in fp64), targets [B,Lmax] int32.
 * nll [B]; grad [B,T,V] (may be NULL) receives d nll_b / d logits  (NOT divided by B).
 * ------------------------------------------------------------------------------------ */
static inline double lse2(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}

ORC_API void orc_ctc_loss_grad(const float* logits, const int32_t* targets,
                               const int32_t* in_len, const int32_t* tgt_len,
                               int B, int T, int V, int Lmax, int blank,
                               double* nll, double* grad) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        int Tb = in_len ? in_len[b] : T;
        int L = tgt_len ? tgt_len[b] : Lmax;
        int S = 2 * L + 1;
        const int32_t* lab = targets + (size_t)b * Lmax;
        double* g = grad ? grad + (size_t)b * T * V : NULL;
        if (g) memset(g, 0, sizeof(double) * (size_t)T * V);
        if (Tb <= 0) { nll[b] = (L == 0) ? 0.0 : INFINITY; continue; }
        double* lp = (double*)malloc(sizeof(double) * (size_t)Tb * V);
        double* al = (double*)malloc(sizeof(double) * (size_t)Tb * S);
        double* be = (double*)malloc(sizeof(double) * (size_t)Tb * S);
        for (int t = 0; t < Tb; ++t) {
            const float* z = logits + ((size_t)b * T + t) * V;
            double mx = z[0];
            for (int v = 1; v < V; ++v) if (z[v] > mx) mx = z[v];
            double den = 0.0;
            for (int v = 0; v < V; ++v) den += exp((double)z[v] - mx);
            double lz = mx + log(den);
            for (int v = 0; v < V; ++v) lp[(size_t)t * V + v] = (double)z[v] - lz;
        }
#define LBL(s) (((s) & 1) ? lab[(s) >> 1] : blank)
        for (int s = 0; s < S; ++s) al[s] = -INFINITY;
        al[0] = lp[blank];
        if (S > 1) al[1] = lp[LBL(1)];
        for (int t = 1; t < Tb; ++t) {
            const double* pa = al + (size_t)(t - 1) * S;
            double* ca = al + (size_t)t * S;
            for (int s = 0; s < S; ++s) {
                double a = pa[s];
                if (s >= 1) a = lse2(a, pa[s - 1]);
                if (s >= 2 && LBL(s) != blank && LBL(s) != LBL(s - 2)) a = lse2(a, pa[s - 2]);
                ca[s] = a + lp[(size_t)t * V + LBL(s)];
            }
        }
        const double* la = al + (size_t)(Tb - 1) * S;
        double ll = la[S - 1];
        if (S > 1) ll = lse2(ll, la[S - 2]);
        nll[b] = -ll;
        if (g && ll != -INFINITY) {
            double* lb = be + (size_t)(Tb - 1) * S;
            for (int s = 0; s < S; ++s) lb[s] = -INFINITY;
            lb[S - 1] = lp[(size_t)(Tb - 1) * V + blank];
            if (S > 1) lb[S - 2] = lp[(size_t)(Tb - 1) * V + LBL(S - 2)];
            for (int t = Tb - 2; t >= 0; --t) {
                const double* nb = be + (size_t)(t + 1) * S;
                double* cb = be + (size_t)t * S;
                for (int s = 0; s < S; ++s) {
                    double a = nb[s];
                    if (s + 1 < S) a = lse2(a, nb[s + 1]);
                    if (s + 2 < S && LBL(s + 2) != blank && LBL(s + 2) != LBL(s))
                        a = lse2(a, nb[s + 2]);
                    cb[s] = a + lp[(size_t)t * V + LBL(s)];
                }
            }
            for (int t = 0; t < Tb; ++t) {
                double* gt = g + (size_t)t * V;
                for (int v = 0; v < V; ++v) gt[v] = exp(lp[(size_t)t * V + v]);
                for (int s = 0; s < S; ++s) {
                    int v = LBL(s);
                    double x = al[(size_t)t * S + s] + be[(size_t)t * S + s];
                    if (x == -INFINITY) continue;
                    gt[v] -= exp(x - lp[(size_t)t * V + v] - ll);
                }
            }
        }
#undef LBL
        free(lp); free(al); free(be);
    }
}

/* ------------------------------------------------------------------------------------
 * Whole step, the CPU arm bench.py times: sample -> collapse -> edit distance -> reward ->
 * baseline -> PG loss/grad, plus CTC loss/grad, combined as
 *   loss = w_pg L_pg + w_ctc mean_b nll_b ;  dlogits = w_pg g_pg + (w_ctc / B) g_ctc
 * scratch is allocated here (this is a baseline, not a product).
 * ------------------------------------------------------------------------------------ */
ORC_API double orc_pg_ctc_step(const float* logits, const int32_t* targets,
                               const int32_t* in_len, const int32_t* tgt_len,
                               const float* uniforms, uint64_t seed,
                               int B, int T, int V, int K, int Lmax, int blank,
                               int reward_mode, int baseline_mode, double baseline_value,
                               double w_pg, double w_ctc,
                               float* rewards, double* nll, float* dlogits) {
    size_t nBK = (size_t)B * K, nBTV = (size_t)B * T * V;
    uint8_t* samples = (uint8_t*)malloc(nBK * T);
    uint8_t* hyps = (uint8_t*)malloc(nBK * T);
    int32_t* hyp_len = (int32_t*)malloc(nBK * sizeof(int32_t));
    int32_t* dist = (int32_t*)malloc(nBK * sizeof(int32_t));
    double* logp = (double*)malloc(nBK * sizeof(double));
    double* adv = (double*)malloc(nBK * sizeof(double));
    double* gpg = (double*)malloc(nBTV * sizeof(double));
    double* gctc = (double*)malloc(nBTV * sizeof(double));
    orc_softmax_sample(logits, in_len, uniforms, seed, B, T, V, K, samples, logp);
    orc_collapse_score(samples, in_len, targets, tgt_len, B, T, K, Lmax, blank, hyps, hyp_len, dist);
    double lpg;
    if (reward_mode == 2)
        lpg = orc_pg_togo_loss_grad(logits, in_len, samples, targets, tgt_len, B, T, V, K, Lmax, blank,
                                    baseline_mode, baseline_value, rewards, NULL, NULL, gpg);
    else
        lpg = orc_pg_loss_grad(logits, in_len, samples, logp, dist, tgt_len, B, T, V, K, Lmax,
                               reward_mode, baseline_mode, baseline_value, rewards, adv, gpg);
    orc_ctc_loss_grad(logits, targets, in_len, tgt_len, B, T, V, Lmax, blank, nll, gctc);
    double lctc = 0.0;
    for (int b = 0; b < B; ++b) lctc += nll[b];
    lctc /= B;
    if (dlogits)
        for (size_t i = 0; i < nBTV; ++i)
            dlogits[i] = (float)(w_pg * gpg[i] + (w_ctc / B) * gctc[i]);
    free(samples); free(hyps); free(hyp_len); free(dist);
    free(logp); free(adv); free(gpg); free(gctc);
    return w_pg * lpg + w_ctc * lctc;
}
