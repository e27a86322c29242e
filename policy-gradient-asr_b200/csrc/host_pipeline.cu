// Host-buffer pipeline around pgasr_pg_ctc_step: what upstream's call site sees when model outputs and
// transcripts live in HOST memory (upstream: model_out.detach().cpu().numpy() before the metric/reward code,
// model.py:317-320; criterion(model_out, t), model.py:235).  `depth` steps are in flight at once: the H2D copy of
// step n+1 and the D2H copy of step n-1 run on their own streams (and copy engines) while step n's kernel runs.
//
//   copy-in stream : logits H2D -> ev_in[slot]
//   compute stream : wait ev_in[slot] ; fused kernel -> ev_k[slot]
//   copy-out stream: wait ev_k[slot] ; dlogits D2H -> ev_out[slot]
// Small inputs (targets, lengths) and small outputs (loss, rewards, nll) are packed into one pinned, device-mapped
// staging block per slot that the kernel accesses directly (zero copy): a step costs ONE DMA per direction.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>

#include "pgasr_common.cuh"

struct pgasr_host_pipeline {
    int B, T, V, K, Lmax, depth;
    cudaStream_t s_in, s_k, s_out;
    void* workspace;
    size_t workspace_bytes;
    struct Slot {
        float* logits_d; float* dlogits_d;
        char* small_in_d; char* small_in_h;       // targets [B*Lmax] i32 | in_len [B] | tgt_len [B]
        float* small_out_d; float* small_out_h;   // loss [4] | rewards [B*K] | nll [B]
        cudaEvent_t ev_in, ev_k, ev_out;
#ifdef PGASR_TIMING
        cudaEvent_t ev_h0, ev_k0, ev_o0;             // starts of the H2D, kernel and D2H phases (timing build)
#endif
        long long ticket;                          // step occupying the slot (-1: free)
        bool collected;                            // small outputs already handed to the caller
        float* loss_h; float* rewards_h; float* nll_h;
        cudaStream_t st; void* ws;                 // per-slot mode: the slot's own stream and step workspace
    }* slots;
    bool per_slot;                                 // one in-order stream per slot instead of one stream per engine
    long long next_ticket;
    size_t small_in_bytes, small_out_floats;
#ifdef PGASR_TIMING
    cudaEvent_t ev_base;
    double host_ns_collect, host_ns_enqueue;
#endif
};

namespace {

using pgasr::cuda_fail;

void destroy(pgasr_host_pipeline* p) {
    if (!p) return;
    if (p->slots) {
        for (int i = 0; i < p->depth; ++i) {
            auto& s = p->slots[i];
            if (s.ev_out) cudaEventSynchronize(s.ev_out);
            cudaFree(s.logits_d); cudaFree(s.dlogits_d);
            cudaFreeHost(s.small_in_h); cudaFreeHost(s.small_out_h);
            if (s.ev_in) cudaEventDestroy(s.ev_in);
            if (s.ev_k) cudaEventDestroy(s.ev_k);
            if (s.ev_out) cudaEventDestroy(s.ev_out);
            if (s.st) cudaStreamDestroy(s.st);
            cudaFree(s.ws);
        }
        delete[] p->slots;
    }
    cudaFree(p->workspace);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_k) cudaStreamDestroy(p->s_k);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
}

// hand the packed small outputs of a finished slot to the caller's buffers
int collect(pgasr_host_pipeline* p, pgasr_host_pipeline::Slot& s) {
    if (s.ticket < 0 || s.collected) return PGASR_OK;
    PGASR_CUDA_TRY(cudaEventSynchronize(s.ev_out));
    const size_t BK = (size_t)p->B * p->K;
    if (s.loss_h) s.loss_h[0] = s.small_out_h[0];
    if (s.rewards_h) std::memcpy(s.rewards_h, s.small_out_h + 4, BK * sizeof(float));
    if (s.nll_h) std::memcpy(s.nll_h, s.small_out_h + 4 + BK, (size_t)p->B * sizeof(float));
    s.collected = true;
    return PGASR_OK;
}

}  // namespace

extern "C" int pgasr_host_create(int B, int T, int V, int K, int Lmax, int depth, pgasr_host_pipeline** out) {
    if (!out || B <= 0 || T <= 0 || V <= 0 || K <= 0 || Lmax <= 0 || depth <= 0 || depth > 16)
        return PGASR_ERR_INVALID_ARG;
    *out = nullptr;
    if (K > 64 || V > pgasr::kMaxV) return PGASR_ERR_UNSUPPORTED;
    const int dc = pgasr_device_check();
    if (dc != PGASR_OK) return dc;
    const size_t ws = pgasr_pg_ctc_step_workspace_bytes(B, T, V, K, Lmax);
    if (ws == 0) return PGASR_ERR_UNSUPPORTED;
    auto* p = new (std::nothrow) pgasr_host_pipeline();
    if (!p) return PGASR_ERR_INVALID_ARG;
    std::memset(p, 0, sizeof(*p));
    p->B = B; p->T = T; p->V = V; p->K = K; p->Lmax = Lmax; p->depth = depth;
    p->workspace_bytes = ws;
    p->small_in_bytes = ((size_t)B * Lmax + 2 * (size_t)B) * sizeof(int32_t);
    p->small_out_floats = 4 + (size_t)B * K + B;
    p->slots = new (std::nothrow) pgasr_host_pipeline::Slot[depth];
    if (!p->slots) { destroy(p); return PGASR_ERR_INVALID_ARG; }
    std::memset(p->slots, 0, sizeof(pgasr_host_pipeline::Slot) * depth);
    const size_t nlog = (size_t)B * T * V * sizeof(float);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; return r == cudaSuccess; };
    ok(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_k, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    ok(cudaMalloc(&p->workspace, ws));
    const char* mode = std::getenv("PGASR_HOST_PIPELINE");
    p->per_slot = mode && std::strcmp(mode, "slot") == 0;
    for (int i = 0; i < depth && e == cudaSuccess; ++i) {
        auto& s = p->slots[i];
        s.ticket = -1;
        ok(cudaMalloc(&s.logits_d, nlog));
        ok(cudaMalloc(&s.dlogits_d, nlog));
        if (p->per_slot) {
            ok(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            ok(cudaMalloc(&s.ws, ws));
            if (e == cudaSuccess && pgasr_pg_ctc_step_workspace_init(s.ws, ws, s.st) != PGASR_OK) e = cudaErrorUnknown;
            ok(cudaStreamSynchronize(s.st));
        }
        ok(cudaHostAlloc(&s.small_in_h, p->small_in_bytes, cudaHostAllocMapped));
        ok(cudaHostAlloc(&s.small_out_h, p->small_out_floats * sizeof(float), cudaHostAllocMapped));
        if (e == cudaSuccess) {                        // device-side aliases of the two staging blocks
            ok(cudaHostGetDevicePointer(reinterpret_cast<void**>(&s.small_in_d), s.small_in_h, 0));
            ok(cudaHostGetDevicePointer(reinterpret_cast<void**>(&s.small_out_d), s.small_out_h, 0));
        }
#ifdef PGASR_TIMING
        ok(cudaEventCreate(&s.ev_in)); ok(cudaEventCreate(&s.ev_k)); ok(cudaEventCreate(&s.ev_out));
        ok(cudaEventCreate(&s.ev_h0)); ok(cudaEventCreate(&s.ev_k0)); ok(cudaEventCreate(&s.ev_o0));
#else
        ok(cudaEventCreateWithFlags(&s.ev_in, cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&s.ev_k, cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&s.ev_out, cudaEventDisableTiming));
#endif
    }
    if (e == cudaSuccess) {
        const int rc = pgasr_pg_ctc_step_workspace_init(p->workspace, ws, p->s_k);
        if (rc != PGASR_OK) { destroy(p); return rc; }
        ok(cudaStreamSynchronize(p->s_k));
#ifdef PGASR_TIMING
        ok(cudaEventCreate(&p->ev_base));
        ok(cudaEventRecord(p->ev_base, p->s_k));
#endif
    }
    if (e != cudaSuccess) {
        destroy(p);
        return cuda_fail(e);
    }
    *out = p;
    return PGASR_OK;
}

extern "C" int pgasr_host_destroy(pgasr_host_pipeline* p) {
    destroy(p);
    return PGASR_OK;
}

extern "C" int pgasr_host_submit(pgasr_host_pipeline* p, const float* logits_h, const int32_t* targets_h,
                                 const int32_t* in_len_h, const int32_t* tgt_len_h, uint64_t seed, int blank,
                                 int reward_mode, int baseline_mode, float baseline_value, float w_pg, float w_ctc,
                                 float* loss_h, float* dlogits_h, float* rewards_h, float* nll_h,
                                 int64_t* ticket) {
    if (!p || !logits_h || !targets_h || !loss_h || !dlogits_h) return PGASR_ERR_INVALID_ARG;
    const long long n = p->next_ticket;
    auto& s = p->slots[n % p->depth];
#ifdef PGASR_TIMING
    const auto h0 = std::chrono::steady_clock::now();
#endif
    int rc = collect(p, s);                        // blocks only when step n - depth has not finished yet
    if (rc != PGASR_OK) return rc;
#ifdef PGASR_TIMING
    const auto h1 = std::chrono::steady_clock::now();
#endif
    const int B = p->B, T = p->T, V = p->V, K = p->K, Lmax = p->Lmax;
    const size_t nlog = (size_t)B * T * V * sizeof(float);
    // pack the small inputs (absent lengths mean "full length")
    int32_t* si = reinterpret_cast<int32_t*>(s.small_in_h);
    std::memcpy(si, targets_h, (size_t)B * Lmax * sizeof(int32_t));
    int32_t* il = si + (size_t)B * Lmax;
    int32_t* tl = il + B;
    for (int b = 0; b < B; ++b) il[b] = in_len_h ? in_len_h[b] : T;
    for (int b = 0; b < B; ++b) tl[b] = tgt_len_h ? tgt_len_h[b] : Lmax;
    // No stream-side wait guards the slot's device buffers: collect() above has already blocked the HOST until the
    // D2H copies of the step that used this slot finished, and those were ordered after its kernel.  (Every
    // cross-stream wait is a semaphore on a copy-engine channel; the redundant ones cost engine time.)
    cudaStream_t q_in = p->per_slot ? s.st : p->s_in, q_k = p->per_slot ? s.st : p->s_k, q_out = p->per_slot ? s.st : p->s_out;
    void* wsp = p->per_slot ? s.ws : p->workspace;
#ifdef PGASR_TIMING
    PGASR_CUDA_TRY(cudaEventRecord(s.ev_h0, q_in));
#endif
    PGASR_CUDA_TRY(cudaMemcpyAsync(s.logits_d, logits_h, nlog, cudaMemcpyHostToDevice, q_in));
    if (!p->per_slot) {
        PGASR_CUDA_TRY(cudaEventRecord(s.ev_in, q_in));
        PGASR_CUDA_TRY(cudaStreamWaitEvent(q_k, s.ev_in, 0));
    }
#ifdef PGASR_TIMING
    if (p->per_slot) PGASR_CUDA_TRY(cudaEventRecord(s.ev_in, q_in));
    PGASR_CUDA_TRY(cudaEventRecord(s.ev_k0, q_k));
#endif
    // The small inputs and outputs are NOT copied: the kernel reads the transcripts and lengths from, and writes the
    // loss / rewards / nll to, the slot's pinned staging block through its device mapping (zero copy).  A 26 KB DMA
    // queued next to the 3.8 MB one cost ~15 us of copy-engine time per direction per step (measured,
    // tools/interference_probe.py: 101 -> 131 us/step); the kernel-side cost is one PCIe round trip at CTA start.
    const int32_t* tg_d = reinterpret_cast<const int32_t*>(s.small_in_d);
    rc = pgasr_pg_ctc_step(s.logits_d, tg_d, tg_d + (size_t)B * Lmax, tg_d + (size_t)B * Lmax + B, nullptr, seed, B,
                           T, V, K, Lmax, blank, reward_mode, baseline_mode, baseline_value, w_pg, w_ctc,
                           s.small_out_d, s.dlogits_d, s.small_out_d + 4, nullptr, nullptr, nullptr,
                           s.small_out_d + 4 + (size_t)B * K, nullptr, wsp, p->workspace_bytes, q_k);
    if (rc != PGASR_OK) return rc;
    if (!p->per_slot) {
        PGASR_CUDA_TRY(cudaEventRecord(s.ev_k, q_k));
        PGASR_CUDA_TRY(cudaStreamWaitEvent(q_out, s.ev_k, 0));
    }
#ifdef PGASR_TIMING
    if (p->per_slot) PGASR_CUDA_TRY(cudaEventRecord(s.ev_k, q_k));
    PGASR_CUDA_TRY(cudaEventRecord(s.ev_o0, q_out));
#endif
    PGASR_CUDA_TRY(cudaMemcpyAsync(dlogits_h, s.dlogits_d, nlog, cudaMemcpyDeviceToHost, q_out));
    PGASR_CUDA_TRY(cudaEventRecord(s.ev_out, q_out));
    s.ticket = n;
    s.collected = false;
    s.loss_h = loss_h; s.rewards_h = rewards_h; s.nll_h = nll_h;
    p->next_ticket = n + 1;
    if (ticket) *ticket = n;
#ifdef PGASR_TIMING
    const auto h2 = std::chrono::steady_clock::now();
    p->host_ns_collect += std::chrono::duration<double, std::nano>(h1 - h0).count();
    p->host_ns_enqueue += std::chrono::duration<double, std::nano>(h2 - h1).count();
#endif
    return PGASR_OK;
}

extern "C" int pgasr_host_wait(pgasr_host_pipeline* p, int64_t ticket) {
    if (!p) return PGASR_ERR_INVALID_ARG;
    for (int i = 0; i < p->depth; ++i) {
        auto& s = p->slots[i];
        if (s.ticket < 0) continue;
        if (ticket < 0 || s.ticket <= ticket) {    // outputs complete in submission order
            const int rc = collect(p, s);
            if (rc != PGASR_OK) return rc;
        }
    }
    return PGASR_OK;
}

extern "C" int pgasr_host_pin(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return PGASR_ERR_INVALID_ARG;
    PGASR_CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return PGASR_OK;
}

extern "C" int pgasr_host_unpin(void* ptr) {
    if (!ptr) return PGASR_ERR_INVALID_ARG;
    PGASR_CUDA_TRY(cudaHostUnregister(ptr));
    return PGASR_OK;
}

#ifdef PGASR_TIMING
// timing build only: ms since pipeline creation of {H2D start, H2D end, kernel start, kernel end, D2H start, D2H end}
// of the step currently held by slot `slot` (call after pgasr_host_wait)
extern "C" __attribute__((visibility("default"))) int pgasr_host_debug_host_ns(pgasr_host_pipeline* p, double* out2) {
    out2[0] = p->host_ns_collect; out2[1] = p->host_ns_enqueue;
    p->host_ns_collect = p->host_ns_enqueue = 0.0;
    return 0;
}
extern "C" __attribute__((visibility("default"))) int pgasr_host_debug_times(pgasr_host_pipeline* p, int slot, float* ms6) {
    auto& s = p->slots[slot];
    cudaEvent_t ev[6] = {s.ev_h0, s.ev_in, s.ev_k0, s.ev_k, s.ev_o0, s.ev_out};
    for (int i = 0; i < 6; ++i)
        if (cudaEventElapsedTime(&ms6[i], p->ev_base, ev[i]) != cudaSuccess) return -5;
    return 0;
}
#endif
