// Context, device matrices and root tables shared by every entry point of libgenstark_b200.so.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <cuda_runtime.h>
#include "fp128.cuh"
#include "comm.h"
#include "../../include/genstark_b200.h"

namespace gs {


struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    // root tables for w_G
    int log_g = 24, log_lo = 12;
    fp* tw_lo = nullptr;
    fp* tw_hi = nullptr;
    fp* tw_small = nullptr;     // w_1024^i
    u128 root_g = 0;
    // full twiddle tables of the NTT passes, built on first use per shape (ntt_host.cuh)
    std::map<unsigned long long, fp*> tw_tables;
    size_t tw_table_bytes = 0;
    fp* ntt2_xs = nullptr;          // CTA-private running-product tiles of the two-pass LDE (ntt2.cuh), 2 * SMs * 64 KiB
    // scratch arena (grow on demand)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned mailbox for small device->host results
    void* mailbox = nullptr;
    size_t mailbox_bytes = 0;
    int sm_count = 148;
    unsigned* counters = nullptr;   // zeroed device words for "last block finishes" kernels (each use leaves them zero)
    // coset-sharded multi-GPU prover: one process per GPU, NCCL communicator over all ranks of the box
    int rank = 0, world = 1;
    NcclComm comm = nullptr;
    unsigned long long launches = 0;   // kernels launched through this context (bench.py gpu_launches)
    uint32_t prove_epoch = 0;          // one per prove ATTEMPT on this context (the mailbox handshake of prover.cuh)
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    // per-kernel-class timing with CUDA events on the launching stream (bench.py roofline)
    bool profiling = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    std::vector<std::string> prof_names;
    std::map<int, std::pair<unsigned long long, double>> prof_acc;   // class -> (launch groups, ms)
    std::string prof_json;
    int prof_class(const char* name) {
        for (size_t i = 0; i < prof_names.size(); ++i) if (prof_names[i] == name) return (int)i;
        prof_names.push_back(name); return (int)prof_names.size() - 1;
    }
    cudaEvent_t prof_event() {
        if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    // call after the stream is synchronized
    void prof_collect() {
        for (auto& r : prof_recs) {
            float ms = 0; if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { auto& acc = prof_acc[r.cls]; acc.first++; acc.second += ms; }
            prof_pool.push_back(r.a); prof_pool.push_back(r.b);
        }
        prof_recs.clear();
    }

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        last_error = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what) {
        return fail(GS_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    int ensure_scratch(size_t bytes) {
        if (bytes <= scratch_bytes) return GS_OK;
        if (scratch) { cudaStreamSynchronize(stream); cudaFree(scratch); scratch = nullptr; scratch_bytes = 0; }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&scratch, want);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
        scratch_bytes = want;
        return GS_OK;
    }
    // w_G^e on the host
    u128 root_pow(u128 e) const { return h_pow(root_g, e); }
    // primitive root of order 2^log_n from the same family
    u128 root_of_order(int log_n) const { return h_pow(root_g, (u128)1 << (log_g - log_n)); }
};

struct ProfScope {
    Ctx* c; int idx = -1;
    ProfScope(Ctx* ctx, const char* name) : c(ctx) {
        if (!c->profiling) return;
        Ctx::ProfRec r; r.cls = c->prof_class(name); r.a = c->prof_event(); r.b = c->prof_event();
        cudaEventRecord(r.a, c->stream);
        c->prof_recs.push_back(r); idx = (int)c->prof_recs.size() - 1;
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof_recs[idx].b, c->stream); }
};

#define GS_CUDA(ctx, call)                                              \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return (ctx)->cuda_fail(_e, #call);      \
    } while (0)

struct Mat {
    Ctx* ctx;
    fp* data;
    long long rows, cols;
    bool owns;
};

static inline int ctx_init_tables(Ctx* c) {
    const int lo_n = 1 << c->log_lo;
    const int hi_n = 1 << (c->log_g > c->log_lo ? c->log_g - c->log_lo : 0);
    std::vector<fp> lo(lo_n), hi(hi_n), sm(1024);
    c->root_g = h_root_of_unity(c->log_g);
    u128 acc = 1;
    for (int i = 0; i < lo_n; ++i) { lo[i] = fp_from_u128(acc); acc = h_mul(acc, c->root_g); }
    u128 step = h_pow(c->root_g, (u128)lo_n);
    acc = 1;
    for (int i = 0; i < hi_n; ++i) { hi[i] = fp_from_u128(acc); acc = h_mul(acc, step); }
    u128 w1024 = c->root_of_order(10);
    acc = 1;
    for (int i = 0; i < 1024; ++i) { sm[i] = fp_from_u128(acc); acc = h_mul(acc, w1024); }
    GS_CUDA(c, cudaMalloc(&c->tw_lo, lo_n * sizeof(fp)));
    GS_CUDA(c, cudaMalloc(&c->tw_hi, hi_n * sizeof(fp)));
    GS_CUDA(c, cudaMalloc(&c->tw_small, 1024 * sizeof(fp)));
    GS_CUDA(c, cudaMemcpy(c->tw_lo, lo.data(), lo_n * sizeof(fp), cudaMemcpyHostToDevice));
    GS_CUDA(c, cudaMemcpy(c->tw_hi, hi.data(), hi_n * sizeof(fp), cudaMemcpyHostToDevice));
    GS_CUDA(c, cudaMemcpy(c->tw_small, sm.data(), 1024 * sizeof(fp), cudaMemcpyHostToDevice));
    return GS_OK;
}

}  // namespace gs
