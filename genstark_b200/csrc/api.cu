// extern "C" surface of libgenstark_b200.so (declared in include/genstark_b200.h).
#include <memory>
#include "core.cuh"
#include "ntt_host.cuh"
#include "pointwise.cuh"
#include "prover.cuh"
#include "../../include/genstark_b200.h"

using namespace gs;

struct gs_ctx : public Ctx {};
struct gs_mat : public Mat {};
struct gs_stark : public Stark { std::vector<uint8_t> proof; std::string times_json; };

static thread_local std::string g_null_error;

static int ilog2_exact(long long n) {
    if (n <= 0 || (n & (n - 1))) return -1;
    int k = 0;
    while ((1ll << k) < n) ++k;
    return k;
}

extern "C" {

int gs_ctx_create(int device, gs_ctx** out) {
    if (!out) return GS_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        g_null_error = "no CUDA device: libgenstark_b200 has no CPU fallback";
        return GS_E_CUDA;
    }
    if (device < 0 || device >= n) { g_null_error = "bad device index"; return GS_E_ARG; }
    gs_ctx* c = new gs_ctx();
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_null_error = cudaGetErrorString(e); delete c; return GS_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    int rc = ctx_init_tables(c);
    if (rc != GS_OK) { g_null_error = c->last_error; delete c; return rc; }
    c->mailbox_bytes = 8 << 20;
    if (cudaHostAlloc(&c->mailbox, c->mailbox_bytes, cudaHostAllocDefault) != cudaSuccess) {
        g_null_error = "cudaHostAlloc(mailbox)"; delete c; return GS_E_CUDA;
    }
    *out = c;
    return GS_OK;
}

void gs_ctx_destroy(gs_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->tw_lo); cudaFree(c->tw_hi); cudaFree(c->tw_small);
    if (c->scratch) cudaFree(c->scratch);
    if (c->mailbox) cudaFreeHost(c->mailbox);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char* gs_last_error(gs_ctx* c) { return c ? c->last_error.c_str() : g_null_error.c_str(); }

int gs_ctx_sync(gs_ctx* c) {
    if (!c) return GS_E_ARG;
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    return GS_OK;
}

uint64_t gs_ctx_launch_count(gs_ctx* c) { return c ? c->launches : 0; }

int gs_field_supported(const uint8_t* modulus_le, size_t nbytes) {
    if (!modulus_le || nbytes != 16) return GS_E_UNSUPPORTED;
    fp p; memcpy(&p, modulus_le, 16);
    return (p.v[0] == P0 && p.v[1] == P1 && p.v[2] == P2 && p.v[3] == P3) ? GS_OK : GS_E_UNSUPPORTED;
}

int gs_field_root_of_unity(int log2_order, uint8_t out16[16]) {
    if (!out16 || log2_order < 0 || log2_order > 32) return GS_E_ARG;
    fp r = fp_from_u128(h_root_of_unity(log2_order));
    memcpy(out16, &r, 16);
    return GS_OK;
}

// scalar field operations run on the host (they are scalars in the reference too)
int gs_field_scalar_op(int op, const uint8_t a16[16], const uint8_t b16[16], uint8_t out16[16]) {
    if (!a16 || !b16 || !out16) return GS_E_ARG;
    fp a, b; memcpy(&a, a16, 16); memcpy(&b, b16, 16);
    u128 x = fp_to_u128(a), y = fp_to_u128(b), r;
    if (x >= HP || y >= HP) return GS_E_ARG;
    switch (op) {
        case 0: r = h_add(x, y); break;
        case 1: r = h_sub(x, y); break;
        case 2: r = h_mul(x, y); break;
        case 3: r = h_mul(x, h_inv(y)); break;          // div, inv(0) = 0
        case 4: r = h_pow(x, y); break;                 // exp (exponent reduced by the caller)
        default: return GS_E_ARG;
    }
    fp o = fp_from_u128(r); memcpy(out16, &o, 16);
    return GS_OK;
}

// ---- matrices --------------------------------------------------------------------------------
int gs_mat_alloc(gs_ctx* c, int64_t rows, int64_t cols, gs_mat** out) {
    if (!c || !out || rows <= 0 || cols <= 0) return c ? c->fail(GS_E_ARG, "bad matrix shape") : GS_E_ARG;
    cudaSetDevice(c->device);
    gs_mat* m = new gs_mat();
    m->ctx = c; m->rows = rows; m->cols = cols; m->owns = true; m->data = nullptr;
    cudaError_t e = cudaMalloc(&m->data, (size_t)rows * cols * sizeof(fp));
    if (e != cudaSuccess) { delete m; return c->cuda_fail(e, "cudaMalloc(matrix)"); }
    *out = m;
    return GS_OK;
}

int gs_mat_from_bytes(gs_ctx* c, const void* bytes, int64_t rows, int64_t cols, gs_mat** out) {
    if (!bytes) return c ? c->fail(GS_E_ARG, "null buffer") : GS_E_ARG;
    int rc = gs_mat_alloc(c, rows, cols, out);
    if (rc != GS_OK) return rc;
    cudaError_t e = cudaMemcpyAsync((*out)->data, bytes, (size_t)rows * cols * sizeof(fp), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { gs_mat_free(*out); *out = nullptr; return c->cuda_fail(e, "H2D copy"); }
    return GS_OK;
}

int gs_mat_to_bytes(gs_ctx* c, const gs_mat* m, void* out) {
    if (!c || !m || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    GS_CUDA(c, cudaMemcpyAsync(out, m->data, (size_t)m->rows * m->cols * sizeof(fp), cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    return GS_OK;
}

int gs_mat_shape(const gs_mat* m, int64_t* rows, int64_t* cols) {
    if (!m) return GS_E_ARG;
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    return GS_OK;
}

void* gs_mat_device_ptr(gs_mat* m) { return m ? m->data : nullptr; }

void gs_mat_free(gs_mat* m) {
    if (!m) return;
    if (m->owns && m->data) { cudaSetDevice(m->ctx->device); cudaFree(m->data); }
    delete m;
}

// ---- K1 ----------------------------------------------------------------------------------------
int gs_interpolate_roots(gs_ctx* c, const gs_mat* values, gs_mat** polys) {
    if (!c || !values || !polys) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const int log_n = ilog2_exact(values->cols);
    if (log_n < 1) return c->fail(GS_E_ARG, "domain size must be a power of two >= 2");
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, values->rows, values->cols, polys);
    if (rc != GS_OK) return rc;
    const size_t bytes = (size_t)values->rows * values->cols * sizeof(fp);
    rc = c->ensure_scratch(bytes);
    if (rc == GS_OK)
        rc = ntt_run(c, values->data, values->cols, (*polys)->data, values->cols, (fp*)c->scratch, values->cols,
                     (int)values->rows, log_n, 0, true);
    if (rc != GS_OK) { gs_mat_free(*polys); *polys = nullptr; }
    return rc;
}

int gs_eval_polys_at_roots(gs_ctx* c, const gs_mat* polys, int log2_domain, gs_mat** evals) {
    if (!c || !polys || !evals) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const int log_t = ilog2_exact(polys->cols);
    if (log_t < 1) return c->fail(GS_E_ARG, "polynomial length must be a power of two >= 2");
    if (log2_domain < log_t) return c->fail(GS_E_ARG, "domain smaller than the polynomial");
    cudaSetDevice(c->device);
    const long long n = 1ll << log2_domain;
    int rc = gs_mat_alloc(c, polys->rows, n, evals);
    if (rc != GS_OK) return rc;
    rc = c->ensure_scratch((size_t)polys->rows * n * sizeof(fp));
    if (rc == GS_OK)
        rc = ntt_run(c, polys->data, polys->cols, (*evals)->data, n, (fp*)c->scratch, n,
                     (int)polys->rows, log_t, log2_domain - log_t, false);
    if (rc != GS_OK) { gs_mat_free(*evals); *evals = nullptr; }
    return rc;
}

// ---- K2: element-wise --------------------------------------------------------------------------
int gs_vec_binary(gs_ctx* c, int op, const gs_mat* a, const gs_mat* b, const uint8_t* scalar16, gs_mat** out) {
    if (!c || !a || !out || (!b && !scalar16)) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    if (b && (b->rows != a->rows || b->cols != a->cols)) return c->fail(GS_E_ARG, "shape mismatch");
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, a->rows, a->cols, out);
    if (rc != GS_OK) return rc;
    fp s; if (scalar16) memcpy(&s, scalar16, 16);
    rc = vec_binary(c, op, a->data, b ? b->data : nullptr, (b || !scalar16) ? nullptr : &s, (*out)->data, a->rows * a->cols);
    if (rc != GS_OK) { gs_mat_free(*out); *out = nullptr; }
    return rc;
}

int gs_debug_modmul_probe(gs_ctx* c, int blocks, int iters, float* ms_out) {
    if (!c || !ms_out) return GS_E_ARG;
    cudaSetDevice(c->device);
    int rc = c->ensure_scratch(64);
    if (rc != GS_OK) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    modmul_probe_kernel<<<blocks, 256, 0, c->stream>>>((fp*)c->scratch, iters);    // warm-up
    cudaEventRecord(e0, c->stream);
    modmul_probe_kernel<<<blocks, 256, 0, c->stream>>>((fp*)c->scratch, iters);
    cudaEventRecord(e1, c->stream);
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->launches += 2;
    return GS_OK;
}

// ---- fused prover ------------------------------------------------------------------------------
int gs_stark_create(gs_ctx* c, const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                    gs_stark** out) {
    if (!c || !air_blob || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    *out = nullptr;
    if (hash_alg != HASH_SHA256 && hash_alg != HASH_BLAKE2S) return c->fail(GS_E_ARG, "Hash algorithm %d is not supported", hash_alg);
    if (exe_queries < 1 || exe_queries > 128) return c->fail(GS_E_ARG, "Execution sample size must be an integer between 1 and 128");
    if (fri_queries < 1 || fri_queries > 64) return c->fail(GS_E_ARG, "FRI sample size must be an integer between 1 and 64");
    cudaSetDevice(c->device);
    std::unique_ptr<gs_stark> S(new gs_stark());
    S->ctx = c; S->hash_alg = hash_alg; S->exe_queries = exe_queries; S->fri_queries = fri_queries;
    {
        int code = GS_OK;
        const std::string err = parse_air(air_blob, blob_len, S.get(), &code);
        if (code != GS_OK) return c->fail(code, "%s", err.c_str());
    }
    const uint32_t n_static = (uint32_t)S->statics.size();
    // device copies of the evaluation program
    int rc;
    const size_t ib = S->evaluation.instrs.size() * 16, cb = std::max<size_t>(S->evaluation.consts.size(), 1) * 16;
    if ((rc = S->d_instrs.ensure(c, ib))) return rc;
    if ((rc = S->d_consts.ensure(c, cb))) return rc;
    std::vector<fp> cs(std::max<size_t>(S->evaluation.consts.size(), 1), fp_zero());
    for (size_t i = 0; i < S->evaluation.consts.size(); ++i) cs[i] = fp_from_u128(S->evaluation.consts[i]);
    GS_CUDA(c, cudaMemcpy(S->d_instrs.p, S->evaluation.instrs.data(), ib, cudaMemcpyHostToDevice));
    GS_CUDA(c, cudaMemcpy(S->d_consts.p, cs.data(), cb, cudaMemcpyHostToDevice));
    // cyclic registers: k~ over the subgroup of order len, evaluated on the subgroup of order E*len
    S->cyc_off.assign(n_static, 0); S->cyc_mask.assign(n_static, 0);
    size_t total = 0;
    for (size_t k = 0; k < n_static; ++k) if (S->statics[k].kind == 0) { S->cyc_off[k] = total; total += S->statics[k].values.size() << S->log_e; }
    if ((rc = S->d_cyc.ensure(c, std::max<size_t>(total, 1) * sizeof(fp)))) return rc;
    for (size_t k = 0; k < n_static; ++k) {
        const StaticReg& sr = S->statics[k];
        if (sr.kind != 0) continue;
        const size_t L = sr.values.size(), EL = L << S->log_e;
        S->cyc_mask[k] = (unsigned)(EL - 1);
        int log_l = 0; while ((1u << log_l) < L) ++log_l;
        std::vector<fp> table(EL);
        if (log_l < 2) {
            // tiny cycles on the host: evaluate the interpolant of (g_L^s, v_s) at g_EL^i
            std::vector<u128> xs(L), ys(sr.values);
            const u128 gl = c->root_of_order(log_l), gel = c->root_of_order(log_l + S->log_e);
            u128 a = 1; for (size_t s2 = 0; s2 < L; ++s2) { xs[s2] = a; a = h_mul(a, gl); }
            const std::vector<u128> poly = h_interpolate(xs, ys);
            a = 1; for (size_t i = 0; i < EL; ++i) { table[i] = fp_from_u128(h_eval_poly(poly, a)); a = h_mul(a, gel); }
            GS_CUDA(c, cudaMemcpy(S->d_cyc.as<fp>() + S->cyc_off[k], table.data(), EL * sizeof(fp), cudaMemcpyHostToDevice));
        } else {
            std::vector<fp> vals(L); for (size_t i = 0; i < L; ++i) vals[i] = fp_from_u128(sr.values[i]);
            DevBuf tmp_v, tmp_p, tmp_w;
            if ((rc = tmp_v.ensure(c, L * sizeof(fp))) || (rc = tmp_p.ensure(c, L * sizeof(fp))) || (rc = tmp_w.ensure(c, EL * sizeof(fp)))) return rc;
            GS_CUDA(c, cudaMemcpy(tmp_v.p, vals.data(), L * sizeof(fp), cudaMemcpyHostToDevice));
            rc = ntt_run(c, tmp_v.as<fp>(), L, tmp_p.as<fp>(), L, tmp_w.as<fp>(), L, 1, log_l, 0, true);
            if (rc == GS_OK) rc = ntt_run(c, tmp_p.as<fp>(), L, S->d_cyc.as<fp>() + S->cyc_off[k], EL, tmp_w.as<fp>(), EL, 1, log_l, S->log_e, false);
            cudaStreamSynchronize(c->stream);
            tmp_v.release(); tmp_p.release(); tmp_w.release();
            if (rc != GS_OK) return rc;
        }
    }
    *out = S.release();
    return GS_OK;
}

/* Stark.generateExecutionTrace (genstark.d.ts:109): host only, no device needed.  out: R x T x 16 bytes */
int gs_air_generate_trace(const uint8_t* air_blob, size_t blob_len, const uint8_t* init_state16, const uint8_t* input_traces,
                          uint8_t* out_trace) {
    if (!air_blob || !init_state16 || !out_trace) return GS_E_ARG;
    Stark S; int code = GS_OK;
    const std::string err = parse_air(air_blob, blob_len, &S, &code);
    if (code != GS_OK) { g_null_error = err; return code; }
    if ((S.n_secret + S.n_public) > 0 && !input_traces) { g_null_error = "input register traces required"; return GS_E_ARG; }
    std::vector<u128> init(S.R);
    for (int r = 0; r < S.R; ++r) { fp v; memcpy(&v, init_state16 + 16 * r, 16); init[r] = fp_to_u128(v); }
    generate_trace(&S, init.data(), (const fp*)input_traces, (fp*)out_trace);
    return GS_OK;
}

void gs_stark_destroy(gs_stark* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    delete s;
}

int gs_stark_set_debug(gs_stark* s, int keep_intermediates) {
    if (!s) return GS_E_ARG;
    s->keep_intermediates = keep_intermediates != 0;
    return GS_OK;
}

int gs_stark_prove(gs_stark* s, const uint8_t* assertions, int n_assertions, const uint8_t* init_state16,
                   const uint8_t* input_traces, const uint8_t* shapes_blob, size_t shapes_len,
                   const uint8_t** proof_out, size_t* proof_len) {
    return gs_stark_prove_ex(s, assertions, n_assertions, init_state16, input_traces, shapes_blob, shapes_len, 0, proof_out, proof_len);
}

int gs_stark_prove_ex(gs_stark* s, const uint8_t* assertions, int n_assertions, const uint8_t* init_state16,
                      const uint8_t* input_traces, const uint8_t* shapes_blob, size_t shapes_len, int flags,
                      const uint8_t** proof_out, size_t* proof_len) {
    if (!s || !assertions || !init_state16 || !proof_out || !proof_len) return s ? s->ctx->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    Ctx* c = s->ctx;
    if ((s->n_secret + s->n_public) > 0 && !input_traces) return c->fail(GS_E_ARG, "input register traces required");
    std::vector<Assertion> as(n_assertions > 0 ? n_assertions : 0);
    for (int i = 0; i < n_assertions; ++i) {
        const uint8_t* p = assertions + 24 * (size_t)i;
        memcpy(&as[i].reg, p, 4); memcpy(&as[i].step, p + 4, 4);
        fp v; memcpy(&v, p + 8, 16); as[i].value = fp_to_u128(v);
    }
    std::vector<u128> init(s->R);
    for (int r = 0; r < s->R; ++r) { fp v; memcpy(&v, init_state16 + 16 * r, 16); init[r] = fp_to_u128(v); if (init[r] >= HP) return c->fail(GS_E_ARG, "non-canonical initial state"); }
    int rc = stark_prove(s, as.data(), n_assertions, init.data(), (const fp*)input_traces, shapes_blob, shapes_len, s->proof, flags);
    if (rc != GS_OK) return rc;
    *proof_out = s->proof.data(); *proof_len = s->proof.size();
    return GS_OK;
}

int gs_stark_last_timing(gs_stark* s, float* device_ms, double* host_ms) {
    if (!s) return GS_E_ARG;
    if (device_ms) *device_ms = s->last_device_ms;
    if (host_ms) *host_ms = s->last_host_ms;
    return GS_OK;
}

int gs_ctx_profile(gs_ctx* c, int on) {
    if (!c) return GS_E_ARG;
    c->profiling = on != 0;
    if (on) { c->prof_acc.clear(); }
    return GS_OK;
}

const char* gs_ctx_profile_report(gs_ctx* c) {
    if (!c) return "";
    cudaStreamSynchronize(c->stream);
    c->prof_collect();
    std::string j = "{";
    bool first = true;
    for (auto& kv : c->prof_acc) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s\"%s\": {\"groups\": %llu, \"ms\": %.5f}", first ? "" : ", ", c->prof_names[kv.first].c_str(), kv.second.first, kv.second.second);
        j += buf; first = false;
    }
    c->prof_json = j + "}";
    return c->prof_json.c_str();
}

int gs_timer_begin(gs_ctx* c) {
    if (!c) return GS_E_ARG;
    if (!c->timer_a) { cudaEventCreate(&c->timer_a); cudaEventCreate(&c->timer_b); }
    GS_CUDA(c, cudaEventRecord(c->timer_a, c->stream));
    return GS_OK;
}

int gs_timer_end(gs_ctx* c, float* ms) {
    if (!c || !ms || !c->timer_a) return GS_E_ARG;
    GS_CUDA(c, cudaEventRecord(c->timer_b, c->stream));
    GS_CUDA(c, cudaEventSynchronize(c->timer_b));
    GS_CUDA(c, cudaEventElapsedTime(ms, c->timer_a, c->timer_b));
    return GS_OK;
}

/* in-place transform into a caller-provided output (no allocation inside the timed region) */
int gs_ntt_into(gs_ctx* c, const gs_mat* src, gs_mat* dst, gs_mat* work, int inverse) {
    if (!c || !src || !dst) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const int log_t = ilog2_exact(src->cols), log_n = ilog2_exact(dst->cols);
    if (log_t < 1 || log_n < log_t || dst->rows != src->rows) return c->fail(GS_E_ARG, "shape mismatch");
    if (work && (work->rows != dst->rows || work->cols != dst->cols)) return c->fail(GS_E_ARG, "work shape mismatch");
    cudaSetDevice(c->device);
    return ntt_run(c, src->data, src->cols, dst->data, dst->cols, work ? work->data : nullptr, dst->cols, (int)src->rows, log_t,
                   log_n - log_t, inverse != 0);
}

const char* gs_stark_stage_times(gs_stark* s) {
    if (!s) return "";
    std::string j = "[";
    for (size_t i = 0; i < s->last_times.items.size(); ++i) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s[\"%s\", %.4f]", i ? ", " : "", s->last_times.items[i].first.c_str(), s->last_times.items[i].second);
        j += buf;
    }
    s->times_json = j + "]";
    return s->times_json.c_str();
}

/* which: 0 = P(x) evaluations (R x N), 1 = C(x) (N), 2 = L(x) (N), 3 = trace polynomials (R x T) */
int gs_stark_read_intermediate(gs_stark* s, int which, void* out, size_t out_bytes) {
    if (!s || !out) return GS_E_ARG;
    Ctx* c = s->ctx;
    const size_t N = (size_t)1 << (s->log_t + s->log_e), T = (size_t)1 << s->log_t;
    const void* src = nullptr; size_t bytes = 0;
    switch (which) {
        case 0: src = s->d_pe.p; bytes = (size_t)s->R * N * 16; break;
        case 1: src = s->d_c.p; bytes = N * 16; break;
        case 2: src = s->d_l.p; bytes = N * 16; break;
        case 3: src = s->d_poly.p; bytes = (size_t)s->R * T * 16; break;
        default: return c->fail(GS_E_ARG, "unknown intermediate");
    }
    if (!src || out_bytes < bytes) return c->fail(GS_E_ARG, "intermediate not available (enable gs_stark_set_debug) or buffer too small");
    GS_CUDA(c, cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return GS_OK;
}

}  // extern "C"
