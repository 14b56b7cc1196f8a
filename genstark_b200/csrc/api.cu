// extern "C" surface of libgenstark_b200.so (declared in include/genstark_b200.h).
#include "core.cuh"
#include "ntt_host.cuh"
#include "pointwise.cuh"
#include "../../include/genstark_b200.h"

using namespace gs;

struct gs_ctx : public Ctx {};
struct gs_mat : public Mat {};

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
    c->mailbox_bytes = 1 << 20;
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

}  // extern "C"
