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
struct gs_digests { Ctx* ctx; uint32_t* data; long long n; };
struct gs_tree { Ctx* ctx; uint32_t* nodes; long long n; int alg; };
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
    if (cudaMalloc(&c->counters, 256) != cudaSuccess || cudaMemset(c->counters, 0, 256) != cudaSuccess) { g_null_error = "cudaMalloc(counters)"; delete c; return GS_E_CUDA; }
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
    for (auto& kv : c->tw_tables) cudaFree(kv.second);
    if (c->ntt2_xs) cudaFree(c->ntt2_xs);
    c->tw_tables.clear();
    if (c->scratch) cudaFree(c->scratch);
    if (c->counters) cudaFree(c->counters);
    if (c->mailbox) cudaFreeHost(c->mailbox);
    if (c->comm) nccl().CommDestroy(c->comm);
    cudaStreamDestroy(c->stream);
    delete c;
}

int gs_comm_unique_id(uint8_t out128[128]) {
    if (!out128) return GS_E_ARG;
    if (!nccl().load()) { g_null_error = nccl().error; return GS_E_UNSUPPORTED; }
    NcclUniqueId id;
    const int rc = nccl().GetUniqueId(&id);
    if (rc != 0) { g_null_error = std::string("ncclGetUniqueId: ") + nccl().GetErrorString(rc); return GS_E_CUDA; }
    memcpy(out128, &id, 128);
    return GS_OK;
}

int gs_ctx_comm_init(gs_ctx* c, int rank, int world, const uint8_t id128[128]) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world || (world & (world - 1))) return c ? c->fail(GS_E_ARG, "bad rank / world size (power of two required)") : GS_E_ARG;
    if (world == 1) { c->rank = 0; c->world = 1; return GS_OK; }
    if (!nccl().load()) return c->fail(GS_E_UNSUPPORTED, "%s", nccl().error.c_str());
    cudaSetDevice(c->device);
    NcclUniqueId id; memcpy(&id, id128, 128);
    const int rc = nccl().CommInitRank(&c->comm, world, id, rank);
    if (rc != 0) return c->fail(GS_E_CUDA, "ncclCommInitRank: %s", nccl().GetErrorString(rc));
    c->rank = rank; c->world = world;
    return GS_OK;
}

/* the coset-sharding index map (host logic of the multi-GPU prover): position i = q*E + j is owned by rank j / (E/world) */
int gs_shard_map(int world, int rank, int log2_e, int64_t index, int to_local, int64_t* out_index, int* out_owner) {
    int log_w = 0; while ((1 << log_w) < world) ++log_w;
    if (world < 1 || (1 << log_w) != world || log_w > log2_e || rank < 0 || rank >= world || index < 0) return GS_E_ARG;
    Shard sh; sh.rank = rank; sh.world = world; sh.log_e = log2_e; sh.log_el = log2_e - log_w;
    if (to_local) { if (out_owner) *out_owner = sh.owner(index); if (out_index) *out_index = sh.to_local(index); }
    else { if (out_owner) *out_owner = rank; if (out_index) *out_index = sh.to_global(index); }
    return GS_OK;
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
    // rows * cols * 16 must not wrap: nothing on this path is larger than 2^36 elements (1 TiB), far beyond the device
    if (rows > (1ll << 36) || cols > (1ll << 36) || (unsigned __int128)rows * (unsigned __int128)cols > ((unsigned __int128)1 << 36))
        return c->fail(GS_E_ARG, "matrix of %lld x %lld elements is too large", (long long)rows, (long long)cols);
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
    if (log2_domain > c->log_g) return c->fail(GS_E_UNSUPPORTED, "domain 2^%d exceeds 2^%d", log2_domain, c->log_g);
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

// expVectorElements(a, e) with a non-negative exponent below 2^128 (16 little-endian bytes); the Python / TypeScript wrappers turn a
// negative exponent into the inverse first, as galois does
int gs_vec_exp(gs_ctx* c, const gs_mat* a, const uint8_t* exponent16, gs_mat** out) {
    if (!c || !a || !exponent16 || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, a->rows, a->cols, out);
    if (rc != GS_OK) return rc;
    fp e; memcpy(&e, exponent16, 16);
    const long long n = a->rows * a->cols;
    long long blocks = (n + 255) / 256; const long long cap = (long long)c->sm_count * 16; if (blocks > cap) blocks = cap;
    vec_exp_kernel<<<(unsigned)(blocks > 0 ? blocks : 1), 256, 0, c->stream>>>(a->data, e, (*out)->data, n);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { gs_mat_free(*out); *out = nullptr; return c->cuda_fail(err, "vec_exp_kernel"); }
    c->launches++;
    return GS_OK;
}
// mulMatrixByVector(m, v): m is rows x cols, v holds cols elements (any shape); out is a vector of `rows` elements
int gs_mat_mul_vector(gs_ctx* c, const gs_mat* m, const gs_mat* v, gs_mat** out) {
    if (!c || !m || !v || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    if (v->rows * v->cols != m->cols) return c->fail(GS_E_ARG, "vector length %lld does not match the matrix (%lld columns)", v->rows * v->cols, m->cols);
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, 1, m->rows, out);
    if (rc != GS_OK) return rc;
    mat_vec_kernel<<<(unsigned)((m->rows + 127) / 128), 128, 0, c->stream>>>(m->data, v->data, (*out)->data, m->rows, m->cols);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { gs_mat_free(*out); *out = nullptr; return c->cuda_fail(err, "mat_vec_kernel"); }
    c->launches++;
    return GS_OK;
}

// divVectorElements(a, b) = a * inv(b) with inv(0) = 0   (CompositionPolynomial.ts:117, BoundaryConstraints.ts:92)
int gs_vec_div(gs_ctx* c, const gs_mat* a, const gs_mat* b, gs_mat** out) {
    if (!c || !a || !b || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    if (b->rows != a->rows || b->cols != a->cols) return c->fail(GS_E_ARG, "shape mismatch");
    cudaSetDevice(c->device);
    const long long n = a->rows * a->cols;
    int rc = gs_mat_alloc(c, a->rows, a->cols, out);
    if (rc != GS_OK) return rc;
    rc = c->ensure_scratch((size_t)n * sizeof(fp));
    if (rc == GS_OK) rc = batch_inverse(c, b->data, (*out)->data, (fp*)c->scratch, n);
    if (rc == GS_OK) rc = vec_binary(c, VOP_MUL, a->data, (*out)->data, nullptr, (*out)->data, n);
    if (rc != GS_OK) { gs_mat_free(*out); *out = nullptr; }
    return rc;
}

int gs_vec_combine_many(gs_ctx* c, const gs_mat* const* vectors, int count, const uint8_t* coefficients16, gs_mat** out) {
    if (!c || !vectors || !coefficients16 || !out || count < 1 || count > 64) return c ? c->fail(GS_E_ARG, "1..64 vectors required") : GS_E_ARG;
    cudaSetDevice(c->device);
    CombineParams P; P.m = count;
    if (!vectors[0]) return c->fail(GS_E_ARG, "null vector");
    const long long n = vectors[0]->rows * vectors[0]->cols;
    for (int m = 0; m < count; ++m) {
        if (!vectors[m] || vectors[m]->rows * vectors[m]->cols != n) return c->fail(GS_E_ARG, "shape mismatch");
        P.v[m] = vectors[m]->data; memcpy(&P.k[m], coefficients16 + 16 * m, 16);
    }
    int rc = gs_mat_alloc(c, 1, n, out);
    if (rc != GS_OK) return rc;
    rc = c->ensure_scratch(sizeof P);
    if (rc != GS_OK) { gs_mat_free(*out); *out = nullptr; return rc; }
    cudaError_t e = cudaMemcpyAsync(c->scratch, &P, sizeof P, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        combine_many_kernel<<<grid_for(c, n, 256), 256, 0, c->stream>>>((const CombineParams*)c->scratch, (*out)->data, n);
        c->launches++;
        e = cudaStreamSynchronize(c->stream);
    }
    if (e != cudaSuccess) { gs_mat_free(*out); *out = nullptr; return c->cuda_fail(e, "combine_many_kernel"); }
    return GS_OK;
}

int gs_power_series(gs_ctx* c, const uint8_t base16[16], int64_t n, gs_mat** out) {
    if (!c || !base16 || !out || n < 1) return c ? c->fail(GS_E_ARG, "bad argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, 1, n, out);
    if (rc != GS_OK) return rc;
    fp b; memcpy(&b, base16, 16);
    long long threads = n < 65536 ? n : 65536;
    power_series_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(b, (*out)->data, n);
    c->launches++;
    return GS_OK;
}

int gs_pluck_vector(gs_ctx* c, const gs_mat* v, int64_t skip, int64_t times, gs_mat** out) {
    if (!c || !v || !out || skip < 0 || times < 1) return c ? c->fail(GS_E_ARG, "bad argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, 1, times, out);
    if (rc != GS_OK) return rc;
    pluck_kernel<<<(unsigned)((times + 255) / 256), 256, 0, c->stream>>>(v->data, v->rows * v->cols, skip, (*out)->data, times);
    c->launches++;
    return GS_OK;
}

int gs_transpose_vector(gs_ctx* c, const gs_mat* v, int columns, int64_t step, gs_mat** out) {
    if (!c || !v || !out || columns < 1 || step < 1) return c ? c->fail(GS_E_ARG, "bad argument") : GS_E_ARG;
    const long long len = v->rows * v->cols;
    if (len % ((long long)columns * step)) return c->fail(GS_E_ARG, "length not divisible by columns * step");
    cudaSetDevice(c->device);
    const long long rows = len / ((long long)columns * step);
    int rc = gs_mat_alloc(c, rows, columns, out);
    if (rc != GS_OK) return rc;
    transpose_vector_kernel<<<(unsigned)((rows * columns + 255) / 256), 256, 0, c->stream>>>(v->data, rows, columns, step, (*out)->data);
    c->launches++;
    return GS_OK;
}

// one FRI layer: interpolateQuarticBatch(transposeVector(domain,4,4^depth), transposeVector(v,4)) evaluated at x*
int gs_fri_fold(gs_ctx* c, const gs_mat* v, int log2_domain, int depth, const uint8_t special_x16[16], gs_mat** column) {
    if (!c || !v || !special_x16 || !column) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const long long L = v->rows * v->cols;
    if (depth < 0 || depth > 30 || log2_domain < 2 || log2_domain > c->log_g) return c->fail(GS_E_ARG, "bad depth / domain");
    if (L < 4 || (L & 3) || 2 * depth > log2_domain || L != (1ll << (log2_domain - 2 * depth))) return c->fail(GS_E_ARG, "layer length must be domain / 4^depth");
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, 1, L >> 2, column);
    if (rc != GS_OK) return rc;
    rc = c->ensure_scratch(64);
    if (rc != GS_OK) { gs_mat_free(*column); *column = nullptr; return rc; }
    GS_CUDA(c, cudaMemcpyAsync(c->scratch, special_x16, 16, cudaMemcpyHostToDevice, c->stream));
    FriFoldParams F; F.v = v->data; F.out = (*column)->data; F.quarter = L >> 2; F.special_x = (const fp*)c->scratch;
    F.tw_lo = c->tw_lo; F.tw_hi = c->tw_hi; F.log_g = c->log_g; F.log_lo = c->log_lo;
    F.x_shift = 2 * depth + (c->log_g - log2_domain);
    F.log_e = 0; F.log_el = 0; F.j0 = 0;          // single GPU: local row index = global row index
    F.iota_inv = fp_from_u128(h_inv(c->root_of_order(2))); F.quarter_inv = fp_from_u128(h_inv(4));
    fri_fold_kernel<<<grid_for(c, L >> 2, 256), 256, 0, c->stream>>>(F);
    c->launches++;
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    return GS_OK;
}

// ---- K5: hashing and Merkle trees ----------------------------------------------------------------
static int digests_alloc(gs_ctx* c, long long n, gs_digests** out) {
    gs_digests* d = new gs_digests(); d->ctx = c; d->n = n; d->data = nullptr;
    cudaError_t e = cudaMalloc(&d->data, (size_t)n * 32);
    if (e != cudaSuccess) { delete d; return c->cuda_fail(e, "cudaMalloc(digests)"); }
    *out = d; return GS_OK;
}

int gs_hash_merge_vector_rows(gs_ctx* c, int alg, const gs_mat* const* mats, int count, gs_digests** out) {
    if (!c || !mats || !out || count < 1) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    HashCols hc; hc.ncols = 0;
    const long long n = mats[0]->cols;
    for (int m = 0; m < count; ++m) {
        if (!mats[m] || mats[m]->cols != n) return c->fail(GS_E_ARG, "all vectors must have the same length");
        for (long long r = 0; r < mats[m]->rows; ++r) {
            if (hc.ncols >= GS_MAX_HASH_COLS) return c->fail(GS_E_UNSUPPORTED, "more than %d columns", GS_MAX_HASH_COLS);
            hc.col[hc.ncols++] = mats[m]->data + r * n;
        }
    }
    int rc = digests_alloc(c, n, out);
    if (rc != GS_OK) return rc;
    rc = hash_columns(c, alg, hc, n, (*out)->data);
    if (rc != GS_OK) { gs_digests_free(*out); *out = nullptr; }
    return rc;
}

int gs_hash_digest_values(gs_ctx* c, int alg, const gs_mat* rows, gs_digests** out) {
    if (!c || !rows || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    int rc = digests_alloc(c, rows->rows, out);
    if (rc != GS_OK) return rc;
    rc = hash_rows(c, alg, rows->data, (int)(rows->cols * 16), rows->rows, (*out)->data);
    if (rc != GS_OK) { gs_digests_free(*out); *out = nullptr; }
    return rc;
}

int gs_digests_to_bytes(gs_ctx* c, const gs_digests* d, void* out) {
    if (!c || !d || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    GS_CUDA(c, cudaMemcpyAsync(out, d->data, (size_t)d->n * 32, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    return GS_OK;
}
int64_t gs_digests_count(const gs_digests* d) { return d ? d->n : 0; }
void gs_digests_free(gs_digests* d) { if (!d) return; cudaSetDevice(d->ctx->device); cudaFree(d->data); delete d; }

int gs_merkle_create(gs_ctx* c, int alg, const gs_digests* leaves, gs_tree** out) {
    if (!c || !leaves || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const long long n = leaves->n;
    if (n < 2 || (n & (n - 1))) return c->fail(GS_E_ARG, "leaf count must be a power of two >= 2");
    cudaSetDevice(c->device);
    gs_tree* t = new gs_tree(); t->ctx = c; t->n = n; t->alg = alg; t->nodes = nullptr;
    cudaError_t e = cudaMalloc(&t->nodes, (size_t)2 * n * 32);
    if (e != cudaSuccess) { delete t; return c->cuda_fail(e, "cudaMalloc(tree)"); }
    cudaMemsetAsync(t->nodes, 0, 32, c->stream);
    cudaMemcpyAsync(t->nodes + 8 * n, leaves->data, (size_t)n * 32, cudaMemcpyDeviceToDevice, c->stream);
    int rc = merkle_commit(c, alg, nullptr, t->nodes, n);
    if (rc != GS_OK) { cudaFree(t->nodes); delete t; return rc; }
    *out = t;
    return GS_OK;
}

/* test hook: hash.mergeVectorRows + MerkleTree.create the way the prover commits (leaves and tree in the same launches) */
int gs_debug_commit_columns(gs_ctx* c, int alg, const gs_mat* const* mats, int count, gs_tree** out) {
    if (!c || !mats || !out || count < 1) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    HashCols hc; memset(&hc, 0, sizeof hc);
    const long long n = mats[0] ? mats[0]->cols : 0;
    if (n < 1 || (n & (n - 1))) return c->fail(GS_E_ARG, "leaf count must be a power of two");
    for (int m = 0; m < count; ++m) {
        if (!mats[m] || mats[m]->cols != n) return c->fail(GS_E_ARG, "all vectors must have the same length");
        for (long long r = 0; r < mats[m]->rows; ++r) {
            if (hc.ncols >= GS_MAX_HASH_COLS) return c->fail(GS_E_UNSUPPORTED, "more than %d columns", GS_MAX_HASH_COLS);
            hc.col[hc.ncols++] = mats[m]->data + r * n;
        }
    }
    gs_tree* t = new gs_tree(); t->ctx = c; t->n = n; t->alg = alg; t->nodes = nullptr;
    cudaError_t e = cudaMalloc(&t->nodes, (size_t)2 * n * 32);
    if (e != cudaSuccess) { delete t; return c->cuda_fail(e, "cudaMalloc(tree)"); }
    cudaMemsetAsync(t->nodes, 0, 32, c->stream);
    int rc = merkle_commit(c, alg, &hc, t->nodes, n);
    if (rc != GS_OK) { cudaFree(t->nodes); delete t; return rc; }
    *out = t;
    return GS_OK;
}

/* test hook: the 2n digests of a tree as stored (slot 0 unused and zero, nodes[1] = root, leaves at [n, 2n)) */
int gs_debug_tree_nodes(gs_ctx* c, const gs_tree* t, void* out, size_t out_bytes) {
    if (!c || !t || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const size_t bytes = (size_t)2 * t->n * 32;
    if (out_bytes < bytes) return c->fail(GS_E_ARG, "buffer too small for %lld digests", 2 * t->n);
    cudaSetDevice(c->device);
    GS_CUDA(c, cudaMemcpyAsync(out, t->nodes, bytes, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    return GS_OK;
}

int gs_merkle_root(gs_ctx* c, const gs_tree* t, uint8_t out32[32]) {
    if (!c || !t || !out32) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    GS_CUDA(c, cudaMemcpyAsync(out32, t->nodes + 8, 32, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    return GS_OK;
}

// proveBatch(indexes): blob = u32 n_values | u32 n_columns | u32 depth | values (32 B leaf digests, input order) |
//                             per column: u32 length, then length x 32 B
int gs_merkle_prove_batch(gs_ctx* c, const gs_tree* t, const uint32_t* indexes, int count, uint8_t* out, size_t out_cap, size_t* out_len) {
    if (!c || !t || !indexes || !out_len || count < 1) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    std::vector<uint32_t> idx(indexes, indexes + count);
    BatchProof bp; std::string err;
    if (merkle_prove_plan(idx, (uint64_t)t->n, bp, err) != 0) return c->fail(GS_E_ARG, "%s", err.c_str());
    std::vector<unsigned long long> addr;
    for (uint32_t i : idx) { addr.push_back((unsigned long long)(uintptr_t)(t->nodes + 8ull * (t->n + i))); addr.push_back((unsigned long long)(uintptr_t)(t->nodes + 8ull * (t->n + i) + 4)); }
    for (auto& col : bp.node_ids) for (uint32_t id : col) { addr.push_back((unsigned long long)(uintptr_t)(t->nodes + 8ull * id)); addr.push_back((unsigned long long)(uintptr_t)(t->nodes + 8ull * id + 4)); }
    const size_t nch = addr.size();
    if (nch * 16 > c->mailbox_bytes) return c->fail(GS_E_ARG, "too many indexes");
    int rc = c->ensure_scratch(nch * 24);
    if (rc != GS_OK) return rc;
    uint8_t* d_addr = (uint8_t*)c->scratch; uint8_t* d_out = d_addr + nch * 8;
    GS_CUDA(c, cudaMemcpyAsync(d_addr, addr.data(), nch * 8, cudaMemcpyHostToDevice, c->stream));
    gather_chunks_kernel<<<(unsigned)((nch + 255) / 256), 256, 0, c->stream>>>((const unsigned long long*)d_addr, (int)nch, (uint4*)d_out);
    c->launches++;
    GS_CUDA(c, cudaMemcpyAsync(c->mailbox, d_out, nch * 16, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<uint8_t> blob;
    auto put32 = [&](uint32_t v) { const uint8_t* p = (const uint8_t*)&v; blob.insert(blob.end(), p, p + 4); };
    put32((uint32_t)count); put32((uint32_t)bp.node_ids.size()); put32((uint32_t)bp.depth);
    const uint8_t* mb = (const uint8_t*)c->mailbox;
    blob.insert(blob.end(), mb, mb + (size_t)count * 32);
    size_t k = (size_t)count * 32;
    for (auto& col : bp.node_ids) { put32((uint32_t)col.size()); blob.insert(blob.end(), mb + k, mb + k + col.size() * 32); k += col.size() * 32; }
    *out_len = blob.size();
    if (!out || out_cap < blob.size()) return c->fail(GS_E_ARG, "output buffer too small (%zu bytes needed)", blob.size());
    memcpy(out, blob.data(), blob.size());
    return GS_OK;
}
void gs_tree_free(gs_tree* t) { if (!t) return; cudaSetDevice(t->ctx->device); cudaFree(t->nodes); delete t; }

static int run_probe(gs_ctx* c, int kind, int blocks, int iters, float* ms_out) {
    if (!c || !ms_out) return GS_E_ARG;
    if (blocks < 1 || blocks > (1 << 20) || iters < 1 || iters > (1 << 24)) return c->fail(GS_E_ARG, "probe shape out of range");
    cudaSetDevice(c->device);
    int rc = c->ensure_scratch(64);
    if (rc != GS_OK) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int pass = 0; pass < 2; ++pass) {          // pass 0 = warm-up
        if (pass == 1) cudaEventRecord(e0, c->stream);
        if (kind == 0) modmul_probe_kernel<<<blocks, 256, 0, c->stream>>>((fp*)c->scratch, iters);
        else butterfly_probe_kernel<<<blocks, 256, 0, c->stream>>>((fp*)c->scratch, iters);
    }
    cudaEventRecord(e1, c->stream);
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->launches += 2;
    return GS_OK;
}
/* blocks x 256 threads x 4 chains x iters squarings (fp_sqr); *mismatches = disagreements of fp_sqr with fp_mul(a, a) on edge and chain values */
int gs_debug_sqr_probe(gs_ctx* c, int blocks, int iters, float* ms_out, int* mismatches) {
    if (!c || !ms_out || !mismatches) return GS_E_ARG;
    if (blocks < 1 || blocks > (1 << 20) || iters < 1 || iters > (1 << 24)) return c->fail(GS_E_ARG, "probe shape out of range");
    cudaSetDevice(c->device);
    int rc = c->ensure_scratch(256);
    if (rc != GS_OK) return rc;
    unsigned* d_bad = (unsigned*)((uint8_t*)c->scratch + 128);
    GS_CUDA(c, cudaMemsetAsync(d_bad, 0, 4, c->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    sqr_probe_kernel<<<blocks, 256, 0, c->stream>>>((fp*)c->scratch, iters, d_bad);          // warm-up + check
    cudaEventRecord(e0, c->stream);
    sqr_probe_kernel<<<blocks, 256, 0, c->stream>>>((fp*)c->scratch, iters, nullptr);
    cudaEventRecord(e1, c->stream);
    GS_CUDA(c, cudaMemcpyAsync(mismatches, d_bad, 4, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->launches += 2;
    return GS_OK;
}
/* blocks x 256 threads x 4 independent chains x iters modular multiplications */
int gs_debug_modmul_probe(gs_ctx* c, int blocks, int iters, float* ms_out) { return run_probe(c, 0, blocks, iters, ms_out); }
/* blocks x 256 threads x 2 butterflies (u + v, (u - v) * w) x iters */
int gs_debug_butterfly_probe(gs_ctx* c, int blocks, int iters, float* ms_out) { return run_probe(c, 1, blocks, iters, ms_out); }

// ---- fused prover ------------------------------------------------------------------------------
static int stark_create_impl(gs_ctx* c, const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries, gs_stark** out);
int gs_stark_create(gs_ctx* c, const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                    gs_stark** out) {
    if (!c || !air_blob || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    *out = nullptr;
    // nothing may unwind through the C ABI (an AIR blob with absurd sizes ends in std::bad_alloc / std::length_error)
    try { return stark_create_impl(c, air_blob, blob_len, hash_alg, exe_queries, fri_queries, out); }
    catch (const std::exception& e) { *out = nullptr; return c->fail(GS_E_ARG, "instantiation failed: %s", e.what()); }
}
static int stark_create_impl(gs_ctx* c, const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                             gs_stark** out) {
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
    // u[j] = 1/(w_N^j - 1) over the evaluation domain (boundary constraints by partial fractions, compose.cuh)
    {
        const int log_n = S->log_t + S->log_e;
        if (log_n > c->log_g) return c->fail(GS_E_UNSUPPORTED, "evaluation domain 2^%d exceeds 2^%d", log_n, c->log_g);
        // coset sharding: this rank owns E / world consecutive cosets
        int log_w = 0; while ((1 << log_w) < c->world) ++log_w;
        if (log_w > S->log_e) return c->fail(GS_E_UNSUPPORTED, "more ranks (%d) than cosets (%d)", c->world, 1 << S->log_e);
        S->shard.rank = c->rank; S->shard.world = c->world; S->shard.log_e = S->log_e; S->shard.log_el = S->log_e - log_w;
        const long long N = (1ll << S->log_t) << S->shard.log_el;          // local positions
        if ((rc = S->d_u.ensure(c, (size_t)N * sizeof(fp)))) return rc;
        if ((rc = c->ensure_scratch((size_t)N * sizeof(fp)))) return rc;
        UTableParams U; U.n_loc = N; U.log_n = log_n; U.log_e = S->log_e; U.log_el = S->shard.log_el; U.j0 = S->shard.j0(); U.tw_lo = c->tw_lo; U.tw_hi = c->tw_hi; U.log_g = c->log_g; U.log_lo = c->log_lo; U.out = S->d_u.as<fp>();
        u_table_kernel<<<grid_for(c, N, 256), 256, 0, c->stream>>>(U);
        c->launches++;
        if ((rc = batch_inverse(c, S->d_u.as<fp>(), S->d_u.as<fp>(), (fp*)c->scratch, N))) return rc;
        GS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    S->compose_jit = compose_jit_get(c, *S);   // K2 specialised for this AIR (NVRTC); the interpreting kernel if that is not possible
    trace_prepare(S.get());     // the transition function as native code (hostjit.h); the interpreter if that is not possible
    *out = S.release();
    return GS_OK;
}

/* Stark.generateExecutionTrace (genstark.d.ts:109): host only, no device needed.  out: R x T x 16 bytes */
int gs_air_generate_trace(const uint8_t* air_blob, size_t blob_len, const uint8_t* init_state16, const uint8_t* input_traces,
                          uint8_t* out_trace) {
    if (!air_blob || !init_state16 || !out_trace) return GS_E_ARG;
    try {
    Stark S; int code = GS_OK;
    const std::string err = parse_air(air_blob, blob_len, &S, &code);
    if (code != GS_OK) { g_null_error = err; return code; }
    if ((S.n_secret + S.n_public) > 0 && !input_traces) { g_null_error = "input register traces required"; return GS_E_ARG; }
    std::vector<u128> init(S.R);
    for (int r = 0; r < S.R; ++r) { fp v; memcpy(&v, init_state16 + 16 * r, 16); init[r] = fp_to_u128(v); }
    generate_trace(&S, init.data(), (const fp*)input_traces, (fp*)out_trace);
    return GS_OK;
    } catch (const std::exception& e) { g_null_error = std::string("trace generation failed: ") + e.what(); return GS_E_ARG; }
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
    try {
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
    } catch (const std::exception& e) { cudaStreamSynchronize(c->stream); return c->fail(GS_E_CUDA, "prove failed: %s", e.what()); }
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

int gs_lde_cosets_into(gs_ctx* c, const gs_mat* src, gs_mat* dst, gs_mat* work, int coset_base, int log_e_total) {
    if (!c || !src || !dst || !work) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const int log_t = ilog2_exact(src->cols), log_n = ilog2_exact(dst->cols);
    if (log_t < 2 || log_n < log_t || dst->rows != src->rows || work->rows != dst->rows || work->cols != dst->cols)
        return c->fail(GS_E_ARG, "shape mismatch");
    const int log_e = log_n - log_t;
    if (log_e_total < log_e || log_e_total < 1 || coset_base < 0 || coset_base + (1 << log_e) > (1 << log_e_total))
        return c->fail(GS_E_ARG, "coset range outside the evaluation domain");
    cudaSetDevice(c->device);
    return ntt_run(c, src->data, src->cols, dst->data, dst->cols, work->data, dst->cols, (int)src->rows, log_t, log_e, false,
                   coset_base, log_e_total);
}

namespace gs {
__global__ void fill_random_kernel(fp* __restrict__ out, size_t n, unsigned long long seed) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long w[2];
    for (int k = 0; k < 2; ++k) {                 // SplitMix64 at counter seed + 2i + k
        unsigned long long z = seed + (2ull * i + k + 1ull) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        w[k] = z ^ (z >> 31);
    }
    fp v; v.v[0] = (uint32_t)w[0]; v.v[1] = (uint32_t)(w[0] >> 32); v.v[2] = (uint32_t)w[1]; v.v[3] = (uint32_t)(w[1] >> 32);
    st_fp(out + i, fp_add(v, fp_zero()));         // the modular addition canonicalises (v < 2^128 < 2p)
}
}  // namespace gs

int gs_mat_fill_random(gs_ctx* c, gs_mat* m, uint64_t seed) {
    if (!c || !m) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    const size_t n = (size_t)m->rows * (size_t)m->cols;
    if (!n) return GS_OK;
    fill_random_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(m->data, n, seed);
    GS_CUDA(c, cudaGetLastError());
    c->launches++;
    return GS_OK;
}

const char* gs_stark_last_error(gs_stark* s) { return (s && s->ctx) ? s->ctx->last_error.c_str() : ""; }

const char* gs_stark_compose_backend(gs_stark* s) {
    static thread_local std::string out;
    out = (s && s->compose_jit) ? s->compose_jit->status : std::string("interpreter");
    return out.c_str();
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

// ---- rest of the FiniteField / Hash seam: the small-vector and reshaping methods ---------------------
/* field.prng(seed) (count == 0: one element) / field.prng(seed, count)   (CompositionPolynomial.ts:58, LowDegreeProver.ts:132,194) */
int gs_field_prng(const uint8_t* seed, size_t seed_len, int count, uint8_t* out16) {
    if (!seed || !out16 || count < 0) return GS_E_ARG;
    if (count == 0) { fp v = fp_from_u128(prng_one(seed, seed_len)); memcpy(out16, &v, 16); return GS_OK; }
    const std::vector<u128> v = prng_many(seed, seed_len, count);
    for (int i = 0; i < count; ++i) { fp f = fp_from_u128(v[i]); memcpy(out16 + 16 * (size_t)i, &f, 16); }
    return GS_OK;
}

static u128 elem_in(const uint8_t* p) { fp f; memcpy(&f, p, 16); return fp_to_u128(f); }
static std::vector<u128> elems_in(const uint8_t* p, int n) { std::vector<u128> v(n); for (int i = 0; i < n; ++i) v[i] = elem_in(p + 16 * (size_t)i); return v; }
static void elems_out(const std::vector<u128>& v, uint8_t* p) { for (size_t i = 0; i < v.size(); ++i) { fp f = fp_from_u128(v[i]); memcpy(p + 16 * i, &f, 16); } }

/* field.interpolate(xs, ys): Lagrange, coefficients low -> high (BoundaryConstraints.ts:42, LowDegreeProver.ts:243); host, n <= 4096 */
int gs_poly_interpolate(const uint8_t* xs16, const uint8_t* ys16, int n, uint8_t* out16) {
    if (!xs16 || !ys16 || !out16 || n < 1 || n > 4096) return GS_E_ARG;
    elems_out(h_interpolate(elems_in(xs16, n), elems_in(ys16, n)), out16);
    return GS_OK;
}
/* field.evalPolyAt(poly, x)   (BoundaryConstraints.ts:59-60, LowDegreeProver.ts:248) */
int gs_poly_eval_at(const uint8_t* poly16, int n, const uint8_t x16[16], uint8_t out16[16]) {
    if (!poly16 || !x16 || !out16 || n < 1) return GS_E_ARG;
    fp f = fp_from_u128(h_eval_poly(elems_in(poly16, n), elem_in(x16)));
    memcpy(out16, &f, 16);
    return GS_OK;
}
/* field.mulPolys(a, b)   (BoundaryConstraints.ts:30); out: na + nb - 1 coefficients */
int gs_poly_mul(const uint8_t* a16, int na, const uint8_t* b16, int nb, uint8_t* out16) {
    if (!a16 || !b16 || !out16 || na < 1 || nb < 1 || (long long)na * nb > (1ll << 26)) return GS_E_ARG;
    const std::vector<u128> a = elems_in(a16, na), b = elems_in(b16, nb);
    std::vector<u128> o((size_t)na + nb - 1, 0);
    for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) o[i + j] = h_add(o[i + j], h_mul(a[i], b[j]));
    elems_out(o, out16);
    return GS_OK;
}

/* field.combineVectors(a, b) = sum a[i]*b[i]   (CompositionPolynomial.ts:168,188; LinearCombination.ts:85) */
int gs_vec_combine(gs_ctx* c, const gs_mat* a, const gs_mat* b, uint8_t out16[16]) {
    if (!c || !a || !b || !out16) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    const long long n = a->rows * a->cols;
    if (b->rows * b->cols != n) return c->fail(GS_E_ARG, "shape mismatch");
    cudaSetDevice(c->device);
    long long blocks = (n + 255) / 256; if (blocks > 1024) blocks = 1024;
    int rc = c->ensure_scratch((size_t)blocks * sizeof(fp));
    if (rc != GS_OK) return rc;
    dot_partial_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(a->data, b->data, n, (fp*)c->scratch);
    c->launches++;
    GS_CUDA(c, cudaMemcpyAsync(c->mailbox, c->scratch, (size_t)blocks * sizeof(fp), cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    u128 acc = 0;
    for (long long i = 0; i < blocks; ++i) acc = h_add(acc, elem_in((const uint8_t*)c->mailbox + 16 * i));
    fp f = fp_from_u128(acc); memcpy(out16, &f, 16);
    return GS_OK;
}

/* field.interpolateQuarticBatch(xSets, ySets): rows x 4 each -> rows x 4 coefficients   (LowDegreeProver.ts:137,191) */
int gs_quartic_interpolate_batch(gs_ctx* c, const gs_mat* xs, const gs_mat* ys, gs_mat** polys) {
    if (!c || !xs || !ys || !polys) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    if (xs->cols != 4 || ys->cols != 4 || xs->rows != ys->rows) return c->fail(GS_E_ARG, "two rows x 4 matrices expected");
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, xs->rows, 4, polys);
    if (rc != GS_OK) return rc;
    quartic_interpolate_kernel<<<(unsigned)((xs->rows + 127) / 128), 128, 0, c->stream>>>(xs->data, ys->data, xs->rows, (*polys)->data);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { gs_mat_free(*polys); *polys = nullptr; return c->cuda_fail(e, "quartic_interpolate_kernel"); }
    return GS_OK;
}
/* field.evalQuarticBatch(polys, x): x = one element per row (xs, length rows) or one scalar (x16)   (LowDegreeProver.ts:140,195) */
int gs_quartic_eval_batch(gs_ctx* c, const gs_mat* polys, const gs_mat* xs, const uint8_t* x16, gs_mat** out) {
    if (!c || !polys || !out || (!xs && !x16)) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    if (polys->cols != 4 || (xs && xs->rows * xs->cols != polys->rows)) return c->fail(GS_E_ARG, "rows x 4 polynomials and one x per row expected");
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, 1, polys->rows, out);
    if (rc != GS_OK) return rc;
    fp x = fp_zero(); if (x16) memcpy(&x, x16, 16);
    quartic_eval_kernel<<<(unsigned)((polys->rows + 255) / 256), 256, 0, c->stream>>>(polys->data, xs ? xs->data : nullptr, x, xs ? 1 : 0, polys->rows, (*out)->data);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { gs_mat_free(*out); *out = nullptr; return c->cuda_fail(e, "quartic_eval_kernel"); }
    return GS_OK;
}

/* field.newMatrixFromVectors(vs) / vectorsToMatrix: stack equally long vectors (or matrices with equal column counts) as rows
   (BoundaryConstraints.ts:84-85) */
int gs_mat_stack(gs_ctx* c, const gs_mat* const* parts, int count, gs_mat** out) {
    if (!c || !parts || !out || count < 1) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    long long rows = 0; const long long cols = parts[0] ? parts[0]->cols : 0;
    for (int i = 0; i < count; ++i) { if (!parts[i] || parts[i]->cols != cols) return c->fail(GS_E_ARG, "parts must have equal column counts"); rows += parts[i]->rows; }
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, rows, cols, out);
    if (rc != GS_OK) return rc;
    long long r = 0;
    for (int i = 0; i < count; ++i) {
        GS_CUDA(c, cudaMemcpyAsync((*out)->data + r * cols, parts[i]->data, (size_t)parts[i]->rows * cols * sizeof(fp), cudaMemcpyDeviceToDevice, c->stream));
        r += parts[i]->rows;
    }
    return GS_OK;
}
/* field.matrixRowsToVectors(m)[row0 .. row0+nrows): a copy of consecutive rows   (Stark.ts:114, BoundaryConstraints.ts:73) */
int gs_mat_rows(gs_ctx* c, const gs_mat* m, int64_t row0, int64_t nrows, gs_mat** out) {
    if (!c || !m || !out || row0 < 0 || nrows < 1 || row0 + nrows > m->rows) return c ? c->fail(GS_E_ARG, "row range out of bounds") : GS_E_ARG;
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, nrows, m->cols, out);
    if (rc != GS_OK) return rc;
    GS_CUDA(c, cudaMemcpyAsync((*out)->data, m->data + row0 * m->cols, (size_t)nrows * m->cols * sizeof(fp), cudaMemcpyDeviceToDevice, c->stream));
    return GS_OK;
}
/* field.transposeMatrix(m)   (LowDegreeProver.ts:181) */
int gs_mat_transpose(gs_ctx* c, const gs_mat* m, gs_mat** out) {
    if (!c || !m || !out) return c ? c->fail(GS_E_ARG, "null argument") : GS_E_ARG;
    cudaSetDevice(c->device);
    int rc = gs_mat_alloc(c, m->cols, m->rows, out);
    if (rc != GS_OK) return rc;
    dim3 grid((unsigned)((m->cols + 15) / 16), (unsigned)((m->rows + 15) / 16));
    if (grid.y > 65535) { gs_mat_free(*out); *out = nullptr; return c->fail(GS_E_UNSUPPORTED, "more than 2^20 rows: transpose the other way round"); }
    transpose_matrix_kernel<<<grid, 256, 0, c->stream>>>(m->data, m->rows, m->cols, (*out)->data);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { gs_mat_free(*out); *out = nullptr; return c->cuda_fail(e, "transpose_matrix_kernel"); }
    return GS_OK;
}
/* field.joinMatrixRows(m) and its inverse: the same elements under another shape (row-major), no copy   (LowDegreeProver.ts:182) */
int gs_mat_reshape(gs_mat* m, int64_t rows, int64_t cols) {
    if (!m || rows < 1 || cols < 1 || rows * cols != m->rows * m->cols) return GS_E_ARG;
    m->rows = rows; m->cols = cols;
    return GS_OK;
}
/* Vector.getValue(i) / Matrix.getValue(row, col)   (Stark.ts:290,357; LowDegreeProver.ts:141-142) */
int gs_mat_get(gs_ctx* c, const gs_mat* m, int64_t row, int64_t col, uint8_t out16[16]) {
    if (!c || !m || !out16 || row < 0 || col < 0 || row >= m->rows || col >= m->cols) return c ? c->fail(GS_E_ARG, "index out of range") : GS_E_ARG;
    cudaSetDevice(c->device);
    GS_CUDA(c, cudaMemcpyAsync(out16, m->data + row * m->cols + col, 16, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    return GS_OK;
}

}  // extern "C"
