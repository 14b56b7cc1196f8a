// Host-side planning and launch of the K1 passes (see ntt.cuh for the decomposition).
#pragma once
#include "core.cuh"
#include <cstdlib>
#include "ntt.cuh"
#include "ntt2.cuh"

namespace gs {

struct NttPlan {
    int n_pass = 0;
    int log_r[3] = {0, 0, 0};
};

// radices <= 256; as few passes over HBM as possible
static inline NttPlan ntt_plan(int log_n, bool pruned) {
    NttPlan p;
    if (pruned && log_n <= 8) {
        // the coset multiply lives in a column pass, so an LDE always has one
        p.n_pass = 2; p.log_r[0] = (log_n + 1) / 2; p.log_r[1] = log_n / 2;
    } else if (log_n <= 8) { p.n_pass = 1; p.log_r[0] = log_n; }
    else if (log_n <= 16) { p.n_pass = 2; p.log_r[0] = (log_n + 1) / 2; p.log_r[1] = log_n / 2; }
    else {
        p.n_pass = 3;
        p.log_r[0] = (log_n + 2) / 3; p.log_r[1] = (log_n + 1) / 3; p.log_r[2] = log_n / 3;
    }
    return p;
}

static inline int ntt_minb() {          // experiment knob: GS_NTT_MINB = 2 (default) | 3 | 4 resident CTAs per SM
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_NTT_MINB"); v = e ? atoi(e) : 2; if (v < 2 || v > 4) v = 2; }
    return v;
}
template <int A, int B, int M, bool TAB>
static inline cudaError_t launch_pass_m(const NttPassParams& P, dim3 grid, int threads, size_t smem, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ntt_pass_kernel<A, B, M, TAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set = true;
    }
    ntt_pass_kernel<A, B, M, TAB><<<grid, threads, smem, s>>>(P);
    return cudaGetLastError();
}
template <int A, int B>
static inline cudaError_t launch_pass_t(const NttPassParams& P, dim3 grid, int threads, size_t smem, cudaStream_t s) {
    // table variant only when every factor the pass needs is stored (the final pass has none)
    const bool tab = !P.final_pass && P.tw_inter != nullptr && (P.coset_log_ntot == 0 || P.tw_coset != nullptr);
    if (tab) return launch_pass_m<A, B, 2, true>(P, grid, threads, smem, s);
    switch (ntt_minb()) {
        case 3: return launch_pass_m<A, B, 3, false>(P, grid, threads, smem, s);
        case 4: return launch_pass_m<A, B, 4, false>(P, grid, threads, smem, s);
        default: return launch_pass_m<A, B, 2, false>(P, grid, threads, smem, s);
    }
}

static inline cudaError_t launch_pass(int log_r, const NttPassParams& P, dim3 grid, int threads, size_t smem, cudaStream_t s) {
    switch (log_r) {
        case 0: case 1: return launch_pass_t<1, 0>(P, grid, threads, smem, s);
        case 2: return launch_pass_t<2, 0>(P, grid, threads, smem, s);
        case 3: return launch_pass_t<3, 0>(P, grid, threads, smem, s);
        case 4: return launch_pass_t<4, 0>(P, grid, threads, smem, s);
        case 5: return launch_pass_t<3, 2>(P, grid, threads, smem, s);
        case 6: return launch_pass_t<3, 3>(P, grid, threads, smem, s);
        case 7: return launch_pass_t<4, 3>(P, grid, threads, smem, s);
        default: return launch_pass_t<4, 4>(P, grid, threads, smem, s);
    }
}

static inline int pass_log_r1(int log_r) {
    static const int t[9] = {1, 1, 2, 3, 4, 3, 3, 4, 4};
    return t[log_r];
}

// Full twiddle tables.  The two-level root table costs one modular multiplication per looked-up power; with 180 GB of
// HBM the powers a pass needs are simply stored: [k][col] for the inter-pass twiddles of a column pass (2^log_nsub
// entries, shared by every prefix / coset, L2-resident up to 2^22) and [j][pos] for the coset factors of the pruned LDE
// pass ((E-1)*T entries, streamed once per transform next to the coefficients).  GS_NTT_TABLES=0 keeps the lookups;
// GS_NTT_TABLE_MB caps the memory spent (default 4096).  Never built under stream capture (first prove runs uncaptured).
static inline bool ntt_tables_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_NTT_TABLES"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v != 0;
}
static inline size_t ntt_table_cap() {
    static size_t v = 0;
    if (!v) { const char* e = getenv("GS_NTT_TABLE_MB"); v = (size_t)(e ? atoll(e) : 4096) << 20; }
    return v;
}
static inline const fp* ntt_table(Ctx* c, unsigned long long key, size_t entries, const NttPassParams& P, int kind, int arg) {
    if (!ntt_tables_enabled()) return nullptr;
    auto it = c->tw_tables.find(key);
    if (it != c->tw_tables.end()) return it->second;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(c->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
    const size_t bytes = entries * sizeof(fp);
    if (c->tw_table_bytes + bytes > ntt_table_cap()) return nullptr;
    fp* t = nullptr;
    if (cudaMalloc(&t, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const unsigned blocks = (unsigned)((entries + 255) / 256);
    if (kind == 0) tw_inter_table_kernel<<<blocks, 256, 0, c->stream>>>(P, arg, t);
    else tw_coset_table_kernel<<<blocks, 256, 0, c->stream>>>(P, arg, t);
    if (cudaGetLastError() != cudaSuccess) { cudaFree(t); return nullptr; }
    c->launches++;
    c->tw_tables[key] = t;
    c->tw_table_bytes += bytes;
    return t;
}

// ------------------------------------------------------------------------------------------------ two-pass form (ntt2.cuh)
static inline int ntt2_mode() {          // GS_NTT2 = 1 (default) | 0: keep the three-pass kernels of ntt.cuh for every size
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_NTT2"); v = e ? atoi(e) : 1; }
    return v;
}
static inline const fp* ntt2_table(Ctx* c, unsigned long long key, const Ntt2Params& P, int kind, fp scale, int has_scale) {
    auto it = c->tw_tables.find(key);
    if (it != c->tw_tables.end()) return it->second;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(c->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
    const size_t entries = (size_t)1 << P.log_t, bytes = entries * sizeof(fp);
    fp* t = nullptr;
    if (cudaMalloc(&t, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const unsigned blocks = (unsigned)((entries + 255) / 256);
    if (kind == 0) tw2_inter_table_kernel<<<blocks, 256, 0, c->stream>>>(P, scale, has_scale, t);
    else tw2_step_table_kernel<<<blocks, 256, 0, c->stream>>>(P, t);
    if (cudaGetLastError() != cudaSuccess) { cudaFree(t); return nullptr; }
    c->launches++;
    c->tw_tables[key] = t;
    c->tw_table_bytes += bytes;
    return t;
}
template <int LOG_R, bool LDE>
static inline cudaError_t ntt2_launch_pass1(const Ntt2Params& P, unsigned grid, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(ntt2_pass1_kernel<LOG_R, LDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT2_SMEM); attr_set = true; }
    ntt2_pass1_kernel<LOG_R, LDE><<<grid, 256, NTT2_SMEM, s>>>(P);
    return cudaGetLastError();
}
static inline int ntt2_tma_mode() {      // GS_NTT2_TMA bit 0 (default on): pass 2 prefetches its tiles with bulk copies (TMA engine); bit 1: the plain pass 1 fetches its
                                         // tiles through a tensor map (cp.async.bulk.tensor); 0 = plain loads
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_NTT2_TMA"); v = e ? atoi(e) : 1; }
    return v;
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*GsTensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline GsTensorMapEncodeTiled tensor_map_encoder() {
    static GsTensorMapEncodeTiled fn = nullptr; static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr; cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess) fn = (GsTensorMapEncodeTiled)p;
        else cudaGetLastError();
    }
    return fn;
}
// the source of a plain pass 1 as a [rows][R][m * 4 x u32] tensor, boxes of C * 4 words x min(R, 256) rows
template <int LOG_R>
static inline bool ntt2_pass1_tensor_map(CUtensorMap* tm, const fp* src, long long src_row_stride, int rows, int log_m) {
    GsTensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    constexpr int R = 1 << LOG_R, C = 1 << (12 - LOG_R);
    const cuuint64_t dims[3] = {(cuuint64_t)4 << log_m, (cuuint64_t)R, (cuuint64_t)rows};
    const cuuint64_t strides[2] = {(cuuint64_t)16 << log_m, (cuuint64_t)src_row_stride * 16};
    const cuuint32_t box[3] = {(cuuint32_t)C * 4, (cuuint32_t)(R < 256 ? R : 256), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<fp*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <int LOG_R>
static inline cudaError_t ntt2_launch_pass1_tma(const Ntt2Params& P, const CUtensorMap& tm, unsigned grid, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(ntt2_pass1_tma_kernel<LOG_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT2_SMEM + 128); attr_set = true; }
    ntt2_pass1_tma_kernel<LOG_R><<<grid, 256, NTT2_SMEM + 128, s>>>(P, tm);
    return cudaGetLastError();
}
template <int LOG_R>
static inline cudaError_t ntt2_launch_pass2_tma(const Ntt2Params& P, unsigned grid, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(ntt2_pass2_tma_kernel<LOG_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT2_SMEM); attr_set = true; }
    ntt2_pass2_tma_kernel<LOG_R><<<grid, 256, NTT2_SMEM, s>>>(P);
    return cudaGetLastError();
}
static inline bool ntt2_shfl_mode() {     // GS_NTT2_SHFL=1: 1024-point final passes exchange steps 2 -> 3 through warp shuffles (A/B only)
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_NTT2_SHFL"); v = (e && atoi(e) != 0) ? 1 : 0; }
    return v != 0;
}
template <int LOG_R>
static inline cudaError_t ntt2_launch_pass2(const Ntt2Params& P, unsigned grid, cudaStream_t s) {
    if (LOG_R == 10 && ntt2_shfl_mode()) {
        static bool attr_set_s = false;
        if (!attr_set_s) { cudaFuncSetAttribute(ntt2_pass2_shfl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT2_SMEM); attr_set_s = true; }
        ntt2_pass2_shfl_kernel<<<grid, 256, NTT2_SMEM, s>>>(P);
        return cudaGetLastError();
    }
    if (ntt2_tma_mode() & 1) return ntt2_launch_pass2_tma<LOG_R>(P, grid, s);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(ntt2_pass2_kernel<LOG_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NTT2_SMEM); attr_set = true; }
    ntt2_pass2_kernel<LOG_R><<<grid, 256, NTT2_SMEM, s>>>(P);
    return cudaGetLastError();
}

// returns GS_OK when the transform was launched, 1 when this shape is left to ntt_run's three-pass plan
static inline int ntt2_try(Ctx* c, const fp* src, long long src_stride, fp* dst, long long dst_stride, fp* work, long long work_stride,
                           int rows, int log_t, int log_e, bool inverse, int coset_base, int log_e_total) {
    if (!ntt2_mode() || log_t < 16 || log_t > 20 || work == nullptr) return 1;
    const bool lde = log_e_total > 0;
    const int lr1 = (log_t + 1) / 2, lr2 = log_t / 2;
    Ntt2Params P;
    memset(&P, 0, sizeof P);
    P.tw_small = c->tw_small; P.tw_lo = c->tw_lo; P.tw_hi = c->tw_hi; P.log_g = c->log_g; P.log_lo = c->log_lo;
    P.inverse = inverse ? 1 : 0;
    P.log_t = log_t; P.log_m = lr2; P.log_ntot = log_t + (lde ? log_e_total : 0);
    P.n_cosets = 1 << log_e; P.log_cosets = log_e; P.coset_base = coset_base;
    P.log_r1 = lr1;
    // tables (built on the first, uncaptured use of a shape)
    fp scale = fp_one();
    if (inverse) scale = fp_from_u128(h_inv((u128)1 << log_t));
    P.tw_inter = ntt2_table(c, 0x3000000ull | ((unsigned long long)log_t << 8) | (inverse ? 1 : 0), P, 0, scale, inverse ? 1 : 0);
    if (!P.tw_inter) return 1;
    if (lde) {
        P.tw_step = ntt2_table(c, 0x4000000ull | ((unsigned long long)log_t << 8) | (unsigned long long)log_e_total, P, 1, scale, 0);
        if (!P.tw_step) return 1;
    }
    const unsigned max_grid = 2u * (unsigned)c->sm_count;
    if (lde && !c->ntt2_xs) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(c->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return 1; }
        if (cudaMalloc(&c->ntt2_xs, (size_t)max_grid * 4096 * sizeof(fp)) != cudaSuccess) { cudaGetLastError(); c->ntt2_xs = nullptr; return 1; }
    }
    // ---- pass 1: src -> work[row][jl][k_1][n_2]
    P.src = src; P.src_row_stride = src_stride;
    P.dst = work; P.dst_row_stride = work_stride;
    P.xs = c->ntt2_xs;
    const int log_c1 = 12 - lr1;
    P.units = (unsigned)rows << (lr2 - log_c1 + log_e);
    unsigned grid = P.units < max_grid ? P.units : max_grid;
    cudaError_t e;
    {
        ProfScope ps(c, lde ? "ntt_column_coset" : "ntt_column");
        if (lde) e = lr1 == 10 ? ntt2_launch_pass1<10, true>(P, grid, c->stream) : lr1 == 9 ? ntt2_launch_pass1<9, true>(P, grid, c->stream) : ntt2_launch_pass1<8, true>(P, grid, c->stream);
        else {
            CUtensorMap tm;
            const bool tma1 = (ntt2_tma_mode() & 2) && (lr1 == 10 ? ntt2_pass1_tensor_map<10>(&tm, src, src_stride, rows, lr2) : lr1 == 9 ? ntt2_pass1_tensor_map<9>(&tm, src, src_stride, rows, lr2)
                                                                                                                                        : ntt2_pass1_tensor_map<8>(&tm, src, src_stride, rows, lr2));
            if (tma1) e = lr1 == 10 ? ntt2_launch_pass1_tma<10>(P, tm, grid, c->stream) : lr1 == 9 ? ntt2_launch_pass1_tma<9>(P, tm, grid, c->stream) : ntt2_launch_pass1_tma<8>(P, tm, grid, c->stream);
            else e = lr1 == 10 ? ntt2_launch_pass1<10, false>(P, grid, c->stream) : lr1 == 9 ? ntt2_launch_pass1<9, false>(P, grid, c->stream) : ntt2_launch_pass1<8, false>(P, grid, c->stream);
        }
        if (e != cudaSuccess) return c->cuda_fail(e, "ntt2_pass1_kernel");
        c->launches++;
    }
    // ---- pass 2: work -> dst, natural order
    P.src = work; P.src_row_stride = work_stride;
    P.dst = dst; P.dst_row_stride = dst_stride;
    const int log_c2 = 12 - lr2;
    P.units = (unsigned)rows << (lr1 + log_e - log_c2);
    grid = P.units < max_grid ? P.units : max_grid;
    {
        ProfScope ps(c, lde ? "ntt_final_lde" : (inverse ? "ntt_final_inv" : "ntt_final_fwd"));
        e = lr2 == 10 ? ntt2_launch_pass2<10>(P, grid, c->stream) : lr2 == 9 ? ntt2_launch_pass2<9>(P, grid, c->stream) : ntt2_launch_pass2<8>(P, grid, c->stream);
        if (e != cudaSuccess) return c->cuda_fail(e, "ntt2_pass2_kernel");
        c->launches++;
    }
    return GS_OK;
}

// Transform `rows` vectors.
//   log_t : log2 of the input length per row (the size of the DFTs actually computed)
//   log_e : log2 of the pruned leading radix (0 = plain transform; >0 = evaluate on the domain of size
//           2^(log_t+log_e), i.e. E coset transforms written interleaved / natural order)
//   inverse: use w^-1 and scale by 2^-log_t (only for log_e == 0)
// src: rows x 2^log_t (row stride src_stride), dst: rows x 2^(log_t+log_e) (row stride dst_stride).
// work: rows x 2^(log_t+log_e) scratch (may be null when a single pass suffices).
static inline int ntt_run(Ctx* c, const fp* src, long long src_stride, fp* dst, long long dst_stride,
                          fp* work, long long work_stride, int rows, int log_t, int log_e, bool inverse,
                          int coset_base = 0, int log_e_total = -1) {
    // sharded LDE: this call computes 2^log_e cosets starting at coset_base of a domain with 2^log_e_total cosets
    if (log_e_total < 0) log_e_total = log_e;
    const int log_ntot = log_t + log_e;
    if (log_t + log_e_total > c->log_g) return c->fail(GS_E_UNSUPPORTED, "domain 2^%d exceeds root table 2^%d", log_t + log_e_total, c->log_g);
    if (log_ntot > c->log_g) return c->fail(GS_E_UNSUPPORTED, "domain 2^%d exceeds root table 2^%d", log_ntot, c->log_g);
    if (log_t < 1) {
        // length-1 polynomials: constant on every coset
        return c->fail(GS_E_UNSUPPORTED, "transform length must be >= 2");
    }
    if (log_e_total > 0 && log_t < 2) return c->fail(GS_E_UNSUPPORTED, "LDE needs at least 4 coefficients");
    if (log_e_total > 0 && inverse) return c->fail(GS_E_UNSUPPORTED, "inverse coset transform");
    {
        const int rc2 = ntt2_try(c, src, src_stride, dst, dst_stride, work, work_stride, rows, log_t, log_e, inverse, coset_base, log_e_total);
        if (rc2 != 1) return rc2;
    }
    NttPlan plan = ntt_plan(log_t, log_e_total > 0);
    if (plan.n_pass > 1 && work == nullptr) return c->fail(GS_E_ARG, "work buffer required");
    int log_m = log_t;
    int log_npre = log_e;                     // prefixes so far (coset digit)
    int digits[4]; int nd = 0;
    if (log_e > 0) digits[nd++] = log_e;
    for (int p = 0; p < plan.n_pass; ++p) {
        const int lr = plan.log_r[p];
        const bool fin = (p == plan.n_pass - 1);
        log_m -= lr;
        NttPassParams P;
        memset(&P, 0, sizeof P);
        P.tw_lo = c->tw_lo; P.tw_hi = c->tw_hi; P.tw_small = c->tw_small;
        P.log_g = c->log_g; P.log_lo = c->log_lo;
        P.inverse = inverse ? 1 : 0;
        P.final_pass = fin ? 1 : 0;
        P.log_ntot = log_ntot;
        P.log_npre = log_npre;
        P.coset_log_ntot = (p == 0 && log_e_total > 0 && plan.n_pass > 1) ? log_t + log_e_total : 0;
        P.coset_base = coset_base;
        // source / destination of this pass
        const bool first = (p == 0);
        P.src = first ? src : work;
        P.src_row_stride = first ? src_stride : work_stride;
        if (fin) { P.dst = dst; P.dst_row_stride = dst_stride; }
        else { P.dst = work; P.dst_row_stride = work_stride; }
        int log_c;
        dim3 grid;
        if (!fin) {
            P.log_m = log_m;
            P.log_nsub = lr + log_m;
            P.src_prefix_stride = (first && log_e_total > 0) ? 0 : (1ll << P.log_nsub);
            P.log_t = log_t;
            // inter-pass tables up to 2^21 entries (32 MiB) stay L2-resident; a 2^23-entry one streams from DRAM next to the data and
            // measured slower than the lookups (forward 2^23: 0.69 vs 0.66 ms)
            if (P.log_nsub <= 21)
                P.tw_inter = ntt_table(c, 0x1000000ull | ((unsigned long long)P.log_nsub << 16) | ((unsigned long long)lr << 8) | (inverse ? 1 : 0),
                                       (size_t)1 << P.log_nsub, P, 0, lr);
            if (P.coset_log_ntot > 0 && P.tw_inter)
                P.tw_coset = ntt_table(c, 0x2000000ull | ((unsigned long long)log_t << 16) | ((unsigned long long)log_e_total << 8),
                                       (size_t)((1 << log_e_total) - 1) << log_t, P, 1, (1 << log_e_total) - 1);
            log_c = 12 - lr; if (log_c > log_m) log_c = log_m; if (log_c > 5) log_c = 5;
            grid = dim3(1u << (log_npre + log_m - log_c), rows);
        } else {
            P.log_m = 0;
            // the final pass reads the first pass' source directly when it is the only pass
            P.log_d0 = nd > 0 ? digits[0] : 0;
            P.log_d1 = nd > 1 ? digits[1] : 0;
            P.log_d2 = nd > 2 ? digits[2] : 0;
            if (nd > 3) return c->fail(GS_E_UNSUPPORTED, "too many prefix digits");
            log_c = 12 - lr; if (log_c > P.log_d0) log_c = P.log_d0; if (log_c > 4) log_c = 4;
            if (inverse) { P.has_scale = 1; P.scale = fp_from_u128(h_inv((u128)1 << log_t)); }
            grid = dim3(1u << (log_npre - log_c), rows);
        }
        P.log_c = log_c;
        const int lr1 = pass_log_r1(lr), lr2 = lr - lr1;
        int g1 = (1 << lr2) << log_c, g2 = (lr2 > 0) ? ((1 << lr1) << log_c) : 0;
        int threads = g1 > g2 ? g1 : g2;
        if (threads > 256) threads = 256;
        if (threads < 32) threads = 32;
        size_t smem = ((size_t)(1 << lr) + (size_t)(1 << lr) * ((1 << log_c) + 1)) * sizeof(fp);
        ProfScope ps(c, fin ? (log_e > 0 ? "ntt_final_lde" : (inverse ? "ntt_final_inv" : "ntt_final_fwd"))
                               : (P.coset_log_ntot > 0 ? "ntt_column_coset" : "ntt_column"));
        cudaError_t e = launch_pass(lr, P, grid, threads, smem, c->stream);
        if (e != cudaSuccess) return c->cuda_fail(e, "ntt_pass_kernel");
        c->launches++;
        digits[nd++] = lr;
        log_npre += lr;
    }
    return GS_OK;
}

}  // namespace gs
