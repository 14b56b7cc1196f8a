// Host-only translation unit (g++): AIR parsing and execution-trace generation.
#define GS_HOSTAIR_IMPL
#include "hostair.h"
#include "hostjit.h"
#include "verifier.h"
#include "hoststark64.h"

extern "C" int gs_stark_verify(const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                               const uint8_t* assertions, int n_assertions, const uint8_t* proof, size_t proof_len,
                               const uint8_t* public_traces, char* err_buf, size_t err_cap) {
    using namespace gs;
    auto fail = [&](int code, const std::string& m) { if (err_buf && err_cap) { snprintf(err_buf, err_cap, "%s", m.c_str()); } return code; };
    if (!air_blob || !assertions || !proof) return fail(GS_E_ARG, "null argument");
    try {
    AirHost A; int code = GS_OK;
    const std::string perr = parse_air(air_blob, blob_len, &A, &code);
    if (code != GS_OK) return fail(code, perr);
    std::vector<VAssertion> as(n_assertions > 0 ? n_assertions : 0);
    for (int i = 0; i < n_assertions; ++i) {
        const uint8_t* p = assertions + 24 * (size_t)i;
        memcpy(&as[i].reg, p, 4); memcpy(&as[i].step, p + 4, 4);
        fp v; memcpy(&v, p + 8, 16); as[i].value = fp_to_u128(v);
    }
    const std::string err = stark_verify(A, hash_alg, exe_queries, fri_queries, as, proof, proof_len, (const fp*)public_traces);
    if (!err.empty()) return fail(GS_E_STARK, err);
    if (err_buf && err_cap) err_buf[0] = 0;
    return GS_OK;
    } catch (const std::exception& e) { return fail(GS_E_STARK, std::string("Verification failed: ") + e.what()); }
}

/* which trace generator the last generate_trace of this thread used: "jit <hash>" or "interpreter (<reason>)" */
extern "C" const char* gs_trace_backend(void) { return gs::trace_backend_status(); }

/* hash.digest(buffer)   (lib/utils/index.ts:37) -- host */
extern "C" int gs_hash_digest(int alg, const uint8_t* msg, size_t len, uint8_t out32[32]) {
    if ((alg != 0 && alg != 1) || (!msg && len) || !out32) return GS_E_ARG;
    using namespace gs;
    HostHash H{alg};
    const Digest d = H.digest(msg, len);
    memcpy(out32, d.data(), 32);
    return GS_OK;
}
/* MerkleTree.verifyBatch(root, indexes, proof, hash)   (Stark.ts:206; LowDegreeProver.ts:86,109,116) -- host.
   proof blob = what gs_merkle_prove_batch writes: u32 n_values, u32 n_columns, u32 depth, values (32 B each),
   then per column u32 length + nodes.  Returns 1 (valid), 0 (invalid) or a negative status. */
extern "C" int gs_merkle_verify_batch(int alg, const uint8_t root32[32], const uint32_t* indexes, int count, const uint8_t* proof, size_t proof_len) {
    if ((alg != 0 && alg != 1) || !root32 || !indexes || count < 1 || !proof || proof_len < 12) return GS_E_ARG;
    using namespace gs;
    uint32_t nv, nc, depth; memcpy(&nv, proof, 4); memcpy(&nc, proof + 4, 4); memcpy(&depth, proof + 8, 4);
    if ((int)nv != count || depth > 32) return 0;
    size_t off = 12;
    if (proof_len < off + (size_t)nv * 32) return GS_E_ARG;
    // the column count comes from the blob: every column costs at least its 4-byte length, so a count the remaining bytes
    // cannot hold is malformed (and must not size an allocation: 2^32 - 1 columns used to end in std::bad_alloc)
    if ((size_t)nc > (proof_len - off - (size_t)nv * 32) / 4) return GS_E_ARG;
    try {
    std::vector<Digest> values(nv);
    for (uint32_t i = 0; i < nv; ++i) { memcpy(values[i].data(), proof + off, 32); off += 32; }
    std::vector<std::vector<Digest>> nodes(nc);
    for (uint32_t k = 0; k < nc; ++k) {
        if (proof_len < off + 4) return GS_E_ARG;
        uint32_t ln; memcpy(&ln, proof + off, 4); off += 4;
        if (proof_len < off + (size_t)ln * 32) return GS_E_ARG;
        nodes[k].resize(ln);
        for (uint32_t j = 0; j < ln; ++j) { memcpy(nodes[k][j].data(), proof + off, 32); off += 32; }
    }
    Digest root; memcpy(root.data(), root32, 32);
    HostHash H{alg};
    return verify_batch(root, std::vector<uint32_t>(indexes, indexes + count), values, nodes, (int)depth, H) ? 1 : 0;
    } catch (const std::exception&) { return GS_E_ARG; }          // nothing may unwind through the C ABI
}

// ---- prime fields of at most 64 bits: Stark.prove / Stark.verify on the host (hoststark64.h).  Elements cross the ABI as the same
// 16-byte little-endian values as everywhere else (assertion values, initial state, input traces); the proof uses the field's own
// element size.  Refuses the 128-bit STARK field: that one has a GPU path and no CPU fallback.
namespace {
thread_local std::vector<uint8_t> g_small_proof;
bool small_elem(const gs::small::Field& F, const uint8_t* p16, uint64_t* v) {
    for (int i = 8; i < 16; ++i) if (p16[i]) return false;
    uint64_t x = 0; for (int i = 0; i < 8; ++i) x |= (uint64_t)p16[i] << (8 * i);
    *v = x; return x < F.p;
}
int small_setup(const uint8_t* air_blob, size_t blob_len, const uint8_t* assertions, int n_assertions, gs::small::Air& A,
                std::vector<gs::small::Assertion64>& as, std::string& err) {
    int code = GS_OK;
    err = gs::small::parse_air64(air_blob, blob_len, A, &code);
    if (code != GS_OK) return code;
    as.resize(n_assertions > 0 ? n_assertions : 0);
    for (int i = 0; i < n_assertions; ++i) {
        const uint8_t* p = assertions + 24 * (size_t)i;
        memcpy(&as[i].reg, p, 4); memcpy(&as[i].step, p + 4, 4);
        if (!small_elem(A.F, p + 8, &as[i].value)) { err = "non-canonical assertion value"; return GS_E_ARG; }
    }
    return GS_OK;
}
bool small_traces(const gs::small::Air& A, const uint8_t* blob, int count, std::vector<gs::small::Vec>& out) {
    const size_t T = (size_t)1 << A.log_t;
    out.assign(count, gs::small::Vec(T));
    for (int k = 0; k < count; ++k) for (size_t s = 0; s < T; ++s) if (!small_elem(A.F, blob + 16 * ((size_t)k * T + s), &out[k][s])) return false;
    return true;
}
}  // namespace

extern "C" int gs_host_stark_prove(const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                                   const uint8_t* assertions, int n_assertions, const uint8_t* init_state16, const uint8_t* input_traces,
                                   const uint8_t* shapes_blob, size_t shapes_len, const uint8_t** proof_out, size_t* proof_len,
                                   char* err_buf, size_t err_cap) {
    using namespace gs;
    auto fail = [&](int code, const std::string& m) { if (err_buf && err_cap) snprintf(err_buf, err_cap, "%s", m.c_str()); return code; };
    if (!air_blob || !assertions || !init_state16 || !proof_out || !proof_len) return fail(GS_E_ARG, "null argument");
    if ((hash_alg != 0 && hash_alg != 1) || exe_queries < 1 || exe_queries > 128 || fri_queries < 1 || fri_queries > 64) return fail(GS_E_ARG, "bad security options");
    try {
    small::Air A; std::vector<small::Assertion64> as; std::string err;
    int rc = small_setup(air_blob, blob_len, assertions, n_assertions, A, as, err);
    if (rc != GS_OK) return fail(rc, err);
    small::Vec init(A.R);
    for (int r = 0; r < A.R; ++r) if (!small_elem(A.F, init_state16 + 16 * r, &init[r])) return fail(GS_E_ARG, "non-canonical initial state");
    std::vector<small::Vec> in;
    const int n_in = A.n_secret + A.n_public;
    if (n_in > 0 && (!input_traces || !small_traces(A, input_traces, n_in, in))) return fail(GS_E_ARG, "input register traces required (canonical 16-byte elements)");
    err = small::prove(A, hash_alg, exe_queries, fri_queries, as, init, in, shapes_blob, shapes_len, g_small_proof);
    if (!err.empty()) return fail(GS_E_STARK, err);
    *proof_out = g_small_proof.data(); *proof_len = g_small_proof.size();
    if (err_buf && err_cap) err_buf[0] = 0;
    return GS_OK;
    } catch (const std::exception& e) { return fail(GS_E_STARK, std::string("prove failed: ") + e.what()); }
}

extern "C" int gs_host_stark_verify(const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                                    const uint8_t* assertions, int n_assertions, const uint8_t* proof, size_t proof_len,
                                    const uint8_t* public_traces, char* err_buf, size_t err_cap) {
    using namespace gs;
    auto fail = [&](int code, const std::string& m) { if (err_buf && err_cap) snprintf(err_buf, err_cap, "%s", m.c_str()); return code; };
    if (!air_blob || !assertions || !proof) return fail(GS_E_ARG, "null argument");
    try {
    small::Air A; std::vector<small::Assertion64> as; std::string err;
    int rc = small_setup(air_blob, blob_len, assertions, n_assertions, A, as, err);
    if (rc != GS_OK) return fail(rc, err);
    std::vector<small::Vec> pub;
    if (A.n_public > 0 && (!public_traces || !small_traces(A, public_traces, A.n_public, pub))) return fail(GS_E_ARG, "public input traces required");
    err = small::verify(A, hash_alg, exe_queries, fri_queries, as, proof, proof_len, pub);
    if (!err.empty()) return fail(GS_E_STARK, err);
    if (err_buf && err_cap) err_buf[0] = 0;
    return GS_OK;
    } catch (const std::exception& e) { return fail(GS_E_STARK, std::string("Verification failed: ") + e.what()); }
}
