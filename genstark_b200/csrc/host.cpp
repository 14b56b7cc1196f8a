// Host-only translation unit (g++): AIR parsing and execution-trace generation.
#define GS_HOSTAIR_IMPL
#include "hostair.h"
#include "hostjit.h"
#include "verifier.h"

extern "C" int gs_stark_verify(const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                               const uint8_t* assertions, int n_assertions, const uint8_t* proof, size_t proof_len,
                               const uint8_t* public_traces, char* err_buf, size_t err_cap) {
    using namespace gs;
    auto fail = [&](int code, const std::string& m) { if (err_buf && err_cap) { snprintf(err_buf, err_cap, "%s", m.c_str()); } return code; };
    if (!air_blob || !assertions || !proof) return fail(GS_E_ARG, "null argument");
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
}

/* which trace generator the last generate_trace of this thread used: "jit <hash>" or "interpreter (<reason>)" */
extern "C" const char* gs_trace_backend(void) { return gs::trace_backend_status(); }
