// Sanitizer build of the host-only half of the library (host.cpp: verifiers, AIR parser, trace generator, small-field prover) for
// tests/fuzz/fuzz_host_asan.py.  gs_air_generate_trace itself lives in api.cu (CUDA build); this is its host-only equivalent.
#include "../../genstark_b200/csrc/host.cpp"
extern "C" int asan_generate_trace(const uint8_t* air_blob, size_t blob_len, const uint8_t* init_state16, const uint8_t* input_traces, uint8_t* out_trace) {
    using namespace gs;
    try {
        AirHost S; int code = GS_OK;
        const std::string err = parse_air(air_blob, blob_len, &S, &code);
        if (code != GS_OK) return code;
        if ((S.n_secret + S.n_public) > 0 && !input_traces) return GS_E_ARG;
        std::vector<u128> init(S.R);
        for (int r = 0; r < S.R; ++r) { fp v; memcpy(&v, init_state16 + 16 * r, 16); init[r] = fp_to_u128(v); }
        generate_trace(&S, init.data(), (const fp*)input_traces, (fp*)out_trace);
        return GS_OK;
    } catch (const std::exception&) { return GS_E_ARG; }
}
