// Run-time specialised K2: the AIR's evaluation function emitted as straight-line CUDA C++ (one register-resident
// value per definition, constants as immediates, column pointers at fixed parameter offsets), compiled for sm_100a
// with NVRTC once per distinct program and launched in place of the interpreting compose_kernel.  It is the device
// counterpart of hostjit.h and of what the reference does at instantiate(): air-assembly generates JavaScript for
// the constraint evaluator (called from lib/components/CompositionPolynomial.ts:76 as
// context.evaluateTransitionConstraints).  Everything around the constraints -- field arithmetic, ComposeParams,
// D(x), boundary part, linear combination -- is the same text the library itself is compiled from (GS_DUAL_SOURCE).
// No NVRTC, GS_COMPOSE_JIT=0, or any failure => the interpreting kernel (identical results).
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include "core.cuh"
#include "compose.cuh"
#include "hostcrypto.h"
#include "jitcache.h"

namespace gs {

struct NvrtcApi {
    void* handle = nullptr;
    int (*CreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(void*, size_t*) = nullptr;
    int (*GetCUBIN)(void*, char*) = nullptr;
    int (*GetProgramLogSize)(void*, size_t*) = nullptr;
    int (*GetProgramLog)(void*, char*) = nullptr;
    int (*DestroyProgram)(void**) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        if (!error.empty()) return false;
        const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* nm : names) { handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (handle) break; }
        if (!handle) { error = "cannot load libnvrtc"; return false; }
#define GS_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(handle, sym)); if (!field) { error = std::string("NVRTC symbol missing: ") + sym; handle = nullptr; return false; }
        GS_SYM(CreateProgram, "nvrtcCreateProgram") GS_SYM(CompileProgram, "nvrtcCompileProgram") GS_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        GS_SYM(GetCUBIN, "nvrtcGetCUBIN") GS_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize") GS_SYM(GetProgramLog, "nvrtcGetProgramLog")
        GS_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef GS_SYM
        return true;
    }
};
static inline NvrtcApi& nvrtc_api() { static NvrtcApi a; return a; }

// driver entry points through the runtime (no link-time dependency on libcuda)
struct DriverApi {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    bool ok = false, tried = false;
    bool load() {
        if (tried) return ok;
        tried = true;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult st;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn;
        };
        ok = get("cuModuleLoadData", (void**)&ModuleLoadData) && get("cuModuleGetFunction", (void**)&ModuleGetFunction) &&
             get("cuLaunchKernel", (void**)&LaunchKernel);
        return ok;
    }
};
static inline DriverApi& driver_api() { static DriverApi a; return a; }

struct ComposeJit {
    CUfunction fn = nullptr;
    std::string status;              // "nvrtc <hash>" or "interpreter (<reason>)"
};

static inline std::string devjit_fp_literal(u128 v) {
    char b[96];
    snprintf(b, sizeof b, "{{0x%xu, 0x%xu, 0x%xu, 0x%xu}}", (uint32_t)v, (uint32_t)(v >> 32), (uint32_t)(v >> 64), (uint32_t)(v >> 96));
    return b;
}

// CUDA C++ of the specialised kernel; "" when the program uses something the generator does not handle
static inline std::string compose_emit_source(const AirHost& A) {
    const HostProgram& pr = A.evaluation;
    std::ostringstream o;
    o << "typedef unsigned int uint32_t;\ntypedef unsigned long long uint64_t;\n"
         "#define GS_HD __device__ __forceinline__\n#define GS_D __device__ __forceinline__\n#define GS_ALIGN16 __align__(16)\n"
      << "#define GS_MAX_COLS " << GS_MAX_COLS << "\n"
      << GS_FP_TYPE_SRC << "\n" << GS_FP_BASIC_SRC << "\n" << GS_FP_DEVICE_SRC << "\n" << GS_FP_LDST_SRC << "\n"
      << GS_COMPOSE_PARAMS_SRC << "\n" << GS_COMPOSE_DEVICE_SRC << "\n";
    // GS_COMPOSE_MINB = resident CTAs per SM the register allocation is held to (experiment knob).  Default: none -- MiMC's evaluator
    // takes 80 registers either way (3 / 4 / 5: 0.444 / 0.443 / 0.456 ms), while a cap of 85 registers spills Poseidon's 12-register
    // evaluator (config 5: compose 3.9 -> 4.9 ms)
    int minb = 0; if (const char* e = getenv("GS_COMPOSE_MINB")) { minb = atoi(e); if (minb < 1 || minb > 8) minb = 0; }
    o << "extern \"C\" __global__ void __launch_bounds__(256" << (minb ? ", " + std::to_string(minb) : std::string()) << ") gs_compose_jit(const ComposeParams* __restrict__ Pp) {\n"
         "  const ComposeParams& P = *Pp;\n"
         "  const long long stride = (long long)gridDim.x * blockDim.x;\n"
         "  const unsigned long long lmask = (unsigned long long)P.n_loc - 1ull;\n"
         "  const unsigned E = 1u << P.log_e; const unsigned EL = 1u << P.log_el;\n"
         "  for (long long il = (long long)blockIdx.x * blockDim.x + threadIdx.x; il < P.n_loc; il += stride) {\n"
         "    const long long i = ((il >> P.log_el) << P.log_e) + P.j0 + (il & (EL - 1));\n"
         "    const long long inext = (il + EL) & (long long)lmask;\n"
         "    const fp x = root_pow(P, (unsigned long long)i);\n"
         "    const unsigned ie = (unsigned)i & (E - 1);\n"
         "    fp acc = fp_zero();\n";
    std::vector<std::string> cur(pr.n_slots + 1);
    for (size_t k = 0; k < pr.instrs.size(); ++k) {
        const uint32_t op = pr.instrs[k][0], d = pr.instrs[k][1], a = pr.instrs[k][2], b = pr.instrs[k][3];
        const std::string v = "v" + std::to_string(k);
        if (op != OP_OUT && d > (uint32_t)pr.n_slots) return "";
        auto in = [&](uint32_t s) -> const std::string& { static const std::string none; return s <= (uint32_t)pr.n_slots ? cur[s] : none; };
        switch (op) {
            case OP_CONST: if (a >= pr.consts.size()) return ""; o << "    const fp " << v << " = " << devjit_fp_literal(pr.consts[a]) << ";\n"; break;
            case OP_CUR: if ((int)a >= A.R) return ""; o << "    const fp " << v << " = ld_fp(P.trace[" << a << "] + il);\n"; break;
            case OP_NEXT: if ((int)a >= A.R) return ""; o << "    const fp " << v << " = ld_fp(P.trace[" << a << "] + inext);\n"; break;
            case OP_STATIC:
                if (a >= A.statics.size()) return "";
                if (A.statics[a].kind == 0) o << "    const fp " << v << " = ld_fp(P.stat[" << a << "] + ((unsigned long long)i & P.stat_mask[" << a << "]));\n";
                else o << "    const fp " << v << " = ld_fp(P.stat[" << a << "] + il);\n";
                break;
            case OP_ADD: if (in(a).empty() || in(b).empty()) return ""; o << "    const fp " << v << " = d_add(" << in(a) << ", " << in(b) << ");\n"; break;
            case OP_SUB: if (in(a).empty() || in(b).empty()) return ""; o << "    const fp " << v << " = d_sub(" << in(a) << ", " << in(b) << ");\n"; break;
            case OP_MUL:
                if (in(a).empty() || in(b).empty()) return "";
                if (in(a) == in(b)) o << "    const fp " << v << " = d_sqr(" << in(a) << ");\n";           // 10 limb products instead of 16
                else o << "    const fp " << v << " = d_mul(" << in(a) << ", " << in(b) << ");\n";
                break;
            case OP_NEG: if (in(a).empty()) return ""; o << "    const fp " << v << " = d_neg(" << in(a) << ");\n"; break;
            case OP_INV: if (in(a).empty()) return ""; o << "    const fp " << v << " = d_inv(" << in(a) << ");\n"; break;
            case OP_OUT: if (in(a).empty() || (int)d >= A.K) return ""; o << "    acc = compose_out(P, i, ie, " << d << "u, " << in(a) << ", acc);\n"; break;
            default: return "";
        }
        if (op != OP_OUT) cur[d] = v;
    }
    o << "    compose_tail(P, il, i, ie, x, acc);\n  }\n}\n";
    return o.str();
}

// source -> cubin for sm_100a (disk cache keyed by the source hash); "" + *why on failure.  Needs no device.
static inline std::string compose_compile_cubin(const std::string& src, std::string* key_out, std::string* why) {
    uint8_t dg[32]; sha256_bytes((const uint8_t*)src.data(), src.size(), dg);
    char hex[33]; for (int i = 0; i < 16; ++i) snprintf(hex + 2 * i, 3, "%02x", dg[i]);
    *key_out = hex;
    std::string dir_why;
    const std::string dir = jit_cache_dir(&dir_why);          // "" => compile every time, nothing is cached
    const std::string path = dir.empty() ? std::string() : dir + "/compose_" + hex + "_sm100a.cubin";
    if (!path.empty() && jit_file_trusted(path)) {
        if (FILE* f = fopen(path.c_str(), "rb")) {
            std::string bin; char buf[65536]; size_t n;
            while ((n = fread(buf, 1, sizeof buf, f)) > 0) bin.append(buf, n);
            fclose(f);
            if (!bin.empty()) return bin;
        }
    }
    NvrtcApi& N = nvrtc_api();
    if (!N.load()) { *why = N.error; return ""; }
    void* prog = nullptr;
    if (N.CreateProgram(&prog, src.c_str(), "gs_compose_jit.cu", 0, nullptr, nullptr) != 0) { *why = "nvrtcCreateProgram failed"; return ""; }
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    const int rc = N.CompileProgram(prog, 3, opts);
    if (rc != 0) {
        size_t ls = 0; N.GetProgramLogSize(prog, &ls);
        std::string log(ls, 0); if (ls) N.GetProgramLog(prog, &log[0]);
        *why = "nvrtc: " + log.substr(0, 400);
        N.DestroyProgram(&prog);
        return "";
    }
    size_t sz = 0; N.GetCUBINSize(prog, &sz);
    std::string bin(sz, 0);
    if (!sz || N.GetCUBIN(prog, &bin[0]) != 0) { *why = "nvrtcGetCUBIN failed"; N.DestroyProgram(&prog); return ""; }
    N.DestroyProgram(&prog);
    const std::string tmp = path + "." + std::to_string((long)getpid()) + ".tmp";
    if (path.empty()) return bin;
    if (FILE* f = fopen(tmp.c_str(), "wb")) { fwrite(bin.data(), 1, bin.size(), f); fclose(f); if (rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str()); }
    return bin;
}

// the specialised kernel for this AIR on the current device; never fails the caller (status says what happened)
static inline std::shared_ptr<ComposeJit> compose_jit_get(Ctx* c, const AirHost& A) {
    static std::mutex mu;
    static std::map<std::string, std::shared_ptr<ComposeJit>> cache;      // (device, source hash)
    auto interp = [](const std::string& why) { auto j = std::make_shared<ComposeJit>(); j->status = "interpreter (" + why + ")"; return j; };
    if (const char* e = getenv("GS_COMPOSE_JIT")) if (e[0] == '0') return interp("GS_COMPOSE_JIT=0");
    const std::string src = compose_emit_source(A);
    if (src.empty()) return interp("program not supported by the code generator");
    std::string key, why;
    const std::string bin = compose_compile_cubin(src, &key, &why);
    if (bin.empty()) return interp(why);
    std::lock_guard<std::mutex> lock(mu);
    const std::string ckey = std::to_string(c->device) + ":" + key;
    auto it = cache.find(ckey);
    if (it != cache.end()) return it->second;
    auto done = [&](std::shared_ptr<ComposeJit> j) { cache[ckey] = j; return j; };
    DriverApi& D = driver_api();
    if (!D.load()) return done(interp("driver entry points unavailable"));
    CUmodule mod = nullptr;
    CUresult r = D.ModuleLoadData(&mod, bin.data());
    if (r != CUDA_SUCCESS) return done(interp("cuModuleLoadData failed (" + std::to_string((int)r) + ")"));
    auto j = std::make_shared<ComposeJit>();
    r = D.ModuleGetFunction(&j->fn, mod, "gs_compose_jit");
    if (r != CUDA_SUCCESS || !j->fn) return done(interp("gs_compose_jit missing from the module"));
    j->status = "nvrtc " + key;
    return done(j);
}

}  // namespace gs
