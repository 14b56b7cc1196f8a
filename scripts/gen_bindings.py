#!/usr/bin/env python
"""Generates bindings/napi/genstark_b200_addon.cc from include/genstark_b200.h: one N-API export per gs_* entry point.

The marshalling is decided by the C type of each parameter (the header is regular on purpose):
  opaque handle in  (gs_ctx* / const gs_mat* / ...)      <- Napi::External
  opaque handle out (gs_mat** out ...)                    -> returned Napi::External
  handle array      (const gs_mat* const* v, int count)   <- JS array of externals (count taken from the array)
  input bytes       (const uint8_t* / const void* / const uint32_t*) <- Buffer | TypedArray | null
  fixed outputs     (uint8_t out16[16] / out32[32] / out128[128])    -> returned Buffer
  sized outputs     (uint8_t* out / void* out / uint8_t* out16 ...)  <- caller-sized Buffer, filled in place
  scalar outputs    (int* / int64_t* / size_t* / float* / double*)   -> returned in the result
  proof_out/proof_len                                                 -> returned Buffer (copied)
  err_buf/err_cap                                                     -> message of the thrown error
  int status < 0 -> throws; GS_E_STARK is marked so the TypeScript shim rethrows it as StarkError (lib/StarkError.ts).
Run: python scripts/gen_bindings.py   (tests/test_bindings.py checks the result is current and complete)"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'genstark_b200.h')
OUT = os.path.join(ROOT, 'bindings', 'napi', 'genstark_b200_addon.cc')
HANDLES = ('gs_ctx', 'gs_mat', 'gs_digests', 'gs_tree', 'gs_stark')


def declarations(text):
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    out = []
    for m in re.finditer(r'^\s*((?:const\s+)?[\w ]+?[\s\*]+)(gs_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.M | re.S):
        ret, name, params = m.group(1).strip(), m.group(2), ' '.join(m.group(3).split())
        ps = []
        if params and params != 'void':
            for p in params.split(','):
                p = p.strip()
                mm = re.match(r'^(.*?)(\w+)(\[\d+\])?$', p)
                ps.append((mm.group(1).strip(), mm.group(2), mm.group(3) or ''))
        out.append((ret, name, ps))
    return out


def camel(name):
    parts = name[3:].split('_')
    return parts[0] + ''.join(p.capitalize() for p in parts[1:])


def gen_function(ret, name, ps):
    L, pre, call, post, results = [], [], [], [], []
    js = 0          # next JS argument index
    ctx_expr = None
    i = 0
    while i < len(ps):
        ty, pn, arr = ps[i]
        base = ty.replace('const', '').replace('*', '').strip()
        stars = ty.count('*')
        if base in HANDLES and stars == 1:
            pre.append(f'    {base}* {pn} = handle<{base}>(info[{js}]);')
            if base == 'gs_ctx':
                ctx_expr = pn
            call.append(pn); js += 1
        elif base in HANDLES and stars == 2 and 'const' in ty and i + 1 < len(ps) and ps[i + 1][1] == 'count':
            pre.append(f'    std::vector<const {base}*> {pn} = handle_array<{base}>(info[{js}]);')
            call += [f'{pn}.data()', f'(int){pn}.size()']; js += 1; i += 1
        elif base in HANDLES and stars == 2:
            pre.append(f'    {base}* {pn} = nullptr;')
            call.append(f'&{pn}')
            results.append(f'Napi::External<{base}>::New(env, {pn})')
        elif pn == 'proof_out':
            pre.append('    const uint8_t* proof_out = nullptr; size_t proof_len = 0;')
            call += ['&proof_out', '&proof_len']
            results.append('Napi::Buffer<uint8_t>::Copy(env, proof_out, proof_len)'); i += 1
        elif pn == 'err_buf':
            pre.append('    char err_buf[512] = {0};')
            call += ['err_buf', 'sizeof err_buf']; i += 1
        elif arr:                                   # fixed-size array
            n = int(arr[1:-1])
            if 'const' in ty:
                pre.append(f'    const uint8_t* {pn} = bytes_in(info[{js}], {n});'); js += 1
                call.append(pn)
            else:
                pre.append(f'    Napi::Buffer<uint8_t> {pn} = Napi::Buffer<uint8_t>::New(env, {n});')
                call.append(f'{pn}.Data()')
                results.append(pn)
        elif stars == 1 and 'const' in ty and base in ('uint8_t', 'void', 'uint32_t'):
            cast = '' if base == 'uint8_t' else f'(const {base}*)'
            pre.append(f'    const uint8_t* {pn} = bytes_in(info[{js}], 0);'); js += 1
            call.append(f'{cast}{pn}')
            if i + 1 < len(ps) and ps[i + 1][0].strip() == 'size_t' and ps[i + 1][1] in ('nbytes', 'len', 'seed_len', 'blob_len', 'shapes_len', 'proof_len'):
                call.append(f'byte_length(info[{js - 1}])'); i += 1
        elif stars == 1 and base in ('uint8_t', 'void') and 'const' not in ty:
            pre.append(f'    uint8_t* {pn} = bytes_out(info[{js}]);'); js += 1
            call.append(pn if base == 'uint8_t' else f'(void*){pn}')
            if i + 1 < len(ps) and ps[i + 1][0].strip() == 'size_t' and ps[i + 1][1] in ('out_cap', 'out_bytes'):
                call.append(f'byte_length(info[{js - 1}])'); i += 1
        elif stars == 1 and base in ('int', 'int64_t', 'size_t', 'float', 'double'):
            pre.append(f'    {base} {pn} = 0;')
            call.append(f'&{pn}')
            results.append(f'Napi::Number::New(env, (double){pn})')
        elif stars == 0 and base in ('int', 'size_t'):
            pre.append(f'    {base} {pn} = ({base})info[{js}].As<Napi::Number>().Int64Value();'); js += 1
            call.append(pn)
        elif stars == 0 and base in ('int64_t', 'uint64_t'):
            pre.append(f'    {base} {pn} = ({base})int64_in(info[{js}]);'); js += 1
            call.append(pn)
        else:
            raise SystemExit(f'{name}: no marshalling rule for parameter "{ty} {pn}{arr}"')
        i += 1
    L.append(f'static Napi::Value {camel(name)}(const Napi::CallbackInfo& info) {{')
    L.append('    Napi::Env env = info.Env();')
    L += pre
    args = ', '.join(call)
    if ret == 'int':
        L.append(f'    const int rc = {name}({args});')
        stark_arg = next((pn for ty, pn, _ in ps if 'gs_stark' in ty and ty.count('*') == 1), None)
        msg = (f'gs_last_error({ctx_expr})' if ctx_expr else 'err_buf' if any('err_buf' in p for p in pre)
               else f'gs_stark_last_error({stark_arg})' if stark_arg else f'"{name} failed"')
        L.append(f'    if (rc < 0) return fail(env, rc, {msg});')
        if not results:
            L.append('    return Napi::Number::New(env, rc);')
    elif ret == 'void':
        L.append(f'    {name}({args});')
        if not results:
            L.append('    return env.Undefined();')
    elif ret.replace(' ', '') == 'constchar*':
        L.append(f'    const char* text_ = {name}({args});')
        L.append('    return Napi::String::New(env, text_ ? text_ : "");')
    elif ret.replace(' ', '') == 'void*':
        L.append(f'    return Napi::BigInt::New(env, (uint64_t)(uintptr_t){name}({args}));')
    else:       # int64_t / uint64_t
        L.append(f'    return Napi::Number::New(env, (double){name}({args}));')
    if results and ret in ('int', 'void'):
        if len(results) == 1:
            L.append(f'    return {results[0]};')
        else:
            L.append(f'    Napi::Array out = Napi::Array::New(env, {len(results)});')
            for k, r in enumerate(results):
                L.append(f'    out.Set((uint32_t){k}, {r});')
            L.append('    return out;')
    L.append('}')
    return '\n'.join(L)


PRELUDE = '''// GENERATED by scripts/gen_bindings.py from include/genstark_b200.h -- do not edit by hand.
// N-API addon: one export per C entry point of libgenstark_b200.so (SURVEY.md section 8f rank 4; the seam it plugs into is
// /root/reference/lib/Stark.ts:35-58 and genstark.d.ts:62-124).  Build: bindings/napi/binding.gyp (node-addon-api).
#include <napi.h>
#include <cstdint>
#include <vector>
#include "genstark_b200.h"

namespace {
template <typename T> T* handle(const Napi::Value& v) { return (v.IsNull() || v.IsUndefined()) ? nullptr : v.As<Napi::External<T>>().Data(); }
template <typename T> std::vector<const T*> handle_array(const Napi::Value& v) {
    Napi::Array a = v.As<Napi::Array>();
    std::vector<const T*> out(a.Length());
    for (uint32_t i = 0; i < a.Length(); ++i) out[i] = handle<T>(a.Get(i));
    return out;
}
size_t byte_length(const Napi::Value& v) {
    if (v.IsNull() || v.IsUndefined()) return 0;
    if (v.IsTypedArray()) return v.As<Napi::TypedArray>().ByteLength();
    return v.As<Napi::Buffer<uint8_t>>().Length();
}
uint8_t* bytes_out(const Napi::Value& v) {
    if (v.IsNull() || v.IsUndefined()) return nullptr;
    if (v.IsTypedArray()) { Napi::TypedArray t = v.As<Napi::TypedArray>(); return (uint8_t*)t.ArrayBuffer().Data() + t.ByteOffset(); }
    return v.As<Napi::Buffer<uint8_t>>().Data();
}
const uint8_t* bytes_in(const Napi::Value& v, size_t expect) {
    if (expect && byte_length(v) != expect) throw Napi::TypeError::New(v.Env(), "buffer of the wrong length");
    return bytes_out(v);
}
int64_t int64_in(const Napi::Value& v) {
    if (v.IsBigInt()) { bool lossless = true; return v.As<Napi::BigInt>().Int64Value(&lossless); }
    return v.As<Napi::Number>().Int64Value();
}
// status < 0 -> exception; `code` lets the TypeScript side map GS_E_STARK to StarkError and GS_E_ARG to TypeError
Napi::Value fail(Napi::Env env, int rc, const char* message) {
    Napi::Error e = Napi::Error::New(env, message ? message : "libgenstark_b200 error");
    e.Set("code", Napi::Number::New(env, rc));
    e.ThrowAsJavaScriptException();
    return env.Undefined();
}
}  // namespace
'''


def generate():
    decls = declarations(open(HEADER).read())
    parts = [PRELUDE]
    for ret, name, ps in decls:
        parts.append(gen_function(ret, name, ps))
    init = ['static Napi::Object Init(Napi::Env env, Napi::Object exports) {']
    for _, name, _ in decls:
        init.append(f'    exports.Set("{camel(name)}", Napi::Function::New(env, {camel(name)}));')
    init += ['    return exports;', '}', 'NODE_API_MODULE(genstark_b200, Init)', '']
    parts.append('\n'.join(init))
    return '\n\n'.join(parts), [n for _, n, _ in decls]


if __name__ == '__main__':
    text, names = generate()
    if '--check' in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    open(OUT, 'w').write(text)
    print(f'{OUT}: {len(names)} exports, {len(text.splitlines())} lines')
