// Declaration-only stand-in for node-addon-api's <napi.h>: just the surface genstark_b200_addon.cc uses, so that
// tests/test_bindings.py can run `g++ -fsyntax-only` on the generated addon in an image without node.  Not shipped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
namespace Napi {
class Env; class Value; class Object; class Array; class Number; class BigInt; class String; class TypedArray; class ArrayBuffer;
class Env { public: Value Undefined() const; };
class Value {
public:
    bool IsNull() const; bool IsUndefined() const; bool IsTypedArray() const; bool IsBigInt() const;
    class Env Env() const;
    template <typename T> T As() const;
};
class Number : public Value { public: static Number New(class Env, double); int64_t Int64Value() const; };
class BigInt : public Value { public: static BigInt New(class Env, uint64_t); int64_t Int64Value(bool* lossless) const; };
class String : public Value { public: static String New(class Env, const char*); };
class Object : public Value { public: void Set(const char*, const Value&); void Set(uint32_t, const Value&); Value Get(uint32_t) const; };
class Array : public Object { public: static Array New(class Env, size_t); uint32_t Length() const; };
class ArrayBuffer : public Object { public: void* Data(); };
class TypedArray : public Object { public: size_t ByteLength() const; size_t ByteOffset() const; class ArrayBuffer ArrayBuffer() const; };
template <typename T> class Buffer : public Object {
public:
    static Buffer<T> New(class Env, size_t); static Buffer<T> Copy(class Env, const T*, size_t);
    T* Data() const; size_t Length() const;
};
template <typename T> class External : public Value { public: static External<T> New(class Env, T*); T* Data() const; };
class Error : public Object {
public:
    static Error New(class Env, const char*); void ThrowAsJavaScriptException() const;
};
class TypeError : public Error { public: static TypeError New(class Env, const char*); };
class CallbackInfo { public: class Env Env() const; Value operator[](size_t) const; size_t Length() const; };
class Function : public Object { public: static Function New(class Env, Value (*)(const CallbackInfo&)); };
}  // namespace Napi
#define NODE_API_MODULE(name, init) Napi::Object (*napi_init_##name)(Napi::Env, Napi::Object) = init;
