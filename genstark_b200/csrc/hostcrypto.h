// Host-side Fiat-Shamir pieces of the prover: SHA-256, the galois prng, query positions, batch Merkle
// proofs and the proof wire format.  All O(queries * log N) scalar work -- it stays on the host in the
// reference too -- but every byte of it is bit-exactness-critical.
//
//   sha256(bigint|Buffer) idiom         lib/components/QueryIndexGenerator.ts:61-68
//   getPseudorandomIndexes              lib/components/QueryIndexGenerator.ts:32-59
//   field.prng(seed[, n])               call sites CompositionPolynomial.ts:58, LinearCombination.ts:58,
//                                       LowDegreeProver.ts:194 (construction: SURVEY App. C [RECALLED])
//   MerkleTree.proveBatch               @guildofweavers/merkle 0.3.12 (SURVEY App. C [RECALLED])
//   writeMerkleProof / serializeProof   lib/utils/serialization.ts:18-96, lib/Serializer.ts:35-79
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "fp128.cuh"

namespace gs {

// ------------------------------------------------------------------------------------------ SHA-256
struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t len = 0;
    Sha256() { static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19}; memcpy(h, iv, 32); }
    static uint32_t rr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
            0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
            0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
            0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
            0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
            0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
            0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
        uint32_t w[64];
        for (int i = 0; i < 16; ++i) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            uint32_t s0 = rr(w[i - 15], 7) ^ rr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rr(w[i - 2], 17) ^ rr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            uint32_t t1 = hh + (rr(e, 6) ^ rr(e, 11) ^ rr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rr(a, 2) ^ rr(a, 13) ^ rr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const uint8_t* p, size_t n) {
        size_t fill = len & 63;
        len += n;
        if (fill) {
            size_t take = std::min(n, 64 - fill);
            memcpy(buf + fill, p, take); p += take; n -= take;
            if (fill + take == 64) block(buf); else return;
        }
        while (n >= 64) { block(p); p += 64; n -= 64; }
        if (n) memcpy(buf, p, n);
    }
    void final(uint8_t out[32]) {
        uint64_t bits = len * 8;
        uint8_t pad[72] = {0x80};
        size_t fill = len & 63;
        size_t padlen = (fill < 56) ? 56 - fill : 120 - fill;
        uint8_t lenb[8];
        for (int i = 0; i < 8; ++i) lenb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(pad, padlen);
        update(lenb, 8);
        for (int i = 0; i < 8; ++i) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
    }
};

static inline void sha256_bytes(const uint8_t* p, size_t n, uint8_t out[32]) { Sha256 s; s.update(p, n); s.final(out); }

// 256-bit big-endian integer (+ carry byte) for the "state + i" counter construction
struct Be256 {
    uint8_t b[33];   // b[0] = overflow byte, b[1..32] = digest (big-endian)
};
static inline Be256 be_from_digest(const uint8_t d[32]) { Be256 r; r.b[0] = 0; memcpy(r.b + 1, d, 32); return r; }
static inline Be256 be_add_small(Be256 v, uint64_t x) {
    for (int i = 32; i >= 0 && x; --i) { uint64_t s = v.b[i] + (x & 0xFF); v.b[i] = (uint8_t)s; x = (x >> 8) + (s >> 8); }
    return v;
}
// Buffer.from(value.toString(16), 'hex'): minimal big-endian hex; an odd number of digits loses the LAST nibble
static inline std::vector<uint8_t> be_to_node_buffer(const Be256& v) {
    // hex digits
    uint8_t nib[66];
    for (int i = 0; i < 33; ++i) { nib[2 * i] = v.b[i] >> 4; nib[2 * i + 1] = v.b[i] & 15; }
    int first = 0;
    while (first < 66 && nib[first] == 0) ++first;
    int nd = 66 - first;
    if (nd == 0) { nd = 1; first = 65; }          // "0"
    std::vector<uint8_t> out(nd / 2);
    for (int k = 0; k < nd / 2; ++k) out[k] = (uint8_t)((nib[first + 2 * k] << 4) | nib[first + 2 * k + 1]);
    return out;
}
static inline void sha256_of_be(const Be256& v, uint8_t out[32]) {
    std::vector<uint8_t> buf = be_to_node_buffer(v);
    sha256_bytes(buf.data(), buf.size(), out);
}
// digest (big-endian 256-bit integer) mod p
static inline u128 digest_mod_p(const uint8_t d[32]) {
    u128 hi = 0, lo = 0;
    for (int i = 0; i < 16; ++i) hi = (hi << 8) | d[i];
    for (int i = 16; i < 32; ++i) lo = (lo << 8) | d[i];
    return h_add(h_canon(lo), h_mul(h_canon(hi), HC));
}

// field.prng(seed) -> one element; field.prng(seed, n) -> n elements
static inline u128 prng_one(const uint8_t* seed, size_t seed_len) {
    uint8_t d[32]; sha256_bytes(seed, seed_len, d);
    return digest_mod_p(d);
}
static inline std::vector<u128> prng_many(const uint8_t* seed, size_t seed_len, int n) {
    uint8_t st[32]; sha256_bytes(seed, seed_len, st);
    Be256 state = be_from_digest(st);
    std::vector<u128> out(n);
    for (int i = 0; i < n; ++i) { uint8_t d[32]; sha256_of_be(be_add_small(state, (uint64_t)i), d); out[i] = digest_mod_p(d); }
    return out;
}

// getPseudorandomIndexes(seed, count, max, excludeMultiplesOf); max is a power of two here
static inline int pseudorandom_indexes(const uint8_t seed[32], int count, uint64_t max, uint64_t skip, std::vector<uint32_t>& out,
                                       std::string& err) {
    const uint64_t max_count = skip ? max - max / skip : max;
    if (max_count < (uint64_t)count) { err = "Cannot select " + std::to_string(count) + " unique pseudorandom indexes from " + std::to_string(max) + " values"; return -1; }
    uint8_t st[32]; sha256_bytes(seed, 32, st);
    Be256 state = be_from_digest(st);
    out.clear();
    std::map<uint64_t, bool> seen;
    const long long max_iter = (long long)count * 1000;
    for (long long i = 0; i < max_iter && (int)out.size() < count; ++i) {
        uint8_t d[32]; sha256_of_be(be_add_small(state, (uint64_t)i), d);
        // value mod max: max <= 2^32 and a power of two, or general 64-bit
        uint64_t idx;
        if ((max & (max - 1)) == 0) {
            uint64_t low = 0; for (int k = 24; k < 32; ++k) low = (low << 8) | d[k];
            idx = low & (max - 1);
        } else {
            unsigned __int128 r = 0; for (int k = 0; k < 32; ++k) r = ((r << 8) | d[k]) % max;
            idx = (uint64_t)r;
        }
        if (skip && idx % skip == 0) continue;
        if (seen.count(idx)) continue;
        seen[idx] = true;
        out.push_back((uint32_t)idx);
    }
    if ((int)out.size() < count) { err = "Could not generate " + std::to_string(count) + " pseudorandom indexes"; return -1; }
    return 0;
}

// ------------------------------------------------------------------------------- batch Merkle proofs
struct BatchProof {
    std::vector<std::vector<uint8_t>> values;            // raw leaf bytes, input order
    std::vector<std::vector<uint32_t>> node_ids;         // per column: tree node index of each entry
    std::vector<std::vector<std::array<uint8_t, 32>>> nodes;
    int depth = 0;
};

// Index logic of MerkleTree.proveBatch: which tree nodes (global index: leaves at n..2n-1) go in which column.
static inline int merkle_prove_plan(const std::vector<uint32_t>& indexes, uint64_t n, BatchProof& bp, std::string& err) {
    int depth = 0; while ((1ull << depth) < n) ++depth;
    bp.depth = depth;
    std::map<uint32_t, int> index_map;
    for (size_t i = 0; i < indexes.size(); ++i) {
        if (indexes[i] >= n) { err = "Invalid index"; return -1; }
        index_map[indexes[i]] = (int)i;
    }
    if (index_map.size() != indexes.size()) { err = "Repeating indexes detected"; return -1; }
    std::vector<uint32_t> sorted(indexes); std::sort(sorted.begin(), sorted.end());
    std::vector<uint32_t> norm;
    for (uint32_t v : sorted) { uint32_t e = v & ~1u; if (norm.empty() || norm.back() != e) norm.push_back(e); }
    bp.node_ids.assign(norm.size(), {});
    std::vector<uint64_t> next;
    for (size_t i = 0; i < norm.size(); ++i) {
        const uint32_t index = norm[i];
        const bool has1 = index_map.count(index), has2 = index_map.count(index + 1);
        if (has1 && !has2) bp.node_ids[i].push_back((uint32_t)(n + index + 1));
        else if (!has1) bp.node_ids[i].push_back((uint32_t)(n + index));
        next.push_back((index + n) >> 1);
    }
    for (int d = depth - 1; d > 0; --d) {
        std::vector<uint64_t> cur; cur.swap(next);
        for (size_t i = 0; i < cur.size(); ++i) {
            const uint64_t sib = cur[i] ^ 1;
            const size_t col = i;
            if (i + 1 < cur.size() && cur[i + 1] == sib) ++i;
            else bp.node_ids[col].push_back((uint32_t)sib);
            next.push_back(sib >> 1);
        }
    }
    return 0;
}

// serialization.ts:18-96
static inline void write_merkle_proof(std::vector<uint8_t>& out, const BatchProof& p, size_t leaf_size) {
    out.push_back((uint8_t)(p.values.size() == 256 ? 0 : p.values.size()));
    for (auto& v : p.values) out.insert(out.end(), v.begin(), v.end());
    out.push_back((uint8_t)(p.nodes.size() == 256 ? 0 : p.nodes.size()));
    for (auto& col : p.nodes) {
        const int type = (!col.empty() && leaf_size == 32) ? 1 : 0;     // column[0].byteLength === leafSize
        out.push_back((uint8_t)(((col.size() << 1) | type) & 0xFF));
    }
    for (auto& col : p.nodes) for (auto& x : col) out.insert(out.end(), x.begin(), x.end());
    out.push_back((uint8_t)p.depth);
}

// sizeof.ts:55-99 limits
static inline int check_merkle_proof_limits(const BatchProof& p, std::string& err) {
    if (p.values.empty()) { err = "Array cannot be zero-length"; return -1; }
    if (p.values.size() > 256) { err = "Array length (" + std::to_string(p.values.size()) + ") cannot exceed 256"; return -1; }
    if (p.nodes.size() > 256) { err = "Matrix column count (" + std::to_string(p.nodes.size()) + ") cannot exceed 256"; return -1; }
    for (auto& col : p.nodes) if (col.size() >= 127) { err = "Matrix column length (" + std::to_string(col.size()) + ") cannot exceed 127"; return -1; }
    return 0;
}

}  // namespace gs
