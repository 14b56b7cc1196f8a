// Stark.verify (/root/reference/lib/Stark.ts:167-248) natively on the host: O(queries * log N) scalar work,
// as in the reference.  Reads the serialized proof (lib/Serializer.ts:83-144), recomputes the Fiat-Shamir
// coefficients and query positions, checks every batch Merkle proof (MerkleTree.verifyBatch), evaluates the
// constraints at the queried points (CompositionPolynomial.evaluateAt :150-191, LinearCombination.computeOne
// :66-88) and runs the FRI verifier (LowDegreeProver.verify :70-172).  Error texts are the reference's.
#pragma once
#include <functional>
#include "hostair.h"
#include "hostcrypto.h"

namespace gs {

// ---- hashes on the host (digest + merge), sha256 / blake2s256
static inline void h_blake2s(const uint8_t* msg, size_t n, uint8_t out[32]) {
    static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    static const uint8_t SG[10][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
    uint32_t h[8]; memcpy(h, IV, 32); h[0] ^= 0x01010020u;
    size_t off = 0;
    auto rr = [](uint32_t x, int k) { return (x >> k) | (x << (32 - k)); };
    for (;;) {
        const size_t take = n - off > 64 ? 64 : n - off;
        const bool last = (off + take == n);
        uint8_t blk[64] = {0}; memcpy(blk, msg + off, take); off += take;
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; ++i) m[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) | ((uint32_t)blk[4 * i + 3] << 24);
        for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = IV[i]; }
        v[12] ^= (uint32_t)off; v[13] ^= (uint32_t)((uint64_t)off >> 32);
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
            v[a] += v[b] + x; v[d] = rr(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rr(v[b] ^ v[c], 12);
            v[a] += v[b] + y; v[d] = rr(v[d] ^ v[a], 8); v[c] += v[d]; v[b] = rr(v[b] ^ v[c], 7);
        };
        for (int r = 0; r < 10; ++r) {
            const uint8_t* s = SG[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]); G(1, 5, 9, 13, m[s[2]], m[s[3]]); G(2, 6, 10, 14, m[s[4]], m[s[5]]); G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]); G(1, 6, 11, 12, m[s[10]], m[s[11]]); G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
        if (last) break;
    }
    for (int i = 0; i < 8; ++i) { out[4 * i] = (uint8_t)h[i]; out[4 * i + 1] = (uint8_t)(h[i] >> 8); out[4 * i + 2] = (uint8_t)(h[i] >> 16); out[4 * i + 3] = (uint8_t)(h[i] >> 24); }
}
typedef std::array<uint8_t, 32> Digest;
struct HostHash {
    int alg;
    Digest digest(const uint8_t* p, size_t n) const { Digest d; if (alg == 1) h_blake2s(p, n, d.data()); else sha256_bytes(p, n, d.data()); return d; }
    Digest merge(const Digest& a, const Digest& b) const { uint8_t buf[64]; memcpy(buf, a.data(), 32); memcpy(buf + 32, b.data(), 32); return digest(buf, 64); }
};

// ---- wire format reader
struct ParsedBatch { std::vector<std::vector<uint8_t>> values; std::vector<std::vector<Digest>> nodes; int depth = 0; };
struct ProofReader {
    const uint8_t* p; size_t n, off = 0; bool ok = true;
    bool need(size_t k) { if (off + k > n) { ok = false; return false; } return true; }
    uint8_t u8() { if (!need(1)) return 0; return p[off++]; }
    void bytes(void* dst, size_t k) { if (!need(k)) { memset(dst, 0, k); return; } memcpy(dst, p + off, k); off += k; }
    bool batch(ParsedBatch& b, size_t leaf_size) {                       // readMerkleProof, serialization.ts:25-124
        size_t nv = u8(); if (nv == 0) nv = 256;
        b.values.assign(nv, {});
        for (auto& v : b.values) { v.resize(leaf_size); bytes(v.data(), leaf_size); }
        size_t nc = u8(); if (nc == 0) nc = 256;
        std::vector<int> len(nc), type(nc);
        for (size_t i = 0; i < nc; ++i) { uint8_t lt = u8(); len[i] = lt >> 1; type[i] = lt & 1; }
        b.nodes.assign(nc, {});
        for (size_t i = 0; i < nc; ++i) {
            b.nodes[i].resize(len[i]);
            for (int j = 0; j < len[i]; ++j) {
                const size_t sz = (j == 0 && type[i] == 1) ? leaf_size : 32;
                if (sz != 32) { ok = false; return false; }              // a raw leaf in a node column cannot be a digest
                bytes(b.nodes[i][j].data(), 32);
            }
        }
        b.depth = u8();
        return ok;
    }
};

// MerkleTree.verifyBatch with the leaf values already hashed (rehashMerkleProofValues, utils/index.ts:34-45)
static inline bool verify_batch(const Digest& root, const std::vector<uint32_t>& indexes, const std::vector<Digest>& values,
                                const std::vector<std::vector<Digest>>& nodes, int depth, const HostHash& H) {
    if (depth < 0 || depth > 32) return false;
    const uint64_t offset = 1ull << depth;
    std::map<uint32_t, int> index_map;
    for (size_t i = 0; i < indexes.size(); ++i) { if (indexes[i] >= offset) return false; index_map[indexes[i]] = (int)i; }
    if (index_map.size() != indexes.size() || values.size() != indexes.size()) return false;
    std::vector<uint32_t> sorted(indexes); std::sort(sorted.begin(), sorted.end());
    std::vector<uint32_t> norm;
    for (uint32_t v : sorted) { uint32_t e = v & ~1u; if (norm.empty() || norm.back() != e) norm.push_back(e); }
    if (norm.size() != nodes.size()) return false;
    std::map<uint64_t, Digest> v;
    std::vector<uint64_t> next; std::vector<size_t> ptr(norm.size(), 0);
    for (size_t i = 0; i < norm.size(); ++i) {
        const uint32_t index = norm[i];
        auto i1 = index_map.find(index), i2 = index_map.find(index + 1);
        Digest v1, v2;
        if (i1 != index_map.end()) {
            v1 = values[i1->second];
            if (i2 != index_map.end()) { v2 = values[i2->second]; ptr[i] = 0; }
            else { if (nodes[i].empty()) return false; v2 = nodes[i][0]; ptr[i] = 1; }
        } else {
            if (nodes[i].empty() || i2 == index_map.end()) return false;
            v1 = nodes[i][0]; v2 = values[i2->second]; ptr[i] = 1;
        }
        const uint64_t parent = (offset + index) >> 1;
        v[parent] = H.merge(v1, v2);
        next.push_back(parent);
    }
    for (int d = depth - 1; d > 0; --d) {
        std::vector<uint64_t> cur; cur.swap(next);
        for (size_t i = 0; i < cur.size(); ++i) {
            const uint64_t node_index = cur[i], sib_index = node_index ^ 1;
            const size_t col = i;
            Digest sib;
            if (i + 1 < cur.size() && cur[i + 1] == sib_index) { sib = v[sib_index]; ++i; }
            else { if (ptr[col] >= nodes[col].size()) return false; sib = nodes[col][ptr[col]++]; }
            const Digest& node = v[node_index];
            v[node_index >> 1] = (node_index & 1) ? H.merge(sib, node) : H.merge(node, sib);
            next.push_back(node_index >> 1);
        }
    }
    auto it = v.find(1);
    return it != v.end() && it->second == root;
}

static inline u128 read_elem(const uint8_t* p) { fp f; memcpy(&f, p, 16); return fp_to_u128(f); }

// evaluate a flat program at a point (evaluateConstraintsAt)
static inline void eval_program_scalar(const HostProgram& pr, const u128* cur, const u128* nxt, const u128* st, std::vector<u128>& slots, u128* out) {
    for (const auto& ins : pr.instrs) {
        const uint32_t op = ins[0], d = ins[1], a = ins[2], b = ins[3];
        switch (op) {
            case OP_CONST: slots[d] = pr.consts[a]; break;
            case OP_CUR: slots[d] = cur[a]; break;
            case OP_NEXT: slots[d] = nxt[a]; break;
            case OP_STATIC: slots[d] = st[a]; break;
            case OP_ADD: slots[d] = h_add(slots[a], slots[b]); break;
            case OP_SUB: slots[d] = h_sub(slots[a], slots[b]); break;
            case OP_MUL: slots[d] = h_mul(slots[a], slots[b]); break;
            case OP_NEG: slots[d] = h_sub(0, slots[a]); break;
            case OP_INV: slots[d] = h_inv(slots[a]); break;
            case OP_EXP: slots[d] = h_pow(slots[a], pr.consts[b]); break;
            case OP_OUT: out[d] = slots[a]; break;
            default: break;
        }
    }
}
// host Lagrange / Horner (same as prover.cuh's, kept here so the verifier is device-free)
static inline std::vector<u128> v_interpolate(const std::vector<u128>& xs, const std::vector<u128>& ys) {
    const size_t n = xs.size();
    std::vector<u128> root(n + 1, 0); root[0] = 1;
    for (size_t i = 0; i < n; ++i) { for (size_t k = i + 1; k > 0; --k) root[k] = h_sub(root[k - 1], h_mul(root[k], xs[i])); root[0] = h_sub(0, h_mul(root[0], xs[i])); }
    std::vector<u128> out(n, 0), num(n);
    for (size_t i = 0; i < n; ++i) {
        u128 acc = 0;
        for (size_t k = n; k > 0; --k) { acc = h_add(root[k], h_mul(acc, xs[i])); num[k - 1] = acc; }
        u128 den = 0; for (size_t k = n; k > 0; --k) den = h_add(h_mul(den, xs[i]), num[k - 1]);
        const u128 f = h_mul(ys[i], h_inv(den));
        for (size_t k = 0; k < n; ++k) out[k] = h_add(out[k], h_mul(num[k], f));
    }
    return out;
}
static inline u128 v_eval(const std::vector<u128>& poly, u128 x) { u128 acc = 0; for (size_t k = poly.size(); k > 0; --k) acc = h_add(h_mul(acc, x), poly[k - 1]); return acc; }
// interpolate values given on the subgroup generated by g (order L = 2^k): radix-2 inverse FFT on the host
static inline std::vector<u128> v_interpolate_subgroup(const std::vector<u128>& vals, u128 g) {
    const size_t L = vals.size();
    int log_l = 0; while (((size_t)1 << log_l) < L) ++log_l;
    std::vector<u128> v(vals);
    for (size_t i = 0; i < L; ++i) { size_t j = 0; for (int b2 = 0; b2 < log_l; ++b2) j |= ((i >> b2) & 1) << (log_l - 1 - b2); if (j > i) std::swap(v[i], v[j]); }
    const u128 ginv = h_inv(g);
    std::vector<u128> tw(L / 2 ? L / 2 : 1); { u128 a = 1; for (size_t i = 0; i < tw.size(); ++i) { tw[i] = a; a = h_mul(a, ginv); } }
    for (size_t len = 2; len <= L; len <<= 1) {
        const size_t half = len >> 1, stride = L / len;
        for (size_t blk = 0; blk < L; blk += len) for (size_t i = 0; i < half; ++i) {
            const u128 t = h_mul(v[blk + i + half], tw[i * stride]), u = v[blk + i];
            v[blk + i] = h_add(u, t); v[blk + i + half] = h_sub(u, t);
        }
    }
    const u128 linv = h_inv((u128)L);
    for (auto& x : v) x = h_mul(x, linv);
    return v;
}

struct VAssertion { uint32_t reg, step; u128 value; };

// returns "" when the proof verifies, otherwise the reference's error text
static inline std::string stark_verify(const AirHost& A, int hash_alg, int exe_queries, int fri_queries,
                                       const std::vector<VAssertion>& asserts, const uint8_t* proof, size_t proof_len,
                                       const fp* public_traces /* n_public x T or null */) {
    if (asserts.empty()) return "At least one assertion must be provided";
    const HostHash H{hash_alg};
    const int R = A.R, K = A.K, log_t = A.log_t, log_e = A.log_e, log_n = log_t + log_e;
    const uint64_t T = 1ull << log_t, N = 1ull << log_n, E = 1ull << log_e;
    // the same checks the prover makes on its assertions (prover.cuh): a register outside the bank would index past
    // the leaf, a repeated (register, step) pair makes the interpolation divide by zero
    for (size_t i = 0; i < asserts.size(); ++i) {
        if (asserts[i].reg >= (uint32_t)R) return "Invalid assertion: register " + std::to_string(asserts[i].reg) + " is outside of register bank";
        if (asserts[i].step >= T) return "Invalid assertion: step " + std::to_string(asserts[i].step) + " is outside of execution trace";
        for (size_t j = 0; j < i; ++j)
            if (asserts[j].reg == asserts[i].reg && asserts[j].step == asserts[i].step)
                return "Invalid assertion: repeated assertion for register " + std::to_string(asserts[i].reg) + " at step " + std::to_string(asserts[i].step);
    }
    const size_t ev_leaf = (size_t)(R + A.n_secret) * 16, ld_leaf = 64;
    // ---- parse (Serializer.ts:83-144)
    ProofReader rd{proof, proof_len};
    Digest ev_root; rd.bytes(ev_root.data(), 32);
    ParsedBatch ev_proof; if (!rd.batch(ev_proof, ev_leaf)) return "Verification of evaluation Merkle proof failed: malformed proof";
    Digest lc_root; rd.bytes(lc_root.data(), 32);
    ParsedBatch lc_proof; if (!rd.batch(lc_proof, ld_leaf)) return "Verification of low degree failed: malformed proof";
    const int n_comp = rd.u8();
    // the layer count follows from N (LowDegreeProver.ts:179: fold while the column is longer than 256); a proof that
    // claims another count would drive column_length below 4 further down
    {
        int want = 0; for (uint64_t l = N; l > 256; l >>= 2) ++want;
        if (n_comp != want) return "Verification of low degree failed: malformed proof (" + std::to_string(n_comp) + " components, " + std::to_string(want) + " expected)";
    }
    struct Comp { Digest root; ParsedBatch column, poly; };
    std::vector<Comp> comps(n_comp);
    for (auto& c : comps) { rd.bytes(c.root.data(), 32); if (!rd.batch(c.column, ld_leaf) || !rd.batch(c.poly, ld_leaf)) return "Verification of low degree failed: malformed proof"; }
    size_t rem_len = rd.u8(); if (rem_len == 0) rem_len = 256;
    std::vector<u128> remainder(rem_len);
    for (auto& v : remainder) { uint8_t b[16]; rd.bytes(b, 16); v = read_elem(b); }
    if (!rd.ok) return "Verification of low degree failed: malformed proof";
    // input shapes close the proof (Serializer.ts:121-131): the instance was built for them already, but a proof cut short
    // inside this section is still a malformed proof
    const int n_shapes = rd.u8();
    for (int i = 0; i < n_shapes && rd.ok; ++i) { const int rank = rd.u8(); for (int j = 0; j < rank && rd.ok; ++j) { uint8_t b[4]; rd.bytes(b, 4); } }
    if (!rd.ok) return "Verification failed: malformed proof (input shapes)";
    // ---- context: composition / linear-combination parameters (CompositionPolynomial ctor :29-61)
    const u128 w = h_root_of_unity(log_n);
    int max_deg = 1; for (int d : A.degrees) if (d > max_deg) max_deg = d;
    int log_c = 0; while ((1 << log_c) < max_deg) ++log_c;
    const uint64_t comb_degree = T << log_c, comp_degree = std::max(comb_degree - T, T);
    std::vector<uint64_t> group_deg; std::vector<std::vector<int>> group_idx;
    for (int k = 0; k < K; ++k) {
        const uint64_t dg = (uint64_t)A.degrees[k] * T; size_t g = 0;
        for (; g < group_deg.size(); ++g) if (group_deg[g] == dg) break;
        if (g == group_deg.size()) { group_deg.push_back(dg); group_idx.emplace_back(); }
        group_idx[g].push_back(k);
    }
    int d_count = K; for (size_t g = 0; g < group_deg.size(); ++g) if (group_deg[g] < comb_degree) d_count += (int)group_idx[g].size();
    std::vector<uint32_t> b_regs; std::vector<std::vector<u128>> b_xs, b_ys;
    for (const auto& a : asserts) {
        size_t b = 0; for (; b < b_regs.size(); ++b) if (b_regs[b] == a.reg) break;
        if (b == b_regs.size()) { b_regs.push_back(a.reg); b_xs.emplace_back(); b_ys.emplace_back(); }
        b_xs[b].push_back(h_pow(w, (u128)a.step * (u128)E)); b_ys[b].push_back(a.value);
    }
    const int nB = (int)b_regs.size();
    const int b_count = nB * (comp_degree > T ? 2 : 1);
    const int n_ev = R + A.n_secret;
    const uint64_t delta = comp_degree - T;
    const int lc_total = n_ev * (delta > 0 ? 2 : 1);
    const std::vector<u128> coeffs = prng_many(ev_root.data(), 32, d_count + b_count + lc_total);
    std::vector<std::vector<u128>> ipolys(nB), zpolys(nB);
    for (int b = 0; b < nB; ++b) {
        ipolys[b] = v_interpolate(b_xs[b], b_ys[b]);
        std::vector<u128> zp(1, 1);
        for (u128 x : b_xs[b]) { zp.push_back(0); for (size_t k = zp.size() - 1; k > 0; --k) zp[k] = h_sub(zp[k - 1], h_mul(zp[k], x)); zp[0] = h_sub(0, h_mul(zp[0], x)); }
        zpolys[b] = zp;
    }
    // static registers as polynomials: cyclic -> k~ over the subgroup of order L, public inputs -> interpolated trace
    std::vector<std::vector<u128>> cyc_poly(A.statics.size());
    std::vector<std::vector<u128>> pub_poly(A.statics.size());
    {
        int pi = 0;
        for (size_t k = 0; k < A.statics.size(); ++k) {
            if (A.statics[k].kind == 0) {
                const size_t L = A.statics[k].values.size();
                cyc_poly[k] = v_interpolate_subgroup(A.statics[k].values, h_pow(w, (u128)(N / L)));
            } else if (A.statics[k].kind == 2) {
                if (!public_traces) return "public inputs required";
                std::vector<u128> vals(T); for (uint64_t s = 0; s < T; ++s) vals[s] = fp_to_u128(public_traces[(size_t)pi * T + s]);
                pub_poly[k] = v_interpolate_subgroup(vals, h_pow(w, (u128)E));
                ++pi;
            }
        }
    }
    // ---- positions and the evaluation tree (Stark.ts:189-215)
    std::string err;
    std::vector<uint32_t> positions;
    if (pseudorandom_indexes(lc_root.data(), (int)std::min<uint64_t>((uint64_t)exe_queries, N - N / E), N, E, positions, err) != 0) return err;
    std::vector<uint32_t> aug;
    { std::map<uint32_t, bool> seen; for (uint32_t p : positions) for (uint32_t q : {p, (uint32_t)((p + E) % N)}) if (!seen.count(q)) { seen[q] = true; aug.push_back(q); } }
    if (ev_proof.values.size() != aug.size()) return "Verification of evaluation Merkle proof failed";
    std::map<uint32_t, const uint8_t*> leaf_at;
    std::vector<Digest> hashed(aug.size());
    for (size_t i = 0; i < aug.size(); ++i) { leaf_at[aug[i]] = ev_proof.values[i].data(); hashed[i] = H.digest(ev_proof.values[i].data(), ev_leaf); }
    if (!verify_batch(ev_root, aug, hashed, ev_proof.nodes, ev_proof.depth, H) || (1ull << ev_proof.depth) != N) return "Verification of evaluation Merkle proof failed";
    // ---- constraint / linear-combination values at the queried points (Stark.ts:217-233)
    const u128 x_last = h_pow(w, (u128)(T - 1) * (u128)E);
    std::vector<u128> lc_values(positions.size());
    std::vector<u128> slots(A.evaluation.n_slots + 1), sv(A.statics.size() + 1), q(K), cur(R), nxt(R), hv(A.n_secret + 1);
    for (size_t qi = 0; qi < positions.size(); ++qi) {
        const uint32_t step = positions[qi];
        const u128 x = h_pow(w, step);
        const uint8_t* lp = leaf_at[step]; const uint8_t* ln = leaf_at[(uint32_t)((step + E) % N)];
        for (int r = 0; r < R; ++r) { cur[r] = read_elem(lp + 16 * r); nxt[r] = read_elem(ln + 16 * r); }
        for (int s = 0; s < A.n_secret; ++s) hv[s] = read_elem(lp + 16 * (R + s));
        int si = 0;
        for (size_t k = 0; k < A.statics.size(); ++k) {
            if (A.statics[k].kind == 0) sv[k] = v_eval(cyc_poly[k], h_pow(x, (u128)(T / A.statics[k].values.size())));
            else if (A.statics[k].kind == 1) sv[k] = hv[si++];
            else sv[k] = v_eval(pub_poly[k], x);
        }
        eval_program_scalar(A.evaluation, cur.data(), nxt.data(), sv.data(), slots, q.data());
        // degree adjustment + combination (:156-168)
        u128 qc = 0; { int next = K;
            for (int k = 0; k < K; ++k) qc = h_add(qc, h_mul(q[k], coeffs[k]));
            for (size_t g = 0; g < group_deg.size(); ++g) {
                if (group_deg[g] == comb_degree) continue;
                const u128 pw = h_pow(x, (u128)(comb_degree - group_deg[g]));
                for (int k : group_idx[g]) qc = h_add(qc, h_mul(h_mul(q[k], pw), coeffs[next++]));
            } }
        // D = Q / Z (:171-172, ZeroPolynomial.evaluateAt :28-34)
        const u128 z = h_mul(h_sub(h_pow(x, (u128)T), 1), h_inv(h_sub(x, x_last)));
        u128 c = h_mul(qc, h_inv(z));
        // boundary (:175-188)
        const u128 xd = delta ? h_pow(x, (u128)delta) : 1;
        for (int b = 0; b < nB; ++b) {
            const u128 bv = h_mul(h_sub(cur[b_regs[b]], v_eval(ipolys[b], x)), h_inv(v_eval(zpolys[b], x)));
            c = h_add(c, h_mul(bv, coeffs[d_count + b]));
            if (delta) c = h_add(c, h_mul(h_mul(bv, xd), coeffs[d_count + nB + b]));
        }
        // linear combination (LinearCombination.computeOne)
        u128 l = c;
        for (int m = 0; m < n_ev; ++m) {
            const u128 v = read_elem(lp + 16 * m);
            l = h_add(l, h_mul(v, coeffs[d_count + b_count + m]));
            if (delta) l = h_add(l, h_mul(h_mul(v, xd), coeffs[d_count + b_count + n_ev + m]));
        }
        lc_values[qi] = l;
    }
    // ---- low-degree proof (LowDegreeProver.verify :70-172)
    auto aug4 = [](const std::vector<uint32_t>& p, uint64_t column_length) {
        std::vector<uint32_t> out; std::map<uint32_t, bool> seen; const uint32_t row = (uint32_t)(column_length >> 2);
        for (uint32_t v : p) { uint32_t m = v % row; if (!seen.count(m)) { seen[m] = true; out.push_back(m); } }
        return out;
    };
    auto column_values = [&](const ParsedBatch& b, const std::vector<uint32_t>& pos, const std::vector<uint32_t>& augp, uint64_t column_length, std::vector<u128>& out) {
        const uint32_t row = (uint32_t)(column_length >> 2);
        out.clear();
        for (uint32_t p : pos) {
            size_t idx = 0; for (; idx < augp.size(); ++idx) if (augp[idx] == p % row) break;
            if (idx >= b.values.size()) return false;
            out.push_back(read_elem(b.values[idx].data() + 16 * (p / row)));
        }
        return true;
    };
    auto hashed_values = [&](const ParsedBatch& b) { std::vector<Digest> h(b.values.size()); for (size_t i = 0; i < h.size(); ++i) h[i] = H.digest(b.values[i].data(), ld_leaf); return h; };
    uint64_t column_length = N;
    {
        const std::vector<uint32_t> lc_pos = aug4(positions, column_length);
        std::vector<u128> checks;
        if (lc_proof.values.size() != lc_pos.size() || !column_values(lc_proof, positions, lc_pos, column_length, checks)) return "Verification of low degree failed: Verification of linear combination Merkle proof failed";
        if ((1ull << (lc_proof.depth & 63)) != (column_length >> 2) || !verify_batch(lc_root, lc_pos, hashed_values(lc_proof), lc_proof.nodes, lc_proof.depth, H)) return "Verification of low degree failed: Verification of linear combination Merkle proof failed";
        for (size_t i = 0; i < lc_values.size(); ++i) if (lc_values[i] != checks[i]) return "Verification of low degree failed: Verification of linear combination correctness failed";
    }
    Digest p_root = lc_root;
    u128 rou = w;
    uint64_t max_deg_p1 = comp_degree;
    const u128 q4[4] = {1, h_pow(w, (u128)(N / 4)), h_pow(w, (u128)(N / 2)), h_pow(w, (u128)(N / 4 * 3))};
    column_length >>= 2;
    for (int depth = 0; depth < n_comp; ++depth) {
        Comp& cp = comps[depth];
        std::vector<uint32_t> pos;
        if (pseudorandom_indexes(cp.root.data(), fri_queries, column_length, E, pos, err) != 0) return "Verification of low degree failed: " + err;
        const std::vector<uint32_t> augp = aug4(pos, column_length);
        std::vector<u128> colv;
        if (cp.column.values.size() != augp.size() || !column_values(cp.column, pos, augp, column_length, colv)) return "Verification of low degree failed: Verification of column Merkle proof failed at depth " + std::to_string(depth);
        if (column_length < 8) return "Verification of low degree failed: malformed proof (column too short)";
        if ((1ull << (cp.column.depth & 63)) != (column_length >> 2) || !verify_batch(cp.root, augp, hashed_values(cp.column), cp.column.nodes, cp.column.depth, H)) return "Verification of low degree failed: Verification of column Merkle proof failed at depth " + std::to_string(depth);
        if (cp.poly.values.size() != pos.size() || (1ull << (cp.poly.depth & 63)) != column_length || !verify_batch(p_root, pos, hashed_values(cp.poly), cp.poly.nodes, cp.poly.depth, H)) return "Verification of low degree failed: Verification of polynomial Merkle proof failed at depth " + std::to_string(depth);
        const u128 special_x = prng_one(p_root.data(), 32);
        for (size_t i = 0; i < pos.size(); ++i) {
            const u128 xe = h_pow(rou, pos[i]);
            std::vector<u128> xs(4), ys(4);
            for (int j = 0; j < 4; ++j) { xs[j] = h_mul(q4[j], xe); ys[j] = read_elem(cp.poly.values[i].data() + 16 * j); }
            if (v_eval(v_interpolate(xs, ys), special_x) != colv[i]) return "Verification of low degree failed: Degree 4 polynomial didn't evaluate to column value at depth " + std::to_string(depth);
        }
        p_root = cp.root; rou = h_pow(rou, 4); max_deg_p1 /= 4; column_length >>= 2;
    }
    if (max_deg_p1 > remainder.size()) return "Verification of low degree failed: Remainder degree is greater than number of remainder values";
    {
        // remainder tree must match the last column root (:152-160)
        const size_t L = remainder.size(), Q = L >> 2;
        if (Q < 1 || (Q & (Q - 1))) return "Verification of low degree failed: Remainder values do not match Merkle root of the last column";
        std::vector<Digest> level(Q);
        for (size_t i = 0; i < Q; ++i) { uint8_t row[64]; for (int j = 0; j < 4; ++j) { fp f = fp_from_u128(remainder[i + j * Q]); memcpy(row + 16 * j, &f, 16); } level[i] = H.digest(row, 64); }
        while (level.size() > 1) { std::vector<Digest> up(level.size() / 2); for (size_t i = 0; i < up.size(); ++i) up[i] = H.merge(level[2 * i], level[2 * i + 1]); level.swap(up); }
        if (level[0] != p_root) return "Verification of low degree failed: Remainder values do not match Merkle root of the last column";
        // verifyRemainder (:223-252)
        std::vector<size_t> ps; for (size_t i = 0; i < L; ++i) if (i % E) ps.push_back(i);
        if (max_deg_p1 > ps.size()) return "Verification of low degree failed: Remainder degree is greater than number of remainder values";
        std::vector<u128> dom(L); { u128 a = 1; for (size_t i = 0; i < L; ++i) { dom[i] = a; a = h_mul(a, rou); } }
        std::vector<u128> xs(max_deg_p1), ys(max_deg_p1);
        for (size_t i = 0; i < max_deg_p1; ++i) { xs[i] = dom[ps[i]]; ys[i] = remainder[ps[i]]; }
        const std::vector<u128> poly = v_interpolate(xs, ys);
        for (size_t i = max_deg_p1; i < ps.size(); ++i) if (v_eval(poly, dom[ps[i]]) != remainder[ps[i]])
            return "Verification of low degree failed: Remainder is not a valid degree " + std::to_string(max_deg_p1 - 1) + " polynomial";
    }
    return "";
}

}  // namespace gs
