// K5: blake2s-256 / sha-256 leaf and row hashing and Merkle tree construction.
//
// Replaces merkle's hash.mergeVectorRows (lib/Stark.ts:115), hash.digestValues
// (lib/components/LowDegreeProver.ts:45,163,201) and MerkleTree.create (lib/Stark.ts:118,
// LowDegreeProver.ts:46,164,202).  Digests are the standard unkeyed 32-byte outputs.
//
// Layout: a leaf is the concatenation of element i of `ncols` column vectors (16 bytes each, the
// canonical little-endian residue), so thread i reads column[c][i] -- coalesced across the warp.  A
// 4-column FRI row (v[i], v[i+L/4], v[i+2L/4], v[i+3L/4]) is the same thing with the four quarters
// of v as columns, so no transposed copy is ever materialised.
// Tree: nodes[1] = root, nodes[i] = H(nodes[2i] || nodes[2i+1]), leaves at nodes[n..2n).
#pragma once
#include "core.cuh"
#include "ntt.cuh"

namespace gs {

enum { HASH_SHA256 = 0, HASH_BLAKE2S = 1 };
#define GS_MAX_HASH_COLS 64

struct HashCols {
    const fp* col[GS_MAX_HASH_COLS];
    int ncols;
};

// ------------------------------------------------------------------------------------------ blake2s
GS_D uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

#define B2S_IV0 0x6A09E667u
#define B2S_IV1 0xBB67AE85u
#define B2S_IV2 0x3C6EF372u
#define B2S_IV3 0xA54FF53Au
#define B2S_IV4 0x510E527Fu
#define B2S_IV5 0x9B05688Cu
#define B2S_IV6 0x1F83D9ABu
#define B2S_IV7 0x5BE0CD19u

// Pipe balance: one blake2s compression is 320 XOR + 320 rotate (LOP3 / PRMT / SHF: ALU pipe only) + 160 two-input and
// 160 three-input additions.  ptxas already issues the two-input additions on the FMA pipe (IMAD.IADD) but keeps the
// three-input ones as IADD3 on the ALU pipe, which then runs at 93 % while the FMA pipe idles at 20 % (ncu, round 1).
// Written as two multiply-adds by a 1 that ptxas cannot fold (a __constant__ word), they issue on the FMA pipe too:
// ALU 648 / FMA 482 instructions per compression instead of 814 / 158.
__device__ __constant__ uint32_t GS_B2S_ONE = 1u;
GS_D uint32_t b2s_add3(uint32_t a, uint32_t b, uint32_t x, uint32_t one) {
    uint32_t t, r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(b), "r"(one), "r"(a));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(t));
    return r;
}
#define B2S_G(a, b, c, d, x, y)                    \
    a = b2s_add3(a, b, (x), one); d = __byte_perm(d ^ a, 0, 0x1032); \
    c = c + d; b = rotr32(b ^ c, 12);              \
    a = b2s_add3(a, b, (y), one); d = __byte_perm(d ^ a, 0, 0x0321); \
    c = c + d; b = rotr32(b ^ c, 7);

// one compression; sigma is fully unrolled so message words stay in registers.  LAT: the latency-bound callers (tree tops, the
// FRI tail: a handful of warps walking a chain of dependent compressions) take the plain three-input addition -- the
// multiplier is a literal 1 there, ptxas folds the two multiply-adds back into one IADD3 and the dependency chain of a G
// function is two operations shorter; pipe balance only matters when the SM is full.
template <bool LAT = false>
GS_D void blake2s_compress(uint32_t (&h)[8], const uint32_t (&m)[16], uint32_t t0, bool last) {
    const uint32_t one = LAT ? 1u : GS_B2S_ONE;
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = B2S_IV0, v9 = B2S_IV1, v10 = B2S_IV2, v11 = B2S_IV3;
    uint32_t v12 = B2S_IV4 ^ t0, v13 = B2S_IV5, v14 = last ? ~B2S_IV6 : B2S_IV6, v15 = B2S_IV7;
#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    B2S_G(v0, v4, v8, v12, m[s0], m[s1]) B2S_G(v1, v5, v9, v13, m[s2], m[s3])             \
    B2S_G(v2, v6, v10, v14, m[s4], m[s5]) B2S_G(v3, v7, v11, v15, m[s6], m[s7])           \
    B2S_G(v0, v5, v10, v15, m[s8], m[s9]) B2S_G(v1, v6, v11, v12, m[s10], m[s11])         \
    B2S_G(v2, v7, v8, v13, m[s12], m[s13]) B2S_G(v3, v4, v9, v14, m[s14], m[s15])
    B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
#undef B2S_ROUND
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

GS_D void blake2s_init(uint32_t (&h)[8]) {
    h[0] = B2S_IV0 ^ 0x01010020u; h[1] = B2S_IV1; h[2] = B2S_IV2; h[3] = B2S_IV3;
    h[4] = B2S_IV4; h[5] = B2S_IV5; h[6] = B2S_IV6; h[7] = B2S_IV7;
}

// ------------------------------------------------------------------------------------------- sha256
__device__ __constant__ uint32_t SHA256_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

GS_D void sha256_init(uint32_t (&h)[8]) {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
}

// w: 16 big-endian message words (clobbered)
GS_D void sha256_compress(uint32_t (&h)[8], uint32_t (&w)[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        if (i >= 16) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t t1 = hh + S1 + ch + SHA256_K[i] + w[i & 15];
        const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        const uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

GS_D uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// ------------------------------------------------------------------------------ message -> digest
// Hash a message given as `nwords` little-endian 32-bit words produced by `get(w)` (nwords % 4 == 0).
template <int ALG, bool LAT = false, typename Get>
GS_D void hash_words(Get get, int nwords, uint32_t (&out)[8]) {
    uint32_t h[8];
    if (ALG == HASH_BLAKE2S) {
        blake2s_init(h);
        const int nblocks = (nwords + 15) / 16;
        for (int b = 0; b < nblocks; ++b) {
            uint32_t m[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { const int w = b * 16 + i; m[i] = (w < nwords) ? get(w) : 0u; }
            const bool last = (b == nblocks - 1);
            blake2s_compress<LAT>(h, m, last ? (uint32_t)nwords * 4u : (uint32_t)(b + 1) * 64u, last);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = h[i];
    } else {
        sha256_init(h);
        const int nbytes = nwords * 4;
        const int nblocks = (nbytes + 9 + 63) / 64;
        for (int b = 0; b < nblocks; ++b) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = b * 16 + i;
                uint32_t v = 0;
                if (k < nwords) v = bswap32(get(k));
                else if (k == nwords) v = 0x80000000u;
                if (b == nblocks - 1 && i == 15) v = (uint32_t)nbytes * 8u;
                w[i] = v;
            }
            sha256_compress(h, w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = bswap32(h[i]);
    }
}

GS_D void store_digest(uint32_t* dst, const uint32_t (&d)[8]) {
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(d[0], d[1], d[2], d[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

// leaf `row` of at most four columns (one 64-byte block: MiMC leaves, every FRI row)
template <int ALG>
GS_D void hash_leaf4(const HashCols& cols, long long row, uint32_t (&d)[8]) {
    uint32_t m[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fp e = (c < cols.ncols) ? ld_fp(cols.col[c] + row) : fp_zero();
        m[4 * c] = e.v[0]; m[4 * c + 1] = e.v[1]; m[4 * c + 2] = e.v[2]; m[4 * c + 3] = e.v[3];
    }
    if (ALG == HASH_BLAKE2S) {
        // one (final) block, zero padded, counter = message bytes: the message words never leave the registers
        // (the general loop of hash_words indexes them with the block number, which puts them in local memory)
        uint32_t h[8];
        blake2s_init(h);
        blake2s_compress<false>(h, m, (uint32_t)cols.ncols * 16u, true);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = h[q];
    } else {
        auto getm = [&](int w) -> uint32_t { return m[w]; };
        hash_words<ALG>(getm, cols.ncols * 4, d);
    }
}

// leaf i = H(col[0][i] || col[1][i] || ...)   ->  out[i] (32 bytes)
template <int ALG>
__global__ void __launch_bounds__(256) hash_columns_kernel(const HashCols cols, long long n, uint32_t* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t d[8];
        auto get = [&](int w) -> uint32_t { return cols.col[w >> 2][i].v[w & 3]; };
        if (cols.ncols <= 4) {
            // common case (MiMC leaves, every FRI row): one block, words straight from registers
            hash_leaf4<ALG>(cols, i, d);
        } else {
            hash_words<ALG>(get, cols.ncols * 4, d);
        }
        store_digest(out + 8 * i, d);
    }
}

// generic digestValues over a raw byte buffer: row r = bytes [r*row_bytes, (r+1)*row_bytes), row_bytes % 16 == 0
template <int ALG>
__global__ void __launch_bounds__(256) hash_rows_kernel(const uint32_t* __restrict__ buf, int row_words, long long n,
                                                        uint32_t* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t* row = buf + i * row_words;
        uint32_t d[8];
        auto get = [&](int w) -> uint32_t { return row[w]; };
        hash_words<ALG>(get, row_words, d);
        store_digest(out + 8 * i, d);
    }
}

// parents [count, 2*count) <- H(children)
// parents [count + first, count + first + cnt) of the level with `count` parents (first = 0, cnt = count: whole level)
template <int ALG>
__global__ void __launch_bounds__(256) merkle_level_kernel(uint32_t* __restrict__ nodes, long long count, long long first, long long cnt) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += stride) {
        const long long i = count + first + j;
        const uint4* ch = reinterpret_cast<const uint4*>(nodes + 16 * i);     // nodes[2i], nodes[2i+1]
        uint32_t m[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) { uint4 t = ch[q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
        uint32_t d[8];
        auto getm = [&](int w) -> uint32_t { return m[w]; };
        hash_words<ALG>(getm, 16, d);
        store_digest(nodes + 8 * i, d);
    }
}

// K consecutive levels of the throughput-bound part of a tree in ONE launch, every thread busy in every level: a block
// owns S = 256 * 2^(K-1) adjacent parents of the level with `count` parents (nodes[count + p0 ..]); a thread hashes
// 2^(K-1) of them (parent p0 + r * 256 + t: a warp reads 2 KB of consecutive children), the block keeps that level in
// shared memory, then 2^(K-2) nodes per thread of the level above, ... down to one node per thread.  Per level this is
// the work of merkle_level_kernel without reading the level below back from L2 / HBM, and K levels cost one launch.
// LEAF: the children of the first level are the leaves themselves -- the thread hashes rows 2p and 2p+1 of the columns
// (hash_columns_kernel's job), stores the two digests and goes on with their parent, so the leaf digests are written once
// and never read back (evaluation tree of the north-star shape: 268 MB less HBM traffic and 8 launches less).
template <int ALG, int K, bool LEAF>
__global__ void __launch_bounds__(256) merkle_span_kernel(uint32_t* __restrict__ nodes, long long count, const HashCols cols) {
    extern __shared__ __align__(16) unsigned char span_raw[];
    constexpr int S = 256 << (K - 1);
    uint4* src = reinterpret_cast<uint4*>(span_raw);                 // S digests (K > 1)
    uint4* dst = src + (K > 1 ? 2 * S : 0);                          // S / 2 digests; LEAF: at least one 64-byte slot per thread
    const long long p0 = (long long)blockIdx.x * S;
#pragma unroll 1
    for (int r = 0; r < (1 << (K - 1)); ++r) {
        const int j = r * 256 + threadIdx.x;
        const long long i = count + p0 + j;                          // heap index of the parent
        uint32_t m[16];
        const uint4* ch;
        if (LEAF) {
            // the two leaf digests go through a thread-private 64-byte slot of the (still unused) second buffer, so that
            // nothing but the slot index is live across a leaf compression (48 registers instead of 75)
            uint4* slot = dst + 4 * threadIdx.x;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint32_t d[8];
                hash_leaf4<ALG>(cols, 2 * (p0 + j) + h, d);
                store_digest(nodes + 8 * (2 * i + h), d);
                slot[2 * h] = make_uint4(d[0], d[1], d[2], d[3]); slot[2 * h + 1] = make_uint4(d[4], d[5], d[6], d[7]);
            }
            ch = slot;
        } else {
            ch = reinterpret_cast<const uint4*>(nodes + 16 * i);                  // nodes[2i], nodes[2i+1]
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { uint4 t = ch[q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
        uint32_t d[8];
        auto getm = [&](int w) -> uint32_t { return m[w]; };
        hash_words<ALG>(getm, 16, d);
        store_digest(nodes + 8 * i, d);
        if (K > 1) { src[2 * j] = make_uint4(d[0], d[1], d[2], d[3]); src[2 * j + 1] = make_uint4(d[4], d[5], d[6], d[7]); }
    }
    long long cl = count >> 1, pl = p0 >> 1;
#pragma unroll 1
    for (int l = 1; l < K; ++l) {
        __syncthreads();                                             // the level below is complete; its buffer two levels down is free
        const int reps = 1 << (K - 1 - l);
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
            const int j = r * 256 + threadIdx.x;
            uint32_t m[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { uint4 t = src[4 * j + q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
            uint32_t d[8];
            auto getm = [&](int w) -> uint32_t { return m[w]; };
            hash_words<ALG>(getm, 16, d);
            store_digest(nodes + 8 * (cl + pl + j), d);
            if (l < K - 1) { dst[2 * j] = make_uint4(d[0], d[1], d[2], d[3]); dst[2 * j + 1] = make_uint4(d[4], d[5], d[6], d[7]); }
        }
        uint4* t = src; src = dst; dst = t;
        cl >>= 1; pl >>= 1;
    }
}

// LEVELS levels of the subtree below node (top_count + blockIdx.x) inside one block: its 2^LEVELS descendants are
// staged in shared memory once and every intermediate node is written out.  Used for the latency-bound middle of
// a tree (<= 2^16 nodes per level), where one launch per level costs ~6 us regardless of its size.
template <int ALG, int LEVELS>
__global__ void __launch_bounds__(1024) merkle_subtree_kernel(uint32_t* __restrict__ nodes, long long top_count, long long block0) {
    extern __shared__ __align__(16) unsigned char sub_raw[];
    uint4* s = reinterpret_cast<uint4*>(sub_raw);                      // 2^LEVELS digests = 2 x uint4 each
    const long long top = top_count + block0 + blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(nodes + 8 * (top << LEVELS));
    for (int i = threadIdx.x; i < (2 << LEVELS); i += blockDim.x) s[i] = src[i];
    __syncthreads();
    for (int l = LEVELS - 1; l >= 0; --l) {
        const int cnt = 1 << l;
        uint32_t d[8]; bool active = false;
        for (int j = threadIdx.x; j < cnt; j += blockDim.x) {          // cnt <= blockDim.x: at most one iteration
            uint32_t m[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { uint4 t = s[4 * j + q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
            auto getm = [&](int w) -> uint32_t { return m[w]; };
            hash_words<ALG>(getm, 16, d);
            active = true;
        }
        __syncthreads();
        if (active) {
            const int j = threadIdx.x;
            s[2 * j] = make_uint4(d[0], d[1], d[2], d[3]); s[2 * j + 1] = make_uint4(d[4], d[5], d[6], d[7]);
            store_digest(nodes + 8 * ((top << l) + j), d);
        }
        __syncthreads();
    }
}

// (Tried and dropped, round 1: hashing one node on four lanes -- column / diagonal steps with width-4 shuffles -- for the
// narrow levels.  merkle_build stayed at 0.708 ms: a level is bound by the ~240-deep dependency chain of one compression,
// which four lanes do not shorten, not by the ~1100 instructions a single lane issues.)
// Top of a tree in ONE launch: the level with `level_nodes` nodes (heap indices level_nodes .. 2*level_nodes) is cut into
// x* = field.prng(root) = SHA-256(root) as a big-endian integer mod p (LowDegreeProver.ts:194) computed on the device, so a FRI
// layer can be folded without a host round trip
GS_D fp fri_challenge_dev(const uint32_t* root) {
    uint32_t d[8];
    auto get = [&](int w) -> uint32_t { return root[w]; };
    hash_words<HASH_SHA256>(get, 8, d);
    // d[] holds the digest bytes as little-endian words of the byte string; big-endian integer: byte 0 is most significant
    uint32_t be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) be[i] = bswap32(d[i]);          // be[0] = most significant 32 bits
    fp hi, lo;
    hi.v[3] = be[0]; hi.v[2] = be[1]; hi.v[1] = be[2]; hi.v[0] = be[3];
    lo.v[3] = be[4]; lo.v[2] = be[5]; lo.v[1] = be[6]; lo.v[0] = be[7];
    // canonical residues of the two halves, then lo + hi * 2^128 = lo + hi * (9*2^32 - 1)
    const fp zero = fp_zero();
    hi = fp_add(hi, zero); lo = fp_add(lo, zero);                 // fp_add canonicalises (adds 2^128 - p when >= p)
    fp c9; c9.v[0] = 0xFFFFFFFFu; c9.v[1] = 8u; c9.v[2] = 0; c9.v[3] = 0;
    return fp_add(lo, fp_mul(hi, c9));
}

// 512-node subtrees, one per block, reduced in shared memory (every intermediate node is written out); the block that
// finishes last then reduces the subtree roots to the root.  Replaces a launch per level where a level is a handful of
// dependent ~1 us compressions and the launch gap costs more than the work.  *counter must be zero and is left zero.
// what the block that writes the root does with it besides storing it in the tree: the FRI challenge x* = prng(root), and the
// root + epoch flag straight into the pinned (host-mapped) mailbox the host polls -- instead of a one-thread launch and two
// 32-byte / 4-byte device-to-host copies per layer
struct RootSink { fp* challenge_out; uint32_t* mb_root; uint32_t* mb_flag; const uint32_t* epoch; };
// LEAF: the level with `level_nodes` nodes is the leaf level and is hashed here from the rows of `cols` (<= 4 columns), so a
// tree of at most 2^17 leaves -- the third and later FRI layers -- is committed by this one launch.
template <int ALG, bool LEAF = false>
__global__ void __launch_bounds__(256) merkle_top_kernel(uint32_t* __restrict__ nodes, int level_nodes, unsigned* counter, const RootSink sink, const HashCols cols) {
    __shared__ uint4 s[1024];                                       // 512 digests
    __shared__ int is_last;
    const int leaves = level_nodes < 512 ? level_nodes : 512;
    int top = level_nodes / leaves + blockIdx.x;                    // heap index of this block's subtree root
    int n = leaves;
    for (int pass = 0; pass < 2; ++pass) {
        // stage the n descendants of `top` (contiguous in the heap layout); .cg: pass 1 reads what other blocks just wrote
        int log_n = 0; while ((1 << log_n) < n) ++log_n;
        const uint4* src = reinterpret_cast<const uint4*>(nodes + 8 * ((long long)top << log_n));
        if (LEAF && pass == 0) {
#pragma unroll 1
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const long long row = (long long)blockIdx.x * leaves + i;
                uint32_t d[8];
                hash_leaf4<ALG>(cols, row, d);
                s[2 * i] = make_uint4(d[0], d[1], d[2], d[3]); s[2 * i + 1] = make_uint4(d[4], d[5], d[6], d[7]);
                store_digest(nodes + 8 * ((long long)level_nodes + row), d);
            }
        } else {
            for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) s[i] = __ldcg(src + i);
        }
        __syncthreads();
        for (int l = log_n - 1; l >= 0; --l) {
            const int cnt = 1 << l;
            uint32_t d[8];
            const int j = threadIdx.x;
            if (j < cnt) {
                uint32_t m[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) { uint4 t = s[4 * j + q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
                auto getm = [&](int w) -> uint32_t { return m[w]; };
                hash_words<ALG, true>(getm, 16, d);
            }
            __syncthreads();
            if (j < cnt) {
                s[2 * j] = make_uint4(d[0], d[1], d[2], d[3]); s[2 * j + 1] = make_uint4(d[4], d[5], d[6], d[7]);
                store_digest(nodes + 8 * (((long long)top << l) + j), d);
            }
            __syncthreads();
        }
        if (pass == 1 || gridDim.x == 1) {
            // the block that wrote the root also derives the FRI challenge from it (thread 0 holds the root it just stored):
            // saves the one-thread launch that used to sit between the tree and the fold
            const uint32_t* root = reinterpret_cast<const uint32_t*>(&s[0]);
            if (sink.mb_root && threadIdx.x < 8) { sink.mb_root[threadIdx.x] = root[threadIdx.x]; __threadfence_system(); }
            if (sink.challenge_out && threadIdx.x == 0) st_fp(sink.challenge_out, fri_challenge_dev(root));
            if (sink.mb_root) {
                __syncthreads();                          // all eight root words are fenced before the flag
                if (threadIdx.x == 0) { *sink.mb_flag = *sink.epoch; __threadfence_system(); }
            }
            return;
        }
        // the last block to get here owns the remaining gridDim.x subtree roots
        __threadfence();
        if (threadIdx.x == 0) {
            const unsigned ticket = atomicAdd(counter, 1u);
            is_last = (ticket == gridDim.x - 1);
            if (is_last) *counter = 0;
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        top = 1; n = (int)gridDim.x;
    }
}

// all levels from `count` parents down to the root inside one block
template <int ALG>
__global__ void __launch_bounds__(1024) merkle_tail_kernel(uint32_t* __restrict__ nodes, int count) {
    for (int c = count; c >= 1; c >>= 1) {
        for (int j = threadIdx.x; j < c; j += blockDim.x) {
            const int i = c + j;
            const uint4* ch = reinterpret_cast<const uint4*>(nodes + 16 * i);
            uint32_t m[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { uint4 t = ch[q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
            uint32_t d[8];
            auto getm = [&](int w) -> uint32_t { return m[w]; };
            hash_words<ALG, true>(getm, 16, d);
            store_digest(nodes + 8 * i, d);
        }
        __syncthreads();
    }
}

__global__ void fri_challenge_kernel(const uint32_t* __restrict__ root, fp* __restrict__ out) { st_fp(out, fri_challenge_dev(root)); }

// sharded tree: this rank's slice of every level from `count_local` parents per rank down to its sub-tree root
// (node 2^log_w + rank), inside one block
template <int ALG>
__global__ void __launch_bounds__(1024) merkle_tail_range_kernel(uint32_t* __restrict__ nodes, int count_local, int log_w, int rank) {
    for (int cl = count_local; cl >= 1; cl >>= 1) {
        const long long base = ((long long)cl << log_w) + (long long)rank * cl;
        for (int j = threadIdx.x; j < cl; j += blockDim.x) {
            const long long i = base + j;
            const uint4* ch = reinterpret_cast<const uint4*>(nodes + 16 * i);
            uint32_t m[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { uint4 t = ch[q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
            uint32_t d[8];
            auto getm = [&](int w) -> uint32_t { return m[w]; };
            hash_words<ALG>(getm, 16, d);
            store_digest(nodes + 8 * i, d);
        }
        __syncthreads();
    }
}

static inline unsigned grid_for(Ctx* c, long long n, int threads, int waves = 8) {
    long long blocks = (n + threads - 1) / threads;
    const long long cap = (long long)c->sm_count * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static inline int hash_columns(Ctx* c, int alg, const HashCols& cols, long long n, uint32_t* out) {
    if (cols.ncols < 1 || cols.ncols > GS_MAX_HASH_COLS) return c->fail(GS_E_ARG, "1..%d columns per leaf", GS_MAX_HASH_COLS);
    const unsigned g = grid_for(c, n, 256);
    ProfScope ps(c, "hash_columns");
    if (alg == HASH_BLAKE2S) hash_columns_kernel<HASH_BLAKE2S><<<g, 256, 0, c->stream>>>(cols, n, out);
    else if (alg == HASH_SHA256) hash_columns_kernel<HASH_SHA256><<<g, 256, 0, c->stream>>>(cols, n, out);
    else return c->fail(GS_E_ARG, "unknown hash algorithm");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->cuda_fail(e, "hash_columns_kernel");
    c->launches++;
    return GS_OK;
}

// Sharded commit over peer memory (NVLink / NVSwitch): a rank hashes the rows of its cosets and stores every digest straight
// into the memory of the rank that owns that leaf RANGE (leaves are dealt to the ranks in W contiguous ranges), through
// pointers obtained once with cudaIpcOpenMemHandle.  The transfer rides under the hashing arithmetic; what is left of the
// exchange is one barrier.  Replaces hash -> local buffer -> ncclSend/ncclRecv all-to-all (0.34 ms for the evaluation tree
// at 2 GPUs, 165 - 190 GB/s).  The digests land in the owner's STAGING buffer in the sender's order -- block [sender][k], the
// layout the all-to-all produced -- so a warp stores 1 KB of consecutive bytes per instruction pair; storing each digest at its
// final leaf position instead (32 isolated bytes every E * 32) was measured at 93 GB/s per rank on 8 GPUs.  The owner puts its
// range in leaf order with permute_digests_kernel after the barrier (local, 64 bytes of traffic per leaf).
// local row il of this rank: block s = il / blk goes to rank s, at [rank][il - s * blk].
struct PeerStage { uint32_t* base[8]; };     // per rank: its staging buffer (W * blk digests)
template <int ALG>
__global__ void __launch_bounds__(256) hash_columns_scatter_kernel(const HashCols cols, long long n_loc, const PeerStage peers, long long blk, int rank) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long il = (long long)blockIdx.x * blockDim.x + threadIdx.x; il < n_loc; il += stride) {
        uint32_t d[8];
        if (cols.ncols <= 4) {
            hash_leaf4<ALG>(cols, il, d);
        } else {
            auto get = [&](int w) -> uint32_t { return cols.col[w >> 2][il].v[w & 3]; };
            hash_words<ALG>(get, cols.ncols * 4, d);
        }
        const long long s = il / blk;
        store_digest(peers.base[s] + 8 * ((long long)rank * blk + (il - s * blk)), d);
    }
}
static inline int hash_columns_scatter(Ctx* c, int alg, const HashCols& cols, long long n_loc, const PeerStage& peers, long long blk, int rank) {
    if (cols.ncols < 1 || cols.ncols > GS_MAX_HASH_COLS) return c->fail(GS_E_ARG, "1..%d columns per leaf", GS_MAX_HASH_COLS);
    const unsigned g = grid_for(c, n_loc, 256);
    ProfScope ps(c, "hash_columns");
    if (alg == HASH_BLAKE2S) hash_columns_scatter_kernel<HASH_BLAKE2S><<<g, 256, 0, c->stream>>>(cols, n_loc, peers, blk, rank);
    else if (alg == HASH_SHA256) hash_columns_scatter_kernel<HASH_SHA256><<<g, 256, 0, c->stream>>>(cols, n_loc, peers, blk, rank);
    else return c->fail(GS_E_ARG, "unknown hash algorithm");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->cuda_fail(e, "hash_columns_scatter_kernel");
    c->launches++;
    return GS_OK;
}
struct PeerTrees { uint32_t* base[8]; };     // per rank: the tree a commit builds there (sub-tree roots are stored into all of them)

static inline int hash_rows(Ctx* c, int alg, const void* buf, int row_bytes, long long n, uint32_t* out) {
    if (row_bytes <= 0 || row_bytes % 16) return c->fail(GS_E_ARG, "row size must be a positive multiple of 16 bytes");
    const unsigned g = grid_for(c, n, 256);
    if (alg == HASH_BLAKE2S) hash_rows_kernel<HASH_BLAKE2S><<<g, 256, 0, c->stream>>>((const uint32_t*)buf, row_bytes / 4, n, out);
    else if (alg == HASH_SHA256) hash_rows_kernel<HASH_SHA256><<<g, 256, 0, c->stream>>>((const uint32_t*)buf, row_bytes / 4, n, out);
    else return c->fail(GS_E_ARG, "unknown hash algorithm");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->cuda_fail(e, "hash_rows_kernel");
    c->launches++;
    return GS_OK;
}

// nodes: 2n digests with the leaves already at [n, 2n).  log_w > 0: this rank builds only its 1/W slice of every
// level down to its sub-tree root (node W + rank); the caller gathers the W roots and finishes the top.
// sink (single-GPU trees only): what the block that writes the root also does with it (RootSink); *sink_done says whether it did
static inline int merkle_build_range(Ctx* c, int alg, uint32_t* nodes, long long n, int log_w, int rank, const RootSink* sink = nullptr, bool* sink_done = nullptr) {
    if (alg != HASH_BLAKE2S && alg != HASH_SHA256) return c->fail(GS_E_ARG, "unknown hash algorithm");
    long long count = n >> 1;                  // parents of the current level (whole level)
    ProfScope ps(c, "merkle_build");
    while ((count >> log_w) > 32768) {         // throughput-bound levels: one launch each
        const long long cl = count >> log_w;
        const unsigned g = grid_for(c, cl, 256);
        if (alg == HASH_BLAKE2S) merkle_level_kernel<HASH_BLAKE2S><<<g, 256, 0, c->stream>>>(nodes, count, rank * cl, cl);
        else merkle_level_kernel<HASH_SHA256><<<g, 256, 0, c->stream>>>(nodes, count, rank * cl, cl);
        c->launches++;
        count >>= 1;
    }
    if (log_w == 0 && count >= 1) {
        // the rest of the tree in one launch; the level with 2 * count nodes holds at most 2^16 of them here
        const int level_nodes = (int)(2 * count);
        const unsigned blocks = level_nodes <= 512 ? 1u : (unsigned)(level_nodes / 512);
        RootSink none; memset(&none, 0, sizeof none);
        const RootSink& sk = sink ? *sink : none;
        HashCols nocols; memset(&nocols, 0, sizeof nocols);
        if (alg == HASH_BLAKE2S) merkle_top_kernel<HASH_BLAKE2S><<<blocks, 256, 0, c->stream>>>(nodes, level_nodes, c->counters, sk, nocols);
        else merkle_top_kernel<HASH_SHA256><<<blocks, 256, 0, c->stream>>>(nodes, level_nodes, c->counters, sk, nocols);
        if (sink_done) *sink_done = sink != nullptr;
        c->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return c->cuda_fail(e, "merkle_top_kernel");
        return GS_OK;
    }
    if ((count >> log_w) >= 2048) {            // latency-bound middle: 11 levels per block in shared memory
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(merkle_subtree_kernel<HASH_BLAKE2S, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            cudaFuncSetAttribute(merkle_subtree_kernel<HASH_SHA256, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            attr = true;
        }
        const long long blocks = count >> 10;  // the level with 2*count nodes splits into 2048-node subtrees
        const long long bl = blocks >> log_w;
        if (alg == HASH_BLAKE2S) merkle_subtree_kernel<HASH_BLAKE2S, 11><<<(unsigned)bl, 1024, 64 * 1024, c->stream>>>(nodes, blocks, rank * bl);
        else merkle_subtree_kernel<HASH_SHA256, 11><<<(unsigned)bl, 1024, 64 * 1024, c->stream>>>(nodes, blocks, rank * bl);
        c->launches++;
        count = blocks >> 1;
    }
    if (log_w == 0) {
        if (count >= 1) {
            if (alg == HASH_BLAKE2S) merkle_tail_kernel<HASH_BLAKE2S><<<1, 1024, 0, c->stream>>>(nodes, (int)count);
            else merkle_tail_kernel<HASH_SHA256><<<1, 1024, 0, c->stream>>>(nodes, (int)count);
            c->launches++;
        }
    } else {
        const int cl = (int)(count >> log_w);
        if (cl < 1) return c->fail(GS_E_ARG, "tree too small to shard");
        if (alg == HASH_BLAKE2S) merkle_tail_range_kernel<HASH_BLAKE2S><<<1, 1024, 0, c->stream>>>(nodes, cl, log_w, rank);
        else merkle_tail_range_kernel<HASH_SHA256><<<1, 1024, 0, c->stream>>>(nodes, cl, log_w, rank);
        c->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->cuda_fail(e, "merkle kernels");
    return GS_OK;
}
static inline int merkle_build(Ctx* c, int alg, uint32_t* nodes, long long n, const RootSink* sink = nullptr, bool* sink_done = nullptr) {
    return merkle_build_range(c, alg, nodes, n, 0, 0, sink, sink_done);
}

// One commit on a single GPU: leaves (rows of `cols`; nullptr: the digests already sit at nodes[n, 2n)) and the whole tree
// above them.  The throughput-bound levels (more than 2^16 parents) go in spans of up to three levels per launch
// (merkle_span_kernel), the first of them hashing the leaves itself when a leaf is a single block; everything from 2^17
// nodes down is the one-launch tree top (which hashes the leaves itself for trees that small).  The evaluation tree of
// the north-star shape (2^23 leaves): 3 launches instead of 9 (leaf kernel + 7 levels + top), a FRI layer of 2^19 rows:
// 2 instead of 5.  GS_MERKLE_FUSE=0: the separate leaf kernel and one launch per level (the round-1 structure, kept for A/B).
static inline bool merkle_fuse_enabled() {
    static const bool on = !(getenv("GS_MERKLE_FUSE") && atoi(getenv("GS_MERKLE_FUSE")) == 0);
    return on;
}
template <int ALG, int K, bool LEAF>
static inline void merkle_span_launch(Ctx* c, uint32_t* nodes, long long count, const HashCols& cols) {
    constexpr int S = 256 << (K - 1);
    constexpr int second = (LEAF && 16 * S < 256 * 64) ? 256 * 64 : (K > 1 ? 16 * S : 0);
    constexpr int smem = (K > 1 ? 32 * S : 0) + second;
    static bool attr = false;
    if (!attr && smem >= 48 * 1024) { cudaFuncSetAttribute(merkle_span_kernel<ALG, K, LEAF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    merkle_span_kernel<ALG, K, LEAF><<<(unsigned)(count / S), 256, smem, c->stream>>>(nodes, count, cols);
}
template <int ALG>
static inline int merkle_commit_alg(Ctx* c, const HashCols* cols, uint32_t* nodes, long long n, const RootSink* sink, bool* sink_done) {
    // the tree top takes the level with 2^(TOP_LOG + 1) nodes; GS_MERKLE_TOP_LOG (13..17) and GS_MERKLE_SPAN (1..3 levels per
    // launch) are measurement knobs
    static const int TOP_LOG = []() { const char* e = getenv("GS_MERKLE_TOP_LOG"); const int v = e ? atoi(e) : 16; return v < 13 ? 13 : (v > 17 ? 17 : v); }();
    static const int SPAN = []() { const char* e = getenv("GS_MERKLE_SPAN"); const int v = e ? atoi(e) : 3; return v < 1 ? 1 : (v > 3 ? 3 : v); }();
    HashCols hc; memset(&hc, 0, sizeof hc);
    if (cols) hc = *cols;
    bool leaf = cols != nullptr;                             // leaves still to be hashed by the next launch
    long long count = n >> 1;
    int rem = 0; while ((count >> rem) > (1ll << TOP_LOG)) ++rem;      // levels with more than 2^TOP_LOG parents
    int groups = (rem + SPAN - 1) / SPAN;
    for (int g = 0; g < groups; ++g) {
        const int k = (rem + (groups - g) - 1) / (groups - g);          // balanced: 4 levels -> 2 + 2, 6 -> 3 + 3
        if (leaf) {
            if (k == 1) merkle_span_launch<ALG, 1, true>(c, nodes, count, hc);
            else if (k == 2) merkle_span_launch<ALG, 2, true>(c, nodes, count, hc);
            else merkle_span_launch<ALG, 3, true>(c, nodes, count, hc);
        } else {
            if (k == 1) merkle_span_launch<ALG, 1, false>(c, nodes, count, hc);
            else if (k == 2) merkle_span_launch<ALG, 2, false>(c, nodes, count, hc);
            else merkle_span_launch<ALG, 3, false>(c, nodes, count, hc);
        }
        c->launches++;
        leaf = false; count >>= k; rem -= k;
    }
    const int level_nodes = (int)(2 * count);                // <= 2^(TOP_LOG + 1)
    const unsigned blocks = level_nodes <= 512 ? 1u : (unsigned)(level_nodes / 512);
    RootSink none; memset(&none, 0, sizeof none);
    const RootSink& sk = sink ? *sink : none;
    if (leaf) merkle_top_kernel<ALG, true><<<blocks, 256, 0, c->stream>>>(nodes, level_nodes, c->counters, sk, hc);
    else merkle_top_kernel<ALG, false><<<blocks, 256, 0, c->stream>>>(nodes, level_nodes, c->counters, sk, hc);
    c->launches++;
    if (sink_done) *sink_done = sink != nullptr;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->cuda_fail(e, "merkle commit kernels");
    return GS_OK;
}
static inline int merkle_commit(Ctx* c, int alg, const HashCols* cols, uint32_t* nodes, long long n, const RootSink* sink = nullptr, bool* sink_done = nullptr) {
    if (alg != HASH_BLAKE2S && alg != HASH_SHA256) return c->fail(GS_E_ARG, "unknown hash algorithm");
    if (n < 1 || (n & (n - 1))) return c->fail(GS_E_ARG, "a tree needs a power-of-two number of leaves");
    if (cols && (cols->ncols < 1 || cols->ncols > GS_MAX_HASH_COLS)) return c->fail(GS_E_ARG, "1..%d columns per leaf", GS_MAX_HASH_COLS);
    const bool fuse = merkle_fuse_enabled() && n >= 2;
    // leaves of more than one block (wide traces), or the A/B switch: the leaf kernel on its own
    int rc;
    if (cols && (!fuse || cols->ncols > 4)) { if ((rc = hash_columns(c, alg, *cols, n, nodes + 8 * n))) return rc; cols = nullptr; }
    if (!fuse) return n >= 2 ? merkle_build(c, alg, nodes, n, sink, sink_done) : GS_OK;
    ProfScope ps(c, cols ? "merkle_commit" : "merkle_build");          // merkle_commit = leaf hashing + tree in the same launches
    return alg == HASH_BLAKE2S ? merkle_commit_alg<HASH_BLAKE2S>(c, cols, nodes, n, sink, sink_done)
                               : merkle_commit_alg<HASH_SHA256>(c, cols, nodes, n, sink, sink_done);
}
// top of a sharded tree once the W sub-tree roots sit at nodes[W .. 2W)
static inline int merkle_build_top(Ctx* c, int alg, uint32_t* nodes, int world) {
    if (world < 2) return GS_OK;
    if (alg == HASH_BLAKE2S) merkle_tail_kernel<HASH_BLAKE2S><<<1, 1024, 0, c->stream>>>(nodes, world >> 1);
    else merkle_tail_kernel<HASH_SHA256><<<1, 1024, 0, c->stream>>>(nodes, world >> 1);
    c->launches++;
    return GS_OK;
}

}  // namespace gs
