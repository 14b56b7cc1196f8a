/* ORACLE (test infrastructure, never on the product path) -- plain-C port of the Python files in oracle/.
 *
 * A CPU restatement of genSTARK's prover (lib/Stark.ts:81-163 and lib/components/) over
 * GF(2^128 - 9*2^32 + 1), following the Python oracle function by function, including the reference's
 * unfused data flow (constraints on the composition domain -> iNTT(M) -> NTT(N); two full NTT(N) per
 * asserted register in BoundaryConstraints.evaluateAll, lib/components/BoundaryConstraints.ts:87-88).
 * It exists (a) to check the GPU proof bytes at sizes the Python oracle cannot reach and (b) as the
 * CPU baseline bench.py reports (kind "port": the reference's WASM path cannot run without node).
 * It is pinned to the Python oracle by tests/test_cport.py (identical proof bytes on small cases).
 * PARITY UNPINNED w.r.t. the real reference, like the rest of oracle/ (see oracle/field.py).
 * Loops over domain points are OpenMP-parallel; OMP_NUM_THREADS=1 gives the single-thread figure.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------ field */
static const u128 PM = (((u128)0xFFFFFFFFFFFFFFFFull) << 64) | 0xFFFFFFF700000001ull;
#define C9 (((u64)9 << 32) - 1)            /* 2^128 mod p */

static inline u128 fadd(u128 a, u128 b) { u128 s = a + b; if (s < a || s >= PM) s -= PM; return s; }
static inline u128 fsub(u128 a, u128 b) { return a >= b ? a - b : a + (PM - b); }
static inline u128 fmul(u128 a, u128 b) {
    u64 a0 = (u64)a, a1 = (u64)(a >> 64), b0 = (u64)b, b1 = (u64)(b >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    u128 mid = (p00 >> 64) + (u64)p01 + (u64)p10;
    u128 lo = ((u128)(u64)mid << 64) | (u64)p00;
    u128 hi = p11 + (p01 >> 64) + (p10 >> 64) + (mid >> 64);
    /* x = lo + hi * 2^128 == lo + hi * C9 ; hi * C9 is 164 bits: fold again */
    u128 h0 = (u128)(u64)hi * C9, h1 = (u128)(u64)(hi >> 64) * C9;      /* hi*C9 = h0 + h1 * 2^64 */
    u128 t = h0 + (h1 << 64);                                            /* low 128 bits */
    u64 carry = t < h0;
    u64 top = (u64)(h1 >> 64) + carry;                                   /* bits >= 128 (< 2^38) */
    u128 r = lo + t; u64 c2 = r < lo;
    u128 adj = (u128)(top + c2) * C9;                                    /* < 2^75 */
    u128 r2 = r + adj;
    if (r2 < r) r2 += C9;                                                /* wrapped: tiny value, cannot wrap again */
    if (r2 >= PM) r2 -= PM;
    return r2;
}
static u128 fpow(u128 b, u128 e) { u128 r = 1; while (e) { if (e & 1) r = fmul(r, b); b = fmul(b, b); e >>= 1; } return r; }
static u128 finv(u128 a) { return a == 0 ? 0 : fpow(a, PM - 2); }

static u128 root_of_unity(int log_order) {          /* oracle/field.py: get_root_of_unity */
    u128 pm1 = PM - 1;
    for (u128 i = 2;; ++i) {
        u128 g = fpow(i, pm1 >> log_order);
        if (log_order == 0) return g;
        if (fpow(g, (u128)1 << (log_order - 1)) != 1) return g;
    }
}

static u128* vnew(size_t n) { u128* p = (u128*)malloc(sizeof(u128) * (n ? n : 1)); if (!p) { fprintf(stderr, "oracle: out of memory\n"); exit(2); } return p; }

static u128* power_series(u128 base, size_t n) {
    u128* out = vnew(n);
    /* blocked so it parallelises: out[i] = base^i */
    const size_t B = 4096;
    size_t nb = (n + B - 1) / B;
    u128 step = fpow(base, B);
    u128* starts = vnew(nb);
    u128 a = 1; for (size_t b = 0; b < nb; ++b) { starts[b] = a; a = fmul(a, step); }
#pragma omp parallel for schedule(static)
    for (long b = 0; b < (long)nb; ++b) {
        u128 v = starts[b]; size_t end = (b + 1) * B < n ? (b + 1) * B : n;
        for (size_t i = (size_t)b * B; i < end; ++i) { out[i] = v; v = fmul(v, base); }
    }
    free(starts);
    return out;
}

/* Montgomery batch inversion skipping zeros (inv(0) = 0), oracle/field.py: inv_vector_elements */
static void batch_inverse(const u128* v, u128* out, size_t n) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
        size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) continue;
        u128 acc = 1;
        for (size_t i = lo; i < hi; ++i) { out[i] = acc; if (v[i]) acc = fmul(acc, v[i]); }
        u128 inv = finv(acc);
        for (size_t i = hi; i-- > lo;) {
            u128 x = v[i];
            if (x) { u128 r = fmul(inv, out[i]); inv = fmul(inv, x); out[i] = r; } else out[i] = 0;
        }
    }
}

/* in-place radix-2 DFT over the power series of `root` (order n): out[k] = sum v[j] root^(jk), natural order */
static void fft(u128* v, int log_n, u128 root) {
    size_t n = (size_t)1 << log_n;
    if (n == 1) return;
    /* bit reversal */
    for (size_t i = 0; i < n; ++i) {
        size_t j = 0; for (int b = 0; b < log_n; ++b) j |= ((i >> b) & 1) << (log_n - 1 - b);
        if (j > i) { u128 t = v[i]; v[i] = v[j]; v[j] = t; }
    }
    u128* tw = power_series(root, n / 2);
    for (int s = 1; s <= log_n; ++s) {
        size_t len = (size_t)1 << s, half = len >> 1, stride = n / len;
#pragma omp parallel for schedule(static)
        for (long q = 0; q < (long)(n / 2); ++q) {
            size_t blk = (size_t)q / half, i = (size_t)q % half;
            size_t a = blk * len + i, b = a + half;
            u128 t = fmul(v[b], tw[i * stride]);
            u128 u = v[a];
            v[a] = fadd(u, t); v[b] = fsub(u, t);
        }
    }
    free(tw);
}
/* eval_poly_at_roots: zero-pad poly (len t) to the domain of size 2^log_n generated by `root` */
static u128* eval_poly_at_roots(const u128* poly, size_t t, int log_n, u128 root) {
    size_t n = (size_t)1 << log_n;
    u128* v = vnew(n);
    memcpy(v, poly, t * sizeof(u128));
    memset(v + t, 0, (n - t) * sizeof(u128));
    fft(v, log_n, root);
    return v;
}
static u128* interpolate_roots(const u128* values, int log_n, u128 root) {
    size_t n = (size_t)1 << log_n;
    u128* v = vnew(n);
    memcpy(v, values, n * sizeof(u128));
    fft(v, log_n, finv(root));
    u128 ninv = finv((u128)n);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) v[i] = fmul(v[i], ninv);
    return v;
}
static u128 eval_poly_at(const u128* poly, size_t n, u128 x) { u128 acc = 0; for (size_t k = n; k-- > 0;) acc = fadd(fmul(acc, x), poly[k]); return acc; }

/* Lagrange interpolation (oracle/field.py: interpolate), low -> high; out has n coefficients */
static void lagrange(const u128* xs, const u128* ys, size_t n, u128* out) {
    u128* root = vnew(n + 1); u128* num = vnew(n);
    memset(root, 0, (n + 1) * sizeof(u128)); root[0] = 1;
    for (size_t i = 0; i < n; ++i) {
        for (size_t k = i + 1; k > 0; --k) root[k] = fsub(root[k - 1], fmul(root[k], xs[i]));
        root[0] = fsub(0, fmul(root[0], xs[i]));
    }
    memset(out, 0, n * sizeof(u128));
    for (size_t i = 0; i < n; ++i) {
        u128 acc = 0;
        for (size_t k = n; k > 0; --k) { acc = fadd(root[k], fmul(acc, xs[i])); num[k - 1] = acc; }
        u128 den = eval_poly_at(num, n, xs[i]);
        u128 f = fmul(ys[i], finv(den));
        for (size_t k = 0; k < n; ++k) out[k] = fadd(out[k], fmul(num[k], f));
    }
    free(root); free(num);
}

/* ------------------------------------------------------------------------------------------ hashes */
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha256_block(uint32_t h[8], const uint8_t* p) {
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
    for (int i = 16; i < 64; ++i) w[i] = w[i - 16] + (rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] + (rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10));
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
        uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha256(const uint8_t* msg, size_t n, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t i = 0;
    for (; i + 64 <= n; i += 64) sha256_block(h, msg + i);
    uint8_t buf[128]; size_t rem = n - i; memcpy(buf, msg + i, rem); buf[rem] = 0x80;
    size_t padded = rem + 1 + 8 <= 64 ? 64 : 128;
    memset(buf + rem + 1, 0, padded - rem - 1);
    u64 bits = (u64)n * 8; for (int k = 0; k < 8; ++k) buf[padded - 1 - k] = (uint8_t)(bits >> (8 * k));
    sha256_block(h, buf); if (padded == 128) sha256_block(h, buf + 64);
    for (int k = 0; k < 8; ++k) { out[4 * k] = h[k] >> 24; out[4 * k + 1] = h[k] >> 16; out[4 * k + 2] = h[k] >> 8; out[4 * k + 3] = h[k]; }
}
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static void blake2s(const uint8_t* msg, size_t n, uint8_t out[32]) {
    static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
    uint32_t h[8]; memcpy(h, IV, 32); h[0] ^= 0x01010020;
    size_t off = 0;
    for (;;) {
        size_t take = n - off > 64 ? 64 : n - off;
        int last = (off + take == n);
        if (!last && take < 64) break;
        uint8_t blk[64]; memset(blk, 0, 64); memcpy(blk, msg + off, take);
        off += take;
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; ++i) m[i] = (uint32_t)blk[4 * i] | ((uint32_t)blk[4 * i + 1] << 8) | ((uint32_t)blk[4 * i + 2] << 16) | ((uint32_t)blk[4 * i + 3] << 24);
        for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = IV[i]; }
        v[12] ^= (uint32_t)off; v[13] ^= (uint32_t)((u64)off >> 32);
        if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y) v[a] += v[b] + x; v[d] = rotr(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotr(v[b] ^ v[c], 12); \
                             v[a] += v[b] + y; v[d] = rotr(v[d] ^ v[a], 8); v[c] += v[d]; v[b] = rotr(v[b] ^ v[c], 7);
        for (int r = 0; r < 10; ++r) {
            const uint8_t* s = B2S_SIGMA[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]) G(1, 5, 9, 13, m[s[2]], m[s[3]]) G(2, 6, 10, 14, m[s[4]], m[s[5]]) G(3, 7, 11, 15, m[s[6]], m[s[7]])
            G(0, 5, 10, 15, m[s[8]], m[s[9]]) G(1, 6, 11, 12, m[s[10]], m[s[11]]) G(2, 7, 8, 13, m[s[12]], m[s[13]]) G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef G
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
        if (last) break;
    }
    for (int i = 0; i < 8; ++i) { out[4 * i] = (uint8_t)h[i]; out[4 * i + 1] = (uint8_t)(h[i] >> 8); out[4 * i + 2] = (uint8_t)(h[i] >> 16); out[4 * i + 3] = (uint8_t)(h[i] >> 24); }
}
static int g_alg = 0;   /* 0 sha256, 1 blake2s256 */
static void hdigest(const uint8_t* msg, size_t n, uint8_t out[32]) { if (g_alg) blake2s(msg, n, out); else sha256(msg, n, out); }

static void put_elem(uint8_t* dst, u128 v) { for (int i = 0; i < 16; ++i) dst[i] = (uint8_t)(v >> (8 * i)); }

/* tree: 2n digests, nodes[1] root, leaves at n..2n-1 (oracle/merkle.py) */
static uint8_t* merkle_create(const uint8_t* leaves, size_t n) {
    uint8_t* nodes = (uint8_t*)malloc(64 * n);
    memcpy(nodes + 32 * n, leaves, 32 * n);
    for (size_t lvl = n >> 1; lvl >= 1; lvl >>= 1) {
#pragma omp parallel for schedule(static)
        for (long i = (long)lvl; i < (long)(2 * lvl); ++i) hdigest(nodes + 64 * (size_t)i, 64, nodes + 32 * (size_t)i);
        if (lvl == 1) break;
    }
    return nodes;
}

/* --------------------------------------------------------------------------------- prng / indexes */
typedef struct { uint8_t b[33]; } be256;
static void sha_of_be(const be256* v, uint8_t out[32]) {
    uint8_t nib[66]; for (int i = 0; i < 33; ++i) { nib[2 * i] = v->b[i] >> 4; nib[2 * i + 1] = v->b[i] & 15; }
    int first = 0; while (first < 66 && nib[first] == 0) ++first;
    int nd = 66 - first; if (nd == 0) { nd = 1; first = 65; }
    uint8_t buf[33]; int nb = nd / 2;                 /* odd digit count: the LAST nibble is dropped (App. E.2) */
    for (int k = 0; k < nb; ++k) buf[k] = (uint8_t)((nib[first + 2 * k] << 4) | nib[first + 2 * k + 1]);
    sha256(buf, (size_t)nb, out);
}
static be256 be_add(const uint8_t d[32], u64 x) {
    be256 r; r.b[0] = 0; memcpy(r.b + 1, d, 32);
    for (int i = 32; i >= 0 && x; --i) { u64 s = r.b[i] + (x & 0xFF); r.b[i] = (uint8_t)s; x = (x >> 8) + (s >> 8); }
    return r;
}
static u128 digest_mod_p(const uint8_t d[32]) {
    u128 hi = 0, lo = 0; for (int i = 0; i < 16; ++i) hi = (hi << 8) | d[i]; for (int i = 16; i < 32; ++i) lo = (lo << 8) | d[i];
    if (hi >= PM) hi -= PM;
    if (lo >= PM) lo -= PM;
    return fadd(lo, fmul(hi, (u128)C9));
}
static void prng_many(const uint8_t* seed, size_t seed_len, int n, u128* out) {
    uint8_t st[32]; sha256(seed, seed_len, st);
    for (int i = 0; i < n; ++i) { be256 v = be_add(st, (u64)i); uint8_t d[32]; sha_of_be(&v, d); out[i] = digest_mod_p(d); }
}
static u128 prng_one(const uint8_t* seed, size_t seed_len) { uint8_t d[32]; sha256(seed, seed_len, d); return digest_mod_p(d); }
/* get_pseudorandom_indexes; max is a power of two */
static int pr_indexes(const uint8_t seed[32], int count, u64 max, u64 skip, uint32_t* out) {
    u64 maxc = skip ? max - max / skip : max;
    if (maxc < (u64)count) return -1;
    uint8_t st[32]; sha256(seed, 32, st);
    int got = 0;
    for (long i = 0; i < (long)count * 1000 && got < count; ++i) {
        be256 v = be_add(st, (u64)i); uint8_t d[32]; sha_of_be(&v, d);
        u64 low = 0; for (int k = 24; k < 32; ++k) low = (low << 8) | d[k];
        u64 idx = low & (max - 1);
        if (skip && idx % skip == 0) continue;
        int dup = 0; for (int k = 0; k < got; ++k) if (out[k] == idx) { dup = 1; break; }
        if (dup) continue;
        out[got++] = (uint32_t)idx;
    }
    return got == count ? 0 : -1;
}

/* ------------------------------------------------------------------------------------ byte output */
typedef struct { uint8_t* p; size_t n, cap; } bytes;
static void bput(bytes* b, const void* src, size_t n) {
    if (b->n + n > b->cap) { b->cap = (b->n + n) * 2 + 1024; b->p = (uint8_t*)realloc(b->p, b->cap); }
    memcpy(b->p + b->n, src, n); b->n += n;
}
static void bput8(bytes* b, unsigned v) { uint8_t x = (uint8_t)v; bput(b, &x, 1); }

/* MerkleTree.prove_batch + write_merkle_proof (oracle/merkle.py, oracle/stark.py); values = raw leaf bytes */
static void write_batch_proof(bytes* out, const uint8_t* nodes, size_t n, const uint32_t* idx, int nidx,
                              const uint8_t* values, size_t value_size, size_t leaf_size) {
    int depth = 0; while (((size_t)1 << depth) < n) ++depth;
    /* normalized = sorted unique even-aligned */
    uint32_t* sorted = (uint32_t*)malloc(4 * (size_t)nidx); memcpy(sorted, idx, 4 * (size_t)nidx);
    for (int i = 1; i < nidx; ++i) { uint32_t v = sorted[i]; int j = i - 1; while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; --j; } sorted[j + 1] = v; }
    uint32_t* norm = (uint32_t*)malloc(4 * (size_t)nidx); int nn = 0;
    for (int i = 0; i < nidx; ++i) { uint32_t e = sorted[i] & ~1u; if (nn == 0 || norm[nn - 1] != e) norm[nn++] = e; }
    /* columns of node ids */
    uint32_t** cols = (uint32_t**)calloc((size_t)nn, sizeof(uint32_t*)); int* clen = (int*)calloc((size_t)nn, sizeof(int));
    for (int i = 0; i < nn; ++i) cols[i] = (uint32_t*)malloc(4 * (size_t)(depth + 2));
    u64* cur = (u64*)malloc(8 * (size_t)nn); u64* nxt = (u64*)malloc(8 * (size_t)nn); int ncur = 0;
    for (int i = 0; i < nn; ++i) {
        int has1 = 0, has2 = 0;
        for (int k = 0; k < nidx; ++k) { if (idx[k] == norm[i]) has1 = 1; if (idx[k] == norm[i] + 1) has2 = 1; }
        if (has1 && !has2) cols[i][clen[i]++] = (uint32_t)(n + norm[i] + 1);
        else if (!has1) cols[i][clen[i]++] = (uint32_t)(n + norm[i]);
        cur[ncur++] = (norm[i] + n) >> 1;
    }
    for (int d = depth - 1; d > 0; --d) {
        int nnx = 0;
        for (int i = 0; i < ncur; ++i) {
            u64 sib = cur[i] ^ 1; int col = i;
            if (i + 1 < ncur && cur[i + 1] == sib) ++i;
            else cols[col][clen[col]++] = (uint32_t)sib;
            nxt[nnx++] = sib >> 1;
        }
        u64* t = cur; cur = nxt; nxt = t; ncur = nnx;
    }
    bput8(out, nidx == 256 ? 0 : (unsigned)nidx);
    bput(out, values, (size_t)nidx * value_size);
    bput8(out, nn == 256 ? 0 : (unsigned)nn);
    for (int i = 0; i < nn; ++i) bput8(out, (unsigned)(((clen[i] << 1) | ((clen[i] > 0 && leaf_size == 32) ? 1 : 0)) & 0xFF));
    for (int i = 0; i < nn; ++i) for (int j = 0; j < clen[i]; ++j) bput(out, nodes + 32 * (size_t)cols[i][j], 32);
    bput8(out, (unsigned)depth);
    for (int i = 0; i < nn; ++i) free(cols[i]);
    free(cols); free(clen); free(cur); free(nxt); free(sorted); free(norm);
}

/* ---------------------------------------------------------------------------------------- AIR blob */
enum { OP_CONST = 0, OP_CUR, OP_NEXT, OP_STATIC, OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_INV, OP_EXP, OP_OUT };
typedef struct { uint32_t n_instr, n_const, n_slots, n_out; const uint32_t* instrs; u128* consts; } program;
typedef struct { int kind; uint32_t len; u128* values; } static_reg;
typedef struct { int R, K, log_t, log_e, n_static; static_reg* statics; uint32_t* degrees; program transition, evaluation; } air_t;

static uint32_t rd32(const uint8_t** p) { uint32_t v; memcpy(&v, *p, 4); *p += 4; return v; }
static u128 rd_elem(const uint8_t** p) { u128 v = 0; for (int i = 15; i >= 0; --i) v = (v << 8) | (*p)[i]; *p += 16; return v; }
static void read_program(const uint8_t** p, program* pr) {
    pr->n_instr = rd32(p); pr->n_const = rd32(p); pr->n_slots = rd32(p); pr->n_out = rd32(p);
    uint32_t* ins = (uint32_t*)malloc(16 * (size_t)pr->n_instr + 16); memcpy(ins, *p, 16 * (size_t)pr->n_instr); *p += 16 * (size_t)pr->n_instr; pr->instrs = ins;
    pr->consts = vnew(pr->n_const); for (uint32_t i = 0; i < pr->n_const; ++i) pr->consts[i] = rd_elem(p);
}
static int parse_air(const uint8_t* blob, air_t* a) {
    const uint8_t* p = blob;
    if (rd32(&p) != 0x52494147u) return -1;
    p += 16;  /* modulus (p128 only) */
    a->R = (int)rd32(&p); a->K = (int)rd32(&p); a->log_t = (int)rd32(&p); a->log_e = (int)rd32(&p); a->n_static = (int)rd32(&p);
    a->statics = (static_reg*)calloc((size_t)a->n_static + 1, sizeof(static_reg));
    for (int k = 0; k < a->n_static; ++k) {
        a->statics[k].kind = (int)rd32(&p); a->statics[k].len = rd32(&p);
        a->statics[k].values = vnew(a->statics[k].len);
        for (uint32_t i = 0; i < a->statics[k].len; ++i) a->statics[k].values[i] = rd_elem(&p);
    }
    a->degrees = (uint32_t*)malloc(4 * (size_t)a->K); for (int k = 0; k < a->K; ++k) a->degrees[k] = rd32(&p);
    read_program(&p, &a->transition); read_program(&p, &a->evaluation);
    return 0;
}
static void run_program(const program* pr, const u128* cur, const u128* nxt, const u128* st, u128* slots, u128* out) {
    for (uint32_t i = 0; i < pr->n_instr; ++i) {
        const uint32_t* in = pr->instrs + 4 * i; uint32_t d = in[1], a = in[2], b = in[3];
        switch (in[0]) {
            case OP_CONST: slots[d] = pr->consts[a]; break;
            case OP_CUR: slots[d] = cur[a]; break;
            case OP_NEXT: slots[d] = nxt[a]; break;
            case OP_STATIC: slots[d] = st[a]; break;
            case OP_ADD: slots[d] = fadd(slots[a], slots[b]); break;
            case OP_SUB: slots[d] = fsub(slots[a], slots[b]); break;
            case OP_MUL: slots[d] = fmul(slots[a], slots[b]); break;
            case OP_NEG: slots[d] = fsub(0, slots[a]); break;
            case OP_INV: slots[d] = finv(slots[a]); break;
            case OP_EXP: slots[d] = fpow(slots[a], pr->consts[b]); break;
            case OP_OUT: out[d] = slots[a]; break;
        }
    }
}

/* -------------------------------------------------------------------------------------------- prove */
static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }
static char g_err[512];
const char* oracle_last_error(void) { return g_err; }

typedef struct { uint32_t reg, step; u128 value; } assertion;

/* hash rows of a "transposeVector(v, 4)" matrix: row i = v[i], v[i+q], v[i+2q], v[i+3q] */
static uint8_t* hash_rows4(const u128* v, size_t q) {
    uint8_t* leaves = (uint8_t*)malloc(32 * q);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)q; ++i) {
        uint8_t row[64]; for (int j = 0; j < 4; ++j) put_elem(row + 16 * j, v[(size_t)i + (size_t)j * q]);
        hdigest(row, 64, leaves + 32 * (size_t)i);
    }
    return leaves;
}
static int first_seen_unique(const uint32_t* in, int n, uint32_t* out) {
    int m = 0;
    for (int i = 0; i < n; ++i) { int dup = 0; for (int k = 0; k < m; ++k) if (out[k] == in[i]) { dup = 1; break; } if (!dup) out[m++] = in[i]; }
    return m;
}
static void rows4_values(const u128* v, size_t q, const uint32_t* idx, int n, uint8_t* out) {
    for (int k = 0; k < n; ++k) for (int j = 0; j < 4; ++j) put_elem(out + 64 * (size_t)k + 16 * j, v[idx[k] + (size_t)j * q]);
}

int oracle_prove(const uint8_t* air_blob, int hash_alg, int exe_queries, int fri_queries,
                 const uint8_t* assertions_blob, int n_assert, const uint8_t* init16, const uint8_t* input_traces,
                 const uint8_t* shapes, size_t shapes_len, uint8_t** proof_out, size_t* proof_len, double* stage_ms /* [12] or NULL */) {
    air_t A; g_err[0] = 0;
    if (parse_air(air_blob, &A) != 0) { snprintf(g_err, sizeof g_err, "bad AIR blob"); return -2; }
    g_alg = hash_alg;
    const int R = A.R, K = A.K, log_t = A.log_t, log_e = A.log_e, log_n = log_t + log_e;
    const size_t T = (size_t)1 << log_t, N = (size_t)1 << log_n, E = (size_t)1 << log_e;
    double t0 = now_ms(); int st_i = 0;
#define STAGE() do { if (stage_ms && st_i < 12) { double t = now_ms(); stage_ms[st_i++] = t - t0; t0 = t; } } while (0)
    assertion* as = (assertion*)malloc(sizeof(assertion) * (size_t)n_assert);
    for (int i = 0; i < n_assert; ++i) { const uint8_t* p = assertions_blob + 24 * (size_t)i; memcpy(&as[i].reg, p, 4); memcpy(&as[i].step, p + 4, 4); const uint8_t* q = p + 8; as[i].value = rd_elem(&q); }
    if (N < 128) { snprintf(g_err, sizeof g_err, "Low degree proof failed: Invalid array length"); return -4; }
    /* 1 ---- context */
    const u128 w = root_of_unity(log_n);
    u128* dom = power_series(w, N);                              /* evaluationDomain */
    int max_deg = 1; for (int k = 0; k < K; ++k) if ((int)A.degrees[k] > max_deg) max_deg = (int)A.degrees[k];
    int log_c = 0; while ((1 << log_c) < max_deg) ++log_c;
    const size_t c = (size_t)1 << log_c, M = T * c;
    const int log_m = log_t + log_c;
    const u128 w_t = fpow(w, E), w_m = fpow(w, N / M);
    int n_in = 0, n_secret = 0; for (int k = 0; k < A.n_static; ++k) if (A.statics[k].kind != 0) { ++n_in; if (A.statics[k].kind == 1) ++n_secret; }
    /* input registers: polys, composition-domain evaluations, secret ones over the evaluation domain */
    u128** in_poly = (u128**)calloc((size_t)n_in + 1, sizeof(u128*));
    u128** in_eval_n = (u128**)calloc((size_t)n_in + 1, sizeof(u128*));
    {
        int ii = 0;
        for (int k = 0; k < A.n_static; ++k) if (A.statics[k].kind != 0) {
            u128* tr = vnew(T); const uint8_t* p = input_traces + 16 * T * (size_t)ii;
            for (size_t s = 0; s < T; ++s) tr[s] = rd_elem(&p);
            in_poly[ii] = interpolate_roots(tr, log_t, w_t); free(tr);
            if (A.statics[k].kind == 1) in_eval_n[ii] = eval_poly_at_roots(in_poly[ii], T, log_n, w);
            ++ii;
        }
    }
    STAGE();
    /* 2 ---- execution trace */
    u128** trace = (u128**)malloc(sizeof(u128*) * (size_t)R);
    for (int r = 0; r < R; ++r) trace[r] = vnew(T);
    {
        u128* state = vnew(R); u128* next = vnew(R); u128* slots = vnew(A.transition.n_slots + 1); u128* sv = vnew(A.n_static + 1);
        const uint8_t* p = init16; for (int r = 0; r < R; ++r) state[r] = rd_elem(&p);
        for (size_t s = 0; s < T; ++s) {
            for (int r = 0; r < R; ++r) trace[r][s] = state[r];
            if (s + 1 < T) {
                int ii = 0;
                for (int k = 0; k < A.n_static; ++k) {
                    if (A.statics[k].kind == 0) sv[k] = A.statics[k].values[s % A.statics[k].len];
                    else { const uint8_t* q = input_traces + 16 * (T * (size_t)ii + s); sv[k] = rd_elem(&q); ++ii; }
                }
                run_program(&A.transition, state, NULL, sv, slots, next);
                u128* t = state; state = next; next = t;
            }
        }
        free(state); free(next); free(slots); free(sv);
        for (int i = 0; i < n_assert; ++i) {
            if ((int)as[i].reg >= R || as[i].step >= T) { snprintf(g_err, sizeof g_err, "Failed to generate the execution trace: Invalid assertion"); return -4; }
            if (trace[as[i].reg][as[i].step] != as[i].value) { snprintf(g_err, sizeof g_err, "Failed to generate the execution trace: Assertion at step %u, register %u conflicts with execution trace", as[i].step, as[i].reg); return -4; }
        }
    }
    STAGE();
    /* 3 ---- P(x) and its low-degree extension */
    u128** ppoly = (u128**)malloc(sizeof(u128*) * (size_t)R); u128** pe = (u128**)malloc(sizeof(u128*) * (size_t)R);
    for (int r = 0; r < R; ++r) ppoly[r] = interpolate_roots(trace[r], log_t, w_t);
    STAGE();
    for (int r = 0; r < R; ++r) pe[r] = eval_poly_at_roots(ppoly[r], T, log_n, w);
    STAGE();
    /* 4 ---- leaves and tree */
    const int n_ev = R + n_secret;
    const u128** ev = (const u128**)malloc(sizeof(u128*) * (size_t)n_ev);
    { int q = 0; for (int r = 0; r < R; ++r) ev[q++] = pe[r]; int ii = 0; for (int k = 0; k < A.n_static; ++k) if (A.statics[k].kind != 0) { if (A.statics[k].kind == 1) ev[q++] = in_eval_n[ii]; ++ii; } }
    uint8_t* leaves = (uint8_t*)malloc(32 * N);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; ++i) {
        uint8_t buf[16 * 64];
        for (int q = 0; q < n_ev; ++q) put_elem(buf + 16 * q, ev[q][i]);
        hdigest(buf, 16 * (size_t)n_ev, leaves + 32 * (size_t)i);
    }
    STAGE();
    uint8_t* e_tree = merkle_create(leaves, N); free(leaves);
    const uint8_t* ev_root = e_tree + 32;
    STAGE();
    /* 5 ---- composition polynomial */
    const size_t comb_degree = T << log_c;
    const size_t comp_degree = comb_degree - T > T ? comb_degree - T : T;
    size_t group_deg[64]; int group_n = 0; int group_of[64];
    for (int k = 0; k < K; ++k) { size_t dg = (size_t)A.degrees[k] * T; int g = 0; for (; g < group_n; ++g) if (group_deg[g] == dg) break; if (g == group_n) group_deg[group_n++] = dg; group_of[k] = g; }
    int d_count = K; for (int k = 0; k < K; ++k) if (group_deg[group_of[k]] < comb_degree) ++d_count;
    uint32_t b_regs[64]; int nB = 0; u128* b_xs[64]; u128* b_ys[64]; int b_n[64];
    for (int i = 0; i < n_assert; ++i) {
        int b = 0; for (; b < nB; ++b) if (b_regs[b] == as[i].reg) break;
        if (b == nB) { b_regs[nB] = as[i].reg; b_xs[nB] = vnew((size_t)n_assert); b_ys[nB] = vnew((size_t)n_assert); b_n[nB] = 0; ++nB; }
        b_xs[b][b_n[b]] = fpow(w, (u128)as[i].step * E); b_ys[b][b_n[b]] = as[i].value; ++b_n[b];
    }
    int b_count = nB * (comp_degree > T ? 2 : 1);
    u128* coeffs = vnew((size_t)(d_count + b_count));
    prng_many(ev_root, 32, d_count + b_count, coeffs);
    /* 5.1 transition constraints over the composition domain (air-assembly evaluateTransitionConstraints) */
    u128** t_ev = (u128**)malloc(sizeof(u128*) * (size_t)R);
    for (int r = 0; r < R; ++r) t_ev[r] = eval_poly_at_roots(ppoly[r], T, log_m, w_m);
    u128** s_ev = (u128**)calloc((size_t)A.n_static + 1, sizeof(u128*)); size_t* s_len = (size_t*)calloc((size_t)A.n_static + 1, sizeof(size_t));
    {
        int ii = 0;
        for (int k = 0; k < A.n_static; ++k) {
            if (A.statics[k].kind == 0) {
                size_t L = A.statics[k].len; int log_l = 0; while (((size_t)1 << log_l) < L) ++log_l;
                u128 g_l = fpow(w, N / L);
                u128* poly = interpolate_roots(A.statics[k].values, log_l, g_l);
                s_ev[k] = eval_poly_at_roots(poly, L, log_l + log_c, fpow(w, N / (L * c))); s_len[k] = L * c; free(poly);
            } else { s_ev[k] = eval_poly_at_roots(in_poly[ii], T, log_m, w_m); s_len[k] = M; ++ii; }
        }
    }
    u128** q_ev = (u128**)malloc(sizeof(u128*) * (size_t)(2 * K));
    for (int k = 0; k < K; ++k) q_ev[k] = vnew(M);
    int fail_k = -1; long fail_step = -1;
#pragma omp parallel
    {
        u128 cur[64], nxt[64], sv[64], qv[64]; u128* slots = vnew(A.evaluation.n_slots + 1);
#pragma omp for schedule(static)
        for (long pos = 0; pos < (long)M; ++pos) {
            for (int r = 0; r < R; ++r) { cur[r] = t_ev[r][pos]; nxt[r] = t_ev[r][((size_t)pos + c) % M]; }
            for (int k = 0; k < A.n_static; ++k) sv[k] = s_ev[k][(size_t)pos % s_len[k]];
            run_program(&A.evaluation, cur, nxt, sv, slots, qv);
            if ((size_t)pos % c == 0 && (size_t)pos < M - c) for (int k = 0; k < K; ++k) if (qv[k] != 0) {
#pragma omp critical
                { if (fail_k < 0) { fail_k = k; fail_step = pos / (long)c; } }
            }
            for (int k = 0; k < K; ++k) q_ev[k][pos] = qv[k];
        }
        free(slots);
    }
    if (fail_k >= 0) { snprintf(g_err, sizeof g_err, "Failed to evaluate transition constraints: Constraint %d didn't evaluate to 0 at step %ld", fail_k, fail_step); return -4; }
    /* 5.2 degree adjustment + linear combination (CompositionPolynomial.ts:84-105) */
    int n_qa = K;
    for (int g = 0; g < group_n; ++g) {
        if (group_deg[g] == comb_degree) continue;
        u128* powers = power_series(fpow(w_m, (u128)(comb_degree - group_deg[g])), M);
        for (int k = 0; k < K; ++k) if (group_of[k] == g) {
            u128* adj = vnew(M);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)M; ++i) adj[i] = fmul(q_ev[k][i], powers[i]);
            q_ev[n_qa++] = adj;
        }
        free(powers);
    }
    u128* qc = vnew(M);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)M; ++i) { u128 acc = 0; for (int m = 0; m < n_qa; ++m) acc = fadd(acc, fmul(q_ev[m][i], coeffs[m])); qc[i] = acc; }
    for (int m = 0; m < n_qa; ++m) free(q_ev[m]);
    /* 5.3 composition domain -> evaluation domain (:109-110) */
    u128* qc_poly = interpolate_roots(qc, log_m, w_m); free(qc);
    u128* qe = eval_poly_at_roots(qc_poly, M, log_n, w); free(qc_poly);
    /* 5.4 D = Q / Z (ZeroPolynomial.ts:36-45, CompositionPolynomial.ts:114-120) */
    const u128 x_last = fpow(w, (u128)(T - 1) * E);
    u128* num = vnew(N); u128* zinv = vnew(N);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; ++i) num[i] = fsub(dom[((size_t)i * T) % N], 1);
    batch_inverse(num, zinv, N);
    u128* C = vnew(N);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; ++i) C[i] = fmul(qe[i], fmul(fsub(dom[i], x_last), zinv[i]));
    free(qe); free(num);
    /* 5.5 boundary constraints (BoundaryConstraints.ts:71-95: two full NTTs per asserted register) */
    u128** ba = (u128**)malloc(sizeof(u128*) * (size_t)(2 * nB + 1));
    for (int b = 0; b < nB; ++b) {
        int na = b_n[b];
        u128* ip = vnew((size_t)na); lagrange(b_xs[b], b_ys[b], (size_t)na, ip);
        u128* zp = vnew((size_t)na + 1); memset(zp, 0, sizeof(u128) * ((size_t)na + 1)); zp[0] = 1;
        for (int k = 0; k < na; ++k) { for (int j = k + 1; j > 0; --j) zp[j] = fsub(zp[j - 1], fmul(zp[j], b_xs[b][k])); zp[0] = fsub(0, fmul(zp[0], b_xs[b][k])); }
        u128* iv = eval_poly_at_roots(ip, (size_t)na, log_n, w);
        u128* zv = eval_poly_at_roots(zp, (size_t)na + 1, log_n, w);
        batch_inverse(zv, zinv, N);
        const u128* pr = pe[b_regs[b]];
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i) iv[i] = fmul(fsub(pr[i], iv[i]), zinv[i]);
        ba[b] = iv; free(zv); free(ip); free(zp);
    }
    int n_ba = nB;
    const size_t delta = comp_degree - T;
    u128* xdelta = NULL;
    if (delta > 0) {
        xdelta = power_series(fpow(w, (u128)delta), N);
        for (int b = 0; b < nB; ++b) { u128* adj = vnew(N);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)N; ++i) adj[i] = fmul(ba[b][i], xdelta[i]);
            ba[n_ba++] = adj; }
    }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; ++i) { u128 acc = 0; for (int m = 0; m < n_ba; ++m) acc = fadd(acc, fmul(ba[m][i], coeffs[d_count + m])); C[i] = fadd(C[i], acc); }
    for (int m = 0; m < n_ba; ++m) free(ba[m]);
    STAGE();
    /* 6 ---- linear combination (LinearCombination.ts:36-64) */
    const int lc_n = n_ev * (delta > 0 ? 2 : 1);
    u128* lcc = vnew((size_t)(d_count + b_count + lc_n));
    prng_many(ev_root, 32, d_count + b_count + lc_n, lcc);
    const u128* lk = lcc + d_count + b_count;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; ++i) {
        u128 acc = 0;
        for (int m = 0; m < n_ev; ++m) acc = fadd(acc, fmul(ev[m][i], lk[m]));
        if (delta > 0) for (int m = 0; m < n_ev; ++m) acc = fadd(acc, fmul(fmul(ev[m][i], xdelta[i]), lk[n_ev + m]));
        C[i] = fadd(C[i], acc);
    }
    u128* Lv = C;
    STAGE();
    /* 7 ---- low-degree proof (LowDegreeProver.ts) */
    bytes out = {0, 0, 0};
    bput(&out, ev_root, 32);
    /* layers */
    u128* layer_v[40]; uint8_t* layer_tree[40]; size_t layer_len[40]; int n_layers = 0;
    u128* remainder = NULL; size_t rem_len = 0;
    {
        u128* v = Lv; size_t L = N; size_t maxdeg = comp_degree; int depth = 0;
        const u128 iota = fpow(w, N / 4);
        for (;;) {
            size_t q = L >> 2;
            uint8_t* lv = hash_rows4(v, q); uint8_t* tree = merkle_create(lv, q); free(lv);
            layer_v[n_layers] = v; layer_tree[n_layers] = tree; layer_len[n_layers] = L; ++n_layers;
            if (L <= 256) {
                remainder = v; rem_len = L;
                /* verifyRemainder (:223-252) */
                u128 rou = fpow(w, (u128)1 << (2 * depth));
                u128* rd = power_series(rou, L);
                size_t* pos = (size_t*)malloc(sizeof(size_t) * L); size_t np = 0;
                for (size_t i = 0; i < L; ++i) if (i % E) pos[np++] = i;
                if (maxdeg > np) { snprintf(g_err, sizeof g_err, "Low degree proof failed: remainder too short"); return -4; }
                u128* xs = vnew(maxdeg); u128* ys = vnew(maxdeg); u128* poly = vnew(maxdeg);
                for (size_t i = 0; i < maxdeg; ++i) { xs[i] = rd[pos[i]]; ys[i] = v[pos[i]]; }
                lagrange(xs, ys, maxdeg, poly);
                for (size_t i = maxdeg; i < np; ++i) if (eval_poly_at(poly, maxdeg, rd[pos[i]]) != v[pos[i]]) { snprintf(g_err, sizeof g_err, "Low degree proof failed: Remainder is not a valid degree %zu polynomial", maxdeg - 1); return -4; }
                free(rd); free(pos); free(xs); free(ys); free(poly);
                break;
            }
            /* interpolateQuarticBatch + evalQuarticBatch: Lagrange through the four points of every row */
            const u128 sx = prng_one(tree + 32, 32);
            u128* col = vnew(q);
            /* denominators prod_{m != j} (x_j - x_m) for every row, batch inverted (as galois does) */
            u128* den = vnew(4 * q); u128* deni = vnew(4 * q);
            const size_t step = (size_t)1 << (2 * depth);
            u128 ip[4]; ip[0] = 1; ip[1] = iota; ip[2] = fmul(iota, iota); ip[3] = fmul(ip[2], iota);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)q; ++i) {
                u128 xi = dom[(size_t)i * step], x[4];
                for (int j = 0; j < 4; ++j) x[j] = fmul(xi, ip[j]);
                for (int j = 0; j < 4; ++j) { u128 d = 1; for (int m = 0; m < 4; ++m) if (m != j) d = fmul(d, fsub(x[j], x[m])); den[4 * (size_t)i + j] = d; }
            }
            batch_inverse(den, deni, 4 * q);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)q; ++i) {
                u128 xi = dom[(size_t)i * step], x[4], acc = 0;
                for (int j = 0; j < 4; ++j) x[j] = fmul(xi, ip[j]);
                for (int j = 0; j < 4; ++j) {
                    u128 nm = 1; for (int m = 0; m < 4; ++m) if (m != j) nm = fmul(nm, fsub(sx, x[m]));
                    acc = fadd(acc, fmul(fmul(v[(size_t)i + (size_t)j * q], nm), deni[4 * (size_t)i + j]));
                }
                col[i] = acc;
            }
            free(den); free(deni);
            v = col; L = q; maxdeg /= 4; ++depth;
        }
    }
    STAGE();
    /* queries */
    uint32_t exe_pos[128], tmp[512], aug[512];
    int n_exe = (int)((size_t)exe_queries < N - N / E ? (size_t)exe_queries : N - N / E);
    const uint8_t* lc_root = layer_tree[0] + 32;
    if (pr_indexes(lc_root, n_exe, N, E, exe_pos) != 0) { snprintf(g_err, sizeof g_err, "Low degree proof failed: could not generate indexes"); return -4; }
    /* evProof */
    {
        int m = 0; for (int i = 0; i < n_exe; ++i) { tmp[m++] = exe_pos[i]; tmp[m++] = (uint32_t)((exe_pos[i] + E) % N); }
        int na = first_seen_unique(tmp, m, aug);
        uint8_t* vals = (uint8_t*)malloc((size_t)na * 16 * (size_t)n_ev);
        for (int k = 0; k < na; ++k) for (int q = 0; q < n_ev; ++q) put_elem(vals + ((size_t)k * n_ev + q) * 16, ev[q][aug[k]]);
        write_batch_proof(&out, e_tree, N, aug, na, vals, 16 * (size_t)n_ev, 16 * (size_t)n_ev);
        free(vals);
    }
    bput(&out, lc_root, 32);
    {
        size_t q = N >> 2; for (int i = 0; i < n_exe; ++i) tmp[i] = (uint32_t)(exe_pos[i] % q);
        int na = first_seen_unique(tmp, n_exe, aug);
        uint8_t* vals = (uint8_t*)malloc(64 * (size_t)na); rows4_values(layer_v[0], q, aug, na, vals);
        write_batch_proof(&out, layer_tree[0], q, aug, na, vals, 64, 64); free(vals);
    }
    bput8(&out, (unsigned)(n_layers - 1));
    for (int d = 0; d + 1 < n_layers; ++d) {
        size_t cl = layer_len[d + 1], cq = cl >> 2, pq = layer_len[d] >> 2;
        uint32_t pos[64];
        if (pr_indexes(layer_tree[d + 1] + 32, fri_queries, cl, E, pos) != 0) { snprintf(g_err, sizeof g_err, "Low degree proof failed: could not generate indexes"); return -4; }
        for (int i = 0; i < fri_queries; ++i) tmp[i] = (uint32_t)(pos[i] % cq);
        int na = first_seen_unique(tmp, fri_queries, aug);
        bput(&out, layer_tree[d + 1] + 32, 32);
        uint8_t* vals = (uint8_t*)malloc(64 * (size_t)(na > fri_queries ? na : fri_queries));
        rows4_values(layer_v[d + 1], cq, aug, na, vals);
        write_batch_proof(&out, layer_tree[d + 1], cq, aug, na, vals, 64, 64);
        rows4_values(layer_v[d], pq, pos, fri_queries, vals);
        write_batch_proof(&out, layer_tree[d], pq, pos, fri_queries, vals, 64, 64);
        free(vals);
    }
    bput8(&out, rem_len == 256 ? 0 : (unsigned)rem_len);
    for (size_t i = 0; i < rem_len; ++i) { uint8_t e[16]; put_elem(e, remainder[i]); bput(&out, e, 16); }
    if (shapes && shapes_len) bput(&out, shapes, shapes_len); else bput8(&out, 0);
    STAGE();
    *proof_out = out.p; *proof_len = out.n;
    /* (memory of the big vectors is released by the caller's process exit or oracle_free; this is a test tool) */
    for (int i = 0; i < n_layers; ++i) { free(layer_tree[i]); if (i > 0) free(layer_v[i]); }
    free(Lv); free(zinv); free(dom); free(e_tree); if (xdelta) free(xdelta); free(coeffs); free(lcc);
    for (int r = 0; r < R; ++r) { free(trace[r]); free(ppoly[r]); free(pe[r]); free(t_ev[r]); }
    for (int k = 0; k < A.n_static; ++k) free(s_ev[k]);
    for (int i = 0; i < n_in; ++i) { free(in_poly[i]); if (in_eval_n[i]) free(in_eval_n[i]); }
    free(trace); free(ppoly); free(pe); free(t_ev); free(s_ev); free(s_len); free(q_ev); free(ba); free(ev); free(as); free(in_poly); free(in_eval_n);
    return 0;
}

void oracle_free(void* p) { free(p); }

/* ORACLE (test infrastructure): the transforms of galois interpolateRoots / evalPolyAtRoots alone (lib/Stark.ts:106,109), for the
 * element-by-element K1 parity tests at sizes the Python restatement cannot reach.  in: t = 2^log_t little-endian 16-byte residues;
 * out: 2^log_n.  inverse != 0: interpolateRoots over the domain of size t (log_n == log_t); else evaluation of the zero-padded
 * polynomial over the domain of size 2^log_n generated by `root` (16 bytes LE, the field's root of unity of that order). */
int oracle_transform(const uint8_t* in, int log_t, int log_n, int inverse, const uint8_t* root_le, uint8_t* out) {
    size_t t = (size_t)1 << log_t, n = (size_t)1 << log_n;
    u128 root; memcpy(&root, root_le, 16);
    u128* v;
    if (inverse) { if (log_n != log_t) return -1; v = interpolate_roots((const u128*)in, log_t, root); }
    else v = eval_poly_at_roots((const u128*)in, t, log_n, root);
    memcpy(out, v, n * 16);
    free(v);
    return 0;
}
int oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
