// Fiat-Shamir on the device: the composition / linear-combination coefficients field.prng(evRoot, n)
// (/root/reference/lib/components/CompositionPolynomial.ts:58-60, LinearCombination.ts:58-59, lib/Stark.ts:129) are drawn from the
// evaluation root where it already is -- in HBM, right behind the Merkle build -- and folded with the E-periodic factors into the
// tables compose.cuh reads (cd_tab, pf_tab, lk_tab).  The host used to do this between two captured graphs (root -> host ->
// 70 SHA-256 -> fold -> upload); now the whole prove is enqueued without a round trip and the host only reads the root at the end,
// for the proof bytes.  Same construction as hostcrypto.h: prng_many (SHA-256 of the big-endian state + i, with galois' odd-length
// hex quirk), which tests/ pin through the proof bytes.
#pragma once
#include "hash.cuh"

namespace gs {

struct CoeffParams {
    const uint32_t* root;            // 32 bytes in HBM
    int K, nB, n_lc, n_pf, E;
    int d_count, b_count;            // layout of the draw: [0, K) d_k, [K, d_count) d'_k, then b_b, (b'_b), then kappa_j, (kappa'_j)
    int has_delta, comp_gt_t;        // delta = compDeg - T > 0; compDeg > T (second boundary coefficient present)
    const int* pow_idx; const int* adj_idx;                  // per constraint: power slot (-1: none), index of d'_k in the draw (-1: none)
    const fp* inv_num; const fp* pow_tab; const fp* delta_tab;      // E, slots x E, E
    const fp* pf_coef; const int* pf_owner;                  // per partial-fraction term
    fp* cd_tab; fp* pf_tab; fp* lk_tab;                      // K x E, n_pf x E, n_lc x E
    int count;                       // d_count + b_count + lc_total
};

// SHA-256 of a byte string of at most 55 bytes (one block); digest as 32 bytes, big-endian words in order
GS_D void sha256_short(const uint8_t* msg, int len, uint8_t (&out)[32]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = 0;
    for (int i = 0; i < len; ++i) w[i >> 2] |= (uint32_t)msg[i] << (24 - 8 * (i & 3));
    w[len >> 2] |= 0x80u << (24 - 8 * (len & 3));
    w[15] = (uint32_t)len * 8u;
    uint32_t h[8];
    sha256_init(h);
    sha256_compress(h, w);
    for (int i = 0; i < 8; ++i) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}

// 32-byte big-endian digest mod p = lo + hi * 2^128 = lo + hi * (9*2^32 - 1)
GS_D fp digest_mod_p_dev(const uint8_t (&d)[32]) {
    fp hi, lo;
    for (int k = 0; k < 4; ++k) {
        hi.v[3 - k] = ((uint32_t)d[4 * k] << 24) | ((uint32_t)d[4 * k + 1] << 16) | ((uint32_t)d[4 * k + 2] << 8) | d[4 * k + 3];
        lo.v[3 - k] = ((uint32_t)d[16 + 4 * k] << 24) | ((uint32_t)d[16 + 4 * k + 1] << 16) | ((uint32_t)d[16 + 4 * k + 2] << 8) | d[16 + 4 * k + 3];
    }
    const fp zero = fp_zero();
    hi = fp_add(hi, zero); lo = fp_add(lo, zero);                 // canonical residues of the halves
    fp c9; c9.v[0] = 0xFFFFFFFFu; c9.v[1] = 8u; c9.v[2] = 0; c9.v[3] = 0;
    return fp_add(lo, fp_mul(hi, c9));
}

// element i of field.prng(seed, n): SHA-256 of Buffer.from((state + i).toString(16), 'hex'), state = SHA-256(seed) as a
// big-endian integer; an odd number of hex digits loses the LAST nibble (hostcrypto.h: be_to_node_buffer)
GS_D fp prng_element_dev(const uint8_t (&state)[32], unsigned i) {
    uint8_t b[33];
    b[0] = 0;
    for (int k = 0; k < 32; ++k) b[k + 1] = state[k];
    unsigned x = i;
    for (int k = 32; k >= 0 && x; --k) { const unsigned s = b[k] + (x & 0xFFu); b[k] = (uint8_t)s; x = (x >> 8) + (s >> 8); }
    // minimal hex digits
    int first = 0;                                                // index of the first non-zero nibble among 66
    while (first < 66 && (((first & 1) ? (b[first >> 1] & 15) : (b[first >> 1] >> 4)) == 0)) ++first;
    int nd = 66 - first;
    if (nd == 0) { nd = 1; first = 65; }
    uint8_t buf[33];
    const int nbytes = nd / 2;
    for (int k = 0; k < nbytes; ++k) {
        const int n0 = first + 2 * k, n1 = n0 + 1;
        const unsigned hi = (n0 & 1) ? (b[n0 >> 1] & 15u) : (b[n0 >> 1] >> 4);
        const unsigned lo = (n1 & 1) ? (b[n1 >> 1] & 15u) : (b[n1 >> 1] >> 4);
        buf[k] = (uint8_t)((hi << 4) | lo);
    }
    uint8_t d[32];
    sha256_short(buf, nbytes, d);
    return digest_mod_p_dev(d);
}

// one block; thread i draws coefficient i, then the threads fold the tables
__global__ void __launch_bounds__(256) derive_coeffs_kernel(const CoeffParams P) {
    extern __shared__ __align__(16) unsigned char coeff_smem[];
    fp* coef = reinterpret_cast<fp*>(coeff_smem);                 // P.count entries
    __shared__ uint8_t state[32];
    if (threadIdx.x == 0) {
        uint8_t seed[32];
        for (int k = 0; k < 8; ++k) { const uint32_t w = P.root[k]; seed[4 * k] = (uint8_t)w; seed[4 * k + 1] = (uint8_t)(w >> 8); seed[4 * k + 2] = (uint8_t)(w >> 16); seed[4 * k + 3] = (uint8_t)(w >> 24); }
        uint8_t d[32];
        sha256_short(seed, 32, d);
        for (int k = 0; k < 32; ++k) state[k] = d[k];
    }
    __syncthreads();
    uint8_t st[32];
    for (int k = 0; k < 32; ++k) st[k] = state[k];
    for (int i = threadIdx.x; i < P.count; i += blockDim.x) coef[i] = prng_element_dev(st, (unsigned)i);
    __syncthreads();
    const int E = P.E;
    for (int idx = threadIdx.x; idx < P.K * E; idx += blockDim.x) {
        const int k = idx / E, j = idx - k * E;
        fp v = coef[k];
        if (P.pow_idx[k] >= 0) v = fp_add(v, fp_mul(coef[P.adj_idx[k]], P.pow_tab[P.pow_idx[k] * E + j]));
        P.cd_tab[idx] = fp_mul(v, P.inv_num[j]);
    }
    for (int idx = threadIdx.x; idx < P.n_pf * E; idx += blockDim.x) {
        const int a = idx / E, j = idx - a * E, b = P.pf_owner[a];
        fp v = coef[P.d_count + b];
        if (P.has_delta) v = fp_add(v, fp_mul(P.comp_gt_t ? coef[P.d_count + P.nB + b] : fp_zero(), P.delta_tab[j]));
        P.pf_tab[idx] = fp_mul(v, P.pf_coef[a]);
    }
    for (int idx = threadIdx.x; idx < P.n_lc * E; idx += blockDim.x) {
        const int q = idx / E, j = idx - q * E;
        fp v = coef[P.d_count + P.b_count + q];
        if (P.has_delta) v = fp_add(v, fp_mul(coef[P.d_count + P.b_count + P.n_lc + q], P.delta_tab[j]));
        P.lk_tab[idx] = v;
    }
}

}  // namespace gs
