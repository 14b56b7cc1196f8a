/* libgenstark_b200.so -- C ABI of the B200 STARK proving hot path.
 *
 * Drop-in boundary for GuildOfWeavers/genSTARK.  genSTARK reaches all numeric work
 * through two objects: the galois `FiniteField` it gets from air-assembly as `context.field`
 * (lib/Stark.ts:37-43) and the merkle `Hash` (lib/Stark.ts:49-53).  A binding (N-API addon, see
 * INTEGRATION.md) implements those two interfaces on top of the entry points below; `Vector` and
 * `Matrix` become device-resident handles (gs_mat), and `gs_stark_prove` is the fused whole-prover
 * crossing that replaces the body of Stark.prove (lib/Stark.ts:81-163).
 *
 * Conventions: every function returns 0 on success or a negative gs_status; the message is available
 * from gs_last_error(ctx).  Field elements cross the boundary as canonical residues, 16 bytes,
 * little-endian 32-bit limbs -- the bytes `Vector.toBuffer()` / `copyValue()` yield in the reference
 * (lib/utils/serialization.ts:131-147).  Handles are owned by the caller and freed explicitly
 * (the reference never frees: vectors live in the WASM arena, lib/Stark.ts:346-354).
 * One CUDA stream per context; calls block until results needed on the host are there.
 */
#ifndef GENSTARK_B200_H
#define GENSTARK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gs_ctx gs_ctx;     /* one device + stream + root tables              */
typedef struct gs_mat gs_mat;     /* rows x cols field elements, row-major, in HBM  */
typedef struct gs_digests gs_digests; /* n 32-byte digests in HBM (hash.mergeVectorRows / digestValues output) */
typedef struct gs_tree gs_tree;       /* MerkleTree: 2n digests, nodes[1] = root                            */
typedef struct gs_stark gs_stark; /* one AIR + security options + device buffers    */

enum gs_status {
    GS_OK = 0,
    GS_E_CUDA = -1,
    GS_E_ARG = -2,
    GS_E_UNSUPPORTED = -3,        /* e.g. a modulus other than 2^128 - 9*2^32 + 1: isOptimized=false */
    GS_E_STARK = -4,              /* protocol failure; shim rethrows as StarkError (lib/StarkError.ts) */
    GS_E_NOMEM = -5
};

/* ---- context -------------------------------------------------------------------------------- */
int gs_ctx_create(int device, gs_ctx** out);
void gs_ctx_destroy(gs_ctx* ctx);
const char* gs_last_error(gs_ctx* ctx);
int gs_ctx_sync(gs_ctx* ctx);
/* multi-GPU (one process per GPU): coset-sharded proving.  Rank 0 obtains an id (gs_comm_unique_id), every rank
 * receives it out of band and joins.  A Stark created on a context with world > 1 shards the E cosets of the
 * evaluation domain over the ranks; every rank must call gs_stark_prove with the same arguments and gets the same proof. */
int gs_comm_unique_id(uint8_t out128[128]);
int gs_ctx_comm_init(gs_ctx* ctx, int rank, int world, const uint8_t id128[128]);
/* the sharding map itself (host logic): to_local = 0: local index of `rank` -> global position; to_local = 1: global
 * position -> (owner rank, local index there).  E = 2^log2_e cosets dealt to the ranks in contiguous ranges. */
int gs_shard_map(int world, int rank, int log2_e, int64_t index, int to_local, int64_t* out_index, int* out_owner);
/* number of kernels launched through this context so far */
uint64_t gs_ctx_launch_count(gs_ctx* ctx);
/* createPrimeField(modulus): 0 when the modulus has the native fast path (isOptimized) */
int gs_field_supported(const uint8_t* modulus_le, size_t nbytes);
/* field.getRootOfUnity(2^log2_order) -> 16 bytes (host) */
int gs_field_root_of_unity(int log2_order, uint8_t out16[16]);
/* field.add/sub/mul/div/exp on scalars (host): op 0 add, 1 sub, 2 mul, 3 div (inv(0)=0), 4 exp */
int gs_field_scalar_op(int op, const uint8_t a16[16], const uint8_t b16[16], uint8_t out16[16]);

/* ---- Vector / Matrix handles (galois newVectorFrom / newMatrixFrom / toBuffer) ---------------- */
int gs_mat_alloc(gs_ctx* ctx, int64_t rows, int64_t cols, gs_mat** out);
int gs_mat_from_bytes(gs_ctx* ctx, const void* le_bytes, int64_t rows, int64_t cols, gs_mat** out);
int gs_mat_to_bytes(gs_ctx* ctx, const gs_mat* m, void* out_le_bytes);
int gs_mat_shape(const gs_mat* m, int64_t* rows, int64_t* cols);
void* gs_mat_device_ptr(gs_mat* m);
void gs_mat_free(gs_mat* m);

/* ---- polynomials over roots of unity (K1) ------------------------------------------------------
 * gs_interpolate_roots : field.interpolateRoots(domain, values)   lib/Stark.ts:106,
 *                        lib/components/CompositionPolynomial.ts:109.  values: rows x n, n = 2^k;
 *                        the domain is the power series of getRootOfUnity(n).
 * gs_eval_polys_at_roots: field.evalPolysAtRoots / evalPolyAtRoots  lib/Stark.ts:109,
 *                        CompositionPolynomial.ts:110.  polys: rows x t, zero-padded to the domain of
 *                        size 2^log2_domain >= t, natural-order output. */
int gs_interpolate_roots(gs_ctx* ctx, const gs_mat* values, gs_mat** polys);
int gs_eval_polys_at_roots(gs_ctx* ctx, const gs_mat* polys, int log2_domain, gs_mat** evals);

/* ---- element-wise vector operations (K2) --------------------------------------------------------
 * field.addVectorElements / subVectorElements / mulVectorElements (a, b) with b a vector of the same
 * shape or a scalar (pass b = NULL and scalar16): op 0 add, 1 sub, 2 mul.
 * Call sites: CompositionPolynomial.ts:98,120,136,145; LinearCombination.ts:50,63; ZeroPolynomial.ts:41-42 */
int gs_vec_binary(gs_ctx* ctx, int op, const gs_mat* a, const gs_mat* b, const uint8_t* scalar16, gs_mat** out);

/* field.divVectorElements(a, b) = a * inv(b), inv(0) = 0          CompositionPolynomial.ts:117, BoundaryConstraints.ts:92 */
int gs_vec_div(gs_ctx* ctx, const gs_mat* a, const gs_mat* b, gs_mat** out);
/* field.expVectorElements(a, e), 0 <= e < 2^128, and field.mulMatrixByVector(m, v): not called from lib/, used by the plain Poseidon
 * implementation of /root/reference/examples/poseidon/utils.ts:31-45 that the example compares its STARK with */
int gs_vec_exp(gs_ctx* ctx, const gs_mat* a, const uint8_t* exponent16, gs_mat** out);
int gs_mat_mul_vector(gs_ctx* ctx, const gs_mat* m, const gs_mat* v, gs_mat** out);
/* field.combineManyVectors(vectors, coefficients) -> vector        CompositionPolynomial.ts:105,142; LinearCombination.ts:60 */
int gs_vec_combine_many(gs_ctx* ctx, const gs_mat* const* vectors, int count, const uint8_t* coefficients16, gs_mat** out);
/* field.getPowerSeries(base, n)                                     CompositionPolynomial.ts:94,132; LowDegreeProver.ts:233 */
int gs_power_series(gs_ctx* ctx, const uint8_t base16[16], int64_t n, gs_mat** out);
/* field.pluckVector(v, skip, times): out[i] = v[(i*skip) mod len]   ZeroPolynomial.ts:40 */
int gs_pluck_vector(gs_ctx* ctx, const gs_mat* v, int64_t skip, int64_t times, gs_mat** out);
/* field.transposeVector(v, columns, step) -> rows x columns matrix  LowDegreeProver.ts:42,162,190,198 */
int gs_transpose_vector(gs_ctx* ctx, const gs_mat* v, int columns, int64_t step, gs_mat** out);
/* one FRI layer (K4): evalQuarticBatch(interpolateQuarticBatch(xs, rows), x*) with xs = transposeVector(domain, 4, 4^depth)
 * and rows = transposeVector(v, 4); v has length domain / 4^depth.   LowDegreeProver.ts:190-195 */
int gs_fri_fold(gs_ctx* ctx, const gs_mat* v, int log2_domain, int depth, const uint8_t special_x16[16], gs_mat** column);

/* ---- the rest of the FiniteField seam: small-vector, batch-cubic and reshaping methods ---------------------------
 * field.prng(seed) (count = 0 -> one element) / field.prng(seed, count)   CompositionPolynomial.ts:58; LinearCombination.ts:58,82;
 *                                                                         LowDegreeProver.ts:132,194   (host; seed = the bytes hashed) */
int gs_field_prng(const uint8_t* seed, size_t seed_len, int count, uint8_t* out16);
/* field.interpolate(xs, ys) -> n coefficients, low -> high (Lagrange)       BoundaryConstraints.ts:42; LowDegreeProver.ts:243   (host) */
int gs_poly_interpolate(const uint8_t* xs16, const uint8_t* ys16, int n, uint8_t* out16);
/* field.evalPolyAt(poly, x)                                               BoundaryConstraints.ts:59-60; LowDegreeProver.ts:248 (host) */
int gs_poly_eval_at(const uint8_t* poly16, int n, const uint8_t x16[16], uint8_t out16[16]);
/* field.mulPolys(a, b) -> na + nb - 1 coefficients                          BoundaryConstraints.ts:30                           (host) */
int gs_poly_mul(const uint8_t* a16, int na, const uint8_t* b16, int nb, uint8_t* out16);
/* field.combineVectors(a, b) = sum a[i] * b[i]                              CompositionPolynomial.ts:168,188; LinearCombination.ts:85 */
int gs_vec_combine(gs_ctx* ctx, const gs_mat* a, const gs_mat* b, uint8_t out16[16]);
/* field.interpolateQuarticBatch(xSets, ySets): rows x 4 each -> rows x 4     LowDegreeProver.ts:137,191 */
int gs_quartic_interpolate_batch(gs_ctx* ctx, const gs_mat* xs, const gs_mat* ys, gs_mat** polys);
/* field.evalQuarticBatch(polys, x): x one element per row (xs) or one scalar (xs = NULL, x16)   LowDegreeProver.ts:140,195 */
int gs_quartic_eval_batch(gs_ctx* ctx, const gs_mat* polys, const gs_mat* xs, const uint8_t* x16, gs_mat** out);
/* field.newMatrixFromVectors(vs): parts stacked as rows                     BoundaryConstraints.ts:84-85 */
int gs_mat_stack(gs_ctx* ctx, const gs_mat* const* parts, int count, gs_mat** out);
/* field.matrixRowsToVectors(m): a copy of rows [row0, row0 + nrows)          Stark.ts:114; BoundaryConstraints.ts:73; LinearCombination.ts:39 */
int gs_mat_rows(gs_ctx* ctx, const gs_mat* m, int64_t row0, int64_t nrows, gs_mat** out);
/* field.transposeMatrix(m)                                                LowDegreeProver.ts:181 */
int gs_mat_transpose(gs_ctx* ctx, const gs_mat* m, gs_mat** out);
/* field.joinMatrixRows(m): the same elements as one row (or any rows x cols with the same product), no copy   LowDegreeProver.ts:182 */
int gs_mat_reshape(gs_mat* m, int64_t rows, int64_t cols);
/* Vector.getValue(i) / Matrix.getValue(row, col)                           Stark.ts:290,357-358; LowDegreeProver.ts:141-142 */
int gs_mat_get(gs_ctx* ctx, const gs_mat* m, int64_t row, int64_t col, uint8_t out16[16]);

/* ---- Hash / MerkleTree (K5); alg 0 = sha256, 1 = blake2s256 --------------------------------------------
 * hash.mergeVectorRows(vectors): digest i = H(v0[i] || v1[i] || ...) over every row of every matrix   lib/Stark.ts:115
 * hash.digestValues(buffer, valueSize): one digest per row of a row-major matrix               LowDegreeProver.ts:45
 * MerkleTree.create / .root / .proveBatch                                   lib/Stark.ts:118,150; LowDegreeProver.ts:46,52
 * proveBatch blob: u32 n_values, u32 n_columns, u32 depth, n_values x 32 B (leaf digests in input order), then per
 * column u32 length + length x 32 B -- the {values, nodes, depth} of BatchMerkleProof (serialization.ts:31-35) */
int gs_hash_merge_vector_rows(gs_ctx* ctx, int alg, const gs_mat* const* mats, int count, gs_digests** out);
int gs_hash_digest_values(gs_ctx* ctx, int alg, const gs_mat* rows, gs_digests** out);
int gs_digests_to_bytes(gs_ctx* ctx, const gs_digests* d, void* out);
int64_t gs_digests_count(const gs_digests* d);
void gs_digests_free(gs_digests* d);
int gs_merkle_create(gs_ctx* ctx, int alg, const gs_digests* leaves, gs_tree** out);
int gs_merkle_root(gs_ctx* ctx, const gs_tree* tree, uint8_t out32[32]);
int gs_merkle_prove_batch(gs_ctx* ctx, const gs_tree* tree, const uint32_t* indexes, int count, uint8_t* out,
                          size_t out_cap, size_t* out_len);
void gs_tree_free(gs_tree* tree);
/* hash.digest(buffer)                                                       lib/utils/index.ts:37   (host) */
int gs_hash_digest(int alg, const uint8_t* msg, size_t len, uint8_t out32[32]);
/* MerkleTree.verifyBatch(root, indexes, proof, hash): 1 valid, 0 invalid, < 0 malformed; proof = the blob of
 * gs_merkle_prove_batch                                                     Stark.ts:206; LowDegreeProver.ts:86,109,116   (host) */
int gs_merkle_verify_batch(int alg, const uint8_t root32[32], const uint32_t* indexes, int count, const uint8_t* proof, size_t proof_len);

/* ---- fused prover: the body of Stark.prove in one crossing (lib/Stark.ts:81-163) -------------------
 * gs_stark_create  <->  new Stark(schema, component, options)            lib/Stark.ts:35-58
 *   air_blob: flattened AirModule (genstark_b200/air.py: pack_air); hash_alg 0 = sha256, 1 = blake2s256
 *   (HASH_ALGORITHMS, lib/Stark.ts:19); query counts validated as buildSecurityOptions does (:318-344).
 * gs_stark_prove   <->  stark.prove(assertions, inputs, seed)            lib/Stark.ts:81-163
 *   assertions: n x { u32 register, u32 step, 16-byte value }; init_state16: the first trace row
 *   (R x 16 bytes, from the AIR's init with inputs/seed); input_traces: one T-length column per input
 *   register (register order), or NULL; shapes_blob: serialized iShapes (u8 count, then u8 rank +
 *   rank x u32le each) appended to the proof.  *proof_out is the serialized proof
 *   (lib/Serializer.ts:35-79), owned by the handle and valid until the next call.
 *   GS_E_STARK carries the reference's StarkError texts (lib/Stark.ts:101,143; CompositionPolynomial.ts:79). */
int gs_stark_create(gs_ctx* ctx, const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries,
                    int fri_queries, gs_stark** out);
void gs_stark_destroy(gs_stark* s);
/* Stark.generateExecutionTrace (lib/Stark.ts:252-257): host only; out_trace: R x T x 16 bytes, row = register */
int gs_air_generate_trace(const uint8_t* air_blob, size_t blob_len, const uint8_t* init_state16,
                          const uint8_t* input_traces, uint8_t* out_trace);
/* which generator the calling thread's last trace generation used: "jit <hash>" (the transition function compiled
 * to native code at first use, as air-assembly compiles it to JavaScript at instantiate()) or
 * "interpreter (<reason>)".  GS_TRACE_JIT=0 forces the interpreter; results are identical. */
const char* gs_trace_backend(void);
int gs_stark_prove(gs_stark* s, const uint8_t* assertions, int n_assertions, const uint8_t* init_state16,
                   const uint8_t* input_traces, const uint8_t* shapes_blob, size_t shapes_len,
                   const uint8_t** proof_out, size_t* proof_len);
/* flags bit 0: reuse the execution trace already resident in HBM from the previous call (skips host trace
 * generation and the host->device copy: the "inputs resident" measurement leg of bench.py) */
int gs_stark_prove_ex(gs_stark* s, const uint8_t* assertions, int n_assertions, const uint8_t* init_state16,
                      const uint8_t* input_traces, const uint8_t* shapes_blob, size_t shapes_len, int flags,
                      const uint8_t** proof_out, size_t* proof_len);
/* CUDA-event time of the device part and host wall clock of the last prove */
int gs_stark_last_timing(gs_stark* s, float* device_ms, double* host_ms);
/* stark.verify(assertions, proof, publicInputs)  lib/Stark.ts:167-248 -- host only (O(queries log N), as in the
 * reference).  proof: the serialized proof; public_traces: one T-length column per PUBLIC input register or NULL.
 * Returns GS_OK, or GS_E_STARK with the reference's StarkError text in err_buf. */
int gs_stark_verify(const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                    const uint8_t* assertions, int n_assertions, const uint8_t* proof, size_t proof_len,
                    const uint8_t* public_traces, char* err_buf, size_t err_cap);
/* per-stage host milliseconds of the last prove as JSON [[name, ms], ...] (Logger, lib/utils/Logger.ts) */
/* which constraint-evaluation kernel this instance launches: "nvrtc <hash>" (the AIR's evaluation function compiled
 * to sm_100a code at creation, as air-assembly compiles it to JavaScript) or "interpreter (<reason>)".
 * GS_COMPOSE_JIT=0 forces the interpreting kernel; results are identical. */
/* message of the last failure on the context this instance lives on (what a binding throws for a negative status) */
const char* gs_stark_last_error(gs_stark* s);
/* Prime fields of at most 64 bits (BASELINE config 1: Foo over 2^32 - 3*2^25 + 1): Stark.prove / Stark.verify on the HOST -- the
 * counterpart of the reference's own fallback to unoptimised arithmetic for fields without a WASM backend (lib/Stark.ts:41-43).
 * Same arguments as gs_stark_prove / gs_stark_verify (16-byte little-endian elements at the boundary); the 128-bit field is
 * refused here: it has the GPU path and no CPU one.  The proof buffer is owned by the library until the thread's next call. */
int gs_host_stark_prove(const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                        const uint8_t* assertions, int n_assertions, const uint8_t* init_state16, const uint8_t* input_traces,
                        const uint8_t* shapes_blob, size_t shapes_len, const uint8_t** proof_out, size_t* proof_len,
                        char* err_buf, size_t err_cap);
int gs_host_stark_verify(const uint8_t* air_blob, size_t blob_len, int hash_alg, int exe_queries, int fri_queries,
                         const uint8_t* assertions, int n_assertions, const uint8_t* proof, size_t proof_len,
                         const uint8_t* public_traces, char* err_buf, size_t err_cap);
const char* gs_stark_compose_backend(gs_stark* s);
const char* gs_stark_stage_times(gs_stark* s);
/* test hooks: keep C(x) and read device-resident intermediates back (0 P evals, 1 C, 2 L, 3 P polys) */
int gs_stark_set_debug(gs_stark* s, int keep_intermediates);
int gs_stark_read_intermediate(gs_stark* s, int which, void* out, size_t out_bytes);

/* ---- measurement helpers ------------------------------------------------------------------------ */
/* CUDA events on the context stream around every kernel class; report is JSON {class: {groups, ms}} */
int gs_ctx_profile(gs_ctx* ctx, int on);
const char* gs_ctx_profile_report(gs_ctx* ctx);
int gs_timer_begin(gs_ctx* ctx);
int gs_timer_end(gs_ctx* ctx, float* ms);
/* transform src (rows x t) into dst (rows x n >= t) with caller-provided work (rows x n): n == t -> NTT or
 * (inverse) iNTT; n > t -> LDE.  No allocation: used to time K1 alone. */
int gs_ntt_into(gs_ctx* ctx, const gs_mat* src, gs_mat* dst, gs_mat* work, int inverse);
/* a rank's share of a coset-sharded LDE: src (rows x t) -> dst (rows x t*cosets), the `cosets` cosets starting at
 * coset_base of the evaluation domain of size t * 2^log_e_total, local position q*cosets + (j - coset_base) (section 8e) */
int gs_lde_cosets_into(gs_ctx* ctx, const gs_mat* src, gs_mat* dst, gs_mat* work, int coset_base, int log_e_total);
/* fills m with uniform canonical residues from a counter-based generator (SplitMix64 of seed + 2*index, two draws
 * -> 128 bits, reduced mod p): the synthetic NTT inputs of SURVEY.md section 8d, generated on the device */
int gs_mat_fill_random(gs_ctx* ctx, gs_mat* m, uint64_t seed);
/* runs blocks x 256 threads x (4*iters) dependent modular multiplications; returns kernel ms */
int gs_debug_modmul_probe(gs_ctx* ctx, int blocks, int iters, float* ms_out);
/* squaring chains (dedicated 10-product squaring) at the same shape; *mismatches counts disagreements with a * a */
int gs_debug_sqr_probe(gs_ctx* ctx, int blocks, int iters, float* ms_out, int* mismatches);
/* same for the NTT's instruction mix: blocks x 256 threads x 2 butterflies (u + v, (u - v) * w) x iters */
int gs_debug_butterfly_probe(gs_ctx* ctx, int blocks, int iters, float* ms_out);
/* test hooks for the commit kernels: mergeVectorRows + MerkleTree.create (Stark.ts:114-118) in the launches the prover
 * uses for a commit (leaf hashing fused into the tree kernels), and the 2n stored digests of a tree (slot 0 unused) */
int gs_debug_commit_columns(gs_ctx* ctx, int alg, const gs_mat* const* mats, int count, gs_tree** out);
int gs_debug_tree_nodes(gs_ctx* ctx, const gs_tree* tree, void* out, size_t out_bytes);

#ifdef __cplusplus
}
#endif
#endif
