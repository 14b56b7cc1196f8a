// The fused prover: body of Stark.prove (/root/reference/lib/Stark.ts:81-163) as one host routine that
// enqueues K1..K6 on the context stream and keeps every O(N) object in HBM.  The host only sees
// 32-byte Merkle roots (for Fiat-Shamir), the <= 256-value FRI remainder and the queried rows.
#pragma once
#include <array>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include "core.cuh"
#include "ntt_host.cuh"
#include "pointwise.cuh"
#include "hash.cuh"
#include "batchinv.cuh"
#include "fri.cuh"
#include "compose.cuh"
#include "coeffs.cuh"
#include "devjit.cuh"
#include "hostcrypto.h"
#include "hostfield.h"
#include "hostair.h"

namespace gs {

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(Ctx* c, size_t bytes) {
        if (bytes <= cap) return GS_OK;
        if (p) { cudaStreamSynchronize(c->stream); cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return c->cuda_fail(e, "cudaMalloc(prover buffer)");
        cap = bytes;
        return GS_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// A captured launch sequence.  Buffers and kernel arguments of a Stark instance are fixed between proves (values that
// change -- coefficients, assertion data -- live in device memory at fixed addresses), so the ~80 small launches of a
// prove are captured once per instance and replayed with one cudaGraphLaunch.
struct GraphSlot { cudaGraphExec_t exec = nullptr; unsigned long long key = 0, launches = 0; };

struct StageTimes {                 // milliseconds, host wall clock around each stage (sync'ed)
    std::vector<std::pair<std::string, double>> items;
};

struct Stark : public AirHost {
    Ctx* ctx = nullptr;
    // options
    int hash_alg = HASH_SHA256, exe_queries = 80, fri_queries = 40;
    // device-resident program + cyclic tables
    DevBuf d_instrs, d_consts, d_cyc, d_u;   // d_u: u[j] = 1/(w_N^j - 1), built once per instance
    std::vector<size_t> cyc_off;      // per static register (cycle kind): element offset into d_cyc
    std::vector<unsigned> cyc_mask;
    // per-prove buffers (grow only)
    DevBuf d_trace, d_poly, d_pe, d_in_trace, d_in_poly, d_in_e, d_work, d_tree, d_zb, d_zbs, d_l, d_c, d_fri, d_fri_trees,
           d_params, d_small, d_idx, d_gather;
    void* h_trace = nullptr; size_t h_trace_bytes = 0;     // pinned
    StageTimes last_times;
    bool keep_intermediates = false;  // stage-level parity tests read P/C/L back
    bool trace_resident = false;      // d_trace / d_in_trace hold the last proved trace
    bool use_graphs = true;           // CUDA graphs for the launch-bound chains (off while profiling / stage timing)
    unsigned long long proves_done = 0;
    GraphSlot g_commit, g_fri;
    std::shared_ptr<ComposeJit> compose_jit;   // K2 specialised for this AIR's constraints (devjit.cuh); fn == null => interpreter
    Shard shard;                      // coset sharding over the ranks of the context (world == 1: everything local)
    DevBuf d_dig_loc, d_dig_all;      // commit boundary: local digests / all-gathered digests before the permutation
    DevBuf d_fri_rep;                 // sharded prover: the gathered FRI layer and the replicated layers behind it
    // peer memory (cudaIpc): every rank's d_tree / d_fri_trees as seen from this rank, re-exchanged when an allocation moves
    struct PeerMap { void* key = nullptr; std::vector<void*> ptr; bool ok = false; } peer_tree, peer_fri, peer_sync, peer_stage;
    DevBuf d_stage;                   // staging buffer the peers store their digest blocks into (never reallocated inside a prove)
    DevBuf d_ipc, d_sync;             // d_sync: [0, 8) barrier flags written by the peers, [8] local barrier counter, [9] timeout flag
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    DevBuf d_epoch;                         // prove counter, copied behind each FRI root so the host can poll for it
    float last_device_ms = 0;         // CUDA-event time from the first enqueue to the last kernel of prove()
    double last_host_ms = 0;          // wall clock of the whole call
    ~Stark() {
        for (DevBuf* b : {&d_instrs, &d_consts, &d_cyc, &d_u, &d_trace, &d_poly, &d_pe, &d_in_trace, &d_in_poly, &d_in_e, &d_work, &d_tree,
                          &d_zb, &d_zbs, &d_l, &d_c, &d_fri, &d_fri_trees, &d_params, &d_small, &d_idx, &d_gather}) b->release();
        d_sync.release(); d_stage.release();
        if (h_trace) cudaFreeHost(h_trace);
        if (g_commit.exec) cudaGraphExecDestroy(g_commit.exec);
        if (g_fri.exec) cudaGraphExecDestroy(g_fri.exec);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev2) cudaEventDestroy(ev2);
        d_epoch.release(); d_dig_loc.release(); d_dig_all.release(); d_fri_rep.release(); d_ipc.release();
        for (PeerMap* m : {&peer_tree, &peer_fri, &peer_sync, &peer_stage}) for (size_t r = 0; r < m->ptr.size(); ++r) if (m->ptr[r] && (int)r != (ctx ? ctx->rank : 0)) cudaIpcCloseMemHandle(m->ptr[r]);
    }
};

// Montgomery's trick with the reference's inv(0) = 0 (zeros are skipped)
static inline std::vector<u128> h_batch_inverse(const std::vector<u128>& v) {
    const size_t n = v.size();
    std::vector<u128> pre(n), out(n, 0);
    u128 acc = 1;
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (v[i] != 0) acc = h_mul(acc, v[i]); }
    u128 inv = h_inv(acc);
    for (size_t i = n; i > 0; --i) {
        if (v[i - 1] == 0) continue;
        out[i - 1] = h_mul(inv, pre[i - 1]);
        inv = h_mul(inv, v[i - 1]);
    }
    return out;
}

// Lagrange interpolation on the host (BoundaryConstraints.ts:42, LowDegreeProver.ts:243), low -> high
static inline std::vector<u128> h_interpolate(const std::vector<u128>& xs, const std::vector<u128>& ys) {
    const size_t n = xs.size();
    std::vector<u128> root(n + 1, 0); root[0] = 1;
    for (size_t i = 0; i < n; ++i) {            // root *= (x - xs[i])
        for (size_t k = i + 1; k > 0; --k) root[k] = h_sub(root[k - 1], h_mul(root[k], xs[i]));
        root[0] = h_sub(0, h_mul(root[0], xs[i]));
    }
    // denominators first, inverted together (one field inversion for all n points; inv(0) = 0 kept per element)
    std::vector<u128> out(n, 0), num(n), den(n, 0);
    for (size_t i = 0; i < n; ++i) {
        u128 acc = 0;
        for (size_t k = n; k > 0; --k) { acc = h_add(root[k], h_mul(acc, xs[i])); num[k - 1] = acc; }
        for (size_t k = n; k > 0; --k) den[i] = h_add(h_mul(den[i], xs[i]), num[k - 1]);
    }
    const std::vector<u128> den_inv = h_batch_inverse(den);
    for (size_t i = 0; i < n; ++i) {
        u128 acc = 0;
        for (size_t k = n; k > 0; --k) { acc = h_add(root[k], h_mul(acc, xs[i])); num[k - 1] = acc; }
        const u128 f = h_mul(ys[i], den_inv[i]);
        for (size_t k = 0; k < n; ++k) out[k] = h_add(out[k], h_mul(num[k], f));
    }
    return out;
}
static inline u128 h_eval_poly(const std::vector<u128>& poly, u128 x) {
    u128 acc = 0;
    for (size_t k = poly.size(); k > 0; --k) acc = h_add(h_mul(acc, x), poly[k - 1]);
    return acc;
}

struct Assertion { uint32_t reg, step; u128 value; };

// Merkle commit boundary of the sharded prover: every rank hashed the rows it owns; all-gather the digests over
// NVLink (NCCL) and put them in leaf order.  world == 1: the caller hashed straight into the tree.
static inline int commit_gather(Stark* S, const uint32_t* d_local, long long n_loc, uint32_t* leaves_out) {
    Ctx* c = S->ctx;
    const Shard& sh = S->shard;
    int log_w = 0; while ((1 << log_w) < sh.world) ++log_w;
    int rc;
    if ((rc = S->d_dig_all.ensure(c, (size_t)(n_loc << log_w) * 32))) return rc;
    { ProfScope ps(c, "nccl_allgather_digests");
      const int nr = nccl().AllGather(d_local, S->d_dig_all.p, (size_t)n_loc * 32, GS_NCCL_UINT8, c->comm, c->stream);
      if (nr != 0) return c->fail(GS_E_CUDA, "ncclAllGather: %s", nccl().GetErrorString(nr)); }
    const long long total = (n_loc << log_w) * 2;
    permute_digests_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(S->d_dig_all.as<uint4>(), reinterpret_cast<uint4*>(leaves_out), n_loc, sh.log_el, log_w);
    c->launches++;
    return GS_OK;
}

static inline bool shard_peer_enabled() {      // GS_SHARD_PEER=0: digests travel through ncclSend/ncclRecv as in round 1
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_SHARD_PEER"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v != 0;
}
// (Re)build the table of peer pointers for one allocation: cudaIpcGetMemHandle here, handles all-gathered with NCCL, opened on
// every rank.  Collective: every rank calls it at the same point of the same prove.  Never called under stream capture (the
// first prove of an instance runs uncaptured and allocations only move when they grow).  ok == false => NCCL path.
static inline int peer_map_update(Stark* S, Stark::PeerMap& m, void* base) {
    Ctx* c = S->ctx;
    if (m.key == base && !m.ptr.empty()) return GS_OK;
    const int W = c->world;
    for (size_t r = 0; r < m.ptr.size(); ++r) if (m.ptr[r] && (int)r != c->rank) cudaIpcCloseMemHandle(m.ptr[r]);
    m.ptr.assign(W, nullptr); m.ok = false; m.key = base;
    int rc;
    if ((rc = S->d_ipc.ensure(c, (size_t)(W + 1) * 128))) return rc;
    struct Slot { cudaIpcMemHandle_t h; int valid; char pad[128 - sizeof(cudaIpcMemHandle_t) - sizeof(int)]; };
    static_assert(sizeof(Slot) == 128, "slot size");
    Slot mine; memset(&mine, 0, sizeof mine);
    mine.valid = (cudaIpcGetMemHandle(&mine.h, base) == cudaSuccess) ? 1 : 0;
    if (!mine.valid) cudaGetLastError();
    uint8_t* d = S->d_ipc.as<uint8_t>();
    GS_CUDA(c, cudaMemcpyAsync(d, &mine, 128, cudaMemcpyHostToDevice, c->stream));
    const int nr = nccl().AllGather(d, d + 128, 128, GS_NCCL_UINT8, c->comm, c->stream);
    if (nr != 0) return c->fail(GS_E_CUDA, "ncclAllGather(ipc handles): %s", nccl().GetErrorString(nr));
    std::vector<Slot> all(W);
    GS_CUDA(c, cudaMemcpyAsync(all.data(), d + 128, (size_t)W * 128, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    bool ok = true;
    for (int r = 0; r < W; ++r) ok = ok && all[r].valid;
    for (int r = 0; r < W && ok; ++r) {
        if (r == c->rank) { m.ptr[r] = base; continue; }
        if (cudaIpcOpenMemHandle(&m.ptr[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); m.ptr[r] = nullptr; ok = false; }
    }
    // every rank must take the same path: agree on the outcome (sum of the failures)
    unsigned bad = ok ? 0u : 1u;
    GS_CUDA(c, cudaMemcpyAsync(d, &bad, 4, cudaMemcpyHostToDevice, c->stream));
    const int nr2 = nccl().AllReduce(d, d, 1, GS_NCCL_UINT32, GS_NCCL_SUM, c->comm, c->stream);
    if (nr2 != 0) return c->fail(GS_E_CUDA, "ncclAllReduce: %s", nccl().GetErrorString(nr2));
    GS_CUDA(c, cudaMemcpyAsync(&bad, d, 4, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    m.ok = (bad == 0);
    return GS_OK;
}
// Cross-rank barrier on the stream, over peer memory: every rank bumps its local epoch, stores it into its slot of every
// peer's flag array (system scope) and spins until all W slots of its own array have reached the epoch -- a few microseconds
// on NVSwitch, where a one-word NCCL all-reduce costs ~30.  Stream order puts it behind this rank's hashing kernel, whose peer
// stores are complete at kernel end, so past the barrier every rank's leaf range is whole.  Optionally carries the sub-tree
// roots: this rank's root (node W + rank) is stored into every peer's tree first, which replaces the 32-byte all-gather.
// A peer that never arrives (a rank died) ends the spin after ~2^27 polls and raises the timeout word; the proof then fails
// the parity / verification checks instead of hanging the box.
struct PeerSync { uint32_t* flags[8]; uint32_t* local; int rank, world; };
__global__ void peer_barrier_kernel(const PeerSync ps, const PeerTrees trees, int with_roots) {
    __shared__ uint32_t epoch;
    const int t = threadIdx.x;
    if (t == 0) epoch = ++ps.local[8];
    __syncthreads();
    if (with_roots && t < ps.world * 8) {
        const int peer = t >> 3, w = t & 7;
        const uint32_t v = trees.base[ps.rank][8 * (ps.world + ps.rank) + w];
        if (peer != ps.rank) trees.base[peer][8 * (ps.world + ps.rank) + w] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < ps.world) {
        volatile uint32_t* theirs = ps.flags[t] + ps.rank;
        *theirs = epoch;
        __threadfence_system();
        volatile uint32_t* mine = ps.local + t;
        unsigned polls = 0;
        while ((int)(*mine - epoch) < 0) { if (++polls > (1u << 27)) { ps.local[9] = 1u; break; } }
    }
    __threadfence_system();
}
static inline int shard_barrier(Stark* S, const PeerTrees* trees) {
    Ctx* c = S->ctx;
    if (S->peer_sync.ok) {
        ProfScope ps(c, "peer_barrier");
        PeerSync sy; memset(&sy, 0, sizeof sy);
        for (int r = 0; r < c->world; ++r) sy.flags[r] = (uint32_t*)S->peer_sync.ptr[r];
        sy.local = S->d_sync.as<uint32_t>(); sy.rank = c->rank; sy.world = c->world;
        PeerTrees none; memset(&none, 0, sizeof none);
        peer_barrier_kernel<<<1, 64, 0, c->stream>>>(sy, trees ? *trees : none, trees ? 1 : 0);
        GS_CUDA(c, cudaGetLastError());
        c->launches++;
        return GS_OK;
    }
    ProfScope ps(c, "nccl_barrier");
    uint32_t* w = S->d_ipc.as<uint32_t>();
    const int nr = nccl().AllReduce(w, w, 1, GS_NCCL_UINT32, GS_NCCL_SUM, c->comm, c->stream);
    if (nr != 0) return c->fail(GS_E_CUDA, "ncclAllReduce(barrier): %s", nccl().GetErrorString(nr));
    if (trees) {
        ProfScope ps2(c, "nccl_allgather_roots");
        uint32_t* tree = trees->base[c->rank];
        const int nr2 = nccl().AllGather(tree + 8 * (c->world + c->rank), tree + 8 * c->world, 32, GS_NCCL_UINT8, c->comm, c->stream);
        if (nr2 != 0) return c->fail(GS_E_CUDA, "ncclAllGather: %s", nccl().GetErrorString(nr2));
    }
    return GS_OK;
}
// split-tree commit with the digests already in place (hash_columns_scatter): barrier, sub-tree, roots, top
static inline int commit_split_tree_peer(Stark* S, long long n, uint32_t* tree, const PeerTrees& trees) {
    Ctx* c = S->ctx;
    const Shard& sh = S->shard;
    int log_w = 0; while ((1 << log_w) < sh.world) ++log_w;
    int rc;
    if ((rc = shard_barrier(S, nullptr))) return rc;                       // every rank's block has landed in the staging buffer
    // leaves of this rank's range, in leaf order: leaf[(q - q0) * E + s * El + jl] = stage[s][(q - q0) * El + jl]
    const long long blk = ((n >> sh.log_e) >> log_w) << sh.log_el;
    uint32_t* leaves = tree + 8 * (n + (long long)sh.rank * (n >> log_w));
    const long long total = (blk << log_w) * 2;
    permute_digests_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(S->d_stage.as<uint4>(), reinterpret_cast<uint4*>(leaves), blk, sh.log_el, log_w);
    c->launches++;
    if ((rc = merkle_build_range(c, S->hash_alg, tree, n, log_w, sh.rank))) return rc;
    if ((rc = shard_barrier(S, &trees))) return rc;                        // sub-tree roots everywhere
    return merkle_build_top(c, S->hash_alg, tree, sh.world);
}

// Merkle commit of the sharded prover for a large tree: instead of gathering all n digests everywhere and building the
// tree W times, the digests are exchanged all-to-all into contiguous leaf ranges (1/W of the volume per rank), every
// rank builds the sub-tree over its range, the W sub-tree roots are all-gathered (W x 32 bytes) and the top log2(W)
// levels are computed everywhere.  d_local: the n/W digests this rank produced, layout [q][local coset].
static inline int commit_split_tree(Stark* S, const uint32_t* d_local, long long n, uint32_t* tree) {
    Ctx* c = S->ctx;
    const Shard& sh = S->shard;
    int log_w = 0; while ((1 << log_w) < sh.world) ++log_w;
    const int W = sh.world, El = 1 << sh.log_el;
    const long long rows = n >> sh.log_e;                 // q values of this commit
    const long long rows_r = rows >> log_w;                // q range owned as LEAVES by one rank
    const long long blk = rows_r * El;                     // digests sent to each peer
    int rc;
    if ((rc = S->d_dig_all.ensure(c, (size_t)(blk * W) * 32))) return rc;
    { ProfScope ps(c, "nccl_alltoall_digests");
      int nr = nccl().GroupStart();
      for (int peer = 0; peer < W && nr == 0; ++peer) {
          nr = nccl().Send(d_local + 8 * (peer * blk), (size_t)blk * 32, GS_NCCL_UINT8, peer, c->comm, c->stream);
          if (nr == 0) nr = nccl().Recv(S->d_dig_all.as<uint32_t>() + 8 * (peer * blk), (size_t)blk * 32, GS_NCCL_UINT8, peer, c->comm, c->stream);
      }
      const int ne = nccl().GroupEnd();
      if (nr != 0 || ne != 0) return c->fail(GS_E_CUDA, "NCCL all-to-all: %s", nccl().GetErrorString(nr ? nr : ne)); }
    // leaves of this rank's range, in leaf order: leaf[(q - q0) * E + s * El + jl] = recv[s][(q - q0) * El + jl]
    uint32_t* leaves = tree + 8 * (n + (long long)sh.rank * (n >> log_w));
    const long long total = (blk << log_w) * 2;
    permute_digests_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(S->d_dig_all.as<uint4>(), reinterpret_cast<uint4*>(leaves), blk, sh.log_el, log_w);
    c->launches++;
    if ((rc = merkle_build_range(c, S->hash_alg, tree, n, log_w, sh.rank))) return rc;
    { ProfScope ps(c, "nccl_allgather_roots");
      const int nr = nccl().AllGather(tree + 8 * (W + sh.rank), tree + 8 * W, 32, GS_NCCL_UINT8, c->comm, c->stream);
      if (nr != 0) return c->fail(GS_E_CUDA, "ncclAllGather: %s", nccl().GetErrorString(nr)); }
    return merkle_build_top(c, S->hash_alg, tree, W);
}
// owner of tree node `id` in a split tree (nodes above the sub-tree roots are everywhere: rank 0 speaks for them)
static inline int split_tree_owner(unsigned long long id, int log_w) {
    if (id < (2ull << log_w)) return 0;
    int lvl = 63 - __builtin_clzll(id);                  // level with 2^lvl nodes
    return (int)((id - (1ull << lvl)) >> (lvl - log_w));
}

// run `fn` (which only enqueues work on the context stream) directly, or capture it once and replay it
template <typename F>
static inline int run_region(Stark* S, GraphSlot& slot, unsigned long long key, bool allow_graph, F&& fn) {
    Ctx* c = S->ctx;
    if (!allow_graph) return fn();
    if (!slot.exec || slot.key != key) {
        if (slot.exec) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
        const unsigned long long l0 = c->launches;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); S->use_graphs = false; return fn(); }
        const int rc = fn();
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (rc != GS_OK) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
        if (e != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); cudaGetLastError(); S->use_graphs = false; c->launches = l0; return fn(); }
        const cudaError_t e2 = cudaGraphInstantiate(&slot.exec, g, 0);
        cudaGraphDestroy(g);
        if (e2 != cudaSuccess) { slot.exec = nullptr; cudaGetLastError(); S->use_graphs = false; c->launches = l0; return fn(); }
        slot.key = key; slot.launches = c->launches - l0; c->launches = l0;
    }
    const cudaError_t e = cudaGraphLaunch(slot.exec, c->stream);
    if (e != cudaSuccess) return c->cuda_fail(e, "cudaGraphLaunch");
    c->launches += slot.launches;
    return GS_OK;
}

struct FriLayer {
    fp* v; long long len; uint32_t* tree; uint8_t root[32];
    bool split;        // sharded prover: the tree is split into W sub-trees (one per rank) below the level with W nodes
    bool replicated;   // sharded prover: from this layer on every rank holds the whole vector and tree (see shard_gather_log)
};

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Build a BatchMerkleProof for `indexes` of a device-resident tree (n leaves): plan on the host, gather
// the needed digests on the device, copy them back.
static inline int device_merkle_proof(Stark* S, const uint32_t* tree, uint64_t n, const std::vector<uint32_t>& indexes,
                                      BatchProof& bp, std::string& err) {
    Ctx* c = S->ctx;
    if (merkle_prove_plan(indexes, n, bp, err) != 0) return c->fail(GS_E_STARK, "%s", err.c_str());
    std::vector<uint32_t> flat;
    for (auto& col : bp.node_ids) flat.insert(flat.end(), col.begin(), col.end());
    bp.nodes.assign(bp.node_ids.size(), {});
    if (flat.empty()) return GS_OK;
    int rc = S->d_idx.ensure(c, flat.size() * 4); if (rc) return rc;
    rc = S->d_gather.ensure(c, flat.size() * 32); if (rc) return rc;
    if (flat.size() * 32 > c->mailbox_bytes) return c->fail(GS_E_STARK, "batch proof too large");
    GS_CUDA(c, cudaMemcpyAsync(S->d_idx.p, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, c->stream));
    const int threads = 128, total = (int)flat.size() * 2;
    gather_digests_kernel<<<(total + threads - 1) / threads, threads, 0, c->stream>>>(tree, S->d_idx.as<unsigned>(), (int)flat.size(),
                                                                                       S->d_gather.as<uint32_t>());
    c->launches++;
    GS_CUDA(c, cudaMemcpyAsync(c->mailbox, S->d_gather.p, flat.size() * 32, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    const uint8_t* src = (const uint8_t*)c->mailbox;
    size_t k = 0;
    for (size_t col = 0; col < bp.node_ids.size(); ++col) {
        bp.nodes[col].resize(bp.node_ids[col].size());
        for (size_t j = 0; j < bp.node_ids[col].size(); ++j, ++k) memcpy(bp.nodes[col][j].data(), src + 32 * k, 32);
    }
    return GS_OK;
}

// gather rows (idx) of `ncols` column vectors into raw value buffers
static inline int device_gather_rows(Stark* S, const std::vector<const fp*>& cols, const std::vector<uint32_t>& idx,
                                     std::vector<std::vector<uint8_t>>& values) {
    Ctx* c = S->ctx;
    const size_t nq = idx.size(), nc = cols.size();
    values.assign(nq, {});
    if (!nq) return GS_OK;
    int rc = S->d_idx.ensure(c, nq * 4); if (rc) return rc;
    rc = S->d_gather.ensure(c, nq * nc * 16); if (rc) return rc;
    if (nq * nc * 16 > c->mailbox_bytes) return c->fail(GS_E_STARK, "too many queried values");
    GatherCols gc; gc.ncols = (int)nc;
    for (size_t i = 0; i < nc; ++i) gc.col[i] = cols[i];
    GS_CUDA(c, cudaMemcpyAsync(S->d_idx.p, idx.data(), nq * 4, cudaMemcpyHostToDevice, c->stream));
    const int threads = 128, total = (int)(nq * nc);
    gather_rows_kernel<<<(total + threads - 1) / threads, threads, 0, c->stream>>>(gc, S->d_idx.as<unsigned>(), (int)nq, S->d_gather.as<fp>());
    c->launches++;
    GS_CUDA(c, cudaMemcpyAsync(c->mailbox, S->d_gather.p, nq * nc * 16, cudaMemcpyDeviceToHost, c->stream));
    GS_CUDA(c, cudaStreamSynchronize(c->stream));
    const uint8_t* src = (const uint8_t*)c->mailbox;
    for (size_t q = 0; q < nq; ++q) values[q].assign(src + q * nc * 16, src + (q + 1) * nc * 16);
    return GS_OK;
}

static inline std::vector<uint32_t> first_seen_unique(const std::vector<uint32_t>& v) {
    std::vector<uint32_t> out; std::map<uint32_t, bool> seen;
    for (uint32_t x : v) if (!seen.count(x)) { seen[x] = true; out.push_back(x); }
    return out;
}

// The tail of the FRI layer chain in ONE launch (LowDegreeProver.ts:176-221 for the layers of at most 2^GS_FRI_TAIL_LOG values).
// Those layers are pure latency: per layer a row-hash launch, a tree launch (one dependent compression per level), a one-thread
// SHA-256 for the challenge and a fold launch -- 27-40 us each for a few thousand hashes.  One block walks them all: row hashes,
// the tree level by level, x* = prng(root), the fold, with block barriers in between; roots, epoch flags and the remainder are
// written straight into the pinned mailbox (host-mapped), system-fenced, so the host plans the queries behind them as before.
// Same buffers and layouts as the per-layer kernels, so the query phase does not know the difference.
struct FriTailParams {
    fp* v; fp* v_next; uint32_t* trees;
    int log_l0, depth0, log_n;
    const fp* tw_lo; const fp* tw_hi; int log_g, log_lo;
    fp iota_inv, quarter_inv;
    uint32_t* mb_root; uint32_t* mb_flag; fp* mb_rem;       // pinned host memory (unified addressing)
    const uint32_t* epoch;
};

// digests of a level live in shared memory (ping-pong between two buffers: one block barrier per tree level, children read
// at shared-memory latency); every node is also stored to the tree in HBM, where the query phase reads authentication paths
template <int ALG>
__global__ void __launch_bounds__(512) fri_tail_kernel(const FriTailParams P) {
    extern __shared__ __align__(16) unsigned char tail_smem[];
    __shared__ fp s_x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    fp* v = P.v; fp* vn = P.v_next; uint32_t* tree = P.trees;
    const unsigned g_mask = (1u << P.log_g) - 1u;
    const uint32_t epoch = *P.epoch;
    const int q0 = (1 << P.log_l0) >> 2;
    uint4* buf_a = reinterpret_cast<uint4*>(tail_smem);           // q0 digests
    uint4* buf_b = buf_a + 2 * q0;                                // q0 / 2 digests
    int depth = P.depth0;
    for (int L = 1 << P.log_l0;; L >>= 2, ++depth) {
        const int Q = L >> 2;
        for (int i = tid; i < Q; i += nthr) {                 // row hashes: H(v[i] || v[i+Q] || v[i+2Q] || v[i+3Q])
            uint32_t m[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const fp e = ld_fp(v + i + c * Q); m[4 * c] = e.v[0]; m[4 * c + 1] = e.v[1]; m[4 * c + 2] = e.v[2]; m[4 * c + 3] = e.v[3]; }
            uint32_t d[8];
            auto getm = [&](int w) -> uint32_t { return m[w]; };
            hash_words<ALG, true>(getm, 16, d);
            buf_a[2 * i] = make_uint4(d[0], d[1], d[2], d[3]); buf_a[2 * i + 1] = make_uint4(d[4], d[5], d[6], d[7]);
            store_digest(tree + 8 * (size_t)(Q + i), d);
        }
        __syncthreads();
        uint4* src = buf_a; uint4* dst = buf_b;
        for (int c = Q >> 1; c >= 1; c >>= 1) {               // tree levels: parents c .. 2c-1 from children in src
            for (int j = tid; j < c; j += nthr) {
                uint32_t m[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) { const uint4 t = src[4 * j + q]; m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
                uint32_t d[8];
                auto getm = [&](int w) -> uint32_t { return m[w]; };
                hash_words<ALG, true>(getm, 16, d);
                dst[2 * j] = make_uint4(d[0], d[1], d[2], d[3]); dst[2 * j + 1] = make_uint4(d[4], d[5], d[6], d[7]);
                store_digest(tree + 8 * (size_t)(c + j), d);
            }
            __syncthreads();
            uint4* tmp = src; src = dst; dst = tmp;
        }
        // src[0..1] = root (Q >= 2 always: L >= 8 on this path)
        const uint32_t* root = reinterpret_cast<const uint32_t*>(src);
        if (tid < 8) { P.mb_root[8 * depth + tid] = root[tid]; __threadfence_system(); }
        if (L <= 256) {
            for (int i = tid; i < L; i += nthr) { st_fp(P.mb_rem + i, ld_fp(v + i)); }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) { P.mb_flag[depth] = epoch; __threadfence_system(); }
            return;
        }
        if (tid == 0) s_x = fri_challenge_dev(root);
        __syncthreads();
        if (tid == 0) { P.mb_flag[depth] = epoch; __threadfence_system(); }
        const fp xs = s_x;
        const int x_shift = 2 * depth + (P.log_g - P.log_n);
        for (int i = tid; i < Q; i += nthr) {
            const fp y0 = ld_fp(v + i), y1 = ld_fp(v + i + Q), y2 = ld_fp(v + i + 2 * Q), y3 = ld_fp(v + i + 3 * Q);
            const unsigned e = (0u - ((unsigned)i << x_shift)) & g_mask;
            fp xinv = ldg_fp(P.tw_lo + (e & ((1u << P.log_lo) - 1u)));
            if (P.log_g > P.log_lo) xinv = fp_mul(xinv, ldg_fp(P.tw_hi + (e >> P.log_lo)));
            const fp t = fp_mul(xs, xinv);
            const fp s02 = fp_add(y0, y2), d02 = fp_sub(y0, y2);
            const fp s13 = fp_add(y1, y3), d13 = fp_mul(fp_sub(y1, y3), P.iota_inv);
            const fp c0 = fp_add(s02, s13), c2 = fp_sub(s02, s13);
            const fp c1 = fp_add(d02, d13), c3 = fp_sub(d02, d13);
            fp acc = fp_add(fp_mul(c3, t), c2);
            acc = fp_add(fp_mul(acc, t), c1);
            acc = fp_add(fp_mul(acc, t), c0);
            st_fp(vn + i, fp_mul(acc, P.quarter_inv));
        }
        __syncthreads();
        tree += (size_t)2 * Q * 8; v = vn; vn += Q;
    }
}
template <int ALG>
static inline cudaError_t fri_tail_launch(const FriTailParams& F, cudaStream_t st) {
    const size_t smem = ((size_t)1 << F.log_l0) / 4 * 48;        // q0 * 32 (buffer A) + q0 * 16 (buffer B)
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(fri_tail_kernel<ALG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
    fri_tail_kernel<ALG><<<1, 512, smem, st>>>(F);
    return cudaGetLastError();
}
// Sharded prover: the first FRI layer of at most 2^this many values is all-gathered (one NCCL call, 16 bytes per value) and the
// rest of the chain runs replicated on every rank with the single-GPU code -- below that size a layer is latency, not work, and
// every sharded layer costs two or three collectives per commit.  GS_SHARD_GATHER_LOG=0: keep every layer sharded.
static inline int shard_gather_log() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_SHARD_GATHER_LOG"); v = e ? atoi(e) : 19; if (v < 0) v = 0; }
    return v;
}
static inline int fri_tail_log() {       // layers of at most 2^this many values go to fri_tail_kernel; GS_FRI_TAIL_LOG=0: none
    static int v = -1;
    if (v < 0) { const char* e = getenv("GS_FRI_TAIL_LOG"); v = e ? atoi(e) : 11; if (v > 14) v = 14; if (v < 0) v = 0; }        // 2^14 values: 4096 rows, 192 KB of digests
    return v;
}

// ---------------------------------------------------------------------------------------- prove
// inputs: initial state (R elements), input register traces (n_input x T, register order), shapes blob
static inline int stark_prove(Stark* S, const Assertion* asserts, int n_assert, const u128* init_state,
                              const fp* input_traces, const uint8_t* shapes_blob, size_t shapes_len,
                              std::vector<uint8_t>& proof_out, int flags = 0) {
    Ctx* c = S->ctx;
    cudaSetDevice(c->device);
    StageTimes& tm = S->last_times; tm.items.clear();
    double t_prev = now_ms();
    auto mark = [&](const char* name, bool sync) {
        if (sync) cudaStreamSynchronize(c->stream);
        double t = now_ms(); tm.items.emplace_back(name, t - t_prev); t_prev = t;
    };
    const bool timing = S->keep_intermediates || getenv("GS_STAGE_TIMES") != nullptr;

    if (n_assert <= 0) return c->fail(GS_E_ARG, "At least one assertion must be provided");
    const int R = S->R, K = S->K, log_t = S->log_t, log_e = S->log_e, log_n = log_t + log_e;
    const long long T = 1ll << log_t, N = 1ll << log_n, E = 1ll << log_e;
    const int n_in = S->n_secret + S->n_public;
    const Shard sh = S->shard;
    const bool sharded = sh.world > 1;
    const int log_el = sh.log_el;
    const long long NL = T << log_el;            // evaluation positions owned by this rank (= N when world == 1)
    if (sharded && S->keep_intermediates) return c->fail(GS_E_UNSUPPORTED, "intermediates are only kept on a single GPU");
    if (log_n > c->log_g) return c->fail(GS_E_UNSUPPORTED, "evaluation domain 2^%d exceeds 2^%d", log_n, c->log_g);
    // getComponentCount quirk (LowDegreeProver.ts:287-291): N < 128 throws RangeError in the reference
    if (N < 128) return c->fail(GS_E_STARK, "Low degree proof failed: Invalid array length");
    int rc;

    const double t_call = now_ms();
    const bool reuse_trace = (flags & 1) && S->trace_resident;
    if (!S->ev0) { cudaEventCreate(&S->ev0); cudaEventCreate(&S->ev1); cudaEventCreate(&S->ev2); }
    // 1-2 ---- execution trace on the host (sequential in steps), checked against the assertions
    const size_t trace_bytes = (size_t)R * T * sizeof(fp);
    if (!reuse_trace) {
    if (S->h_trace_bytes < trace_bytes) {
        if (S->h_trace) cudaFreeHost(S->h_trace);
        GS_CUDA(c, cudaHostAlloc(&S->h_trace, trace_bytes, cudaHostAllocDefault));
        S->h_trace_bytes = trace_bytes;
    }
    {
        fp* tr = (fp*)S->h_trace;
        // the device copy of finished 2^16-step blocks is enqueued while the next block is being generated
        if ((rc = S->d_trace.ensure(c, trace_bytes))) return rc;
        const TraceChunkFn on_chunk = [&](long long s0, long long s1) {
            for (int r = 0; r < R; ++r)
                cudaMemcpyAsync(S->d_trace.as<fp>() + (size_t)r * T + s0, tr + (size_t)r * T + s0, (size_t)(s1 - s0) * sizeof(fp), cudaMemcpyHostToDevice, c->stream);
        };
        generate_trace(S, init_state, input_traces, tr, &on_chunk);
        for (int a = 0; a < n_assert; ++a) {
            if ((int)asserts[a].reg >= R) return c->fail(GS_E_STARK, "Failed to generate the execution trace: Invalid assertion: register %u is outside of register bank", asserts[a].reg);
            if (asserts[a].step >= (uint64_t)T) return c->fail(GS_E_STARK, "Failed to generate the execution trace: Invalid assertion: step %u is outside of execution trace", asserts[a].step);
            if (fp_to_u128(tr[(size_t)asserts[a].reg * T + asserts[a].step]) != asserts[a].value)
                return c->fail(GS_E_STARK, "Failed to generate the execution trace: Assertion at step %u, register %u conflicts with execution trace", asserts[a].step, asserts[a].reg);
        }
    }
    }
    mark("Generated execution trace", false);

    // 3-4 ---- P(x) = iNTT(trace); low-degree extension; leaf hashing; Merkle tree  (one captured region)
    cudaEventRecord(S->ev0, c->stream);
    const int wrows = R > n_in ? R : (n_in > 0 ? n_in : 1);
    const size_t in_bytes = (size_t)n_in * T * sizeof(fp);
    long long fri_tot_v = 0, fri_tot_t = 0;
    for (long long L = N; ; L >>= 2) { fri_tot_t += 2 * (L >> 2); if (L <= 256) break; fri_tot_v += L >> 2; }
    if ((rc = S->d_trace.ensure(c, trace_bytes)) || (rc = S->d_poly.ensure(c, trace_bytes)) || (rc = S->d_pe.ensure(c, (size_t)R * NL * sizeof(fp))) ||
        (rc = S->d_work.ensure(c, (size_t)wrows * NL * sizeof(fp))) || (rc = S->d_tree.ensure(c, (size_t)2 * N * 32)) ||
        (rc = S->d_l.ensure(c, (size_t)NL * sizeof(fp))) || (rc = S->d_fri.ensure(c, (size_t)((fri_tot_v >> (log_e - log_el)) + 4 + 256) * sizeof(fp))) ||
        (rc = S->d_fri_trees.ensure(c, (size_t)fri_tot_t * 32)) || (rc = S->d_params.ensure(c, sizeof(ComposeParams))) ||
        (rc = S->d_small.ensure(c, 1 << 20)) || (rc = S->d_epoch.ensure(c, 64))) return rc;
    if (sharded && (rc = S->d_dig_loc.ensure(c, (size_t)NL * 32))) return rc;
    if (n_in > 0 && ((rc = S->d_in_trace.ensure(c, in_bytes)) || (rc = S->d_in_poly.ensure(c, in_bytes)) || (rc = S->d_in_e.ensure(c, (size_t)n_in * NL * sizeof(fp))))) return rc;
    if (S->keep_intermediates && (rc = S->d_c.ensure(c, (size_t)N * sizeof(fp)))) return rc;
    if (!reuse_trace) {
        if (n_in > 0) GS_CUDA(c, cudaMemcpyAsync(S->d_in_trace.p, input_traces, in_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    std::vector<const fp*> e_cols;            // eVectors: trace rows then secret rows (Stark.ts:113-114)
    for (int r = 0; r < R; ++r) e_cols.push_back(S->d_pe.as<fp>() + (size_t)r * NL);
    std::vector<const fp*> in_cols(S->statics.size(), nullptr);
    {
        int ii = 0;
        for (size_t k = 0; k < S->statics.size(); ++k) if (S->statics[k].kind != 0) in_cols[k] = S->d_in_e.as<fp>() + (size_t)(ii++) * NL;
        for (size_t k = 0; k < S->statics.size(); ++k) if (S->statics[k].kind == 1) e_cols.push_back(in_cols[k]);
    }
    if (e_cols.size() > GS_MAX_HASH_COLS) return c->fail(GS_E_UNSUPPORTED, "more than %d committed registers", GS_MAX_HASH_COLS);
    uint32_t* e_tree = S->d_tree.as<uint32_t>();
    // the sharded chains are captured too (NCCL calls are stream-ordered and capturable; every rank enqueues the same sequence)
    static const bool shard_graphs = !(getenv("GS_SHARD_GRAPHS") && atoi(getenv("GS_SHARD_GRAPHS")) == 0);
    const bool graphs = S->use_graphs && !c->profiling && !timing && S->proves_done > 0 && (!sharded || shard_graphs);
    // any reallocation changes a pointer below and invalidates the captured graphs
    unsigned long long gkey = 1469598103934665603ull;
    for (const DevBuf* b : {&S->d_trace, &S->d_poly, &S->d_pe, &S->d_work, &S->d_tree, &S->d_l, &S->d_fri, &S->d_fri_trees, &S->d_params,
                            &S->d_small, &S->d_in_trace, &S->d_in_poly, &S->d_in_e, &S->d_c, &S->d_u, &S->d_cyc, &S->d_instrs, &S->d_consts, &S->d_epoch,
                            &S->d_dig_loc, &S->d_dig_all, &S->d_fri_rep, &S->d_sync, &S->d_stage})
        gkey = (gkey ^ (unsigned long long)(uintptr_t)b->p) * 1099511628211ull;
    gkey = (gkey ^ (unsigned long long)S->keep_intermediates) * 1099511628211ull;
    // trees with at least 2^14 leaves per ... are split across the ranks; small ones are replicated
    auto split_ok = [&](long long n) { return sharded && n >= (16384ll * sh.world) && (n >> log_e) >= sh.world; };
    const bool e_split = split_ok(N);
    int log_w_all = 0; while ((1 << log_w_all) < sh.world) ++log_w_all;
    bool peers_ok = false;
    if (sharded && shard_peer_enabled() && sh.world <= 8 && e_split) {
        if (!S->d_sync.p) { if ((rc = S->d_sync.ensure(c, 256))) return rc; GS_CUDA(c, cudaMemsetAsync(S->d_sync.p, 0, 256, c->stream)); GS_CUDA(c, cudaStreamSynchronize(c->stream)); }
        if ((rc = S->d_stage.ensure(c, (size_t)NL * 32))) return rc;            // W blocks of NL / W digests (the largest commit)
        if ((rc = peer_map_update(S, S->peer_tree, S->d_tree.p)) || (rc = peer_map_update(S, S->peer_fri, S->d_fri_trees.p)) ||
            (rc = peer_map_update(S, S->peer_stage, S->d_stage.p))) return rc;
        static const bool peer_barrier = !(getenv("GS_SHARD_PEER_BARRIER") && atoi(getenv("GS_SHARD_PEER_BARRIER")) == 0);
        if (peer_barrier && (rc = peer_map_update(S, S->peer_sync, S->d_sync.p))) return rc;
        peers_ok = S->peer_tree.ok && S->peer_fri.ok && S->peer_stage.ok;
    }
    auto commit_region = [&]() -> int {
        int r2;
        if ((r2 = ntt_run(c, S->d_trace.as<fp>(), T, S->d_poly.as<fp>(), T, S->d_work.as<fp>(), T, R, log_t, 0, true))) return r2;
        if (timing) mark("Computed execution trace polynomials P(x)", true);
        // each rank extends onto the cosets it owns (all of them on a single GPU)
        if ((r2 = ntt_run(c, S->d_poly.as<fp>(), T, S->d_pe.as<fp>(), NL, S->d_work.as<fp>(), NL, R, log_t, log_el, false, sh.j0(), log_e))) return r2;
        if (n_in > 0) {      // input registers (secret: committed; public: only feed the constraints)
            if ((r2 = ntt_run(c, S->d_in_trace.as<fp>(), T, S->d_in_poly.as<fp>(), T, S->d_work.as<fp>(), T, n_in, log_t, 0, true))) return r2;
            if ((r2 = ntt_run(c, S->d_in_poly.as<fp>(), T, S->d_in_e.as<fp>(), NL, S->d_work.as<fp>(), NL, n_in, log_t, log_el, false, sh.j0(), log_e))) return r2;
        }
        if (timing) mark("Low-degree extended P(x) polynomials over evaluation domain", true);
        HashCols hc; hc.ncols = (int)e_cols.size();
        for (size_t i = 0; i < e_cols.size(); ++i) hc.col[i] = e_cols[i];
        if (!sharded) { if ((r2 = merkle_commit(c, S->hash_alg, &hc, e_tree, N))) return r2; }       // leaves + tree
        else if (e_split && peers_ok) {
            PeerTrees pt; memset(&pt, 0, sizeof pt);
            PeerStage pst; memset(&pst, 0, sizeof pst);
            for (int r = 0; r < sh.world; ++r) { pt.base[r] = (uint32_t*)S->peer_tree.ptr[r]; pst.base[r] = (uint32_t*)S->peer_stage.ptr[r]; }
            if ((r2 = hash_columns_scatter(c, S->hash_alg, hc, NL, pst, NL >> log_w_all, sh.rank))) return r2;
            if (timing) mark("Serialized evaluations of P(x) and S(x) polynomials", true);
            if ((r2 = commit_split_tree_peer(S, N, e_tree, pt))) return r2;
        }
        else {
            if ((r2 = hash_columns(c, S->hash_alg, hc, NL, S->d_dig_loc.as<uint32_t>()))) return r2;
            if (e_split) { if ((r2 = commit_split_tree(S, S->d_dig_loc.as<uint32_t>(), N, e_tree))) return r2; }
            else if ((r2 = commit_gather(S, S->d_dig_loc.as<uint32_t>(), NL, e_tree + 8 * N))) return r2;
        }
        if (timing && !(e_split && peers_ok)) mark("Serialized evaluations of P(x) and S(x) polynomials", true);
        if (sharded && !e_split && (r2 = merkle_build(c, S->hash_alg, e_tree, N))) return r2;
        GS_CUDA(c, cudaMemcpyAsync(c->mailbox, e_tree + 8, 32, cudaMemcpyDeviceToHost, c->stream));
        return GS_OK;
    };
    if ((rc = run_region(S, S->g_commit, gkey, graphs, commit_region))) return rc;
    // 5 ---- composition polynomial: coefficients, boundary polynomials, fused evaluation.  Everything that does not
    // depend on the evaluation root (degrees, interpolants, partial fractions, the E-periodic tables) is computed
    // on the host while the commit chain is still running on the device; the root is awaited after that.
    int max_deg = 1;
    for (int d : S->degrees) if (d > max_deg) max_deg = d;
    int log_comp = 0; while ((1 << log_comp) < max_deg) ++log_comp;
    const long long comb_degree = T << log_comp;                                    // CompositionPolynomial.ts:196-204
    const long long comp_degree = std::max(comb_degree - T, T);                     // :40
    // constraint groups by degree, first-appearance order (:206-225)
    std::vector<long long> group_deg; std::vector<std::vector<int>> group_idx;
    for (int k = 0; k < K; ++k) {
        const long long dg = (long long)S->degrees[k] * T;
        size_t g = 0; for (; g < group_deg.size(); ++g) if (group_deg[g] == dg) break;
        if (g == group_deg.size()) { group_deg.push_back(dg); group_idx.emplace_back(); }
        group_idx[g].push_back(k);
    }
    int d_count = K;
    for (size_t g = 0; g < group_deg.size(); ++g) if (group_deg[g] < comb_degree) d_count += (int)group_idx[g].size();
    // boundary constraints: registers in first-appearance order (BoundaryConstraints.ts:19-44)
    std::vector<uint32_t> b_regs; std::vector<std::vector<u128>> b_xs, b_ys;
    const u128 w_n = c->root_of_order(log_n);
    for (int a = 0; a < n_assert; ++a) {
        size_t b = 0; for (; b < b_regs.size(); ++b) if (b_regs[b] == asserts[a].reg) break;
        if (b == b_regs.size()) { b_regs.push_back(asserts[a].reg); b_xs.emplace_back(); b_ys.emplace_back(); }
        b_xs[b].push_back(h_pow(w_n, (u128)asserts[a].step * (u128)E));
        b_ys[b].push_back(asserts[a].value);
    }
    const int nB = (int)b_regs.size();
    int b_count = nB; if (comp_degree > T) b_count *= 2;
    // I(x) per asserted register, and the partial-fraction form of 1/Z_b(x) (see compose.cuh)
    std::vector<fp> ipoly; std::vector<u128> pf_coef; std::vector<int> pf_owner; std::vector<unsigned> pf_shift; std::vector<int> ioff(nB), ilen(nB), pfoff(nB), pflen(nB), breg(nB);
    {
        std::vector<std::vector<unsigned>> b_steps(nB);
        for (int a = 0; a < n_assert; ++a) { size_t b = 0; for (; b < b_regs.size(); ++b) if (b_regs[b] == asserts[a].reg) break; b_steps[b].push_back(asserts[a].step); }
        for (int b = 0; b < nB; ++b) {
            const std::vector<u128> ip = h_interpolate(b_xs[b], b_ys[b]);
            ioff[b] = (int)ipoly.size(); ilen[b] = (int)ip.size(); for (u128 v : ip) ipoly.push_back(fp_from_u128(v));
            pfoff[b] = (int)pf_coef.size(); pflen[b] = (int)b_xs[b].size();
            for (size_t k = 0; k < b_xs[b].size(); ++k) {
                u128 den = 1;
                for (size_t m = 0; m < b_xs[b].size(); ++m) if (m != k) {
                    if (b_xs[b][m] == b_xs[b][k]) return c->fail(GS_E_STARK, "Failed to generate the execution trace: repeated assertion for register %u", b_regs[b]);
                    den = h_mul(den, h_sub(b_xs[b][k], b_xs[b][m]));
                }
                // c_k * X_k^-1
                pf_coef.push_back(h_mul(h_inv(den), h_inv(b_xs[b][k])));
                pf_shift.push_back((unsigned)((unsigned long long)b_steps[b][k] * (unsigned long long)E));
                pf_owner.push_back(b);
            }
            breg[b] = (int)b_regs[b];
        }
    }
    // per-constraint power slot and the position of its second coefficient in the stream (:88-100)
    std::vector<int> pow_idx(K, -1), adj_idx(K, -1);
    std::vector<unsigned long long> pow_incr;
    {
        int next = K;
        for (size_t g = 0; g < group_deg.size(); ++g) {
            if (group_deg[g] == comb_degree) continue;                                    // :89
            pow_incr.push_back((unsigned long long)(comb_degree - group_deg[g]));
            for (int k : group_idx[g]) { adj_idx[k] = next++; pow_idx[k] = (int)pow_incr.size() - 1; }
        }
    }
    const int n_lc = (int)e_cols.size();
    const long long delta = comp_degree - T;
    const int lc_total = delta > 0 ? 2 * n_lc : n_lc;
    // E-periodic factors: 1/(x^T - 1) (num_i = w^(i*T) - 1 depends on i mod E, ZeroPolynomial.ts:40-41; inv(0) = 0 at
    // i mod E == 0), x^incr per constraint group and x^delta (every increment is a multiple of T).  The random
    // coefficients are folded with them into E-entry tables here, once per proof (compose.cuh).
    std::vector<u128> inv_num(E);
    {
        const u128 w_e = c->root_of_order(log_e);        // w^T
        u128 acc = 1;
        for (long long j = 0; j < E; ++j) { inv_num[j] = h_sub(acc, 1); acc = h_mul(acc, w_e); }
        inv_num = h_batch_inverse(inv_num);
    }
    std::vector<u128> pow_tab(pow_incr.size() * E), delta_tab(E, 1);
    for (size_t g = 0; g < pow_incr.size(); ++g) {
        if (pow_incr[g] % (unsigned long long)T) return c->fail(GS_E_UNSUPPORTED, "degree increment is not a multiple of the trace length");
        const u128 base = h_pow(w_n, (u128)(pow_incr[g] % (unsigned long long)N));
        u128 a = 1; for (long long j = 0; j < E; ++j) { pow_tab[g * E + j] = a; a = h_mul(a, base); }
    }
    if (delta > 0) {
        const u128 base = h_pow(w_n, (u128)((unsigned long long)delta % (unsigned long long)N));
        u128 a = 1; for (long long j = 0; j < E; ++j) { delta_tab[j] = a; a = h_mul(a, base); }
    }
    uint8_t ev_root[32];
    mark("Built evaluation merkle tree", timing);
    // The random coefficients (one draw covers the composition coefficients, :58-60, and the linear-combination coefficients,
    // which continue the same stream: LinearCombination.ts:58-59, Stark.ts:129) are drawn from the evaluation root ON THE DEVICE
    // (coeffs.cuh) and folded there with the E-periodic factors into cd_tab / pf_tab / lk_tab: no host round trip sits between
    // the commit chain and compose + FRI.  The host reads the root at the end, for the proof bytes.
    std::vector<fp> inv_num_fp(E), pow_tab_fp(pow_tab.size()), delta_tab_fp(E), pf_coef_fp(pf_coef.size());
    for (long long j = 0; j < E; ++j) { inv_num_fp[j] = fp_from_u128(inv_num[j]); delta_tab_fp[j] = fp_from_u128(delta_tab[j]); }
    for (size_t i = 0; i < pow_tab.size(); ++i) pow_tab_fp[i] = fp_from_u128(pow_tab[i]);
    for (size_t i = 0; i < pf_coef.size(); ++i) pf_coef_fp[i] = fp_from_u128(pf_coef[i]);
    const size_t n_cd = (size_t)K * E, n_pf_tab = pf_coef.size() * E, n_lk = (size_t)n_lc * E;
    const std::vector<fp> cd_tab(n_cd, fp_zero()), pf_tab(n_pf_tab, fp_zero()), lk_tab(n_lk, fp_zero());     // filled by derive_coeffs_kernel
    // small-object upload: one packed buffer
    std::vector<uint8_t> small;
    auto put = [&](const void* p, size_t n) { size_t off = (small.size() + 15) & ~(size_t)15; small.resize(off + n); memcpy(small.data() + off, p, n); return off; };
    const int flag_init[2] = {0x7FFFFFFF, 0};
    const size_t o_flag = put(flag_init, 8);        // fixed offset 0: the captured graph copies it back
    const size_t o_cd = put(cd_tab.data(), cd_tab.size() * 16), o_pf = put(pf_tab.data(), pf_tab.size() * 16), o_lk = put(lk_tab.data(), lk_tab.size() * 16);
    const size_t o_ip = put(ipoly.data(), ipoly.size() * 16), o_ps = put(pf_shift.data(), pf_shift.size() * 4);
    const size_t o_io = put(ioff.data(), nB * 4), o_il = put(ilen.data(), nB * 4), o_po = put(pfoff.data(), nB * 4), o_pl = put(pflen.data(), nB * 4);
    const size_t o_br = put(breg.data(), nB * 4);
    const size_t o_inv = put(inv_num_fp.data(), inv_num_fp.size() * 16), o_pow = put(pow_tab_fp.data(), pow_tab_fp.size() * 16),
                 o_dl = put(delta_tab_fp.data(), delta_tab_fp.size() * 16), o_pc = put(pf_coef_fp.data(), pf_coef_fp.size() * 16),
                 o_pw = put(pf_owner.data(), pf_owner.size() * 4), o_pi = put(pow_idx.data(), K * 4), o_ai = put(adj_idx.data(), K * 4);
    if (small.size() + 64 > S->d_small.cap) return c->fail(GS_E_UNSUPPORTED, "too many assertions / constraints for the parameter block");
    uint8_t* ds = S->d_small.as<uint8_t>();
    GS_CUDA(c, cudaMemcpyAsync(ds, small.data(), small.size(), cudaMemcpyHostToDevice, c->stream));
    {
        CoeffParams Q; memset(&Q, 0, sizeof Q);
        Q.root = e_tree + 8;
        Q.K = K; Q.nB = nB; Q.n_lc = n_lc; Q.n_pf = (int)pf_coef.size(); Q.E = (int)E;
        Q.d_count = d_count; Q.b_count = b_count; Q.has_delta = delta > 0 ? 1 : 0; Q.comp_gt_t = comp_degree > T ? 1 : 0;
        Q.pow_idx = (const int*)(ds + o_pi); Q.adj_idx = (const int*)(ds + o_ai);
        Q.inv_num = (const fp*)(ds + o_inv); Q.pow_tab = (const fp*)(ds + o_pow); Q.delta_tab = (const fp*)(ds + o_dl);
        Q.pf_coef = (const fp*)(ds + o_pc); Q.pf_owner = (const int*)(ds + o_pw);
        Q.cd_tab = (fp*)(ds + o_cd); Q.pf_tab = (fp*)(ds + o_pf); Q.lk_tab = (fp*)(ds + o_lk);
        Q.count = d_count + b_count + lc_total;
        if ((size_t)Q.count * 16 > 40 * 1024) return c->fail(GS_E_UNSUPPORTED, "too many random coefficients (%d)", Q.count);
        derive_coeffs_kernel<<<1, 256, (size_t)Q.count * 16, c->stream>>>(Q);
        GS_CUDA(c, cudaGetLastError());
        c->launches++;
    }
    {
        ComposeParams P; memset(&P, 0, sizeof P);
        P.n = N; P.log_n = log_n; P.log_e = log_e; P.n_loc = NL; P.log_el = log_el; P.j0 = sh.j0();
        P.instrs = S->d_instrs.as<uint4>(); P.n_instr = (int)S->evaluation.instrs.size(); P.consts = S->d_consts.as<fp>(); P.n_slots = S->evaluation.n_slots;
        P.n_trace = R; for (int r = 0; r < R; ++r) P.trace[r] = S->d_pe.as<fp>() + (size_t)r * NL;
        P.n_static = (int)S->statics.size();
        for (size_t k = 0; k < S->statics.size(); ++k) {
            if (S->statics[k].kind == 0) { P.stat[k] = S->d_cyc.as<fp>() + S->cyc_off[k]; P.stat_mask[k] = S->cyc_mask[k]; }
            else { P.stat[k] = in_cols[k]; P.stat_mask[k] = 0xFFFFFFFFu; }
        }
        P.n_constraints = K; P.cd_tab = (const fp*)(ds + o_cd);
        P.x_last = fp_from_u128(h_pow(w_n, (u128)(T - 1) * (u128)E));
        P.n_boundary = nB; P.b_reg = (const int*)(ds + o_br); P.b_ipoly_off = (const int*)(ds + o_io); P.b_ipoly_len = (const int*)(ds + o_il);
        P.b_ipoly = (const fp*)(ds + o_ip);
        P.b_pf_off = (const int*)(ds + o_po); P.b_pf_len = (const int*)(ds + o_pl); P.pf_tab = (const fp*)(ds + o_pf); P.b_pf_shift = (const unsigned*)(ds + o_ps);
        P.u_table = S->d_u.as<fp>();
        P.n_lc = n_lc; for (int j = 0; j < n_lc; ++j) P.lc_col[j] = e_cols[j];
        P.lk_tab = (const fp*)(ds + o_lk);
        P.tw_lo = c->tw_lo; P.tw_hi = c->tw_hi; P.log_g = c->log_g; P.log_lo = c->log_lo;
        P.out = S->d_l.as<fp>(); P.c_out = S->keep_intermediates ? S->d_c.as<fp>() : nullptr;
        P.fail_flag = (int*)(ds + o_flag);
        GS_CUDA(c, cudaMemcpyAsync(S->d_params.p, &P, sizeof P, cudaMemcpyHostToDevice, c->stream));
    }
    // 5b-7 ---- compose + the whole FRI layer chain (second captured region)
    std::vector<FriLayer> layers;
    uint8_t* mb = (uint8_t*)c->mailbox;
    // mailbox layout: [0, 32) evaluation root, [32, 40) the compose fail flag, [48, 52) peer-barrier timeout word; [1024, 1792) FRI layer roots; [2048, 2144) one
    // epoch flag per layer; [4096, 8192) remainder; [8192, ...) gathered query data
    const size_t MB_ROOT = 1024, MB_FLAG = 2048, MB_REM = 4096;
    int n_layers = 0;
    // epoch flag: the chain copies it to the mailbox right behind each layer root; the host polls for it
    // (events recorded inside a captured graph cannot be synchronised from the host).  The mailbox belongs to the
    // context, which several Stark instances share, and a prove can fail half way: the epoch counts every ATTEMPT on
    // the context (never a value an earlier prove left behind) and the flag / root slots are cleared first -- nothing
    // in flight writes them: the previous prove was synchronised to its end and this prove's commit chain only writes
    // mailbox[0, 32).
    if (++c->prove_epoch == 0) ++c->prove_epoch;
    const uint32_t epoch = c->prove_epoch;
    memset(mb + MB_ROOT, 0, 32 * 24);
    memset(mb + MB_FLAG, 0, 4 * 24);
    memset(mb + 48, 0, 4);
    GS_CUDA(c, cudaMemcpyAsync(S->d_epoch.p, &epoch, 4, cudaMemcpyHostToDevice, c->stream));
    GS_CUDA(c, cudaMemsetAsync(S->d_epoch.as<uint8_t>() + 32, 0, 16, c->stream));
    auto fri_region = [&]() -> int {
        int rc;   // shadows the outer one on purpose: this lambda may run under stream capture
        layers.clear(); n_layers = 0;
        {
        const unsigned g = grid_for(c, NL, 256);
        const int ns = S->evaluation.n_slots;
        ProfScope ps(c, "compose");
        if (S->compose_jit && S->compose_jit->fn) {
            const ComposeParams* dp = S->d_params.as<ComposeParams>();
            void* args[1] = {(void*)&dp};
            const CUresult r = driver_api().LaunchKernel(S->compose_jit->fn, g, 1, 1, 256, 1, 1, 0, (CUstream)c->stream, args, nullptr);
            if (r != CUDA_SUCCESS) return c->fail(GS_E_CUDA, "cuLaunchKernel(gs_compose_jit): error %d", (int)r);
        }
        else if (ns <= 8) compose_kernel<8><<<g, 256, 0, c->stream>>>(S->d_params.as<ComposeParams>());
        else if (ns <= 32) compose_kernel<32><<<g, 256, 0, c->stream>>>(S->d_params.as<ComposeParams>());
        else if (ns <= 128) compose_kernel<128><<<g, 256, 0, c->stream>>>(S->d_params.as<ComposeParams>());
        else return c->fail(GS_E_UNSUPPORTED, "evaluation program needs %d value slots (max 128)", ns);
        c->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return c->cuda_fail(e, "compose_kernel");
        GS_CUDA(c, cudaMemcpyAsync((uint8_t*)c->mailbox + 32, ds + o_flag, 8, cudaMemcpyDeviceToHost, c->stream));
        }
    // 7 ---- low-degree proof (LowDegreeProver.ts:39-68,176-221)
    const u128 iota_inv = h_inv(c->root_of_order(2));
    const u128 quarter_inv = h_inv(4);
    fp* d_special = S->d_fri.as<fp>();            // slot 0..3 reserved for the challenge
    fp* v_next = S->d_fri.as<fp>() + 4;
    uint32_t* t_next = S->d_fri_trees.as<uint32_t>();
    fp* v_cur = S->d_l.as<fp>();
    // The whole layer chain is enqueued without host round trips: each challenge x* = prng(root_d) is derived on
    // the device; roots are copied to mailbox slots as they appear and the host plans the queries behind them.
    bool in_tail = false, replicated = false;
    for (int depth = 0;; ++depth) {
        const long long L = N >> (2 * depth), Q = L >> 2;
        if (sharded && !replicated && shard_gather_log() > 0 && L <= (1ll << shard_gather_log())) {
            // all-gather this layer's vector, restore position order, continue replicated
            int log_w = 0; while ((1 << log_w) < sh.world) ++log_w;
            const long long LL = L >> log_w;
            if ((rc = S->d_fri_rep.ensure(c, (size_t)(L * 2 + L / 2 + 1024) * sizeof(fp)))) return rc;
            fp* gathered = S->d_fri_rep.as<fp>();
            fp* nat = gathered + L;
            { ProfScope ps(c, "nccl_allgather_layer");
              const int nr = nccl().AllGather(v_cur, gathered, (size_t)LL * 16, GS_NCCL_UINT8, c->comm, c->stream);
              if (nr != 0) return c->fail(GS_E_CUDA, "ncclAllGather: %s", nccl().GetErrorString(nr)); }
            permute_elems_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(gathered), reinterpret_cast<uint4*>(nat), LL, log_el, log_w);
            c->launches++;
            v_cur = nat; v_next = nat + L;
            replicated = true;
        }
        const bool shl = sharded && !replicated;               // this layer is sharded
        const long long QL = shl ? (NL >> (2 * depth)) >> 2 : Q;          // rows of this layer held by this rank
        FriLayer ly; ly.v = v_cur; ly.len = L; ly.tree = t_next; ly.split = false; ly.replicated = replicated; t_next += (size_t)2 * Q * 8;
        if (in_tail || (!shl && fri_tail_log() >= 8 && L <= (1ll << fri_tail_log()))) {
            if (!in_tail) {
                in_tail = true;
                if (depth >= 24) return c->fail(GS_E_UNSUPPORTED, "too many FRI layers");
                FriTailParams F; memset(&F, 0, sizeof F);
                F.v = v_cur; F.v_next = v_next; F.trees = ly.tree;
                int ll = 0; while ((1ll << ll) < L) ++ll;
                F.log_l0 = ll; F.depth0 = depth; F.log_n = log_n;
                F.tw_lo = c->tw_lo; F.tw_hi = c->tw_hi; F.log_g = c->log_g; F.log_lo = c->log_lo;
                F.iota_inv = fp_from_u128(iota_inv); F.quarter_inv = fp_from_u128(quarter_inv);
                F.mb_root = (uint32_t*)(mb + MB_ROOT); F.mb_flag = (uint32_t*)(mb + MB_FLAG); F.mb_rem = (fp*)(mb + MB_REM);
                F.epoch = S->d_epoch.as<uint32_t>();
                ProfScope ps(c, "fri_tail");
                GS_CUDA(c, S->hash_alg == HASH_BLAKE2S ? fri_tail_launch<HASH_BLAKE2S>(F, c->stream) : fri_tail_launch<HASH_SHA256>(F, c->stream));
                c->launches++;
            }
            layers.push_back(ly); ++n_layers;
            if (L <= 256) break;
            v_cur = v_next; v_next += QL;
            continue;
        }
        HashCols hc; hc.ncols = 4; for (int j = 0; j < 4; ++j) hc.col[j] = v_cur + j * QL;
        if (!sharded) { /* single GPU: leaves and tree in the same launches (merkle_commit below) */ }
        else if (!shl) { if ((rc = hash_columns(c, S->hash_alg, hc, Q, ly.tree + 8 * Q))) return rc; }
        else {
            if (split_ok(Q) && peers_ok) {
                ly.split = true;
                PeerTrees pt; memset(&pt, 0, sizeof pt);
                PeerStage pst; memset(&pst, 0, sizeof pst);
                const size_t off = (size_t)(ly.tree - S->d_fri_trees.as<uint32_t>());
                for (int r = 0; r < sh.world; ++r) { pt.base[r] = (uint32_t*)S->peer_fri.ptr[r] + off; pst.base[r] = (uint32_t*)S->peer_stage.ptr[r]; }
                if ((rc = hash_columns_scatter(c, S->hash_alg, hc, QL, pst, QL >> log_w_all, sh.rank))) return rc;
                if ((rc = commit_split_tree_peer(S, Q, ly.tree, pt))) return rc;
            } else {
            if ((rc = hash_columns(c, S->hash_alg, hc, QL, S->d_dig_loc.as<uint32_t>()))) return rc;
            if (split_ok(Q)) { ly.split = true; if ((rc = commit_split_tree(S, S->d_dig_loc.as<uint32_t>(), Q, ly.tree))) return rc; }
            else if ((rc = commit_gather(S, S->d_dig_loc.as<uint32_t>(), QL, ly.tree + 8 * Q))) return rc;
            }
        }
        if (depth >= 24) return c->fail(GS_E_UNSUPPORTED, "too many FRI layers");
        bool have_challenge = false;
        RootSink sink; sink.challenge_out = (L > 256) ? d_special + (depth & 3) : nullptr;
        sink.mb_root = (uint32_t*)(mb + MB_ROOT + 32 * depth); sink.mb_flag = (uint32_t*)(mb + MB_FLAG + 4 * depth); sink.epoch = S->d_epoch.as<uint32_t>();
        if (!sharded) { if ((rc = merkle_commit(c, S->hash_alg, &hc, ly.tree, Q, &sink, &have_challenge))) return rc; }
        else if (!ly.split && (rc = merkle_build(c, S->hash_alg, ly.tree, Q, &sink, &have_challenge))) return rc;
        if (!have_challenge) {          // split (sharded) trees and the multi-launch paths: root and flag by copies
            GS_CUDA(c, cudaMemcpyAsync(mb + MB_ROOT + 32 * depth, ly.tree + 8, 32, cudaMemcpyDeviceToHost, c->stream));
            GS_CUDA(c, cudaMemcpyAsync(mb + MB_FLAG + 4 * depth, S->d_epoch.p, 4, cudaMemcpyDeviceToHost, c->stream));
        }
        layers.push_back(ly); ++n_layers;
        if (L <= 256) {
            if (!shl) GS_CUDA(c, cudaMemcpyAsync(mb + MB_REM, v_cur, L * sizeof(fp), cudaMemcpyDeviceToHost, c->stream));
            else {
                // remainder: gather the per-rank pieces and restore position order
                int log_w = 0; while ((1 << log_w) < sh.world) ++log_w;
                const long long LL = L >> log_w;
                if ((rc = S->d_dig_all.ensure(c, (size_t)L * 16 * 2))) return rc;
                const int nr = nccl().AllGather(v_cur, S->d_dig_all.p, (size_t)LL * 16, GS_NCCL_UINT8, c->comm, c->stream);
                if (nr != 0) return c->fail(GS_E_CUDA, "ncclAllGather: %s", nccl().GetErrorString(nr));
                uint4* nat = S->d_dig_all.as<uint4>() + L;
                permute_elems_kernel<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(S->d_dig_all.as<uint4>(), nat, LL, log_el, log_w);
                c->launches++;
                GS_CUDA(c, cudaMemcpyAsync(mb + MB_REM, nat, L * sizeof(fp), cudaMemcpyDeviceToHost, c->stream));
            }
            break;
        }
        { ProfScope ps(c, "fri_fold");
          if (!have_challenge || !sink.challenge_out) { fri_challenge_kernel<<<1, 1, 0, c->stream>>>(ly.tree + 8, d_special + (depth & 3)); c->launches++; }
          FriFoldParams F; F.v = v_cur; F.out = v_next; F.quarter = QL; F.special_x = d_special + (depth & 3);
          F.log_e = log_e; F.log_el = shl ? log_el : log_e; F.j0 = shl ? sh.j0() : 0;
          F.tw_lo = c->tw_lo; F.tw_hi = c->tw_hi; F.log_g = c->log_g; F.log_lo = c->log_lo;
          F.x_shift = 2 * depth + (c->log_g - log_n); F.iota_inv = fp_from_u128(iota_inv); F.quarter_inv = fp_from_u128(quarter_inv);
          fri_fold_kernel<<<grid_for(c, QL, 256), 256, 0, c->stream>>>(F); }
        c->launches += 1;
        v_cur = v_next; v_next += QL;
    }
        return GS_OK;
    };
    if ((rc = run_region(S, S->g_fri, gkey ^ 0x9E3779B97F4A7C15ull, graphs, fri_region))) return rc;
    if (graphs && layers.empty()) {
        // replayed graph: rebuild the layer table (the pointer arithmetic of the loop above, no launches)
        fp* vc = S->d_l.as<fp>(); fp* vn = S->d_fri.as<fp>() + 4; uint32_t* tn = S->d_fri_trees.as<uint32_t>();
        bool repl = false;
        for (int depth = 0;; ++depth) {
            const long long L = N >> (2 * depth), Q = L >> 2;
            if (sharded && !repl && shard_gather_log() > 0 && L <= (1ll << shard_gather_log())) {
                fp* nat = S->d_fri_rep.as<fp>() + L;
                vc = nat; vn = nat + L; repl = true;
            }
            const bool shl = sharded && !repl;
            const long long QL = shl ? (NL >> (2 * depth)) >> 2 : Q;
            FriLayer ly; ly.v = vc; ly.len = L; ly.tree = tn; ly.split = shl && split_ok(Q); ly.replicated = repl; tn += (size_t)2 * Q * 8;
            layers.push_back(ly); ++n_layers;
            if (L <= 256) break;
            vc = vn; vn += QL;
        }
    }
    cudaEventRecord(S->ev1, c->stream);     // end of the enqueued chain (remainder copy included)

    // ---- query phase, planned on the host while the device runs the layers
    std::string err;
    std::vector<unsigned long long> g_addr;
    auto add_chunk = [&](const void* p) { g_addr.push_back((unsigned long long)(uintptr_t)p); return g_addr.size() - 1; };
    // sharded: tree nodes are contributed by rank 0 only and row values by the rank that owns the row; everything else
    // reads a zero block, and one all-reduce (integer sum) of the gathered buffer gives every rank the full query data
    const void* zero_block = S->d_epoch.as<uint8_t>() + 32;
    int log_w_n = 0; while ((1 << log_w_n) < sh.world) ++log_w_n;
    auto node_ptr = [&](const uint32_t* tree, unsigned long long id, bool split, int half) -> const void* {
        const int owner = (sharded && split) ? split_tree_owner(id, log_w_n) : 0;
        return (!sharded || sh.rank == owner) ? (const void*)(tree + 8ull * id + 4 * half) : zero_block;
    };
    struct PlannedProof { BatchProof bp; std::vector<size_t> node_chunk; std::vector<size_t> value_chunk; int chunks_per_value = 0; };
    auto plan_proof = [&](PlannedProof& pp, const uint32_t* tree, uint64_t n, const std::vector<uint32_t>& idx,
                          const std::vector<const fp*>& cols, bool split, bool repl = false) -> int {
        if (merkle_prove_plan(idx, n, pp.bp, err) != 0) return c->fail(GS_E_STARK, "%s", err.c_str());
        for (auto& col : pp.bp.node_ids) for (uint32_t id : col) { pp.node_chunk.push_back(add_chunk(node_ptr(tree, id, split, 0))); add_chunk(node_ptr(tree, id, split, 1)); }
        pp.chunks_per_value = (int)cols.size();
        for (uint32_t i : idx) {
            const bool mine = !sharded || (repl ? sh.rank == 0 : sh.owner(i) == sh.rank);
            const long long il = (sharded && !repl) ? sh.to_local(i) : (long long)i;
            for (const fp* cp : cols) pp.value_chunk.push_back(add_chunk(mine ? (const void*)(cp + il) : zero_block));
        }
        return GS_OK;
    };
    auto aug4 = [&](const std::vector<uint32_t>& p, long long column_length) {
        std::vector<uint32_t> m(p.size()); const uint32_t row_len = (uint32_t)(column_length >> 2);
        for (size_t i = 0; i < p.size(); ++i) m[i] = p[i] % row_len;
        return first_seen_unique(m);
    };
    int log_w_q = 0; while ((1 << log_w_q) < sh.world) ++log_w_q;
    auto row_cols = [&](const FriLayer& ly) { std::vector<const fp*> cols; for (int j = 0; j < 4; ++j) cols.push_back(ly.v + j * ((ly.len >> (ly.replicated ? 0 : log_w_q)) >> 2)); return cols; };
    PlannedProof lc_pp, ev_pp;
    struct Comp { const uint8_t* root; PlannedProof column, poly; };
    std::vector<Comp> comps(layers.size() - 1);
    for (int d = 0; d < n_layers; ++d) {
        {
            volatile const uint32_t* flag = (volatile const uint32_t*)(mb + MB_FLAG + 4 * d);
            unsigned spins = 0;
            while (*flag != epoch) {
                _mm_pause();
                if ((++spins & 0xFFFF) == 0) {
                    const cudaError_t q = cudaStreamQuery(c->stream);
                    if (q != cudaErrorNotReady && *flag != epoch) {      // chain finished (or failed) without the flag
                        if (q != cudaSuccess) return c->cuda_fail(q, "FRI chain");
                        if (*flag != epoch) return c->fail(GS_E_CUDA, "FRI layer %d never signalled", d);
                    }
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        memcpy(layers[d].root, mb + MB_ROOT + 32 * d, 32);
        if (d == 0) {
            memcpy(ev_root, c->mailbox, 32);          // copied behind the commit chain, long before the first FRI root
            const int* fl = (const int*)((const uint8_t*)c->mailbox + 32);
            if (fl[0] != 0x7FFFFFFF) { cudaStreamSynchronize(c->stream); return c->fail(GS_E_STARK, "Failed to evaluate transition constraints: Constraint %d didn't evaluate to 0 at step %d", fl[0] & 63, fl[0] >> 6); }
            // lcProof and the trace queries depend on root_0 only (LowDegreeProver.ts:50-54, Stark.ts:147-151)
            std::vector<uint32_t> exe_pos;
            if (pseudorandom_indexes(layers[0].root, (int)std::min<long long>(S->exe_queries, N - N / E), (uint64_t)N, (uint64_t)E, exe_pos, err) != 0) {
                cudaStreamSynchronize(c->stream); return c->fail(GS_E_STARK, "Low degree proof failed: %s", err.c_str()); }
            if ((rc = plan_proof(lc_pp, layers[0].tree, (uint64_t)(N >> 2), aug4(exe_pos, N), row_cols(layers[0]), layers[0].split, layers[0].replicated))) return rc;
            std::vector<uint32_t> m;
            for (uint32_t p : exe_pos) { m.push_back(p); m.push_back((uint32_t)((p + E) % N)); }
            if ((rc = plan_proof(ev_pp, e_tree, (uint64_t)N, first_seen_unique(m), e_cols, e_split))) return rc;
        } else {
            const FriLayer& pl = layers[d - 1]; const FriLayer& cl = layers[d];
            std::vector<uint32_t> positions;
            if (pseudorandom_indexes(cl.root, S->fri_queries, (uint64_t)cl.len, (uint64_t)E, positions, err) != 0) {
                cudaStreamSynchronize(c->stream); return c->fail(GS_E_STARK, "Low degree proof failed: %s", err.c_str()); }
            comps[d - 1].root = cl.root;
            if ((rc = plan_proof(comps[d - 1].column, cl.tree, (uint64_t)(cl.len >> 2), aug4(positions, cl.len), row_cols(cl), cl.split, cl.replicated))) return rc;
            if ((rc = plan_proof(comps[d - 1].poly, pl.tree, (uint64_t)(pl.len >> 2), positions, row_cols(pl), pl.split, pl.replicated))) return rc;
        }
    }
    const uint8_t* lc_root = layers[0].root;
    if (timing) mark("Computed low-degree proof (layers)", false);
    {
        const size_t nch = g_addr.size();
        if (MB_REM + 4096 + nch * 16 > c->mailbox_bytes) return c->fail(GS_E_STARK, "query phase needs %zu bytes of mailbox", nch * 16);
        if ((rc = S->d_idx.ensure(c, nch * 8))) return rc;
        if ((rc = S->d_gather.ensure(c, nch * 16))) return rc;
        GS_CUDA(c, cudaMemcpyAsync(S->d_idx.p, g_addr.data(), nch * 8, cudaMemcpyHostToDevice, c->stream));
        { ProfScope ps(c, "gather_queries");
          gather_chunks_kernel<<<(unsigned)((nch + 255) / 256), 256, 0, c->stream>>>(S->d_idx.as<unsigned long long>(), (int)nch, S->d_gather.as<uint4>()); }
        c->launches++;
        if (sharded) {
            const int nr = nccl().AllReduce(S->d_gather.p, S->d_gather.p, nch * 4, GS_NCCL_UINT32, GS_NCCL_SUM, c->comm, c->stream);
            if (nr != 0) return c->fail(GS_E_CUDA, "ncclAllReduce: %s", nccl().GetErrorString(nr));
        }
        GS_CUDA(c, cudaMemcpyAsync(mb + MB_REM + 4096, S->d_gather.p, nch * 16, cudaMemcpyDeviceToHost, c->stream));
        if (sharded && S->peer_sync.ok) GS_CUDA(c, cudaMemcpyAsync(mb + 48, S->d_sync.as<uint32_t>() + 9, 4, cudaMemcpyDeviceToHost, c->stream));
        cudaEventRecord(S->ev2, c->stream);
    }
    // the remainder check runs on the host while the gather and its copy back are in flight
    // verifyRemainder (:223-252)
    std::vector<u128> remainder;
    {
        GS_CUDA(c, cudaEventSynchronize(S->ev1));      // end of the layer chain: the remainder is in the mailbox
        const int depth = n_layers - 1;
        const long long L = layers[depth].len;
        long long max_deg_p1 = comp_degree; for (int d = 0; d < depth; ++d) max_deg_p1 /= 4;
        remainder.resize(L);
        for (long long i = 0; i < L; ++i) remainder[i] = fp_to_u128(((const fp*)(mb + MB_REM))[i]);
        const u128 rou = h_pow(w_n, (u128)1 << (2 * depth));
        std::vector<long long> pos;
        for (long long i = 0; i < L; ++i) if (i % E) pos.push_back(i);
        if (max_deg_p1 > (long long)pos.size()) { cudaStreamSynchronize(c->stream); return c->fail(GS_E_STARK, "Low degree proof failed: remainder too short for degree %lld", max_deg_p1); }
        std::vector<u128> dom(L); { u128 a = 1; for (long long i = 0; i < L; ++i) { dom[i] = a; a = h_mul(a, rou); } }
        std::vector<u128> xs(max_deg_p1), ys(max_deg_p1);
        for (long long i = 0; i < max_deg_p1; ++i) { xs[i] = dom[pos[i]]; ys[i] = remainder[pos[i]]; }
        const std::vector<u128> poly = h_interpolate(xs, ys);
        for (size_t i = (size_t)max_deg_p1; i < pos.size(); ++i)
            if (h_eval_poly(poly, dom[pos[i]]) != remainder[pos[i]]) {
                cudaStreamSynchronize(c->stream);
                return c->fail(GS_E_STARK, "Low degree proof failed: Remainder is not a valid degree %lld polynomial", max_deg_p1 - 1);
            }
    }
    cudaEventSynchronize(S->ev2);
    cudaEventElapsedTime(&S->last_device_ms, S->ev0, S->ev2);
    if (sharded && S->peer_sync.ok && *(const uint32_t*)(mb + 48) != 0) return c->fail(GS_E_CUDA, "peer barrier timed out: a rank of the sharded prover did not arrive");
    S->trace_resident = true;
    S->proves_done++;
    if (c->profiling) c->prof_collect();
    {
        const uint8_t* mb = (const uint8_t*)c->mailbox + 4096 + 4096;
        auto fill = [&](PlannedProof& pp) {
            size_t k = 0;
            pp.bp.nodes.assign(pp.bp.node_ids.size(), {});
            for (size_t col = 0; col < pp.bp.node_ids.size(); ++col) {
                pp.bp.nodes[col].resize(pp.bp.node_ids[col].size());
                for (size_t j = 0; j < pp.bp.node_ids[col].size(); ++j, ++k) memcpy(pp.bp.nodes[col][j].data(), mb + 16 * pp.node_chunk[k], 32);
            }
            const size_t nv = pp.chunks_per_value ? pp.value_chunk.size() / pp.chunks_per_value : 0;
            pp.bp.values.assign(nv, {});
            for (size_t q = 0; q < nv; ++q) {
                pp.bp.values[q].resize((size_t)pp.chunks_per_value * 16);
                for (int j = 0; j < pp.chunks_per_value; ++j) memcpy(pp.bp.values[q].data() + 16 * j, mb + 16 * pp.value_chunk[q * pp.chunks_per_value + j], 16);
            }
        };
        fill(lc_pp); fill(ev_pp);
        for (auto& cp : comps) { fill(cp.column); fill(cp.poly); }
    }
    BatchProof& lc_proof = lc_pp.bp; BatchProof& ev_proof = ev_pp.bp;
    mark("Computed evaluation spot checks and Merkle proofs", false);

    // serialize (Serializer.ts:35-79)
    std::vector<uint8_t>& out = proof_out; out.clear();
    const size_t ev_leaf = e_cols.size() * 16, ld_leaf = 64;
    for (const BatchProof* p : {&ev_proof, &lc_proof}) if (check_merkle_proof_limits(*p, err) != 0) return c->fail(GS_E_STARK, "%s", err.c_str());
    out.insert(out.end(), ev_root, ev_root + 32);
    write_merkle_proof(out, ev_proof, ev_leaf);
    out.insert(out.end(), lc_root, lc_root + 32);
    write_merkle_proof(out, lc_proof, ld_leaf);
    out.push_back((uint8_t)comps.size());
    for (auto& cp : comps) {
        if (check_merkle_proof_limits(cp.column.bp, err) != 0 || check_merkle_proof_limits(cp.poly.bp, err) != 0) return c->fail(GS_E_STARK, "%s", err.c_str());
        out.insert(out.end(), cp.root, cp.root + 32);
        write_merkle_proof(out, cp.column.bp, ld_leaf);
        write_merkle_proof(out, cp.poly.bp, ld_leaf);
    }
    out.push_back((uint8_t)(remainder.size() == 256 ? 0 : remainder.size()));
    for (u128 v : remainder) { fp f = fp_from_u128(v); const uint8_t* b = (const uint8_t*)&f; out.insert(out.end(), b, b + 16); }
    if (shapes_blob && shapes_len) out.insert(out.end(), shapes_blob, shapes_blob + shapes_len);
    else out.push_back(0);
    mark("Serialized proof", false);
    S->last_host_ms = now_ms() - t_call;
    return GS_OK;
}

}  // namespace gs
