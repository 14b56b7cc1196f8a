// Lab probe for K1 design: what rate does a register-resident radix-2^S DIF (the inner loop of ntt_pass_kernel: modular add,
// modular sub, modular multiplication by a twiddle read from shared memory) reach as a function of elements per thread and
// resident warps per SM?  This is the arithmetic roof of the NTT measured on its real instruction mix, not on a synthetic chain.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o scripts/ntt_lab.bin scripts/ntt_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../genstark_b200/csrc/fp128.cuh"
#include "../genstark_b200/csrc/ntt.cuh"
using namespace gs;

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// iters x one radix-2^S DIF on 2^S register-resident elements; between iterations the elements are rotated so that the
// compiler cannot hoist anything.  MINB = resident CTAs per SM the register allocation is asked to allow.
template <int S, int MINB>
__global__ void __launch_bounds__(256, MINB) dif_loop_kernel(fp* out, const fp* tw_g, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fp* s_tw = reinterpret_cast<fp*>(smem_raw);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_tw[i] = tw_g[i];
    __syncthreads();
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    fp x[1 << S];
#pragma unroll
    for (int i = 0; i < (1 << S); ++i) x[i] = fp_from_u64(0x9E3779B97F4A7C15ull * (t + 1 + i));
    for (int it = 0; it < iters; ++it) {
        dif_butterfly<S>(x, s_tw + ((it & 3) << 4), 10 - 2);       // twiddle index stride 4 -> inside the 1024-entry table
        fp tmp = x[0];
#pragma unroll
        for (int i = 0; i + 1 < (1 << S); ++i) x[i] = x[i + 1];
        x[(1 << S) - 1] = tmp;
    }
    fp r = x[0];
#pragma unroll
    for (int i = 1; i < (1 << S); ++i) r = fp_add(r, x[i]);
    st_fp(out + t, r);
}

template <typename F>
static float time_ms(F&& launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); CHECK(cudaDeviceSynchronize());
    cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms;
}

template <int S, int MINB>
static void sweep(int sms, fp* out, const fp* tw, const char* label) {
    auto k = dif_loop_kernel<S, MINB>;
    CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaFuncAttributes fa; CHECK(cudaFuncGetAttributes(&fa, k));
    for (int per_sm = 1; per_sm <= MINB; ++per_sm) {
        // dynamic shared memory sized so that exactly per_sm CTAs fit on an SM (227 KB usable)
        size_t smem = (size_t)(220 * 1024) / per_sm - 1024;
        if (smem < 16 * 1024) smem = 16 * 1024;
        int occ = 0; CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, smem));
        const int blocks = sms * occ;
        const int iters = 600 >> (S - 2);
        float ms = time_ms([&] { k<<<blocks, 256, smem>>>(out, tw, iters); });
        const double bf = (double)blocks * 256 * iters * (S << (S - 1));
        printf("%-26s regs %3d  CTAs/SM %d (warps/SMSP %2d)  %8.4f ms  %7.1f G butterflies/s\n", label, fa.numRegs, occ, occ * 2, ms, bf / (ms * 1e-3) / 1e9);
    }
}

int main() {
    cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    fp *out, *tw;
    CHECK(cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(fp)));
    CHECK(cudaMalloc(&tw, 1024 * sizeof(fp)));
    fp* h = (fp*)malloc(1024 * sizeof(fp));
    for (int i = 0; i < 1024; ++i) { h[i].v[0] = 0x9E3779B9u * (i + 1); h[i].v[1] = 0x85EBCA6Bu * (i + 3); h[i].v[2] = 0xC2B2AE35u * (i + 5); h[i].v[3] = (0x27D4EB2Fu * (i + 7)) >> 1; }
    CHECK(cudaMemcpy(tw, h, 1024 * sizeof(fp), cudaMemcpyHostToDevice));
    printf("# register-resident radix-2^S DIF loops, 256 threads per CTA, twiddles from shared memory\n");
    sweep<4, 2>(sms, out, tw, "radix-16, 16 el/thread");
    sweep<4, 3>(sms, out, tw, "radix-16, regs for 3 CTAs");
    sweep<3, 4>(sms, out, tw, "radix-8, 8 el/thread");
    sweep<3, 5>(sms, out, tw, "radix-8, regs for 5 CTAs");
    sweep<2, 4>(sms, out, tw, "radix-4, 4 el/thread");
    sweep<2, 6>(sms, out, tw, "radix-4, regs for 6 CTAs");
    CHECK(cudaDeviceSynchronize());
    return 0;
}
