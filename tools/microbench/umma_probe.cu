// Probe of the tcgen05 building blocks the feature-branch kernels rely on (run on a B200):
//   1. no-swizzle K-major operands read from "plane" tiles with a row shift (convolution tap)      -> numerics
//   2. N = 128 (two weight sets side by side)                                                        -> numerics
//   3. no-swizzle MN-major A and B (weight-gradient orientation, K = positions)                      -> numerics
//   4. issue-rate of back-to-back MMAs for the shapes above                                          -> cycles / MMA
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I rcf_unsupvideoseg_b200/csrc \
//        tools/microbench/umma_probe.cu -o /tmp/umma_probe && /tmp/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "rcf_umma.cuh"

using namespace umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int Q = 160;                 // positions per plane
constexpr int PS_A = Q * 16;           // plane stride of the activation tile
constexpr int NPL = 16;                // 16 planes (two 8-plane sets back to back: "hi" and "lo", or two tile rows)

// mode 0: K-major, N = 64, shift q0     D[m][n] = sum_k A[q0+m][k] W[n][k]            (K = 64)
// mode 1: K-major, N = 128
// mode 2: MN-major A (M = 128 = 16 planes x 8), MN-major B (N = 64 = 8 planes x 8), K = 32 positions starting at x0
__global__ void __launch_bounds__(128) k_numerics(int mode, const __nv_bfloat16* __restrict__ Ag /*[NPL*8][Q] as planes*/,
                                                  const __nv_bfloat16* __restrict__ Bg, int q0, float* __restrict__ D, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    uint8_t* sA = smem;                       // NPL planes x Q x 16 B
    uint8_t* sB = smem + NPL * PS_A;          // mode 0/1: [8 kchunks][N][16 B]; mode 2: [8 planes][Q][16 B]
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = mode == 1 ? 128 : 64;
    // fill: the global arrays are already in plane order (byte images)
    for (int i = tid; i < NPL * PS_A / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(Ag)[i];
    const int bbytes = mode == 2 ? 8 * PS_A : 8 * N * 16;
    for (int i = tid; i < bbytes / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bg)[i];
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc<128>(&tmem_slot);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        if (mode <= 1) {
            const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
            const uint64_t ahi = make_desc(0, PS_A, 128), bhi = make_desc(0, N * 16, 128);
            for (int s = 0; s < 4; ++s)
                mma_bf16(tmem, desc_at(ahi, smem_u32(sA) + 2 * s * PS_A + q0 * 16), desc_at(bhi, smem_u32(sB) + 2 * s * N * 16), idesc, s > 0);
        } else {
            const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
            const uint64_t ahi = make_desc(0, 128, PS_A), bhi = make_desc(0, 128, PS_A);
            for (int s = 0; s < 2; ++s)      // 16 positions per MMA
                mma_bf16(tmem, desc_at(ahi, smem_u32(sA) + (q0 + 16 * s) * 16), desc_at(bhi, smem_u32(sB) + (16 * s) * 16), idesc, s > 0);
        }
        mma_commit(&bar);
    }
    if (!mbar_wait(&bar, 0)) { if (tid == 0) *err = 1; }
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

// Issue-rate: NMMA back-to-back MMAs into the same accumulator.  variant: 0 K-major N=64, 1 K-major N=128, 2 K-major N=256,
// 3 MN-major A/B N=64, 4 K-major N=64 alternating between two accumulators, 5 K-major N = 64 + N = 128 interleaved
__global__ void __launch_bounds__(128) k_rate(int variant, int nmma, long long* cyc, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t sA = smem_u32(smem), sB = sA + 100 * 1024;
        const int N = variant == 1 ? 128 : variant == 2 ? 256 : 64;
        const uint32_t idesc = variant == 3 ? make_idesc_bf16(128, 64, 1, 1) : make_idesc_bf16(128, N, 0, 0);
        const uint32_t idesc128 = make_idesc_bf16(128, 128, 0, 0);
        const uint64_t ad = variant == 3 ? make_desc(0, 128, PS_A) : make_desc(0, PS_A, 128);
        const uint64_t bd = variant == 3 ? make_desc(0, 128, PS_A) : make_desc(0, N * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint32_t sh = (i & 7) * 16;                  // vary the tap shift like the real kernel does
            if (variant == 4) mma_bf16(tmem + (i & 1) * 64, desc_at(ad, sA + sh), desc_at(bd, sB + (i & 3) * 2048), idesc, 1);
            else if (variant == 5) {
                if (i & 1) mma_bf16(tmem, desc_at(ad, sA + sh), desc_at(bd, sB), idesc, 1);
                else mma_bf16(tmem + 64, desc_at(ad, sA + sh), desc_at(make_desc(0, 128 * 16, 128), sB), idesc128, 1);
            } else mma_bf16(tmem, desc_at(ad, sA + sh), desc_at(bd, sB + (i & 3) * 2048), idesc, 1);
        }
        mma_commit(&bar);
        if (!mbar_wait(&bar, 0)) *err = 1;
        const long long t1 = clock64();
        if (blockIdx.x == 0) cyc[0] = t1 - t0;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s sm_%d%d, %d SMs\n", p.name, p.major, p.minor, p.multiProcessorCount);
    int* derr; CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
    float* dD; CK(cudaMalloc(&dD, 128 * 128 * 4));
    srand(1);
    // logical activation array: 128 "channels" (16 planes x 8) x Q positions
    std::vector<float> A(128 * Q);
    for (auto& v : A) v = bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<__nv_bfloat16> Ap((size_t)NPL * Q * 8);
    for (int c = 0; c < NPL; ++c) for (int q = 0; q < Q; ++q) for (int e = 0; e < 8; ++e)
        Ap[((size_t)c * Q + q) * 8 + e] = __float2bfloat16(A[(c * 8 + e) * Q + q]);
    __nv_bfloat16* dA; CK(cudaMalloc(&dA, Ap.size() * 2)); CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
    int fails = 0;
    for (int mode = 0; mode < 3; ++mode) {
        const int N = mode == 1 ? 128 : 64, q0 = mode == 2 ? 7 : 5;
        std::vector<float> W;            // mode 0/1: W[n][k], k < 64; mode 2: G[n][x], x < Q
        std::vector<__nv_bfloat16> Bp;
        if (mode <= 1) {
            W.resize((size_t)N * 64);
            for (auto& v : W) v = bf((rand() % 2001 - 1000) / 1000.f);
            Bp.resize((size_t)8 * N * 8);
            for (int c = 0; c < 8; ++c) for (int n = 0; n < N; ++n) for (int e = 0; e < 8; ++e)
                Bp[((size_t)c * N + n) * 8 + e] = __float2bfloat16(W[(size_t)n * 64 + c * 8 + e]);
        } else {
            W.resize((size_t)64 * Q);
            for (auto& v : W) v = bf((rand() % 2001 - 1000) / 1000.f);
            Bp.resize((size_t)8 * Q * 8);
            for (int c = 0; c < 8; ++c) for (int q = 0; q < Q; ++q) for (int e = 0; e < 8; ++e)
                Bp[((size_t)c * Q + q) * 8 + e] = __float2bfloat16(W[(size_t)(c * 8 + e) * Q + q]);
        }
        __nv_bfloat16* dB; CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
        const int smem_bytes = NPL * PS_A + (mode == 2 ? 8 * PS_A : 8 * N * 16);
        CK(cudaFuncSetAttribute(k_numerics, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaMemset(dD, 0, 128 * 128 * 4));
        k_numerics<<<1, 128, smem_bytes>>>(mode, dA, dB, q0, dD, derr);
        CK(cudaDeviceSynchronize());
        std::vector<float> D((size_t)128 * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0;
            if (mode <= 1) for (int k = 0; k < 64; ++k) ref += (double)A[k * Q + q0 + m] * W[(size_t)n * 64 + k];
            else for (int x = 0; x < 32; ++x) ref += (double)A[m * Q + q0 + x] * W[(size_t)n * Q + x];
            maxerr = fmax(maxerr, fabs(ref - D[(size_t)m * N + n]));
        }
        int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
        const bool ok = maxerr < 1e-3 && !herr;
        printf("numerics mode %d: max |err| = %.3e  timeout=%d  %s\n", mode, maxerr, herr, ok ? "PASS" : "FAIL");
        fails += !ok;
        CK(cudaFree(dB));
    }
    long long* dcyc; CK(cudaMalloc(&dcyc, 8));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char* names[] = {"K-major N=64", "K-major N=128", "K-major N=256", "MN-major A/B N=64", "K-major N=64, 2 accumulators", "N=64 / N=128 interleaved"};
    for (int grid : {1, 148})
        for (int variant = 0; variant < 6; ++variant) {
            const int nmma = 4096;
            k_rate<<<grid, 128, 200 * 1024>>>(variant, nmma, dcyc, derr);
            CK(cudaDeviceSynchronize());
            long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
            int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
            printf("rate grid=%3d %-32s: %8.2f clk / MMA (timeout=%d)\n", grid, names[variant], (double)c / nmma, herr);
        }
    printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
    return fails;
}
