// Third tcgen05 probe: issue rate with the issue loop written the way the conv kernels write it -- the whole warp runs
// warp-uniform code, one elected lane issues, descriptors advance by compile-time constants in fully unrolled loops.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I rcf_unsupvideoseg_b200/csrc tools/microbench/umma_probe3.cu -o build/umma_probe3
#include <cstdio>
#include <cstdlib>
#include "rcf_umma.cuh"
using namespace umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred;
}

// One "M-tile": 9 taps x 4 K-steps [x NPROD products], SW128 K-major A ([pos][64ch], WP positions per tile row), B = 8 KB / tap.
template <int N, int NPROD, int WP>
__global__ void __launch_bounds__(128) k_rate(int ntiles, long long* cyc, int* err) {
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
    if (warp == 1) {
        const uint32_t sA = smem_u32(smem), sB = sA + 96 * 1024;
        constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint64_t hi = ((uint64_t)(16 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t a0 = hi | ((sA >> 4) & 0x3FFF), b0 = hi | ((sB >> 4) & 0x3FFF);
        const long long t0 = clock64();
        if (elect_one()) {
            for (int t = 0; t < ntiles; ++t) {
                const uint32_t dcol = tmem + (t & 3) * N;
                const uint64_t at = a0 + (uint64_t)((t & 1) * (16384 >> 4));
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int pr = 0; pr < NPROD; ++pr) {
                            const uint32_t aoff = ((tap / 3) * WP + (tap % 3)) * 128 + ks * 32 + (pr == 1 ? 48 * 1024 : 0);
                            const uint32_t boff = tap * 8192 + ks * 32 + (pr == 2 ? 72 * 1024 : 0);
                            mma_bf16(dcol, at + (aoff >> 4), b0 + (boff >> 4), idesc, (tap | ks | pr) != 0);
                        }
            }
            mma_commit(&bar);
        }
        __syncwarp();
        if (!mbar_wait(&bar, 0)) *err = 1;
        const long long t1 = clock64();
        if (blockIdx.x == 0 && (tid & 31) == 0) cyc[0] = t1 - t0;
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N, int NPROD, int WP>
static void run(const char* name, long long* dcyc, int* derr) {
    CK(cudaFuncSetAttribute(k_rate<N, NPROD, WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int grid : {1, 148}) {
        const int ntiles = 64;
        CK(cudaMemset(derr, 0, 4));
        k_rate<N, NPROD, WP><<<grid, 128, 200 * 1024>>>(ntiles, dcyc, derr);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
        int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
        printf("rate grid=%3d %-40s %8.2f clk / MMA, %9.1f clk / M-tile (timeout=%d)\n", grid, name, (double)c / (ntiles * 36 * NPROD), (double)c / ntiles, herr);
    }
}

int main() {
    CK(cudaSetDevice(0));
    int* derr; CK(cudaMalloc(&derr, 4));
    long long* dcyc; CK(cudaMalloc(&dcyc, 8));
    run<64, 1, 64>("N=64, 1 product", dcyc, derr);
    run<64, 2, 64>("N=64, 2 products (A hi, A lo)", dcyc, derr);
    run<64, 3, 64>("N=64, 3 products", dcyc, derr);
    run<128, 1, 64>("N=128, 1 product", dcyc, derr);
    run<256, 1, 64>("N=256, 1 product", dcyc, derr);
    run<32, 1, 64>("N=32, 1 product", dcyc, derr);
    run<16, 1, 64>("N=16, 1 product", dcyc, derr);
    return 0;
}
