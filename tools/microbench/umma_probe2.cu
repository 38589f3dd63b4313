// Second tcgen05 probe: a host-scripted harness.  The host builds a shared-memory byte image and a list of MMAs
// (operand start offsets + descriptor high bits + instruction descriptor); one generic kernel executes the list and
// returns the accumulator; a second generic kernel measures the issue rate of a cyclic list.  Used to compare the
// no-swizzle "plane" layout with the 128-byte-swizzled [position][64 ch] layout (with tap shifts of one position,
// K-major and MN-major), before committing the conv kernels to one of them.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I rcf_unsupvideoseg_b200/csrc \
//        tools/microbench/umma_probe2.cu -o build/umma_probe2
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <functional>
#include <cuda_bf16.h>
#include "rcf_umma.cuh"

using namespace umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct MmaOp { uint32_t a_off, b_off, d_col, accumulate; uint64_t a_hi, b_hi; uint32_t idesc, pad; };
constexpr int MAX_OPS = 64;
struct Script { int nops; int ncols; MmaOp op[MAX_OPS]; };

__global__ void __launch_bounds__(128) k_script(const uint8_t* __restrict__ image, int image_bytes, Script sc, float* __restrict__ D, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < image_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(image)[i];
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t s0 = smem_u32(smem);
        for (int i = 0; i < sc.nops; ++i) {
            const MmaOp& o = sc.op[i];
            mma_bf16(tmem + o.d_col, o.a_hi | (uint64_t)(((s0 + o.a_off) >> 4) & 0x3FFF), o.b_hi | (uint64_t)(((s0 + o.b_off) >> 4) & 0x3FFF), o.idesc, o.accumulate);
        }
        mma_commit(&bar);
    }
    if (!mbar_wait(&bar, 0)) { if (tid == 0) *err = 1; }
    fence_after_sync();
    for (int c0 = 0; c0 < sc.ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * sc.ncols + c0 + j] = __uint_as_float(v[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

__global__ void __launch_bounds__(128) k_rate(Script sc, int reps, long long* cyc, int* err) {
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
        const uint32_t s0 = smem_u32(smem);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
#pragma unroll 1
            for (int i = 0; i < sc.nops; ++i) {
                const MmaOp& o = sc.op[i];
                mma_bf16(tmem + o.d_col, o.a_hi | (uint64_t)(((s0 + o.a_off) >> 4) & 0x3FFF), o.b_hi | (uint64_t)(((s0 + o.b_off) >> 4) & 0x3FFF), o.idesc, 1);
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
static uint64_t desc_hi(uint32_t lbo, uint32_t sbo, int layout, int base_offset = 0) {
    return ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
           ((uint64_t)(base_offset & 7) << 49) | ((uint64_t)layout << 61);
}
static void put_bf16(std::vector<uint8_t>& img, size_t byte, float v) {
    __nv_bfloat16 h = __float2bfloat16(v);
    memcpy(&img[byte], &h, 2);
}
// 128-byte swizzle on absolute offsets within a 1024-aligned image: 16-byte chunk index ^= (offset >> 7) & 7
static size_t sw128(size_t row_base, int ch) { return row_base + ((((ch >> 3) ^ ((row_base >> 7) & 7)) << 4) | ((ch & 7) << 1)); }

int main() {
    CK(cudaSetDevice(0));
    int* derr; CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
    float* dD; CK(cudaMalloc(&dD, 128 * 512 * 4));
    uint8_t* dimg; CK(cudaMalloc(&dimg, 200 * 1024));
    long long* dcyc; CK(cudaMalloc(&dcyc, 8));
    CK(cudaFuncSetAttribute(k_script, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    srand(3);
    const int Q = 256;                              // positions
    std::vector<float> X(Q * 64), X2(Q * 64), W(64 * 64), G(Q * 64);
    for (auto* v : {&X, &X2, &W, &G}) for (auto& e : *v) e = bf((rand() % 2001 - 1000) / 1000.f);
    int fails = 0;
    auto run = [&](const char* name, const std::vector<uint8_t>& img, const Script& sc, std::function<double(int, int)> ref, int M, int N) {
        CK(cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0, 128 * 512 * 4)); CK(cudaMemset(derr, 0, 4));
        k_script<<<1, 128, 200 * 1024>>>(dimg, (int)img.size(), sc, dD, derr);
        CK(cudaDeviceSynchronize());
        std::vector<float> D((size_t)128 * sc.ncols);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) maxerr = fmax(maxerr, fabs(ref(m, n) - D[(size_t)m * sc.ncols + n]));
        const bool ok = maxerr < 1e-3 && !herr;
        printf("numerics %-58s max|err| %.3e timeout=%d %s\n", name, maxerr, herr, ok ? "PASS" : "FAIL");
        fails += !ok;
        return ok;
    };
    // ---- S0/S1: SW128 K-major A = X[q0+m][k] (rows of 128 B, swizzled on absolute offsets), B = W[n][k] -----------------
    for (int variant = 0; variant < 2; ++variant) {
        const int q0 = 5;
        std::vector<uint8_t> img(Q * 128 + 64 * 128, 0);
        for (int q = 0; q < Q; ++q) for (int c = 0; c < 64; ++c) put_bf16(img, sw128((size_t)q * 128, c), X[q * 64 + c]);
        const size_t boff = (size_t)Q * 128;
        for (int n = 0; n < 64; ++n) for (int c = 0; c < 64; ++c) put_bf16(img, sw128(boff + (size_t)n * 128, c), W[n * 64 + c]);
        Script sc{}; sc.nops = 4; sc.ncols = 64;
        for (int s = 0; s < 4; ++s) {
            MmaOp& o = sc.op[s];
            o.a_off = q0 * 128 + s * 32; o.b_off = (uint32_t)boff + s * 32; o.d_col = 0; o.accumulate = s > 0;
            o.a_hi = desc_hi(16, 1024, 2, variant ? (q0 & 7) : 0); o.b_hi = desc_hi(16, 1024, 2);
            o.idesc = make_idesc_bf16(128, 64, 0, 0);
        }
        run(variant ? "SW128 K-major, shift 5 rows, base_offset = 5" : "SW128 K-major, shift 5 rows, base_offset = 0", img, sc,
            [&](int m, int n) { double r = 0; for (int k = 0; k < 64; ++k) r += (double)X[(q0 + m) * 64 + k] * W[n * 64 + k]; return r; }, 128, 64);
    }
    // ---- S2: SW128 MN-major A: rows (set, ci) = two [pos][64ch] tiles (X at 0, X2 at Q*128), K = 32 positions from x0;
    //          MN-major B: G[pos][co];  D[(set,ci)][co] = sum_x Xset[x0+x][ci] * G[x][co] ---------------------------------
    for (int variant = 0; variant < 2; ++variant) {
        const int x0 = variant ? 8 : 7;
        std::vector<uint8_t> img(3 * Q * 128, 0);
        for (int q = 0; q < Q; ++q) for (int c = 0; c < 64; ++c) {
            put_bf16(img, sw128((size_t)q * 128, c), X[q * 64 + c]);
            put_bf16(img, sw128((size_t)(Q + q) * 128, c), X2[q * 64 + c]);
            put_bf16(img, sw128((size_t)(2 * Q + q) * 128, c), G[q * 64 + c]);
        }
        Script sc{}; sc.nops = 2; sc.ncols = 64;
        for (int s = 0; s < 2; ++s) {
            MmaOp& o = sc.op[s];
            o.a_off = (x0 + 16 * s) * 128; o.b_off = (2 * Q + 16 * s) * 128; o.d_col = 0; o.accumulate = s > 0;
            o.a_hi = desc_hi(Q * 128, 1024, 2); o.b_hi = desc_hi(Q * 128, 1024, 2);
            o.idesc = make_idesc_bf16(128, 64, 1, 1);
        }
        run(variant ? "SW128 MN-major A (2 atoms, LBO = tile), MN-major B, x0 = 8" : "SW128 MN-major A (2 atoms, LBO = tile), MN-major B, x0 = 7", img, sc,
            [&](int m, int n) { const std::vector<float>& S = m < 64 ? X : X2; double r = 0;
                                for (int x = 0; x < 32; ++x) r += (double)S[(x0 + x) * 64 + (m & 63)] * G[x * 64 + n]; return r; }, 128, 64);
    }
    // ---- rates -----------------------------------------------------------------------------------------------------------
    struct RateCase { const char* name; Script sc; };
    std::vector<RateCase> cases;
    auto add = [&](const char* name, int nops, std::function<void(int, MmaOp&)> f) { RateCase rc{name, {}}; rc.sc.nops = nops; rc.sc.ncols = 64; for (int i = 0; i < nops; ++i) f(i, rc.sc.op[i]); cases.push_back(rc); };
    const uint32_t BOFF = 128 * 1024;
    for (int N : {64, 128, 256})
        add(N == 64 ? "SW128 K-major N=64 (taps shift by 1 row)" : N == 128 ? "SW128 K-major N=128" : "SW128 K-major N=256", 36, [&, N](int i, MmaOp& o) {
            o.a_off = (i / 4) * 128 + (i % 4) * 32; o.b_off = BOFF + (i % 4) * 32; o.d_col = 0;
            o.a_hi = desc_hi(16, 1024, 2); o.b_hi = desc_hi(16, 1024, 2); o.idesc = make_idesc_bf16(128, N, 0, 0); });
    add("SW128 K-major N=64, aligned starts only", 32, [&](int i, MmaOp& o) {
        o.a_off = (i / 4) * 1024 + (i % 4) * 32; o.b_off = BOFF + (i % 4) * 32; o.d_col = 0;
        o.a_hi = desc_hi(16, 1024, 2); o.b_hi = desc_hi(16, 1024, 2); o.idesc = make_idesc_bf16(128, 64, 0, 0); });
    add("SW128 K-major N=64, 4 accumulators round robin", 36, [&](int i, MmaOp& o) {
        o.a_off = (i / 4) * 128 + (i % 4) * 32; o.b_off = BOFF + (i % 4) * 32; o.d_col = (i % 4) * 64;
        o.a_hi = desc_hi(16, 1024, 2); o.b_hi = desc_hi(16, 1024, 2); o.idesc = make_idesc_bf16(128, 64, 0, 0); });
    add("SW128 MN-major A/B N=64", 32, [&](int i, MmaOp& o) {
        o.a_off = i * 128; o.b_off = BOFF + (i % 8) * 2048; o.d_col = 0;
        o.a_hi = desc_hi(32 * 1024, 1024, 2); o.b_hi = desc_hi(32 * 1024, 1024, 2); o.idesc = make_idesc_bf16(128, 64, 1, 1); });
    add("no swizzle K-major N=64, K chunks adjacent (LBO=128,SBO=256)", 32, [&](int i, MmaOp& o) {
        o.a_off = i * 256; o.b_off = BOFF + (i % 4) * 2048; o.d_col = 0;
        o.a_hi = desc_hi(128, 256, 0); o.b_hi = desc_hi(128, 256, 0); o.idesc = make_idesc_bf16(128, 64, 0, 0); });
    add("no swizzle K-major N=64, planes (LBO=4096,SBO=128)", 32, [&](int i, MmaOp& o) {
        o.a_off = i * 16; o.b_off = BOFF + (i % 4) * 2048; o.d_col = 0;
        o.a_hi = desc_hi(4096, 128, 0); o.b_hi = desc_hi(1024, 128, 0); o.idesc = make_idesc_bf16(128, 64, 0, 0); });
    add("SW128 K-major M=64 N=8 (issue floor)", 32, [&](int i, MmaOp& o) {
        o.a_off = (i % 4) * 32; o.b_off = BOFF + (i % 4) * 32; o.d_col = 0;
        o.a_hi = desc_hi(16, 1024, 2); o.b_hi = desc_hi(16, 1024, 2); o.idesc = make_idesc_bf16(64, 8, 0, 0); });
    add("SW128 K-major M=128 N=16", 32, [&](int i, MmaOp& o) {
        o.a_off = (i % 4) * 32; o.b_off = BOFF + (i % 4) * 32; o.d_col = 0;
        o.a_hi = desc_hi(16, 1024, 2); o.b_hi = desc_hi(16, 1024, 2); o.idesc = make_idesc_bf16(128, 16, 0, 0); });
    add("SW128 K-major M=128 N=32", 32, [&](int i, MmaOp& o) {
        o.a_off = (i % 4) * 32; o.b_off = BOFF + (i % 4) * 32; o.d_col = 0;
        o.a_hi = desc_hi(16, 1024, 2); o.b_hi = desc_hi(16, 1024, 2); o.idesc = make_idesc_bf16(128, 32, 0, 0); });
    for (int grid : {1, 148})
        for (auto& rc : cases) {
            const int reps = 128;
            CK(cudaMemset(derr, 0, 4));
            k_rate<<<grid, 128, 200 * 1024>>>(rc.sc, reps, dcyc, derr);
            CK(cudaDeviceSynchronize());
            long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
            int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
            printf("rate grid=%3d %-62s %8.2f clk / MMA (timeout=%d)\n", grid, rc.name, (double)c / (reps * rc.sc.nops), herr);
        }
    printf(fails ? "PROBE2: %d numerics case(s) FAILED\n" : "PROBE2 numerics OK\n", fails);
    return 0;
}
