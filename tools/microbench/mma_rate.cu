// Issue-rate microbenchmark of the legacy mma.sync shapes on sm_100a (how many MMAs / clk / SM), next to FFMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 2048, CHAINS = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* clk) {
    float d[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int q = 0; q < 4; ++q) d[c][q] = threadIdx.x * 1e-9f;
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 5, b1 = 6;
    const long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 3)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(b0));
            else if (MODE == 6) {       // packed fp32x2 FMA (sm_100 FFMA2): 2 per chain slot = 4 FMAs per lane
                float2* ff = reinterpret_cast<float2*>(d[c]);
                ff[0] = __ffma2_rn(ff[0], make_float2(1.0001f, 1.0002f), make_float2(0.5f, 0.25f));
                ff[1] = __ffma2_rn(ff[1], make_float2(1.0001f, 1.0002f), make_float2(0.5f, 0.25f));
            }
            else if (MODE == 5) {       // fp64 FMA: 2 per chain slot (d[c][0..1] and d[c][2..3] viewed as doubles)
                double* dd = reinterpret_cast<double*>(d[c]);
                dd[0] = fma(dd[0], 1.0001, 0.5);
                dd[1] = fma(dd[1], 1.0001, 0.5);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) d[c][q] = fmaf(d[c][q], 1.0001f, 0.5f);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int c = 0; c < CHAINS; ++c) for (int q = 0; q < 4; ++q) s += d[c][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double macs_per_op, int ctas_per_sm) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* clk;
    cudaMalloc(&out, sizeof(float) * nsm * ctas_per_sm * 256); cudaMalloc(&clk, sizeof(long long) * nsm * ctas_per_sm);
    k<MODE><<<nsm * ctas_per_sm, 256>>>(out, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<nsm * ctas_per_sm, 256>>>(out, clk);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c0; cudaMemcpy(&c0, clk, sizeof(c0), cudaMemcpyDeviceToHost);
    const double ops_per_sm = (double)ITERS * CHAINS * 8 * ctas_per_sm;   // warp-level ops per SM (8 warps per CTA)
    printf("%-28s %d CTA/SM: %8.1f clk/CTA-loop, %.3f warp-ops/clk/SM, %.0f MAC/clk/SM, %.1f T MAC/s (%s)\n", name, ctas_per_sm,
           (double)c0, ops_per_sm / (double)c0, ops_per_sm / (double)c0 * macs_per_op, ops_per_sm * nsm * macs_per_op / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}

int main() {
    for (int c = 1; c <= 2; ++c) {
        run<0>("mma.m16n8k8 tf32", 16 * 8 * 8, c);
        run<3>("mma.m16n8k4 tf32", 16 * 8 * 4, c);
        run<1>("mma.m16n8k16 bf16", 16 * 8 * 16, c);
        run<2>("mma.m16n8k16 f16", 16 * 8 * 16, c);
        run<4>("ffma x4 (per lane)", 32 * 4, c);
        run<5>("dfma x2 (per lane)", 32 * 2, c);
        run<6>("ffma2 x2 (per lane)", 32 * 4, c);
    }
    return 0;
}
