// rcf_forward_fused.cu -- whole forward of the loss core (theta supplied) in ONE launch, ordered so that pass 2
// re-reads the masks from the 126 MB L2 instead of HBM.
//
// Persistent CTAs (one wave: SMs x resident CTAs) pull work items from an atomic ticket in "slot" order: slot t = [pass-1 tiles of frame-direction t,
// pass-2 tiles of frame-direction t - LAG].  The last pass-1 CTA of a frame-direction runs the per-segment solve
// and publishes ready[fd]; a pass-2 CTA of fd waits for that flag.  Because tickets are taken in order, every
// pass-1 tile of fd has been claimed by a RUNNING CTA before any pass-2 tile of fd is claimed, and pass-1 CTAs
// never wait on anything: no deadlock, independent of the hardware's block scheduling order.  Between pass 1 and
// pass 2 of a frame-direction only ~LAG slots (tens of MB) stream through L2, so its masks (6.5 MB at K=4, 480x854)
// are still resident: DRAM traffic of the forward drops from 4K + (12K+8) to (12K+8) bytes per pixel.
// Results are bit-identical to the three-kernel path (same tile bodies, same partial layout, same fixed-order sums).
// STATUS: opt-in (rcf_debug_set_option(RCF_OPT_FUSED_FORWARD, 1)).  On B200 the L2 reuse works (779 MB instead of
// 945 MB of DRAM reads at C2) but the launch is ~6 % slower than the three kernels: with 2 resident CTAs per SM the
// fixed tail of every pass-1 tile (reduce, fence, arrival atomic) is not hidden.  See DESIGN.md section 3.
#include "rcf_loss_dev.cuh"
#include "rcf_moments_dev.cuh"
#include "rcf_segment_dev.cuh"

template <int K, int D, int PX, bool VIS>
__global__ void __launch_bounds__(RCF_BLOCK, (D <= 2) ? 2 : 1) k_fwd_fused(const RcfK a, int total_items) {
    constexpr int NS = rcf_ns(D), SEGD = rcf_segd(D), CF = rcf_cf(D), GM = rcf_gm(K, D);
    constexpr int SM1 = RCF_WARPS * K * NS, SM2 = RCF_WARPS * GM;
    __shared__ float red_raw[SM1 > SM2 ? SM1 : SM2];
    __shared__ float cf[K * CF];
    __shared__ double stat[K * NS];
    __shared__ int s_next[2];     // [0] next item, [1] its ready flag was already seen set
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const int n1 = a.nchunk1, n2 = a.nchunk2, per_slot = n1 + n2;
    int* ticket = a.sync;
    int* arrived = a.sync + 1;
    int* ready = a.sync + 1 + a.nfd;

    // Persistent CTAs: the ticket (and the ready flag) of the NEXT item is fetched while the current tile is being
    // processed, so neither atomic round trip sits on the critical path of a tile.
    if (tid == 0) { s_next[0] = atomicAdd(ticket, 1); s_next[1] = 0; }
    __syncthreads();
    int item = s_next[0];
    int known_ready = 0;
    while (item < total_items) {
        __syncthreads();                              // s_next consumed by everyone; shared buffers free again
        int nxt = 0;
        if (tid == 0) nxt = atomicAdd(ticket, 1);     // result is not needed until the end of this tile
        const int slot = item / per_slot, r = item - slot * per_slot;
        if (r < n1) {
            // ---------------- pass 1 tile ----------------
            const int fd = slot;
            if (fd < a.nfd) {
                moments_tile<K, D, PX>(a, fd, r, reinterpret_cast<float(*)[K * NS]>(red_raw));
                __threadfence();                       // publish this CTA's partials
                __syncthreads();
                if (tid == 0) s_last = (atomicAdd(arrived + fd, 1) == n1 - 1);
                __syncthreads();
                if (s_last) {
                    // last pass-1 tile of this frame-direction: statistics, solve, coefficient pack, then the flag
                    __threadfence();
                    reduce_partials(a.part1 + (size_t)fd * K * NS * n1, K * NS, n1, stat);
                    __syncthreads();
                    const int dir = fd / a.B, b = fd - dir * a.B;
                    if (tid < K) {
                        double* sd = a.segd + ((size_t)fd * K + tid) * SEGD;
                        seg_affine_fwd<D>(stat + tid * NS, sd);
                        float* c = a.coef + ((size_t)fd * K + tid) * CF;
                        c[0] = __ldg(a.theta[dir] + (size_t)b * 2 * K + tid);          // theta [B,2,K]
                        c[1] = __ldg(a.theta[dir] + (size_t)b * 2 * K + K + tid);
                        if constexpr (D > 0) {
                            const double* A = sd + 3 + 3 * D + 2 * D * D;
#pragma unroll
                            for (int i = 0; i < 2 * D; ++i) c[2 + i] = (float)A[i];
#pragma unroll
                            for (int d = 0; d < D; ++d) c[2 + 2 * D + d] = (float)sd[1 + d];
                        }
                        __threadfence();               // coefficient pack visible before the flag
                    }
                    __syncthreads();
                    if (tid == 0) atomicExch(ready + fd, 1);
                }
            }
        } else {
            // ---------------- pass 2 tile ----------------
            const int fd = slot - a.lag;
            if (fd >= 0) {
                if (!known_ready) {
                    if (tid == 0) {
                        while (atomicAdd(ready + fd, 0) == 0) __nanosleep(100);
                        __threadfence();
                    }
                    __syncthreads();
                }
                loss_tile<K, D, PX, VIS>(a, fd, r - n1, cf, reinterpret_cast<float(*)[GM]>(red_raw));
            }
        }
        if (tid == 0) {
            // peek the ready flag of the next item (pass-2 items only) so that its tile can start without a round trip
            int rdy = 0;
            if (nxt < total_items) {
                const int nslot = nxt / per_slot, nr = nxt - nslot * per_slot;
                const int nfd = nslot - a.lag;
                if (nr >= n1 && nfd >= 0) rdy = (*reinterpret_cast<volatile int*>(ready + nfd) != 0);
            }
            s_next[0] = nxt;
            s_next[1] = rdy;
            __threadfence();
        }
        __syncthreads();
        item = s_next[0];
        known_ready = s_next[1];
    }
}

template <int K, int D>
static cudaError_t launch_kd(const RcfK& a, cudaStream_t s) {
    constexpr int VPX = K <= 4 ? 4 : 2;
    const int items = (a.nfd + a.lag) * (a.nchunk1 + a.nchunk2);
    const bool vis = a.vis_gt || a.vis_pred || a.vis_agg || a.vis_res || a.vis_aff;
    static int grid_cache[2] = {0, 0};
    int& grid = grid_cache[vis ? 1 : 0];
    if (grid == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess)
            e = vis ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fwd_fused<K, D, VPX, true>, RCF_BLOCK, 0)
                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fwd_fused<K, D, VPX, false>, RCF_BLOCK, 0);
        if (e != cudaSuccess) return e;
        grid = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int g = grid < items ? grid : items;
    if (vis) k_fwd_fused<K, D, VPX, true><<<g, RCF_BLOCK, 0, s>>>(a, items);
    else k_fwd_fused<K, D, VPX, false><<<g, RCF_BLOCK, 0, s>>>(a, items);
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_k(const RcfK& a, cudaStream_t s) {
    switch (a.D) {
        case 0: return launch_kd<K, 0>(a, s);
        case 2: return launch_kd<K, 2>(a, s);
        case 5: return launch_kd<K, 5>(a, s);
    }
    return cudaErrorInvalidValue;
}

// Vector path, theta supplied.  a.sync (1 + 2*nfd ints) must have been zeroed on the same stream.
cudaError_t rcf_launch_forward_fused(const RcfK& a, cudaStream_t s) {
    switch (a.K) {
        case 1: return launch_k<1>(a, s);
        case 2: return launch_k<2>(a, s);
        case 3: return launch_k<3>(a, s);
        case 4: return launch_k<4>(a, s);
        case 5: return launch_k<5>(a, s);
        case 6: return launch_k<6>(a, s);
        case 7: return launch_k<7>(a, s);
        case 8: return launch_k<8>(a, s);
    }
    return cudaErrorInvalidValue;
}
