// rcf_conv64.cuh -- shared definitions of the tcgen05 implicit-GEMM kernels for the 64 -> 64 channel 3x3 convolution of
// flow_feat_before_agg (reference models/flow_aggregation_head_with_residual.py:89-91) and its gradients.
//
// Tile geometry ("linear padded tile").  A CTA tile is TR x TW output pixels.  Its input window (TR+2) x Wp, Wp = TW+2,
// is brought into shared memory by ONE TMA tiled load (out-of-image elements arrive as zeros = the padding) as rows of
// 64 bf16 channels (128 B, 128-byte swizzle) indexed by the LINEAR position
// q = r * Wp + x.  An M-tile of the GEMM is 128 CONSECUTIVE positions starting at q = Wp + 1 (the first output pixel),
// so the A operand of tap (ty, tx) is the same tile read from a start address shifted by (ty * Wp + tx) rows: no im2col
// copy.  Rows of an M-tile that fall on halo columns are computed and dropped (2 / Wp of the work).  One M-tile per CTA
// tile ((TR-1) * Wp + TW <= 128): tiles are small (<= 27 KB) so three of them fit beside the 144 KB of weights and the
// TMA latency hides behind the MMAs of the two tiles in front.
#pragma once
#include <stdint.h>

#define C64_THREADS 320          /* warps 0-7 epilogue (TMEM lane quarter x column half), warp 8 MMA issue + weights, warp 9 TMA */
#define C64_EPI_WARPS 8
#define C64_MMA_WARP 8
#define C64_TMA_WARP 9
#define C64_TAP_BYTES 16384      /* one tap of packed weights: 128 rows (hi of co 0..63, lo of co 0..63) x 64 k, bf16 */
#define C64_W_BYTES (9 * C64_TAP_BYTES)
#define C64_MAX_POS 216          /* positions per staged tile buffer (27 KB) */
#define C64_ABUF_BYTES (C64_MAX_POS * 128)
#define C64_NA 3                 /* A-tile ring stages (a TMA tile load takes longer than the MMAs of one tile) */
#define C64_NT 4                 /* accumulator stages in tensor memory: 4 x 128 columns */
#define C64_SMEM_BYTES (C64_W_BYTES + C64_NA * C64_ABUF_BYTES + 256)

/* CTA-pair kernel (rcf_conv64_pair.cu): per-CTA weight image = region Y (9 taps x 64 rows) + region Z (9 taps x 32 rows) */
#define C64_PAIR_Y_BYTES (9 * 8192)
#define C64_PAIR_Z_BYTES (9 * 4096)
#define C64_PAIR_IMAGE_BYTES (C64_PAIR_Y_BYTES + C64_PAIR_Z_BYTES)
#define C64_WPACK_TOTAL_BYTES (C64_W_BYTES + 2 * C64_PAIR_IMAGE_BYTES)   /* == RCF_CONV64_WPACK_BYTES */

struct Conv64Geom {
    int nimg, H, W;
    int TW, TR, Wp;              // tile of TR x TW outputs; padded row Wp = TW + 2
    int tiles_x, tiles_y, ntiles;
    int nmt;                     // M-tiles (128 positions) per tile: always 1 (one accumulator stage per tile)
    int npos;                    // (TR + 2) * Wp staged positions
    uint32_t wp_magic;           // ceil(2^32 / Wp): q / Wp == __umulhi(q, wp_magic) for q < 2^16
};

// Picks the tile that minimises the number of MMA rows (then staged positions) for an H x W image.
static inline Conv64Geom conv64_make_geom(int nimg, int H, int W) {
    Conv64Geom best = {};
    long long best_rows = -1, best_pos = 0;
    for (int TW = 4; TW <= 64; ++TW) {
        const int Wp = TW + 2;
        for (int TR = 1; TR <= 64; ++TR) {
            const int npos = (TR + 2) * Wp, span = (TR - 1) * Wp + TW;
            if (span > 128) break;
            const int nmt = (span + 127) / 128;
            if (npos > C64_MAX_POS || 2 * Wp + 2 + 128 * nmt > C64_MAX_POS) continue;
            const int tx = (W + TW - 1) / TW, ty = (H + TR - 1) / TR;
            const long long tiles = (long long)tx * ty, rows = tiles * nmt * 128, pos = tiles * npos;
            if (best_rows < 0 || rows < best_rows || (rows == best_rows && pos < best_pos)) {
                best_rows = rows; best_pos = pos;
                best.TW = TW; best.TR = TR; best.Wp = Wp; best.tiles_x = tx; best.tiles_y = ty; best.nmt = nmt; best.npos = npos;
            }
        }
    }
    best.nimg = nimg; best.H = H; best.W = W;
    best.ntiles = nimg * best.tiles_x * best.tiles_y;
    best.wp_magic = (uint32_t)((0x100000000ull + best.Wp - 1) / best.Wp);
    return best;
}
