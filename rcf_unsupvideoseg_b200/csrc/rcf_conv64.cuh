// rcf_conv64.cuh -- shared definitions of the tcgen05 implicit-GEMM kernels for the 64 -> 64 channel 3x3 convolution of
// flow_feat_before_agg (reference models/flow_aggregation_head_with_residual.py:89-91) and its gradients.
//
// Tile geometry ("linear padded tile").  A CTA tile is TR x TW output pixels.  Its input window (TR+2) x Wp, Wp = TW+2,
// is staged in shared memory as rows of 64 bf16 channels (128 B, 128-byte swizzle) indexed by the LINEAR position
// q = r * Wp + x.  An M-tile of the GEMM is 128 CONSECUTIVE positions starting at q = Wp + 1 (the first output pixel),
// so the A operand of tap (ty, tx) is the same tile read from a start address shifted by (ty * Wp + tx) rows: no im2col
// copy.  Rows of an M-tile that fall on halo columns are computed and dropped (2 / Wp of the work).
#pragma once
#include <stdint.h>

#define C64_THREADS 384          /* warps 0-3 epilogue (TMEM lane quarters), warp 4 MMA issue + weights, warps 5-11 producers */
#define C64_EPI_WARPS 4
#define C64_MMA_WARP 4
#define C64_PROD_WARP0 5
#define C64_PROD_WARPS 7
#define C64_PROD_THREADS (C64_PROD_WARPS * 32)
#define C64_TAP_BYTES 16384      /* one tap of packed weights: 128 rows (hi of co 0..63, lo of co 0..63) x 64 k, bf16 */
#define C64_W_BYTES (9 * C64_TAP_BYTES)
#define C64_MAX_POS 328          /* positions per staged tile buffer */
#define C64_ABUF_BYTES (C64_MAX_POS * 128)
#define C64_SMEM_BYTES (C64_W_BYTES + 2 * C64_ABUF_BYTES + 128)

struct Conv64Geom {
    int nimg, H, W;
    int TW, TR, Wp;              // tile of TR x TW outputs; padded row Wp = TW + 2
    int tiles_x, tiles_y, ntiles;
    int nmt;                     // M-tiles (128 positions) per tile: 1 or 2
    int npos;                    // (TR + 2) * Wp staged positions
};

// Picks the tile that minimises the number of MMA rows (then staged positions) for an H x W image.
static inline Conv64Geom conv64_make_geom(int nimg, int H, int W) {
    Conv64Geom best = {};
    long long best_rows = -1, best_pos = 0;
    for (int TW = 4; TW <= 97; ++TW) {
        const int Wp = TW + 2;
        for (int TR = 1; TR <= 64; ++TR) {
            const int npos = (TR + 2) * Wp, span = (TR - 1) * Wp + TW;
            if (span > 256) break;
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
    return best;
}
