// Shared by the two dense-grid kernels (query_col.cu: layers 1-3 as GEMMs; query_inc.cu: layer 1
// updated incrementally along the column): layout of the per-column table and of the packed
// parameter buffer (surs_ctx::col_weights).
#pragma once
#include "tc_common.cuh"

namespace col {

using namespace tc;

// ---- per-column vectors (fp32).  Row = [MLP lr | MLP hr | C1 lr | C1 hr] -----------------------
//   per MLP: C0 = W0[:, :320] f + b0 (1024), C2 = W2[:, 512:832] f + b2 (256),
//            C3 = W3[:, 256:576] f + b3 (128), C4 = W4[:, 128:448] f + b4 (1, padded to 4)
//   C1 = 0.01 W1 (W0[:, :320] f + b0) + b1 (512): layer 1 of a column whose layer-0 channels are all
//        on the negative side of the leaky ReLU -- the state query_inc.cu starts every column from
constexpr int CV_C0 = 0, CV_C2 = 1024, CV_C3 = 1280, CV_C4 = 1408, CV_STRIDE = 1412;
constexpr int CV_FLOATS = 2 * CV_STRIDE;                 // what the main kernels stage in shared memory
constexpr int CV_BYTES = CV_FLOATS * 4;                  // 11296 B
constexpr int CV_C1 = CV_FLOATS;                         // + 512 m
constexpr int CV_ROW_FLOATS = CV_FLOATS + 2 * 512;       // 3848 floats = 15392 B per column
constexpr int CV_ROW_BYTES = CV_ROW_FLOATS * 4;

// ---- constant vectors of query_col.cu (fp32), per MLP ------------------------------------------
constexpr int GV_WZ0 = 0, GV_WP0 = 1024, GV_B1 = 2048, GV_WZ2 = 2560, GV_WP2 = 2816, GV_WZ3 = 3072, GV_WP3 = 3200,
              GV_W4Y = 3328, GV_WZ4 = 3456, GV_WP4 = 3457, GV_STRIDE = 3460;
constexpr int GV_BYTES = 2 * GV_STRIDE * 4;              // 27680 B

// ---- constant vectors of query_inc.cu: the LR block has no pred_lr terms -----------------------
constexpr int XV_WZ0 = 0, XV_WZ2 = 1024, XV_WZ3 = 1280, XV_W4Y = 1408, XV_WZ4 = 1536, XV_LR_FLOATS = 1540;
constexpr int XV_WP0 = 1540, XV_WP2 = 2564, XV_WP3 = 2820, XV_WP4 = 2948, XV_HR_FLOATS = 2952;
constexpr int XV_FLOATS = XV_LR_FLOATS + XV_HR_FLOATS;   // HR block starts at XV_LR_FLOATS
constexpr int XV_BYTES = XV_FLOATS * 4;                  // 17968 B

// ---- weight streams -----------------------------------------------------------------------------
constexpr int W128_BLK_BYTES = 128 * 128;                // 128 rows x 64 fp16
constexpr size_t MLP_BYTES = (size_t)40 * W_BLK_BYTES + 4 * (size_t)W128_BLK_BYTES;      // query_col.cu main stream
constexpr int TB_CHUNKS_PER_MLP = 8;                     // 4 x C0, 2 x C1, C2, C3 (+ C4 row)
constexpr size_t TB_MLP_BYTES = (size_t)35 * W_BLK_BYTES + 5 * (size_t)W3_BLK_BYTES;     // table stream
constexpr int XW_BLOCKS_PER_MLP = 16 + 4;                // query_inc.cu: layer 2 (8 K blocks x 2 halves), layer 3
constexpr size_t XW_MLP_BYTES = (size_t)XW_BLOCKS_PER_MLP * W128_BLK_BYTES;
constexpr int G_STRIDE = 324;                            // G = 0.01 W1 W0 [512][cin0 padded]

// byte offsets inside surs_ctx::col_weights
constexpr size_t OFF_MAIN = 0;
constexpr size_t OFF_TABLE = OFF_MAIN + 2 * MLP_BYTES;
constexpr size_t OFF_GV = OFF_TABLE + 2 * TB_MLP_BYTES;
constexpr size_t OFF_G = (OFF_GV + GV_BYTES + 1023) / 1024 * 1024;
constexpr size_t OFF_G0 = OFF_G + (size_t)2 * 512 * G_STRIDE * 4;     // 0.01 W1 b0 + b1   [2][512]
constexpr size_t OFF_QSTAR = OFF_G0 + 2 * 512 * 4;                    // 0.01 W1 w_z       [2][512]
constexpr size_t OFF_RSTAR = OFF_QSTAR + 2 * 512 * 4;                 // 0.01 W1 w_p       [2][512]
constexpr size_t OFF_XW = OFF_RSTAR + 2 * 512 * 4;
constexpr size_t OFF_XV = OFF_XW + 2 * XW_MLP_BYTES;
constexpr size_t OFF_W1H = (OFF_XV + XV_BYTES + 255) / 256 * 256;    // fp16 W1^T [2][1024][512] (query_inc.cu events)
constexpr size_t COL_WEIGHTS_BYTES = OFF_W1H + (size_t)2 * 1024 * 512 * 2;
// split-operand copy of the streams (surs_ctx::col_weights_x3): main stream as (hi, lo) pairs, table stream as (hi, hi, lo)
constexpr size_t X3_TABLE_OFF = 2 * OFF_TABLE;
constexpr size_t X3_BYTES = X3_TABLE_OFF + 3 * (OFF_GV - OFF_TABLE);
static_assert(OFF_XW % 1024 == 0 && OFF_TABLE % 1024 == 0, "operand blocks must stay 1024-byte aligned");

}  // namespace col

// per-column table for planes [plane_lo, plane_lo + ncols / R1) -> ctx->col_table (query_col.cu)
// passes: 1 = fp16 operands (SURS_PREC_FP16), 3 = split hi/lo operands (SURS_PREC_FP16X3)
// point0 >= 0: point mode, row r = point point0 + r of io (R1 / plane_lo unused)
int surs_col_build_table(surs_ctx *ctx, const PointIO &io, int R1, int plane_lo, int64_t ncols, cudaStream_t st, int passes = 1, int64_t point0 = -1);
int surs_launch_query_generic_x3(surs_ctx *ctx, const PointIO &io, cudaStream_t st);
// col0: first grid column of the resident table (0 for the whole-grid table of the octree, plane_lo * R1 for a slab's)
// nmlp = 1: only the LR MLP runs and only the LR volume is written (vol32 scatter)
int surs_launch_query_col_indexed(surs_ctx *ctx, const PointIO &io, int R1, int R2, cudaStream_t st, int passes = 1, int64_t col0 = 0, int nmlp = 2);
int surs_launch_query_inc(surs_ctx *ctx, const PointIO &io, int R1, int R2, int plane_lo, int nplanes, cudaStream_t st);
