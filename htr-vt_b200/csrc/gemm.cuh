// Parameters of the sm_100a "tap GEMM" (gemm.cu): one persistent, warp-specialised tcgen05 kernel
// that serves every dense contraction of the encoder: linear fwd/dgrad/wgrad and the 3x3 / 1x1
// stem convolutions (implicit GEMM over shifted TMA windows of the NHWC activation) fwd/dgrad/wgrad.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace htrvt {

struct TapTab {
  int8_t dh[9], dw[9], pw[9], widx[9];
};

enum : int {
  EPI_BF16 = 1 << 0,       // store bf16 (else fp32)
  EPI_BIAS = 1 << 1,       // + bias[n]
  EPI_ACCUM = 1 << 4,      // out += (TMA reduce-add store)
  EPI_STATS = 1 << 5,      // column sum / sum of squares of the stored tile -> stats[cta*4+quadrant][2][N] (+=)
  EPI_RELU = 1 << 7,       // max(x, 0)
  EPI_NOSTORE = 1 << 8,    // measurement aid: skip the epilogue body (main-loop-only timing)
  EPI_RES = 1 << 9,        // kind 0, 16-bit out: + res[n,h,w,col] (same NHWC geometry as the output) before the ReLU
  EPI_F16 = 1 << 10,       // the 16-bit output (and the EPI_RES residual) is IEEE fp16, not bf16 (forward stem tensors)
  EPI_GELU = 1 << 12,      // bf16 out: out = gelu(acc * alpha + bias) (erf form) - timm Mlp fc1 -> act
  EPI_DUAL = 1 << 13,      // with EPI_GELU (kernel built with DUAL): the pre-activation goes to a second output (tmC2)
  EPI_GELU_BWD = 1 << 14,  // bf16 out, kind 0: out = acc * gelu'(res[row, col]) (res = the saved pre-activation u)
  EPI_COLSUM = 1 << 15,    // with EPI_GELU_BWD: stats[N] += column sums of the stored tile (bias gradient)
  EPI_BN_BWD = 1 << 16,    // bf16 out, kind 0 (input gradient of a 3x3 conv): out = g' = acc * relu_mask, and the BatchNorm
                           // backward reduction of the layer in front rides along: stats[0][N] += sum g', stats[1][N] +=
                           // sum g' * (res - bn_mean) * bn_rstd  (res = that layer's raw fp16 conv output)
};

struct GemmP {
  int kind;                // 0: rows = output pixels, K = taps x channels; 1 / 2 / 3: wgrad (K = pixels)
  int Wo, Ho, NB;          // output pixel grid (kind 0) / dY pixel grid (kind 1)
  int tiles_per_row;       // kind 0: ceil(Wo / 128)
  int tiles_m, tiles_n, n_taps, splits;
  int k_chunks;            // kind 0: 64-channel chunks per tap; kind 1: 64-pixel chunks per row
  int a_sh;                // input-row multiplier (vertical stride) for the shifted operand
  int b_tap_stride;        // kind 0: K (K-major B) or N (MN-major B) elements per tap in the weight matrix
  TapTab tap;
  int M_valid, N_valid;    // kind 1: rows (Cout) valid; columns valid
  int a_taps, a_atoms_per_tap;   // KIND 3: the M axis is (tap, 64-channel atom of x): real tap count, Cin / 64
  int flags;
  int a_f16, b_f16;        // operand element formats of the MMA: 0 = bf16, 1 = fp16 (must be equal: a mixed descriptor
                           // is an illegal instruction on B200)
  const float* bias;
  const void* res;         // EPI_RES: bf16 residual, laid out like the output
  float* stats;
  const uint8_t* mask;     // EPI_BN_BWD: ReLU mask bits of the output pixels, 1 byte per 8 channels
  const float* bn_mean;    // EPI_BN_BWD: batch mean / 1 / sqrt(var + eps) per channel
  const float* bn_rstd;
  float alpha;             // scale applied to the accumulator
};

}  // namespace htrvt
