// Shared pieces of the CTC kernels (ctc.cu: one CTA per sequence, latency-optimised; ctc_tput.cu: one warp per
// sequence, throughput-optimised): parameter block and the "packed double" helpers.
#pragma once
#include "common.cuh"

namespace htrvt {

constexpr int kFastRenorm = 4;                 // rows between power-of-two renormalisations of the linear-domain recursion
constexpr int kFastTargetExp = 1023 - 16;      // the row maximum is moved to 2^-16

__device__ __forceinline__ double unpack_pd(uint32_t w) { return __hiloint2double(static_cast<int>(w >> 2), static_cast<int>(w << 30)); }
__device__ __forceinline__ uint32_t pack_pd(double v) {
  const uint32_t hi = static_cast<uint32_t>(__double2hiint(v)), lo = static_cast<uint32_t>(__double2loint(v));
  return __funnelshift_l(lo, hi, 2);        // bits 61..30 (values are in [0, 2): sign and exponent MSB are 0)
}
__device__ __forceinline__ double shfl_up_d(double v, int d) {
  return __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(v), d), __shfl_up_sync(0xffffffffu, __double2loint(v), d));
}
__device__ __forceinline__ double shfl_down_d(double v, int d) {
  return __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(v), d), __shfl_down_sync(0xffffffffu, __double2loint(v), d));
}
// exact 2^-sh as a double (|sh| < 1000)
__device__ __forceinline__ double pow2_neg(int sh) { return __hiloint2double((1023 - sh) << 20, 0); }


struct CtcParams {
  const float* x;            // logits (or log-probs when is_logprob) [.., C] with strides below
  long long x_sb, x_st;      // element strides of the batch and time axes (class axis contiguous)
  float* grad;               // same logical shape as x, own strides; may be null (loss only)
  long long g_sb, g_st;
  const int* targets;        // concatenated (tgt_stride == 0) or padded [B, tgt_stride]
  int tgt_stride;
  const int* input_lengths;  // [B] or null (=> T)
  const int* target_lengths; // [B]
  float* nll;                // [B]
  const float* grad_scale;   // [B] per-sample upstream gradient, or null
  float grad_scale_const;    // used when grad_scale == null
  float* scratch;            // global alpha/beta scratch when they do not fit in shared memory
  int B, T, C, kmax, is_logprob, scratch_in_smem, force_slow, dbg, ovl;
  const int* only;           // nullable: run only sequences with only[b] != 0 (fix-up pass after the throughput kernel)
};


// throughput kernel (ctc_tput.cu): handles every sequence it can on the linear fp64 path and sets flags[b] = 1 for the
// rest (label too long for its register budget, or the fast path's consistency guard tripped)
int ctc_tput_launch(const CtcParams& P, int* flags, int* offsets, cudaStream_t stream);
size_t ctc_tput_smem_bytes(int T, int C, int kmax);

}  // namespace htrvt
