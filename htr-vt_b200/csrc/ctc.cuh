// Shared pieces of the CTC kernels (ctc.cu: one CTA per sequence, latency-optimised; ctc_grp.cu: a group of lanes per
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


// exclusive prefix sum of the label lengths -> start of every sequence's labels in the concatenated target stream
int ctc_offsets_launch(const int* lengths, int B, int* offsets, cudaStream_t stream);

// lane-group throughput kernel (ctc_grp.cu): G lanes per sequence, fp32 linear domain, alpha rows in a per-warp global
// scratch.  Handles every sequence it can and sets flags[b] = 1 for the rest (label longer than 8 G, or the consistency
// guard tripped): those are redone by ctc.cu's kernel in a fix-up launch.  lmax = longest label of the batch (<= 256)
bool ctc_grp_supported(int B, int T, int C, int lmax);
size_t ctc_grp_scratch_bytes(int B, int T, int C, int lmax, int sms);
int ctc_grp_launch(const CtcParams& P, int lmax, int* flags, const int* offsets, void* scratch, int sms,
                   cudaStream_t stream);
int ctc_num_sms();

}  // namespace htrvt
