// CTC loss forward-backward for sm_100a: one CTA per sequence, the alpha and the beta recursion
// each run by ONE warp (warp-per-recursion, neighbour states exchanged with shuffles), log-space
// (base 2: ex2/lg2 MUFU), log-probabilities staged once in shared memory, gradient w.r.t. the
// LOGITS (softmax - posterior) written in a single coalesced pass.
//
// Replaces, for the reference call site model_v1/train.py:21-30 / model_v1/valid.py:32-38:
//   preds.float().permute(1,0,2).log_softmax(2) -> nn.CTCLoss(reduction='none', zero_infinity=True)
// (ATen _ctc_loss / _ctc_loss_backward; blank = 0).  Semantics: SURVEY.md 8a / 9.14-9.16.
#include "common.cuh"

namespace htrvt {

constexpr float kNeg = -1.0e30f;       // finite stand-in for log(0): keeps ex2(a-a) well defined
constexpr int kCtcThreads = 512;

// log2(2^a + 2^b + 2^c): the largest term contributes exactly 1, so only the other two go through the SFU
// (3 MUFU ops per state instead of 4 - the recursion is MUFU-issue bound: 4 SFU lanes per SM sub-partition)
__device__ __forceinline__ float lse3_log2(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  return m + lg2f(1.0f + ex2f(lo - m) + ex2f(mid - m));
}

// alpha recursion.  Lane owns states s = lane*K + j.  A is [T][32*K] (shared or global).
// Every kRenorm steps the running maximum is subtracted and accumulated in double (`off[t]` = offset of
// row t), so stored values stay O(10^1): fp32 log-space then resolves ~4e-6 instead of ~6e-5 at |nll|~10^3.
constexpr int kRenorm = 8;

template <int K>
__device__ __forceinline__ void ctc_alpha(const float* __restrict__ l2p, int ldp, const int* __restrict__ ext, int S,
                                          int Tb, float* A, double* off) {
  const int lane = threadIdx.x & 31;
  const int SP = 32 * K;
  float a[K], p[K];
  int lab[K];
  bool valid[K], skip[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = lane * K + j;
    valid[j] = s < S;
    lab[j] = valid[j] ? ext[s] : 0;
    skip[j] = valid[j] && (s & 1) && s >= 3 && ext[s] != ext[s - 2];
    a[j] = kNeg;
    if (s == 0) a[j] = l2p[0];
    if (s == 1 && S > 1) a[j] = l2p[lab[j]];
    A[s] = a[j];
  }
  double c = 0.0;
  if (lane == 0) off[0] = 0.0;
  if (Tb > 1) {
#pragma unroll
    for (int j = 0; j < K; ++j) p[j] = l2p[ldp + lab[j]];
  }
  for (int t = 1; t < Tb; ++t) {
    float up1 = __shfl_up_sync(0xffffffffu, a[K - 1], 1);
    float up2 = (K >= 2) ? __shfl_up_sync(0xffffffffu, a[K >= 2 ? K - 2 : 0], 1)
                         : __shfl_up_sync(0xffffffffu, a[0], 2);
    if (lane == 0) { up1 = kNeg; up2 = kNeg; }
    if (K == 1 && lane == 1) up2 = kNeg;
    float pn[K];
    if (t + 1 < Tb) {
#pragma unroll
      for (int j = 0; j < K; ++j) pn[j] = l2p[(t + 1) * ldp + lab[j]];
    }
    float n[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float p1 = (j >= 1) ? a[j >= 1 ? j - 1 : 0] : up1;
      const float p2 = (j >= 2) ? a[j >= 2 ? j - 2 : 0] : (j == 1 ? up1 : up2);
      const float v = lse3_log2(a[j], p1, skip[j] ? p2 : kNeg) + p[j];
      n[j] = valid[j] ? fmaxf(v, kNeg) : kNeg;
    }
    if ((t & (kRenorm - 1)) == 0) {
      float m = n[0];
#pragma unroll
      for (int j = 1; j < K; ++j) m = fmaxf(m, n[j]);
      m = warp_max(m);
      if (m > -1.0e29f) {
#pragma unroll
        for (int j = 0; j < K; ++j) n[j] = fmaxf(n[j] - m, kNeg);
        c += static_cast<double>(m);
      }
    }
    if (lane == 0) off[t] = c;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      a[j] = n[j];
      A[t * SP + lane * K + j] = n[j];
      p[j] = pn[j];
    }
  }
}

// beta recursion (backwards in time).  Bt is [T][32*K].
template <int K>
__device__ __forceinline__ void ctc_beta(const float* __restrict__ l2p, int ldp, const int* __restrict__ ext, int S,
                                         int Tb, float* Bt, double* off) {
  const int lane = threadIdx.x & 31;
  const int SP = 32 * K;
  float b[K], p[K];
  int lab[K];
  bool valid[K], skip[K];
  const float* row = l2p + (Tb - 1) * ldp;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = lane * K + j;
    valid[j] = s < S;
    lab[j] = valid[j] ? ext[s] : 0;
    skip[j] = valid[j] && (s & 1) && (s + 2 < S) && ext[s + 2] != ext[s];
    b[j] = kNeg;
    if (s == S - 1) b[j] = row[lab[j]];
    if (s == S - 2 && S > 1) b[j] = row[lab[j]];
    Bt[(Tb - 1) * SP + s] = b[j];
  }
  double c = 0.0;
  if (lane == 0) off[Tb - 1] = 0.0;
  if (Tb > 1) {
#pragma unroll
    for (int j = 0; j < K; ++j) p[j] = l2p[(Tb - 2) * ldp + lab[j]];
  }
  for (int t = Tb - 2; t >= 0; --t) {
    float dn1 = __shfl_down_sync(0xffffffffu, b[0], 1);
    float dn2 = (K >= 2) ? __shfl_down_sync(0xffffffffu, b[K >= 2 ? 1 : 0], 1)
                         : __shfl_down_sync(0xffffffffu, b[0], 2);
    if (lane == 31) { dn1 = kNeg; dn2 = kNeg; }
    if (K == 1 && lane == 30) dn2 = kNeg;
    float pn[K];
    if (t >= 1) {
#pragma unroll
      for (int j = 0; j < K; ++j) pn[j] = l2p[(t - 1) * ldp + lab[j]];
    }
    float n[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float p1 = (j + 1 < K) ? b[j + 1 < K ? j + 1 : 0] : dn1;
      const float p2 = (j + 2 < K) ? b[j + 2 < K ? j + 2 : 0] : (j + 2 == K ? dn1 : dn2);
      const float v = lse3_log2(b[j], p1, skip[j] ? p2 : kNeg) + p[j];
      n[j] = valid[j] ? fmaxf(v, kNeg) : kNeg;
    }
    if ((t & (kRenorm - 1)) == 0) {
      float m = n[0];
#pragma unroll
      for (int j = 1; j < K; ++j) m = fmaxf(m, n[j]);
      m = warp_max(m);
      if (m > -1.0e29f) {
#pragma unroll
        for (int j = 0; j < K; ++j) n[j] = fmaxf(n[j] - m, kNeg);
        c += static_cast<double>(m);
      }
    }
    if (lane == 0) off[t] = c;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      b[j] = n[j];
      Bt[t * SP + lane * K + j] = n[j];
      p[j] = pn[j];
    }
  }
}

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
  int B, T, C, kmax, is_logprob, scratch_in_smem;
};

#define CTC_DISPATCH(FN, ...)                 \
  switch (K) {                                \
    case 1: FN<1>(__VA_ARGS__); break;        \
    case 2: FN<2>(__VA_ARGS__); break;        \
    case 3: FN<3>(__VA_ARGS__); break;        \
    case 4: FN<4>(__VA_ARGS__); break;        \
    case 5: FN<5>(__VA_ARGS__); break;        \
    case 6: FN<6>(__VA_ARGS__); break;        \
    case 7: FN<7>(__VA_ARGS__); break;        \
    case 8: FN<8>(__VA_ARGS__); break;        \
    case 9: FN<9>(__VA_ARGS__); break;        \
    case 13: FN<13>(__VA_ARGS__); break;      \
    default: FN<17>(__VA_ARGS__); break;      \
  }

__host__ __device__ inline int ctc_round_k(int k) { return k <= 9 ? k : (k <= 13 ? 13 : 17); }

__global__ void __launch_bounds__(kCtcThreads, 1) ctc_loss_grad_kernel(const CtcParams P) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kCtcThreads / 32;
  const int T = P.T, C = P.C;
  const int ldp = C | 1;

  float* l2p = smem;                               // [T][ldp] log2-domain log-probabilities
  float* post = l2p + T * ldp;                     // [NW][ldp] per-warp posterior row
  int* ext = reinterpret_cast<int*>(post + NW * ldp);   // [32*kmax] extended label sequence
  float* red = reinterpret_cast<float*>(ext + 32 * P.kmax);   // [40] reductions / broadcast
  double* offA = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(red + 40) + 7) & ~uintptr_t(7));         // [T] renormalisation offsets of alpha rows
  double* offB = offA + T;                                    // [T] ... of beta rows
  float* AB = reinterpret_cast<float*>(offB + T);  // alpha | beta when they fit in shared memory

  // ---- per-sequence metadata -------------------------------------------------------------------
  int Tb = P.input_lengths ? P.input_lengths[b] : T;
  Tb = min(max(Tb, 0), T);
  const int L = P.target_lengths[b];
  int toff;
  if (P.tgt_stride > 0) {
    toff = b * P.tgt_stride;
  } else {
    float part = 0.f;                              // exact for sums < 2^24
    for (int i = tid; i < b; i += kCtcThreads) part += static_cast<float>(P.target_lengths[i]);
    toff = static_cast<int>(block_sum(part, red) + 0.5f);
  }
  const int S = 2 * L + 1;
  const int K = ctc_round_k((S + 31) / 32);
  const int SP = 32 * K;
  const bool fits = (L <= Tb) && (K <= P.kmax) && L >= 0;   // L > Tb can never be aligned
  float* A = P.scratch_in_smem ? AB : P.scratch + static_cast<size_t>(b) * 2 * T * 32 * P.kmax;
  float* Bt = A + static_cast<size_t>(T) * SP;

  const float* xb = P.x + static_cast<long long>(b) * P.x_sb;
  float* gb = P.grad ? P.grad + static_cast<long long>(b) * P.g_sb : nullptr;

  if (fits) {
    for (int s = tid; s < SP; s += kCtcThreads) ext[s] = (s < S && (s & 1)) ? P.targets[toff + (s >> 1)] : 0;
  }
  // ---- phase 0: stage log-softmax rows (log2 domain) ------------------------------------------
  for (int t = warp; t < Tb; t += NW) {
    const float* xr = xb + static_cast<long long>(t) * P.x_st;
    float v[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? xr[c] : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
    for (int c = lane + 256; c < C; c += 32) mx = fmaxf(mx, xr[c]);
    float lse2 = 0.f;
    if (!P.is_logprob) {
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += ex2f((v[i] - mx) * kLog2e);
      for (int c = lane + 256; c < C; c += 32) sum += ex2f((xr[c] - mx) * kLog2e);
      sum = warp_sum(sum);
      lse2 = mx * kLog2e + lg2f(sum);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < C) l2p[t * ldp + c] = v[i] * kLog2e - lse2;
    }
    for (int c = lane + 256; c < C; c += 32) l2p[t * ldp + c] = xr[c] * kLog2e - lse2;
  }
  __syncthreads();

  // ---- phase 1: alpha on warp 0, beta on warp 1, concurrently ---------------------------------
  const bool run = fits && Tb > 0;
  if (run) {
    if (warp == 0) {
      CTC_DISPATCH(ctc_alpha, l2p, ldp, ext, S, Tb, A, offA)
    } else if (warp == 1) {
      CTC_DISPATCH(ctc_beta, l2p, ldp, ext, S, Tb, Bt, offB)
    }
  }
  if (!P.scratch_in_smem) __threadfence_block();
  __syncthreads();

  float ll2r = kNeg;                               // log2 likelihood relative to offA[Tb-1]
  double ll2 = 0.0;
  if (run) {
    const float a1 = A[(Tb - 1) * SP + S - 1];
    const float a2 = S > 1 ? A[(Tb - 1) * SP + S - 2] : kNeg;
    const float m = fmaxf(a1, a2);
    ll2r = m + lg2f(ex2f(a1 - m) + ex2f(a2 - m));
    ll2 = static_cast<double>(ll2r) + offA[Tb - 1];
  } else if (Tb == 0 && L == 0) {
    ll2r = 0.f;
  }
  const bool feasible = ll2r > -1.0e29f;
  if (tid == 0) {
    float out = feasible ? static_cast<float>(-ll2 * 0.6931471805599453) : 0.f;      // zero_infinity=True
    if (L > Tb) out = 0.f;
    else if (K > P.kmax) out = __int_as_float(0x7fc00000);   // provisioning error: be loud
    P.nll[b] = out;
  }
  if (!gb) return;

  // ---- phase 2: posterior collect + gradient rows ----------------------------------------------
  const float gs = P.grad_scale ? P.grad_scale[b] : P.grad_scale_const;
  float* pw = post + warp * ldp;
  for (int t = warp; t < T; t += NW) {
    float* gr = gb + static_cast<long long>(t) * P.g_st;
    if (t >= Tb || !feasible) {
      for (int c = lane; c < C; c += 32) gr[c] = 0.f;
      continue;
    }
    for (int c = lane; c < C; c += 32) pw[c] = 0.f;
    __syncwarp();
    const float* lr = l2p + t * ldp;
    float blank = 0.f;
    const float* ar = A + t * SP;
    const float* br = Bt + t * SP;
    const float rowc = static_cast<float>(offA[t] + offB[t] - ll2);
    for (int s = lane; s < S; s += 32) {
      const int c = ext[s];
      const float v = ex2f((ar[s] + br[s] - lr[c]) + rowc);
      if (s & 1) atomicAdd(&pw[c], v);
      else blank += v;
    }
    blank = warp_sum(blank);
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
      const float pc = (c == 0) ? blank : pw[c];
      gr[c] = (ex2f(lr[c]) - pc) * gs;
    }
    __syncwarp();
  }
}

// nll only (validation: model_v1/valid.py:36-38 needs no gradient) reuses the same kernel with grad == null.

size_t ctc_smem_bytes(int T, int C, int kmax, bool scratch_in_smem) {
  const int ldp = C | 1;
  size_t words = static_cast<size_t>(T) * ldp + (kCtcThreads / 32) * ldp + 32 * kmax + 40 + 4 * T + 2;
  if (scratch_in_smem) words += static_cast<size_t>(2) * T * 32 * kmax;
  return words * 4;
}

}  // namespace htrvt

using namespace htrvt;

extern "C" size_t htrvt_ctc_workspace_bytes(int B, int T, int C, int max_target_len) {
  int lmax = max_target_len < 0 || max_target_len > T ? T : max_target_len;
  const int kmax = ctc_round_k((2 * lmax + 1 + 31) / 32);
  if (ctc_smem_bytes(T, C, kmax, true) <= 227 * 1024) return 0;
  return static_cast<size_t>(B) * 2 * T * 32 * kmax * sizeof(float);
}

extern "C" int htrvt_ctc_loss_grad(const float* x, long long x_stride_b, long long x_stride_t, int is_logprob,
                                   const int* targets, int tgt_stride, const int* input_lengths,
                                   const int* target_lengths, int B, int T, int C, int max_target_len,
                                   float* nll, float* grad, long long g_stride_b, long long g_stride_t,
                                   const float* grad_scale, float grad_scale_const, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream) {
  if (B <= 0 || T <= 0 || C <= 1 || !x || !nll || !target_lengths) return HTRVT_ERR_SHAPE;   // targets may be null when every label is empty
  int lmax = max_target_len < 0 || max_target_len > T ? T : max_target_len;
  CtcParams P;
  P.x = x; P.x_sb = x_stride_b; P.x_st = x_stride_t;
  P.grad = grad; P.g_sb = g_stride_b; P.g_st = g_stride_t;
  P.targets = targets; P.tgt_stride = tgt_stride;
  P.input_lengths = input_lengths; P.target_lengths = target_lengths;
  P.nll = nll; P.grad_scale = grad_scale; P.grad_scale_const = grad_scale_const;
  P.B = B; P.T = T; P.C = C; P.is_logprob = is_logprob;
  P.kmax = ctc_round_k((2 * lmax + 1 + 31) / 32);
  P.scratch_in_smem = ctc_smem_bytes(T, C, P.kmax, true) <= 227 * 1024 ? 1 : 0;
  P.scratch = static_cast<float*>(workspace);
  const size_t smem = ctc_smem_bytes(T, C, P.kmax, P.scratch_in_smem);
  if (smem > 227 * 1024) return HTRVT_ERR_SHAPE;     // T*C slab itself does not fit
  if (!P.scratch_in_smem) {
    const size_t need = static_cast<size_t>(B) * 2 * T * 32 * P.kmax * sizeof(float);
    if (!workspace || workspace_bytes < need) return HTRVT_ERR_WORKSPACE;
  }
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(ctc_loss_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return HTRVT_ERR_LAUNCH;
    configured = 227 * 1024;
  }
  ctc_loss_grad_kernel<<<B, kCtcThreads, smem, stream>>>(P);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
