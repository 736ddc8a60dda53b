// CTC loss forward-backward, THROUGHPUT variant for large batches (B >> number of SMs): one WARP per sequence,
// four sequences per CTA, ~10 KB of shared memory per sequence so that a dozen or more sequences are resident per SM
// and the serial alpha / beta chains of different sequences fill each other's latency.  (ctc.cu gives one CTA - one
// SM - to a sequence: the right shape for B <= 148, but at B = 4096 each SM then walks 28 sequences one after another
// with two busy warps.)
//
//   state     linear domain on the FP64 pipe with exact power-of-two renormalisation, as in ctc.cu's fast path;
//   storage   alpha rows are NOT kept: every 8th row is checkpointed (packed 32-bit words) in shared memory; the
//             backward sweep recomputes the 8 rows of a block from its checkpoint into REGISTERS and consumes them
//             immediately - beta needs no storage at all, the posterior gamma_t(s) = alpha_t(s) beta'_t(s) / Z is formed
//             lane-locally (alpha and beta use the same state -> lane mapping) and scattered to per-class bins;
//   inputs    the logits are staged 8 rows at a time, one block ahead of the sweep, by 16-byte cp.async copies into a
//             two-block ring (5 KB); the log-sum-exp of a row is formed when its block lands (alpha sweep), and each
//             step gathers its K logits per lane from the ring: p = 2^(x log2e - lse2_t), one MUFU per state;
//   guard     Z from the alpha sweep must be positive for a feasible label and every row of posteriors must sum to 1
//             (2e-5): otherwise the sequence is FLAGGED and recomputed by ctc.cu's kernel (which owns the log-space
//             path) in a fix-up launch that exits immediately for unflagged sequences.
// Same call-site semantics as ctc.cu (model_v1/train.py:21-30: log_softmax + nn.CTCLoss(reduction='none',
// zero_infinity=True), blank = 0); gradient w.r.t. the logits (or log-probs) = (softmax - posterior) * grad_scale.
#include "ctc.cuh"

namespace htrvt {

constexpr int kTpWarps = 4;                    // sequences per CTA
constexpr int kTpR = 8;                        // checkpoint interval = rows recomputed per block
constexpr int kTpMaxK = 6;                     // states per lane: S = 2L + 1 <= 32 * 6 (L <= 95)

// shift (in binades) that moves the row maximum to 2^-16; 0 for an all-zero row
template <int K>
__device__ __forceinline__ int tp_row_shift(const double (&n)[K]) {
  uint32_t m = static_cast<uint32_t>(__double2hiint(n[0]));
#pragma unroll
  for (int j = 1; j < K; ++j) m = max(m, static_cast<uint32_t>(__double2hiint(n[j])));
  m = __reduce_max_sync(0xffffffffu, m);
  return m == 0u ? 0 : static_cast<int>(m >> 20) - kFastTargetExp;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1)
    v += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), d),
                          __shfl_xor_sync(0xffffffffu, __double2loint(v), d));
  return v;
}

__device__ long long g_tp_stamps[8];                   // HTRVT_CTC_DEBUG: clock64 at the phase boundaries of sequence 0
#define TP_STAMP(i) do { if (q.dbg && lane == 0) g_tp_stamps[i] = clock64(); } while (0)

struct TpSeq {
  const float* x;            // this sequence's logits / log-probs, row stride x_st
  long long x_st;
  float* g;                  // gradient rows (nullable), row stride g_st
  long long g_st;
  float* lse2;               // smem [T]: log2-domain log-sum-exp per row (0 for log-prob input), filled by the alpha sweep
  uint32_t* ck;              // smem [(T / R) + 1][32 * K]: packed alpha checkpoints (rows 0, R, 2R, ...)
  int* ckoff;                // smem [(T / R) + 1]: their exponent offsets
  uint32_t* bins;            // smem [C]: posterior per class, fixed point 2^30
  float* ring;               // smem [2][R][ldr]: two blocks of R staged logits rows (cp.async, one block ahead)
  int ldr;
  int Tb, S, L, C, vec_ok, is_logprob, dbg;
  float gs;
};

__device__ __forceinline__ void tp_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tp_cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// stage rows [kb*R, min(kb*R + R, Tb)) into ring half `half` (asynchronously; ONE commit group, possibly empty)
__device__ __forceinline__ void tp_stage(const TpSeq& q, int kb, int half) {
  const int lane = threadIdx.x & 31;
  if (kb >= 0 && kb * kTpR < q.Tb) {
    const int t0 = kb * kTpR, nrow = min(kTpR, q.Tb - t0);
    float* dst = q.ring + half * kTpR * q.ldr;
    // row by row (no integer division on this path): lanes cover a row's 16-byte chunks / elements
    const float* src = q.x + static_cast<long long>(t0) * q.x_st;
    uint32_t d = smem_u32(dst);
    if (q.vec_ok) {
      const int c4 = q.C >> 2;
      for (int r = 0; r < nrow; ++r, src += q.x_st, d += q.ldr * 4)
        for (int v = lane; v < c4; v += 32) tp_cp16(d + 16 * v, src + 4 * v);
    } else {
      for (int r = 0; r < nrow; ++r, src += q.x_st, d += q.ldr * 4)
        for (int c = lane; c < q.C; c += 32) tp_cp4(d + 4 * c, src + c);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tp_wait_block() {      // everything but the newest group has landed
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncwarp();
}

// One alpha step for the lane's K states (natural order s = lane*K + j, K even so parity(s) == parity(j)).
//   n[j] = (a[j] + a[s-1] + skip * a[s-2]) * pd[j]
template <int K>
__device__ __forceinline__ void tp_alpha_step(const double (&a)[K], const double (&msk)[K / 2], const double (&pd)[K],
                                              double (&n)[K]) {
  const int lane = threadIdx.x & 31;
  double up1 = shfl_up_d(a[K - 1], 1);               // previous lane's last state = s - 1 of j = 0 and s - 2 of j = 1
  if (lane == 0) up1 = 0.0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double p1 = (j >= 1) ? a[j >= 1 ? j - 1 : 0] : up1;
    double su;
    if (j & 1) {
      const double p2 = (j >= 2) ? a[j >= 2 ? j - 2 : 0] : up1;
      su = fma(msk[j >> 1], p2, a[j] + p1);
    } else {
      su = a[j] + p1;
    }
    n[j] = su * pd[j];
  }
}

// p_t(l_s) * 2^-sh for the lane's K states from a staged row (the power-of-two renormalisation of the PREVIOUS row
// rides on this row's probabilities)
template <int K>
__device__ __forceinline__ void tp_probs(const float* row, const int (&lab)[K], float lse2, const bool (&valid)[K], int sh,
                                         double (&pd)[K]) {
  // the fp32 probability is the SAME value in the alpha sweep, the recomputation and the beta sweep (the posterior is
  // a ratio in which its rounding cancels only then); the renormalisation is an exact power of two applied in fp64
#pragma unroll
  for (int j = 0; j < K; ++j) pd[j] = valid[j] ? static_cast<double>(ex2f(fmaf(row[lab[j]], kLog2e, -lse2))) : 0.0;
  if (sh != 0) {                                       // warp-uniform
    const double sc = pow2_neg(sh);
#pragma unroll
    for (int j = 0; j < K; ++j) pd[j] *= sc;
  }
}

// log2-domain log-sum-exp of the nrow staged rows of a block: four lanes per row
__device__ __forceinline__ void tp_block_lse(const TpSeq& q, const float* blk, int t0, int nrow) {
  const int lane = threadIdx.x & 31, r = lane >> 2, part = lane & 3;
  float l2 = 0.f;
  if (!q.is_logprob) {
    const float* row = blk + min(r, nrow - 1) * q.ldr;
    float mx = -INFINITY;
    for (int c = part; c < q.C; c += 4) mx = fmaxf(mx, row[c]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float mxl = mx * kLog2e;
    float sum = 0.f;
    for (int c = part; c < q.C; c += 4) sum += ex2f(fmaf(row[c], kLog2e, -mxl));
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    l2 = mxl + lg2f(sum);
  }
  if (part == 0 && r < nrow) q.lse2[t0 + r] = l2;
  __syncwarp();
}

// The whole sequence on one warp.  Returns false if the sequence must be handed to the fallback kernel.
template <int K>
__device__ bool tp_sequence(const TpSeq& q, const int* __restrict__ labels, float* nll_out) {
  static_assert((K & 1) == 0 && K <= kTpMaxK, "even K");
  const int lane = threadIdx.x & 31;
  const int SP = 32 * K;
  const int Tb = q.Tb, S = q.S, C = q.C;
  int lab[K];
  bool valid[K];
  double mskA[K / 2], mskB[K / 2];                    // skip-transition masks of the label states (alpha / beta direction)
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = lane * K + j;
    valid[j] = s < S;
    lab[j] = (valid[j] && (s & 1)) ? labels[s >> 1] : 0;
  }
#pragma unroll
  for (int j = 1; j < K; j += 2) {
    const int s = lane * K + j;
    const int prev = (s >= 3 && s < S) ? labels[(s >> 1) - 1] : -1;
    const int next = (s + 2 < S) ? labels[(s >> 1) + 1] : -1;
    mskA[j >> 1] = (valid[j] && s >= 3 && lab[j] != prev) ? 1.0 : 0.0;
    mskB[j >> 1] = (valid[j] && s + 2 < S && lab[j] != next) ? 1.0 : 0.0;
  }
  const int nblk = (Tb + kTpR - 1) / kTpR;
  TP_STAMP(1);

  // ------------------------------- alpha sweep: checkpoints every R rows, likelihood ---------------------------
  double a[K];
  int c = 0;                                          // true alpha = a * 2^c
  int sh = 0;
  tp_stage(q, 0, 0);
  tp_stage(q, 1, 1);
  for (int kb = 0; kb < nblk; ++kb) {
    const int t0 = kb * kTpR, nrow = min(kTpR, Tb - t0);
    const float* blk = q.ring + (kb & 1) * kTpR * q.ldr;
    tp_wait_block();
    tp_block_lse(q, blk, t0, nrow);
    for (int i = 0; i < nrow; ++i) {
      const int t = t0 + i;
      double pd[K];
      tp_probs<K>(blk + i * q.ldr, lab, q.lse2[t], valid, sh, pd);
      c += sh;
      if (t == 0) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int s = lane * K + j;
          a[j] = (s == 0 || (s == 1 && S > 1)) ? pd[j] : 0.0;
        }
      } else {
        double n[K];
        tp_alpha_step<K>(a, mskA, pd, n);
#pragma unroll
        for (int j = 0; j < K; ++j) a[j] = n[j];
      }
      sh = ((t & (kFastRenorm - 1)) == 0) ? tp_row_shift<K>(a) : 0;
      if (i == 0) {                                   // t0 is a checkpoint row
        uint32_t* row = q.ck + kb * SP + lane * K;
#pragma unroll
        for (int j = 0; j < K; ++j) row[j] = pack_pd(a[j]);
        if (lane == 0) q.ckoff[kb] = c;
      }
    }
    __syncwarp();                                     // every lane is done with this half of the ring
    tp_stage(q, kb + 2, kb & 1);
  }
  TP_STAMP(2);
  // start staging the last two blocks for the backward sweep while the likelihood is formed
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  tp_stage(q, nblk - 1, 0);
  tp_stage(q, nblk - 2, 1);
  // Z = alpha_{Tb-1}(S-1) + alpha_{Tb-1}(S-2)
  double zl = 0.0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = lane * K + j;
    if (s == S - 1 || (s == S - 2 && S > 1)) zl += a[j];
  }
  const double Zm = warp_sum_d(zl);
  const int Zoff = c;
  // feasibility by counting (a blank is needed between equal neighbours): L + repeats <= Tb
  int rep = 0;
  for (int i = lane + 1; i < q.L; i += 32) rep += labels[i] == labels[i - 1];
  rep = __reduce_add_sync(0xffffffffu, rep);
  const bool feasible = q.L + rep <= Tb && Tb > 0;
  bool ok = true;
  if (feasible && !(Zm > 0.0)) ok = false;            // mass flushed below the fp64 range: log space decides
  if (!feasible) {
    if (lane == 0) *nll_out = 0.f;                     // zero_infinity=True
    if (q.g)
      for (int t = 0; t < Tb; ++t)
        for (int cc = lane; cc < C; cc += 32) q.g[static_cast<long long>(t) * q.g_st + cc] = 0.f;
  } else if (ok) {
    // log2 Z = exponent + lg2(mantissa), from the packed word (22-bit mantissa: abs err ~2e-7 of a bit)
    const uint32_t w = pack_pd(Zm);
    const double ll2 = static_cast<double>(static_cast<int>(w >> 22) - 1023 + Zoff) +
                       static_cast<double>(lg2f(__uint_as_float(0x3F800000u | ((w & 0x3FFFFFu) << 1))));
    if (lane == 0) *nll_out = static_cast<float>(-ll2 * 0.6931471805599453);
  }
  if (!q.g || !feasible || !ok) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    return ok;
  }

  // ------------------------------- backward sweep: recompute blocks of R alpha rows, beta + posterior ----------
  const double rz = 1.0 / Zm;
  double bf[K];                                       // B_{t+1}(s): beta of the following row INCLUDING its emission
  int cb = 0;                                         // true B = bf * 2^cb
  int shb = 0;                                        // pending renormalisation of bf (folded into the next product)
#pragma unroll
  for (int j = 0; j < K; ++j) bf[j] = 0.0;
  for (int kb = nblk - 1, nb = 0; kb >= 0; --kb, ++nb) {
    const int t0 = kb * kTpR;
    const int nrow = min(kTpR, Tb - t0);
    const float* blk = q.ring + (nb & 1) * kTpR * q.ldr;
    tp_wait_block();
    // ---- recompute alpha rows t0 .. t0 + nrow - 1 into registers ---------------------------------------------
    double ar[kTpR][K];
    int oa[kTpR];
    {
      const uint32_t* row = q.ck + kb * SP + lane * K;
#pragma unroll
      for (int j = 0; j < K; ++j) ar[0][j] = unpack_pd(row[j]);
      oa[0] = q.ckoff[kb];
      int shr = tp_row_shift<K>(ar[0]);
#pragma unroll
      for (int i = 1; i < kTpR; ++i) {
        if (i < nrow) {
          double pd[K];
          tp_probs<K>(blk + i * q.ldr, lab, q.lse2[t0 + i], valid, shr, pd);
          oa[i] = oa[i - 1] + shr;
          tp_alpha_step<K>(ar[i - 1], mskA, pd, ar[i]);
          shr = (((t0 + i) & (kFastRenorm - 1)) == 0) ? tp_row_shift<K>(ar[i]) : 0;
        } else {
          oa[i] = oa[i - 1];
#pragma unroll
          for (int j = 0; j < K; ++j) ar[i][j] = 0.0;
        }
      }
    }
    // ---- beta steps t = t0 + nrow - 1 .. t0, posterior and gradient row of each -------------------------------
#pragma unroll
    for (int i = kTpR - 1; i >= 0; --i) {
      if (i >= nrow) continue;
      const int t = t0 + i;
      const float* xrow = blk + i * q.ldr;
      // beta'_t(s) = B_{t+1}(s) + B_{t+1}(s+1) + skip * B_{t+1}(s+2)   (boundary row: 1 for the two end states)
      double su[K];
      if (t == Tb - 1) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int s = lane * K + j;
          su[j] = (s == S - 1 || (s == S - 2 && S > 1)) ? 1.0 : 0.0;
        }
      } else {
        double dn1 = shfl_down_d(bf[0], 1), dn2 = shfl_down_d(bf[1], 1);
        if (lane == 31) { dn1 = 0.0; dn2 = 0.0; }
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const double p1 = (j + 1 < K) ? bf[j + 1 < K ? j + 1 : 0] : dn1;
          if (j & 1) {
            const double p2 = (j + 2 < K) ? bf[j + 2 < K ? j + 2 : 0] : dn2;
            su[j] = fma(mskB[j >> 1], p2, bf[j] + p1);
          } else {
            su[j] = bf[j] + p1;
          }
        }
      }
      // B_t(s) = beta'_t(s) * p_t(l_s) for the next (earlier) row, with the pending renormalisation folded in: issued
      // first so that the chain of the recursion does not wait for the posterior / gradient work below
      const int cb_row = cb;                           // scale of su
      {
        double pd[K];
        tp_probs<K>(xrow, lab, q.lse2[t], valid, shb, pd);
        cb += shb;
#pragma unroll
        for (int j = 0; j < K; ++j) bf[j] = su[j] * pd[j];
        shb = ((t & (kFastRenorm - 1)) == 0) ? tp_row_shift<K>(bf) : 0;
      }
      // posterior: gamma * 2^30 = (alpha * crow) * beta' with crow = 2^(30 + oa + cb - Zoff) / Zm; + 2^52 leaves the
      // rounded integer in the low word (ctc.cu's trick)
      int kk = oa[i] + cb_row - Zoff + 30;
      if (kk > 1000 || kk < -1000) { ok = false; kk = 0; }
      const double crow = rz * __hiloint2double((1023 + kk) << 20, 0);
      uint32_t blank = 0u, tot = 0u;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const double qv = (ar[i][j] * crow) * su[j] + 4503599627370496.0;
        const uint32_t u = valid[j] ? static_cast<uint32_t>(__double2loint(qv)) : 0u;
        tot += u;
        if (j & 1) { if (u) atomicAdd(q.bins + lab[j], u); }     // (most states carry no posterior mass: u == 0)
        else blank += u;
      }
      blank = __reduce_add_sync(0xffffffffu, blank);
      tot = __reduce_add_sync(0xffffffffu, tot);
      const int dev1 = static_cast<int>(tot) - (1 << 30);
      if (dev1 > 21475 || dev1 < -21475) ok = false;    // the row's posteriors do not sum to 1 (2e-5)
      __syncwarp();
      // gradient row: (softmax - posterior) * gs
      {
        const float l2 = q.lse2[t];
        float* grow = q.g + static_cast<long long>(t) * q.g_st;
        for (int cc = lane; cc < C; cc += 32) {
          const uint32_t pc = (cc == 0) ? blank : q.bins[cc];
          q.bins[cc] = 0u;
          grow[cc] = (ex2f(fmaf(xrow[cc], kLog2e, -l2)) - static_cast<float>(pc) * 9.313225746154785e-10f) * q.gs;
        }
      }
      __syncwarp();
    }
    tp_stage(q, kb - 2, nb & 1);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  TP_STAMP(3);
  // cross-check: the beta sweep's likelihood (start states of row 0) must equal the alpha sweep's
  {
    double zb = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int s = lane * K + j;
      if (s == 0 || (s == 1 && S > 1)) zb += bf[j];
    }
    zb = warp_sum_d(zb);
    const uint32_t wa = pack_pd(Zm), wb = pack_pd(zb);
    const double la = static_cast<double>(static_cast<int>(wa >> 22) + Zoff) +
                      static_cast<double>(lg2f(__uint_as_float(0x3F800000u | ((wa & 0x3FFFFFu) << 1))));
    const double lb = static_cast<double>(static_cast<int>(wb >> 22) + cb) +
                      static_cast<double>(lg2f(__uint_as_float(0x3F800000u | ((wb & 0x3FFFFFu) << 1))));
    if (!(zb > 0.0) || fabs(la - lb) > 3.0e-5) ok = false;
  }
  return ok;
}

// exclusive prefix sum of the label lengths -> start of every sequence's labels in the concatenated target stream
__global__ void __launch_bounds__(1024) tp_offsets_kernel(const int* __restrict__ lengths, int B, int* __restrict__ offs) {
  __shared__ int wsum[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < B ? lengths[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += n;
      }
      wsum[lane] = wi - w;                            // exclusive offset of each warp
    }
    __syncthreads();
    if (i < B) offs[i] = carry + wsum[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += wsum[31] + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kTpWarps * 32, 3) ctc_tput_kernel(const CtcParams P, int* __restrict__ flags,
                                                                   const int* __restrict__ offsets) {
  extern __shared__ __align__(16) unsigned char tp_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kTpWarps + warp;
  if (b >= P.B) return;                               // (whole warps: no CTA-wide barrier below)
  const int T = P.T, C = P.C;
  const int nck = T / kTpR + 1;
  const int SPmax = 32 * P.kmax;
  const int ldr = (C + 3) & ~3;
  // per-warp carve-up (ring first: 16-byte aligned rows)
  const size_t per_warp = (static_cast<size_t>(2 * kTpR) * ldr + static_cast<size_t>(nck) * SPmax + nck + T + C +
                           SPmax / 2 + 8) * 4;
  unsigned char* base = tp_smem + warp * ((per_warp + 15) & ~size_t(15));
  float* ring = reinterpret_cast<float*>(base);
  uint32_t* ck = reinterpret_cast<uint32_t*>(ring + 2 * kTpR * ldr);
  int* ckoff = reinterpret_cast<int*>(ck + static_cast<size_t>(nck) * SPmax);
  float* lse2 = reinterpret_cast<float*>(ckoff + nck);
  uint32_t* bins = reinterpret_cast<uint32_t*>(lse2 + T);
  int* labels = reinterpret_cast<int*>(bins + C);     // [SPmax / 2] label ids

  int Tb = P.input_lengths ? P.input_lengths[b] : T;
  Tb = min(max(Tb, 0), T);
  const int L = P.target_lengths[b];
  const int toff = P.tgt_stride > 0 ? b * P.tgt_stride : offsets[b];
  const int S = 2 * L + 1;
  const int Kf = ((S + 31) / 32 + 1) & ~1;
  float* gb = P.grad ? P.grad + static_cast<long long>(b) * P.g_sb : nullptr;
  if (L < 0 || Kf > P.kmax || Kf > kTpMaxK) {          // not provisioned here: the CTA-per-sequence kernel takes it
    if (lane == 0) flags[b] = 1;
    return;
  }
  if (lane == 0) flags[b] = 0;
  // rows beyond the input length carry no gradient
  if (gb)
    for (int t = Tb; t < T; ++t)
      for (int cc = lane; cc < C; cc += 32) gb[static_cast<long long>(t) * P.g_st + cc] = 0.f;
  if (L > Tb || Tb == 0) {                             // can never be aligned (nll 0 under zero_infinity); Tb == 0 && L == 0: nll 0
    if (lane == 0) P.nll[b] = 0.f;
    if (gb)
      for (int t = 0; t < Tb; ++t)
        for (int cc = lane; cc < C; cc += 32) gb[static_cast<long long>(t) * P.g_st + cc] = 0.f;
    return;
  }
  bool bad_label = false;
  for (int i = lane; i < L; i += 32) {
    int v = P.targets[toff + i];
    if (v < 0 || v >= C) { v = min(max(v, 0), C - 1); bad_label = true; }
    labels[i] = v;
  }
  for (int cc = lane; cc < C; cc += 32) bins[cc] = 0u;
  __syncwarp();
  bad_label = __any_sync(0xffffffffu, bad_label);
  const float* xb = P.x + static_cast<long long>(b) * P.x_sb;
  TpSeq q;
  q.x = xb; q.x_st = P.x_st; q.g = gb; q.g_st = P.g_st; q.lse2 = lse2; q.ck = ck; q.ckoff = ckoff; q.bins = bins;
  q.ring = ring; q.ldr = ldr;
  q.Tb = Tb; q.S = S; q.L = L; q.C = C; q.is_logprob = P.is_logprob;
  q.dbg = P.dbg && b == 0;
  if (q.dbg && lane == 0) g_tp_stamps[0] = clock64();
  q.vec_ok = ((C & 3) == 0) && ((P.x_st & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0);
  q.gs = P.grad_scale ? P.grad_scale[b] : P.grad_scale_const;
  float nll = 0.f;
  bool ok;
  switch (Kf) {
    case 2: ok = tp_sequence<2>(q, labels, &nll); break;
    case 4: ok = tp_sequence<4>(q, labels, &nll); break;
    default: ok = tp_sequence<6>(q, labels, &nll); break;
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    if (!ok) flags[b] = 1;
    else P.nll[b] = bad_label ? __int_as_float(0x7fc00000) : nll;
  }
}

}  // namespace htrvt
extern "C" int htrvt_ctc_tput_debug_stamps(long long* out8) {
  return cudaMemcpyFromSymbol(out8, htrvt::g_tp_stamps, sizeof(long long) * 8) == cudaSuccess ? HTRVT_OK : HTRVT_ERR_LAUNCH;
}
namespace htrvt {

size_t ctc_tput_smem_bytes(int T, int C, int kmax) {
  const int nck = T / kTpR + 1, SPmax = 32 * kmax, ldr = (C + 3) & ~3;
  const size_t per_warp = (static_cast<size_t>(2 * kTpR) * ldr + static_cast<size_t>(nck) * SPmax + nck + T + C +
                           SPmax / 2 + 8) * 4;
  return kTpWarps * ((per_warp + 15) & ~size_t(15));
}

// flags: int [B]; offsets: int [B] scratch (exclusive prefix sum of the label lengths, filled here)
int ctc_tput_launch(const CtcParams& P, int* flags, int* offsets, cudaStream_t stream) {
  const size_t smem = ctc_tput_smem_bytes(P.T, P.C, P.kmax);
  if (smem > 200 * 1024) return HTRVT_ERR_SHAPE;
  if (smem > 48 * 1024 && !HTRVT_ENSURE_SMEM(ctc_tput_kernel, smem)) return HTRVT_ERR_LAUNCH;
  if (P.tgt_stride <= 0) {
    tp_offsets_kernel<<<1, 1024, 0, stream>>>(P.target_lengths, P.B, offsets);
    HTRVT_LAUNCH_CHECK();
  }
  ctc_tput_kernel<<<(P.B + kTpWarps - 1) / kTpWarps, kTpWarps * 32, smem, stream>>>(P, flags, offsets);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

}  // namespace htrvt
