// CTC loss forward-backward, THROUGHPUT kernel for large batches: a GROUP of G lanes per sequence (G = 4 / 8 / 16 / 32
// for labels up to 32 / 64 / 128 / 256), 32 / G sequences per warp in lock step, persistent warps.
//
// Why this shape (profiles/r02_ctc_tput.md: the round's first throughput kernel, one warp per sequence on the fp64 pipe,
// issued 55 k warp instructions per sequence at 0.45 IPC; it is gone): the cost of CTC at B >> #SMs is instruction issue and per-step latency, so the layout minimises
// instructions per (t, s) cell, keeps the loop-carried chain of a step to three FP32 instructions and software-pipelines
// everything else around it:
//   states    K = 16 (or 8) CONSECUTIVE states per lane in registers (state s = 1 + K j + i on lane j: even i = label, odd i =
//             blank), state 0 rides on lane 0 as a scalar.  A step is 2 (blank) / 3 (label) FP32 instructions per state,
//             two shuffles per lane for alpha and one for beta; the 16 updates of a lane are independent (ILP 16);
//   domain    linear fp32 with an exact power-of-two rescale every step.  The scale applied at step t comes from the row
//             maximum of step t-2 (its shuffle reduction completes off the chain) minus the shift already applied at
//             t-1 (deadbeat: the maximum stays within two steps' drift of 2^48).  Emissions e_t(c) = 2^((x_t(c) - max_t)
//             log2 e) are un-normalised: the softmax denominators are only formed in the backward sweep, where the
//             gradient needs them anyway; the row maxima of the forward sweep are kept in shared memory for it;
//   storage   alpha rows go to a per-warp slot of a global scratch that the persistent warp reuses for every sequence
//             pack (written forwards, read backwards by cp.async into a 4-row ring: L2 resident); beta needs no storage:
//             the backward sweep forms gamma_t(s) = alpha_t(s) beta'_t(s) / Z lane-locally (beta' = beta without its own
//             emission; Z and both power-of-two scales are known, so no per-row reduction sits in front of the scatter),
//             scatters it to per-class fixed-point bins (integer shared-memory atomics: deterministic) and writes the
//             gradient row (softmax - posterior) * scale with 16-byte stores;
//   pipeline  forward iteration t: [row max + emission gather of row t+1] | [alpha step t]; backward iteration t:
//             [emissions + denominator of row t-1] | [posterior scatter + beta step of row t] | [gradient of row t+1];
//             ONE warp-level barrier per iteration (emission rows triple-buffered, bins double-buffered);
//   inputs    logits rows are staged by 16-byte cp.async copies into an 8-row ring per group, 7 rows ahead, and pulled
//             into L2 12 rows ahead;
//   guard     sum_s alpha_t(s) beta'_t(s) must reproduce the forward likelihood at every row (6e-5 relative) and the
//             likelihood must be a positive normal number: otherwise the sequence is FLAGGED and redone by ctc.cu's
//             CTA-per-sequence kernel (fp64 linear domain with a log-space fallback) in the fix-up launch.
// Same call-site semantics as ctc.cu (model_v1/train.py:21-30: log_softmax + nn.CTCLoss(reduction='none',
// zero_infinity=True), blank = 0); gradient w.r.t. the logits (or log-probs) = (softmax - posterior) * grad_scale.
#include "ctc.cuh"

namespace htrvt {

constexpr int kGrpWarps = 2;                   // warps per CTA (warps are independent: no CTA-wide barrier)
constexpr int kGrpNR = 4;                      // staged alpha rows per warp (power of two)
constexpr int kGrpNX = 8;                      // staged logits rows per group (power of two, >= kGrpNR): 7 rows ahead
constexpr int kGrpTgt = 48;                    // row maximum is steered to ~2^48
constexpr float kGrpMagic = 12582912.f;        // 1.5 * 2^23: (v + magic) holds round(v) in its low mantissa bits
constexpr int kGrpMagicBits = 0x4B400000;
constexpr float kGrpFix = 4194304.f;           // posterior fixed point 2^22
constexpr float kGrpTolCounts = 252.f;         // 6e-5 * 2^22

__device__ unsigned long long g_grp_flagged = 0ull;   // sequences handed to the fix-up kernel by the guard

__device__ __forceinline__ void grp_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void grp_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void grp_cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void grp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void grp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G>
__device__ __forceinline__ float grp_max(float v, unsigned m) {
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, d));
  return v;
}
template <int G>
__device__ __forceinline__ float grp_sum(float v, unsigned m) {
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(m, v, d);
  return v;
}
// reductions of a register array as balanced trees (a sequential `acc = op(acc, x[i])` chain is N dependent
// instructions: FP addition is not reassociated by the compiler)
template <int N, int STRIDE = 1>
__device__ __forceinline__ float tree_max(const float* x) {
  if constexpr (N == 1) return x[0];
  else return fmaxf(tree_max<N / 2, STRIDE>(x), tree_max<N - N / 2, STRIDE>(x + (N / 2) * STRIDE));
}
template <int N, int STRIDE = 1>
__device__ __forceinline__ float tree_sum(const float* x) {
  if constexpr (N == 1) return x[0];
  else return tree_sum<N / 2, STRIDE>(x) + tree_sum<N - N / 2, STRIDE>(x + (N / 2) * STRIDE);
}

// exact 2^e as a float, e clamped to the normal range
__device__ __forceinline__ float grp_pow2(int e) { return __int_as_float((127 + min(max(e, -126), 127)) << 23); }
// deadbeat rescale: mx = reduced row maximum measured two steps ago, kprev = shift applied since then.
// Returns k (the row is multiplied by 2^-k) steering the maximum to 2^kGrpTgt; 0 for an empty row.
__device__ __forceinline__ int grp_shift(float mx, int kprev) {
  const int ex = (__float_as_int(mx) >> 23) & 0xff;
  return ex == 0 ? 0 : min(max(ex - 127 - kGrpTgt - kprev, -100), 100);
}

// explicit shared-space accesses on 32-bit addresses (no generic-pointer arithmetic in the loops)
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_f4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_z4(uint32_t a) {
  asm volatile("st.shared.v4.s32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0) : "memory");
}
__device__ __forceinline__ void red_s(uint32_t a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct GrpSeq {
  const float* x;            // this sequence's logits / log-probs, row stride x_st
  long long x_st;
  float* g;                  // gradient rows (nullable), row stride g_st
  long long g_st;
  uint32_t ring;             // smem [NR][ldr] staged rows, pad columns [C, ldr) hold -inf
  uint32_t erow;             // smem [3][ldr] emissions of rows t+1, t, t-1 (backward sweep)
  uint32_t bins;             // smem [2][ldr] posterior per class, fixed point 2^22
  uint32_t mrow;             // smem [T] row maximum * log2 e, written by the forward sweep
  uint32_t a0ring;           // smem [NR] {alpha(state 0), scale exponent} of staged alpha rows
  uint32_t aring;            // smem alpha ring of this LANE: row r, quarter u at aring + ((r & 3) * 4 + u) * 512
  float4* arow;              // global alpha scratch of this LANE: row t, quarter u at arow[(t * 4 + u) * 32]
  float2* a0row;             // global {alpha(state 0), scale exponent} of this GROUP: row t at a0row[t * (32 / G)]
  int ldr, C, Tb, L, vec, gvec, is_logprob;
  float gs;
};

constexpr int kGrpPF = 12;                     // rows ahead of the sweeps that are pulled into L2

// One sequence on the G lanes of a group.  Returns false if it must be redone by the fix-up kernel.
// NQ > 0: every lane handles exactly NQ 16-byte chunks of a row (index clamped to the last chunk: duplicates are
// harmless for max / identical stores, `clive` masks sums); NQ == 0: runtime chunk loops.
template <int G, int NQ, int K>
__device__ __forceinline__ bool grp_sequence(const GrpSeq& q, const int* __restrict__ tg, int j, unsigned gm, float* nll_out) {
  const int L = q.L, Tb = q.Tb, C = q.C;
  const int C4v = (C + 3) >> 2;                         // 16-byte chunks of a row that hold classes
  constexpr int NGw = 32 / G;
  constexpr int NR = kGrpNR, NX = kGrpNX;
  constexpr int KL = K / 2;                             // labels per lane
  constexpr int KQ = K / 4;                             // float4 quarters of a lane's alpha row
  static_assert(K == 8 || K == 16, "states per lane");
  constexpr int NQs = NQ > 0 ? NQ : 1;
  const uint32_t ldr4 = static_cast<uint32_t>(q.ldr) * 4u;
  // ---- labels of this lane: k = 8 j + i; byte offset of the emission column (pad column C: emission 0 for k >= L) -------
  uint32_t lab4[KL];
  float mk[KL + 1];                                          // mk[i]: skip transition INTO label k allowed (label k != label k-1)
  {
    int prev = -1;
    if (KL * j - 1 >= 0 && KL * j - 1 < L) prev = min(max(tg[KL * j - 1], 0), C - 1);
#pragma unroll
    for (int i = 0; i < KL + 1; ++i) {
      const int k = KL * j + i;
      int v = -2;
      if (k < L) v = min(max(tg[k], 0), C - 1);
      if (i < KL) lab4[i] = static_cast<uint32_t>(k < L ? v : C) * 4u;
      mk[i] = (k < L && k >= 1 && v != prev) ? 1.f : 0.f;
      prev = v;
    }
  }
  uint32_t coff[NQs];                                   // byte offsets of the lane's chunks, liveness
  bool clive[NQs];
#pragma unroll
  for (int u = 0; u < NQs; ++u) {
    clive[u] = j + G * u < C4v;
    coff[u] = static_cast<uint32_t>(clive[u] ? j + G * u : C4v - 1) * 16u;
  }
  const bool own = K * j + 1 <= 2 * L;                  // the lane holds at least one real state
  const int s0 = 1 + K * j;                             // state of a[0]
  if (!own) {                                           // its slice of the alpha ring reads as zeros
#pragma unroll
    for (int u = 0; u < KQ * NR; ++u) sts_z4(q.aring + u * 512);
  }
  const int pfl = static_cast<int>(threadIdx.x & 31);   // L2 prefetch: lane l pulls 128-byte line l of a row

  auto stage_x = [&](int r) {                           // logits row r -> ring (no commit)
    if (r >= 0 && r < Tb) {
      const float* src = q.x + static_cast<long long>(r) * q.x_st;
      const uint32_t d = q.ring + static_cast<uint32_t>(r & (NX - 1)) * ldr4;
      if (q.vec) {
        if (NQ > 0) {
#pragma unroll
          for (int u = 0; u < NQs; ++u)
            if (clive[u]) grp_cp16(d + coff[u], reinterpret_cast<const char*>(src) + coff[u]);
        } else {
          for (int c4 = j; c4 < C4v; c4 += G) grp_cp16(d + 16 * c4, src + 4 * c4);
        }
      } else {
        for (int c = j; c < C; c += G) grp_cp4(d + 4 * c, src + c);
      }
    }
  };
  auto stage_a = [&](int r) {                           // alpha row r -> ring (no commit)
    if (r >= 0 && r < Tb) {
      if (own) {
        const uint32_t d = q.aring + static_cast<uint32_t>(r & (NR - 1)) * (KQ * 512u);
        const float4* src = q.arow + r * (KQ * 32);
#pragma unroll
        for (int u = 0; u < KQ; ++u) grp_cp16(d + u * 512, src + u * 32);
      }
      if (j == 0) grp_cp8(q.a0ring + static_cast<uint32_t>(r & (NR - 1)) * 8u, q.a0row + r * NGw);
    }
  };
  auto pull_x = [&](int r) {                            // logits row r -> L2
    if (r >= 0 && r < Tb && j * 128 < C * 4)
      prefetch_l2(reinterpret_cast<const char*>(q.x + static_cast<long long>(r) * q.x_st) + j * 128);
  };

  // ================================ alpha sweep =========================================================================
  float a[K];
#pragma unroll
  for (int i = 0; i < K; ++i) a[i] = 0.f;
  float a0 = j == 0 ? grp_pow2(kGrpTgt) : 0.f;          // "alpha_{-1}": all mass in front of state 0, pre-scaled
  int KA = -kGrpTgt;                                    // stored alpha = alpha_e * 2^-KA
  int ksc = 0;                                          // shift applied in the current step
  float sc = 1.f;
  float mx_old = grp_pow2(kGrpTgt);                     // reduced maximum measured one step ago
  float ecb = 0.f, ec[KL];                              // raw emissions of the row the alpha step consumes
#pragma unroll
  for (int i = 0; i < KL; ++i) ec[i] = 0.f;

  // iteration t: [row maximum + raw emissions of row t + 1] | [alpha step of row t]
  auto fwd_iter = [&](bool do_g, bool do_s, int t) {
    float xb = 0.f, xg[KL];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < KL; ++i) xg[i] = 0.f;
    if (do_g) {                                         // loads first
      const uint32_t row = q.ring + static_cast<uint32_t>((t + 1) & (NX - 1)) * ldr4;
      if (NQ > 0) {
        float4 v[NQs];
#pragma unroll
        for (int u = 0; u < NQs; ++u) v[u] = lds_f4(row + coff[u]);
        xb = lds_f(row);
#pragma unroll
        for (int i = 0; i < KL; ++i) xg[i] = lds_f(row + lab4[i]);
#pragma unroll
        for (int u = 0; u < NQs; ++u) m = fmaxf(fmaxf(m, fmaxf(v[u].x, v[u].y)), fmaxf(v[u].z, v[u].w));
      } else {
        xb = lds_f(row);
#pragma unroll
        for (int i = 0; i < KL; ++i) xg[i] = lds_f(row + lab4[i]);
        for (int c4 = j; c4 < C4v; c4 += G) {
          const float4 v = lds_f4(row + 16 * c4);
          m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
      }
      m = grp_max<G>(m, gm);
    }
    if (do_s) {
      const float eb = ecb * sc;
      float e[KL];
#pragma unroll
      for (int i = 0; i < KL; ++i) e[i] = ec[i] * sc;
      KA += ksc;
      float p1 = __shfl_up_sync(gm, a[K - 1], 1, G), p2 = __shfl_up_sync(gm, a[K - 2], 1, G);
      if (j == 0) { p1 = a0; p2 = 0.f; }
#pragma unroll
      for (int i = K - 1; i >= 0; --i) {                // descending: a[i-1], a[i-2] are still the previous row
        const float x1 = i >= 1 ? a[i >= 1 ? i - 1 : 0] : p1;
        if (i & 1) {
          a[i] = (a[i] + x1) * eb;
        } else {
          const float x2 = i >= 2 ? a[i >= 2 ? i - 2 : 0] : p2;
          a[i] = fmaf(mk[i >> 1], x2, a[i] + x1) * e[i >> 1];
        }
      }
      a0 *= eb;
      float4* dst = q.arow + t * (KQ * 32);
      if (own) {
#pragma unroll
        for (int u = 0; u < KQ; ++u) dst[u * 32] = make_float4(a[4 * u], a[4 * u + 1], a[4 * u + 2], a[4 * u + 3]);
      }
      if (j == 0) q.a0row[t * NGw] = make_float2(a0, __int_as_float(KA));
      float mx = fmaxf(a0, tree_max<K>(a));
      mx = grp_max<G>(mx, gm);                          // consumed one step later
      ksc = grp_shift(mx_old, ksc);                     // shift for the next step: measured at t - 1, minus the shift of t
      sc = grp_pow2(-ksc);
      mx_old = mx;
    }
    if (do_g) {
      const float mL = m * kLog2e;
      if (j == 0) sts_f(q.mrow + (t + 1) * 4, mL);
      ecb = ex2f(fmaf(xb, kLog2e, -mL));
#pragma unroll
      for (int i = 0; i < KL; ++i) ec[i] = ex2f(fmaf(xg[i], kLog2e, -mL));
    }
  };

#pragma unroll
  for (int r = 0; r < NX - 1; ++r) { stage_x(r); grp_commit(); }
  for (int r = NX - 1; r < kGrpPF; ++r) pull_x(r);
  for (int t = -1; t < Tb; ++t) {
    grp_wait<NX - 2>();
    __syncwarp(gm);
    stage_x(t + NX);
    grp_commit();
    pull_x(t + kGrpPF);
    if (t >= 0 && t + 1 < Tb) fwd_iter(true, true, t);  // steady state: both stages in one basic block
    else fwd_iter(t + 1 < Tb, t >= 0, t);
  }
  // likelihood (emission domain): Z_e = (alpha(2L) + alpha(2L - 1)) * 2^KA
  float z = (j == 0 && L == 0) ? a0 : 0.f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int s = s0 + i;
    if (s == 2 * L || s == 2 * L - 1) z += a[i];
  }
  z = grp_sum<G>(z, gm);
  const int Kfin = KA;
  const int zbits = __float_as_int(z);
  const int ez = ((zbits >> 23) & 0xff) - 127;
  bool ok = z >= 1.17549435e-38f && z < INFINITY;       // positive, normal, finite (NaN fails)
  grp_wait<0>();
  __syncwarp(gm);
  if (!ok) return false;
  const float lzf = lg2f(z);
  const float rzm = __frcp_rn(__int_as_float((zbits & 0x007fffff) | 0x3f800000));   // 1 / mantissa(z), in (0.5, 1]

  // ================================ backward sweep: beta', posterior, gradient ==========================================
  double corr2 = 0.0;                                   // log2 Z = lzf + Kfin - corr2
  if (!q.g) {                                           // loss only: the softmax denominators of every row
#pragma unroll
    for (int r = 0; r < NX - 1; ++r) { stage_x(Tb - 1 - r); grp_commit(); }
    for (int t = Tb - 1; t >= 0; --t) {
      grp_wait<NX - 2>();
      __syncwarp(gm);
      stage_x(t + 1 - NX);                              // into the slot read one iteration ago
      grp_commit();
      const uint32_t row = q.ring + static_cast<uint32_t>(t & (NX - 1)) * ldr4;
      const float mL = lds_f(q.mrow + t * 4);
      float sum = 0.f;
      for (int c4 = j; c4 < C4v; c4 += G) {
        const float4 v = lds_f4(row + 16 * c4);
        sum += (ex2f(fmaf(v.x, kLog2e, -mL)) + ex2f(fmaf(v.y, kLog2e, -mL))) +
               (ex2f(fmaf(v.z, kLog2e, -mL)) + ex2f(fmaf(v.w, kLog2e, -mL)));
      }
      sum = grp_sum<G>(sum, gm);
      corr2 += q.is_logprob ? static_cast<double>(-mL) : static_cast<double>(lg2f(sum));
    }
    grp_wait<0>();
    __syncwarp(gm);
    if (j == 0) *nll_out = static_cast<float>(-(static_cast<double>(lzf) + Kfin - corr2) * 0.6931471805599453);
    return true;
  }

  float bp[K];                                          // beta'_t (without the emission of row t), scaled by 2^-KB
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int s = s0 + i;
    bp[i] = (s == 2 * L || s == 2 * L - 1) ? grp_pow2(kGrpTgt) : 0.f;
  }
  float bp0 = (j == 0 && L == 0) ? grp_pow2(kGrpTgt) : 0.f;
  int KB = -kGrpTgt, kscb = 0;
  float scb = 1.f, mxb_old = grp_pow2(kGrpTgt);
  float s_new = 1.f, s_mid = 1.f, s_old = 1.f;          // denominators of rows t-1, t, t+1
  float m_new = 0.f, m_mid = 0.f, m_old = 0.f;          // their row maxima * log2 e
  float tot_chk = kGrpFix;                              // sum of the previous row's posteriors in counts (deferred check)
  const float ginv = q.gs * (1.f / kGrpFix);
  const float gc0 = 8388608.f * ginv;                  // (2^23 + n) * -ginv + gc0 = -n * ginv: the int -> float conversion folded into one FFMA
  // rotating buffers: emissions of rows t-1 (written), t (gathered), t+1 (gradient); bins of rows t, t+1
  uint32_t e_nxt = q.erow + static_cast<uint32_t>((Tb + 2) % 3) * ldr4;
  uint32_t e_cur = q.erow + static_cast<uint32_t>(Tb % 3) * ldr4;
  uint32_t e_old = q.erow + static_cast<uint32_t>((Tb + 1) % 3) * ldr4;
  uint32_t b_cur = q.bins + static_cast<uint32_t>(Tb & 1) * ldr4;
  uint32_t b_old = q.bins + static_cast<uint32_t>((Tb + 1) & 1) * ldr4;

  // iteration t: D = emissions + denominator of row t-1 -> e_nxt; S = posterior scatter of row t + beta'_{t-1};
  //              Gd = gradient of row t+1 from e_old / b_old (and clears b_old)
  auto bwd_iter = [&](bool do_d, bool do_s, bool do_g, int t) {
    // ---------------- loads ----------------
    float4 xv[NQs], ev[NQs];
    int4 bv[NQs];
    float mLd = 0.f;
    float an[K], an0 = 0.f;
    float eg[KL], egb = 0.f;
    int KAt = 0;
#pragma unroll
    for (int u = 0; u < NQs; ++u) {
      xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      ev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      bv[u] = make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < K; ++i) an[i] = 0.f;
#pragma unroll
    for (int i = 0; i < KL; ++i) eg[i] = 0.f;
    if (do_d) {
      const uint32_t row = q.ring + static_cast<uint32_t>((t - 1) & (NX - 1)) * ldr4;
      mLd = lds_f(q.mrow + (t - 1) * 4);
      if (NQ > 0) {
#pragma unroll
        for (int u = 0; u < NQs; ++u) xv[u] = lds_f4(row + coff[u]);
      }
    }
    if (do_s) {
      const uint32_t ar = q.aring + static_cast<uint32_t>(t & (NR - 1)) * (KQ * 512u);
#pragma unroll
      for (int u = 0; u < KQ; ++u) {
        const float4 v = lds_f4(ar + u * 512);
        an[4 * u] = v.x; an[4 * u + 1] = v.y; an[4 * u + 2] = v.z; an[4 * u + 3] = v.w;
      }
      const float2 a0k = lds_f2(q.a0ring + static_cast<uint32_t>(t & (NR - 1)) * 8u);
      an0 = j == 0 ? a0k.x : 0.f;
      KAt = __float_as_int(a0k.y);
      egb = lds_f(e_cur);
#pragma unroll
      for (int i = 0; i < KL; ++i) eg[i] = lds_f(e_cur + lab4[i]);
    }
    if (do_g && q.gvec && NQ > 0) {
#pragma unroll
      for (int u = 0; u < NQs; ++u) { ev[u] = lds_f4(e_old + coff[u]); bv[u] = lds_i4(b_old + coff[u]); }
    }
    // ---------------- D: emissions of row t-1 ----------------
    if (do_d) {
      const uint32_t row = q.ring + static_cast<uint32_t>((t - 1) & (NX - 1)) * ldr4;
      float sum = 0.f;
      if (NQ > 0) {
#pragma unroll
        for (int u = 0; u < NQs; ++u) {
          float4 w;
          w.x = ex2f(fmaf(xv[u].x, kLog2e, -mLd));
          w.y = ex2f(fmaf(xv[u].y, kLog2e, -mLd));
          w.z = ex2f(fmaf(xv[u].z, kLog2e, -mLd));
          w.w = ex2f(fmaf(xv[u].w, kLog2e, -mLd));
          const float s4 = (w.x + w.y) + (w.z + w.w);
          sum += clive[u] ? s4 : 0.f;
          sts_f4(e_nxt + coff[u], w);
        }
      } else {
        for (int c4 = j; c4 < C4v; c4 += G) {
          const float4 v = lds_f4(row + 16 * c4);
          float4 w;
          w.x = ex2f(fmaf(v.x, kLog2e, -mLd));
          w.y = ex2f(fmaf(v.y, kLog2e, -mLd));
          w.z = ex2f(fmaf(v.z, kLog2e, -mLd));
          w.w = ex2f(fmaf(v.w, kLog2e, -mLd));
          sum += (w.x + w.y) + (w.z + w.w);
          sts_f4(e_nxt + 16 * c4, w);
        }
      }
      s_new = grp_sum<G>(sum, gm);                      // consumed two iterations later
      m_new = mLd;
    }
    // ---------------- S: posterior scatter of row t, beta'_{t-1} ----------------
    if (do_s) {
      // sum_s alpha beta' = Z_e 2^-(KA_t + KB_t): the normaliser is known without a reduction
      const int E = 22 - ez + (KAt + KB - Kfin);
      if (E < -126 || E > 126) ok = false;           // the row's posterior mass lies > 2^100 below the row maxima: fp64 kernel
      const float rr = rzm * grp_pow2(E);
      float ab[K];
      const float ab0 = an0 * bp0;
#pragma unroll
      for (int i = 0; i < K; ++i) ab[i] = an[i] * bp[i];
#pragma unroll
      for (int i = 0; i < K; i += 2) red_s(b_cur + lab4[i >> 1], __float_as_int(fmaf(ab[i], rr, kGrpMagic)) - kGrpMagicBits);
      const float blank = tree_sum<KL, 2>(ab + 1) + ab0;
      const float tot = tree_sum<KL, 2>(ab) + blank;
      red_s(b_cur, __float_as_int(fmaf(blank, rr, kGrpMagic)) - kGrpMagicBits);
      // deferred guard: the previous row's posteriors must sum to 2^22 counts
      if (!(fabsf(tot_chk - kGrpFix) < kGrpTolCounts)) ok = false;
      tot_chk = grp_sum<G>(tot, gm) * rr;
      const float ebs = egb * scb;
      float bb[K];
#pragma unroll
      for (int i = 0; i < K; ++i) bb[i] = bp[i] * ((i & 1) ? ebs : eg[i >> 1] * scb);
      const float b0 = bp0 * ebs;
      float n1 = __shfl_down_sync(gm, bb[0], 1, G);
      if (j == G - 1) n1 = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const float y1 = i + 1 < K ? bb[i + 1 < K ? i + 1 : 0] : n1;
        if (i & 1) {
          bp[i] = bb[i] + y1;
        } else {
          const float y2 = i + 2 < K ? bb[i + 2 < K ? i + 2 : 0] : n1;
          bp[i] = fmaf(mk[(i >> 1) + 1], y2, bb[i] + y1);
        }
      }
      bp0 = j == 0 ? b0 + bb[0] : 0.f;
      KB += kscb;
      float mx = fmaxf(bp0, tree_max<K>(bp));
      mx = grp_max<G>(mx, gm);                          // consumed one step later
      kscb = grp_shift(mxb_old, kscb);
      scb = grp_pow2(-kscb);
      mxb_old = mx;
    }
    // ---------------- Gd: gradient row t+1 = (e * norm - posterior) * gs ----------------
    if (do_g) {
      const float norm = q.is_logprob ? ex2f(m_old) : __frcp_rn(s_old);
      corr2 += q.is_logprob ? static_cast<double>(-m_old) : static_cast<double>(lg2f(s_old));
      const float ig = norm * q.gs;
      float* grow = q.g + static_cast<long long>(t + 1) * q.g_st;
      if (q.gvec) {
        if (NQ > 0) {
#pragma unroll
          for (int u = 0; u < NQs; ++u) {
            sts_z4(b_old + coff[u]);
            float4 gv;
            gv.x = fmaf(__int_as_float(bv[u].x | 0x4B000000), -ginv, fmaf(ev[u].x, ig, gc0));
            gv.y = fmaf(__int_as_float(bv[u].y | 0x4B000000), -ginv, fmaf(ev[u].y, ig, gc0));
            gv.z = fmaf(__int_as_float(bv[u].z | 0x4B000000), -ginv, fmaf(ev[u].z, ig, gc0));
            gv.w = fmaf(__int_as_float(bv[u].w | 0x4B000000), -ginv, fmaf(ev[u].w, ig, gc0));
            *reinterpret_cast<float4*>(reinterpret_cast<char*>(grow) + coff[u]) = gv;
          }
        } else {
          for (int c4 = j; c4 < C4v; c4 += G) {
            const float4 e4 = lds_f4(e_old + 16 * c4);
            const int4 b4 = lds_i4(b_old + 16 * c4);
            sts_z4(b_old + 16 * c4);
            float4 gv;
            gv.x = fmaf(__int_as_float(b4.x | 0x4B000000), -ginv, fmaf(e4.x, ig, gc0));
            gv.y = fmaf(__int_as_float(b4.y | 0x4B000000), -ginv, fmaf(e4.y, ig, gc0));
            gv.z = fmaf(__int_as_float(b4.z | 0x4B000000), -ginv, fmaf(e4.z, ig, gc0));
            gv.w = fmaf(__int_as_float(b4.w | 0x4B000000), -ginv, fmaf(e4.w, ig, gc0));
            reinterpret_cast<float4*>(grow)[c4] = gv;
          }
        }
      } else {
        for (int c = j; c < C; c += G) {
          int bvs;
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(bvs) : "r"(b_old + 4 * c));
          asm volatile("st.shared.s32 [%0], %1;" ::"r"(b_old + 4 * c), "r"(0) : "memory");
          grow[c] = fmaf(__int_as_float(bvs | 0x4B000000), -ginv, fmaf(lds_f(e_old + 4 * c), ig, gc0));
        }
      }
    }
  };

#pragma unroll
  // commit groups: one per iteration = {logits row t - NX, alpha row t + 1 - NR}; waiting for all but the NR - 2 newest
  // guarantees alpha row t (and logits rows up to t - NX + NR - 1, i.e. row t - 1 long before it is read)
#pragma unroll
  for (int r = NR - 1; r < NX - 1; ++r) stage_x(Tb - 1 - r);
#pragma unroll
  for (int r = 0; r < NR - 1; ++r) { stage_x(Tb - 1 - r); stage_a(Tb - 1 - r); grp_commit(); }
  for (int t = Tb; t >= -1; --t) {
    grp_wait<NR - 2>();
    __syncwarp(gm);
    stage_x(t - NX);
    stage_a(t + 1 - NR);
    grp_commit();
    pull_x(t - kGrpPF);
    if (t - kGrpPF >= 0 && pfl < KQ * 4)
      prefetch_l2(reinterpret_cast<const char*>(q.arow - pfl + (t - kGrpPF) * (KQ * 32)) + pfl * 128);
    if (t >= 1 && t + 1 < Tb) bwd_iter(true, true, true, t);       // steady state: the three stages in one basic block
    else bwd_iter(t >= 1, t >= 0 && t < Tb, t + 1 < Tb, t);
    s_old = s_mid; s_mid = s_new;
    m_old = m_mid; m_mid = m_new;
    const uint32_t er = e_old; e_old = e_cur; e_cur = e_nxt; e_nxt = er;
    const uint32_t br = b_old; b_old = b_cur; b_cur = br;
  }
  if (!(fabsf(tot_chk - kGrpFix) < kGrpTolCounts)) ok = false;   // row 0
  grp_wait<0>();
  __syncwarp(gm);
  if (j == 0) *nll_out = static_cast<float>(-(static_cast<double>(lzf) + Kfin - corr2) * 0.6931471805599453);
  return ok;
}

static __host__ __device__ inline size_t grp_group_floats(int T, int C) {
  const int ldr = (C + 4) & ~3;
  return static_cast<size_t>(kGrpNX + 3 + 2) * ldr + ((T + 3) & ~3) + 2 * kGrpNR;
}
// per warp: the alpha ring (kGrpNR rows of 32 lanes x K floats) + the groups' private buffers
static __host__ __device__ inline size_t grp_warp_floats(int T, int C, int G, int K) {
  return static_cast<size_t>(kGrpNR) * 32 * K + (32 / G) * grp_group_floats(T, C);
}

// resident CTAs per SM the register budget is sized for: K = 16 -> 4 (255 registers), K = 8 -> 7 (146 registers)
constexpr int grp_min_ctas(int K) { return K == 8 ? 7 : 4; }

template <int G, int NQ, int K>
__global__ void __launch_bounds__(kGrpWarps * 32, grp_min_ctas(K)) ctc_grp_kernel(const CtcParams P, int* __restrict__ flags,
                                                               const int* __restrict__ offsets,
                                                               float4* __restrict__ ascr, float2* __restrict__ a0scr) {
  extern __shared__ __align__(16) unsigned char grp_smem[];
  constexpr int NGw = 32 / G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = lane & (G - 1), grp = lane / G;
  const unsigned gm = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
  const int T = P.T, C = P.C;
  const int ldr = (C + 4) & ~3;
  float* wbase = reinterpret_cast<float*>(grp_smem) + warp * grp_warp_floats(T, C, G, K);
  float* gbase = wbase + kGrpNR * 32 * K + grp * grp_group_floats(T, C);
  GrpSeq q;
  float* erow_p = gbase + kGrpNX * ldr;
  int* bins_p = reinterpret_cast<int*>(erow_p + 3 * ldr);
  float* mrow_p = reinterpret_cast<float*>(bins_p + 2 * ldr);
  q.aring = smem_u32(reinterpret_cast<float4*>(wbase) + lane);
  q.ring = smem_u32(gbase);
  q.erow = smem_u32(erow_p);
  q.bins = smem_u32(bins_p);
  q.mrow = smem_u32(mrow_p);
  q.a0ring = smem_u32(mrow_p + ((T + 3) & ~3));
  q.ldr = ldr; q.C = C; q.is_logprob = P.is_logprob;
  for (int r = 0; r < kGrpNX; ++r)
    for (int c = C + j; c < ldr; c += G) gbase[r * ldr + c] = -INFINITY;
  for (int c = j; c < 3 * ldr; c += G) erow_p[c] = 0.f;
  for (int c = j; c < 2 * ldr; c += G) bins_p[c] = 0;
  __syncwarp(gm);
  const int slot = blockIdx.x * kGrpWarps + warp, nslots = gridDim.x * kGrpWarps;
  q.arow = ascr + static_cast<size_t>(slot) * T * (K * 8) + lane;
  q.a0row = a0scr + static_cast<size_t>(slot) * T * NGw + grp;
  const int npacks = (P.B + NGw - 1) / NGw;
  for (int pack = slot; pack < npacks; pack += nslots) {
    const int b = pack * NGw + grp;
    if (b >= P.B) continue;
    int Tb = P.input_lengths ? P.input_lengths[b] : T;
    Tb = min(max(Tb, 0), T);
    const int L = P.target_lengths[b];
    float* gb = P.grad ? P.grad + static_cast<long long>(b) * P.g_sb : nullptr;
    if (L < 0 || L > (K / 2) * G) {                            // not provisioned here: the CTA-per-sequence kernel takes it
      if (j == 0) flags[b] = 1;
      continue;
    }
    const int toff = P.tgt_stride > 0 ? b * P.tgt_stride : offsets[b];
    const int* tg = P.targets + toff;
    // feasibility by counting (a blank is needed between equal neighbours): L + repeats <= Tb; labels outside [0, C)
    int rep = 0, bad = 0;
    for (int i = j; i < L; i += G) {
      const int v = tg[i];
      bad |= (v < 0 || v >= C);
      if (i >= 1) rep += min(max(v, 0), C - 1) == min(max(tg[i - 1], 0), C - 1);
    }
#pragma unroll
    for (int d = G / 2; d >= 1; d >>= 1) {
      rep += __shfl_xor_sync(gm, rep, d);
      bad |= __shfl_xor_sync(gm, bad, d);
    }
    const bool feasible = Tb > 0 && L + rep <= Tb;
    const int Te = feasible ? Tb : 0;
    if (gb)                                              // rows beyond the input length (all rows if infeasible): no gradient
      for (int t = Te; t < T; ++t)
        for (int c = j; c < C; c += G) gb[static_cast<long long>(t) * P.g_st + c] = 0.f;
    if (!feasible) {                                     // zero_infinity=True
      if (j == 0) { flags[b] = 0; P.nll[b] = 0.f; }
      continue;
    }
    q.x = P.x + static_cast<long long>(b) * P.x_sb;
    q.x_st = P.x_st; q.g = gb; q.g_st = P.g_st; q.Tb = Tb; q.L = L;
    q.vec = ((C & 3) == 0) && ((P.x_st & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.x) & 15) == 0);
    q.gvec = gb && ((C & 3) == 0) && ((P.g_st & 3) == 0) && ((reinterpret_cast<uintptr_t>(gb) & 15) == 0);
    q.gs = P.grad_scale ? P.grad_scale[b] : P.grad_scale_const;
    float nll = 0.f;
    bool ok = grp_sequence<G, NQ, K>(q, tg, j, gm, &nll);
    ok = __all_sync(gm, ok);
    if (!ok) {                                           // leave the shared-memory state clean for the next sequence
      for (int c = j; c < 2 * ldr; c += G) bins_p[c] = 0;
    }
    __syncwarp(gm);
    if (j == 0) {
      flags[b] = ok ? 0 : 1;
      if (ok) P.nll[b] = bad ? __int_as_float(0x7fc00000) : nll;
      else atomicAdd(&g_grp_flagged, 1ull);
    }
  }
}

// exclusive prefix sum of the label lengths -> start of every sequence's labels in the concatenated target stream
__global__ void __launch_bounds__(1024) ctc_offsets_kernel(const int* __restrict__ lengths, int B, int* __restrict__ offs) {
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

int ctc_offsets_launch(const int* lengths, int B, int* offsets, cudaStream_t stream) {
  ctc_offsets_kernel<<<1, 1024, 0, stream>>>(lengths, B, offsets);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

int ctc_num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!n[dev]) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

// Configuration: K states per lane (8 or 16) and G lanes per sequence (labels up to (K / 2) G).  K = 16 issues the fewest
// instructions per sequence (4 sequences per warp at L <= 64); K = 8 halves a warp's work per time step and doubles the
// warps - the better trade while the batch cannot fill the SMs with K = 16 warps (measured: tools/ctc_grp_probe.py).
struct GrpCfgHost { int G, K; };
static int g_grp_force_k = -1;                          // HTRVT_CTC_K = 8 / 16 (developer knob), 0 = automatic
static GrpCfgHost grp_config(int B, int lmax, int sms) {
  if (g_grp_force_k < 0) {
    const char* e = getenv("HTRVT_CTC_K");
    g_grp_force_k = e ? atoi(e) : 0;
  }
  auto lanes = [](int lm, int K) {
    const int per = K / 2;
    for (int G = 4; G <= 32; G *= 2)
      if (lm <= per * G) return G;
    return 0;
  };
  const int g16 = lanes(lmax, 16), g8 = lanes(lmax, 8);
  GrpCfgHost c = {g16, 16};
  if (g16 == 0) return c;
  if (g8 != 0) {
    const long long warps16 = (static_cast<long long>(B) * g16 + 31) / 32;
    // measured (T = 128, C = 80, L <= 64): K = 8 wins up to B = 2048 (160 / 181 us against 179 / 188 us at B = 1024 / 2048:
    // per-warp latency), K = 16 from B = 4096 (273 against 281 us: instruction count)
    const bool want8 = g_grp_force_k == 8 || (g_grp_force_k != 16 && warps16 <= 4LL * sms);
    if (want8) { c.G = g8; c.K = 8; }
  }
  return c;
}

static size_t grp_smem_bytes(int T, int C, int G, int K) { return kGrpWarps * grp_warp_floats(T, C, G, K) * sizeof(float); }

// resident CTAs per SM: shared memory (227 KB) and the register file (255 registers x 64 threads fit 4 times)
static int grp_ctas_per_sm(int T, int C, int G, int K) {
  const size_t smem = grp_smem_bytes(T, C, G, K) + 1024;
  int n = static_cast<int>((227 * 1024) / smem);
  const int cap = grp_min_ctas(K);
  return n < 1 ? 0 : (n > cap ? cap : n);
}

static int grp_grid(int B, int T, int C, int G, int K, int sms) {
  const int npacks = (B + 32 / G - 1) / (32 / G);
  const int ctas = (npacks + kGrpWarps - 1) / kGrpWarps;
  const int cap = grp_ctas_per_sm(T, C, G, K) * sms;
  return ctas < cap ? ctas : cap;
}

bool ctc_grp_supported(int B, int T, int C, int lmax) {
  const GrpCfgHost c = grp_config(B, lmax, ctc_num_sms());
  return c.G != 0 && T >= 1 && grp_ctas_per_sm(T, C, c.G, c.K) >= 1;
}

// bytes of alpha scratch the launch needs (per resident warp: T rows of 32 lanes x K floats + {alpha(0), exponent})
size_t ctc_grp_scratch_bytes(int B, int T, int C, int lmax, int sms) {
  const GrpCfgHost c = grp_config(B, lmax, sms);
  if (!c.G) return 0;
  const size_t slots = static_cast<size_t>(grp_grid(B, T, C, c.G, c.K, sms)) * kGrpWarps;
  return slots * T * (static_cast<size_t>(128) * c.K + (32 / c.G) * 8) + 256;
}

template <int G, int NQ, int K>
static int grp_launch_q(const CtcParams& P, int* flags, const int* offsets, void* scratch, int sms, cudaStream_t stream) {
  const size_t smem = grp_smem_bytes(P.T, P.C, G, K);
  if (smem > 48 * 1024 && !HTRVT_ENSURE_SMEM((ctc_grp_kernel<G, NQ, K>), smem)) return HTRVT_ERR_LAUNCH;
  const int grid = grp_grid(P.B, P.T, P.C, G, K, sms);
  const size_t slots = static_cast<size_t>(grid) * kGrpWarps;
  float4* ascr = static_cast<float4*>(scratch);
  float2* a0scr = reinterpret_cast<float2*>(static_cast<char*>(scratch) + slots * P.T * 128 * K);
  ctc_grp_kernel<G, NQ, K><<<grid, kGrpWarps * 32, smem, stream>>>(P, flags, offsets, ascr, a0scr);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// NQ = 16-byte chunks of a logits row per lane: the natural counts of C = 80 / 90 (and C = 228 at G = 32) are compiled,
// anything else takes the runtime-loop instance
template <int K>
static int grp_launch_k(const CtcParams& P, int G, int* flags, const int* offsets, void* scratch, int sms, cudaStream_t stream) {
  const int nq = (((P.C + 3) >> 2) + G - 1) / G;
  switch (G) {
    case 4: return grp_launch_q<4, 0, K>(P, flags, offsets, scratch, sms, stream);
    case 8: return nq == 3 ? grp_launch_q<8, 3, K>(P, flags, offsets, scratch, sms, stream)
                           : grp_launch_q<8, 0, K>(P, flags, offsets, scratch, sms, stream);
    case 16: return nq == 2 ? grp_launch_q<16, 2, K>(P, flags, offsets, scratch, sms, stream)
                            : grp_launch_q<16, 0, K>(P, flags, offsets, scratch, sms, stream);
    case 32: return nq == 1 ? grp_launch_q<32, 1, K>(P, flags, offsets, scratch, sms, stream)
                   : nq == 2 ? grp_launch_q<32, 2, K>(P, flags, offsets, scratch, sms, stream)
                             : grp_launch_q<32, 0, K>(P, flags, offsets, scratch, sms, stream);
    default: return HTRVT_ERR_SHAPE;
  }
}

// flags: int [B]; offsets: int [B] (exclusive prefix sum of the label lengths, filled by the caller when the targets are
// concatenated); scratch: ctc_grp_scratch_bytes, 16-byte aligned
int ctc_grp_launch(const CtcParams& P, int lmax, int* flags, const int* offsets, void* scratch, int sms,
                   cudaStream_t stream) {
  const GrpCfgHost c = grp_config(P.B, lmax, sms);
  if (!c.G) return HTRVT_ERR_SHAPE;
  return c.K == 8 ? grp_launch_k<8>(P, c.G, flags, offsets, scratch, sms, stream)
                  : grp_launch_k<16>(P, c.G, flags, offsets, scratch, sms, stream);
}

}  // namespace htrvt

// sequences the lane-group throughput kernel handed to the fix-up kernel since the library was loaded
extern "C" long long htrvt_ctc_flagged_count(void) {
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, htrvt::g_grp_flagged, sizeof(v)) != cudaSuccess) return -1;
  return static_cast<long long>(v);
}
