// CTC loss forward-backward for sm_100a: one CTA per sequence, the alpha and the beta recursion each run by ONE warp
// (neighbour states exchanged with shuffles), probabilities staged once in shared memory, gradient w.r.t. the LOGITS
// (softmax - posterior) written in coalesced rows.
//   fast path : linear domain on the FP64 pipe with exact power-of-two row renormalisation (below: "Fast path");
//               the other 14 warps collect posterior rows WHILE the recursions run their second half (rows are
//               complete from the middle of the sequence outwards), normalised with the middle row's likelihood;
//   guard     : posterior rows must sum to 1 and the alpha-end / beta-start / middle-row likelihoods must agree,
//               else the sequence is recomputed by the
//   log path  : log2 domain (ex2/lg2 MUFU), renormalised every 8 steps with a float64 offset - any input.
//
// Replaces, for the reference call site model_v1/train.py:21-30 / model_v1/valid.py:32-38:
//   preds.float().permute(1,0,2).log_softmax(2) -> nn.CTCLoss(reduction='none', zero_infinity=True)
// (ATen _ctc_loss / _ctc_loss_backward; blank = 0).  Semantics: SURVEY.md 8a / 9.14-9.16.
#include "common.cuh"
#include "ctc.cuh"
#include <cstdlib>

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


// =================================================================================================
// Fast path: the same recursion in the LINEAR domain on the FP64 pipe.
//   value  = double, kept in registers; a step is 2 DADD + 1 DMUL per state behind two 64-bit shuffles -
//            no MUFU on the dependency chain (the log-space step above costs 3 SFU ops per state);
//   range  = every 4 rows the row maximum is moved to 2^-16 by an exact power of two folded into the next row's
//            probabilities (off the chain); the integer shifts accumulate in off[t], so alpha_t = a * 2^off[t];
//   probs  = softmax values as "packed doubles": bits 61..30 of the IEEE double (10 exponent bits + 22 mantissa
//            bits), one 32-bit word per (t, c): unpack = two shifts; 0 = exact zero;
//   stored = alpha / beta' rows in the same packed format (22-bit mantissa, |rel err| < 2.4e-7); the posterior pass
//            unpacks them back to doubles (two shifts each) and forms alpha * beta' * scale on the FP64 pipe, so the
//            full exponent range survives without any per-state exponent arithmetic.
// Linear fp64 can flush states that sit > ~2^-1000 below their row maximum; log space cannot.  Such a loss is
// only relevant when the other recursion is correspondingly huge there, and then it shows: every row of
// posteriors must sum to 1 (sum_s alpha_t(s) beta_t(s) / p_t(l_s) = Z for all t; flushing only ever removes mass)
// and the beta-side likelihood must equal the alpha-side one.  A sequence failing either check (> 2e-5) is
// recomputed by the log-space path above, so the result never depends on the fast path's range.
// =================================================================================================

// shift (in binades) that moves the row maximum to 2^-16; 0 for an all-zero row
template <int K>
__device__ __forceinline__ int fast_row_shift(const double (&n)[K]) {
  uint32_t m = static_cast<uint32_t>(__double2hiint(n[0]));
#pragma unroll
  for (int j = 1; j < K; ++j) m = max(m, static_cast<uint32_t>(__double2hiint(n[j])));
  m = __reduce_max_sync(0xffffffffu, m);
  return m == 0u ? 0 : static_cast<int>(m >> 20) - kFastTargetExp;
}

// explicit shared-state-space accesses (32-bit addresses): keeps the address arithmetic of the serial loops to
// one add per access instead of the generic-pointer window conversions the compiler re-materialises under predicates
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void reds_add_u32(uint32_t a, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// One recursion for both directions.  REV = false: alpha (time ascending, states s = lane*K + j).
// REV = true: beta = the alpha recursion of the REVERSED label sequence run backwards in time; the lane's states
// are r = S-1-s (S is odd, so r and s have the same parity) and rows are stored at their natural index s.
// K is even: state parity == parity of j, so blank states (even) never carry the skip term - they cost one DADD +
// one DMUL, label states DADD + DFMA (skip mask 0.0 / 1.0) + DMUL, and only ONE neighbour value crosses lanes.
// pk_s: shared address of the packed probabilities [T][ldp] with column C == 0 (what out-of-range states read);
// AB: row storage [T][SP] (shared address when SM, else a global pointer).
template <int K, bool REV, bool SM>
__device__ __forceinline__ void ctc_rec_fast(uint32_t pk_s, int ldp, int C, const int* __restrict__ ext, int S, int Tb,
                                             uint32_t ab_s, uint32_t* ab_g, int SP, uint32_t off_s, uint32_t prog_s) {
  static_assert((K & 1) == 0, "even K");
  const int lane = threadIdx.x & 31;
  double a[K], msk[K / 2];
  uint32_t q[K], lo4[K];
  bool valid[K];
  const int row_step = REV ? -ldp * 4 : ldp * 4;
  uint32_t prow = pk_s + (REV ? (Tb - 1) * ldp * 4 : 0);            // probabilities of the current row
  const int t0 = REV ? Tb - 1 : 0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int r = lane * K + j;
    const int s = REV ? S - 1 - r : r;
    valid[j] = r < S;
    const int lab = valid[j] ? ext[s] : C;
    lo4[j] = static_cast<uint32_t>(lab) * 4u;
    if (j & 1) {
      const int s2 = REV ? s + 2 : s - 2;                            // the state two steps "behind" in this direction
      msk[j >> 1] = (valid[j] && r >= 3 && ext[s] != ext[s2]) ? 1.0 : 0.0;
    }
    a[j] = 0.0;
    if (r == 0 || (r == 1 && S > 1)) a[j] = unpack_pd(lds32(prow + lo4[j]));
  }
  // store address of state j of row t: base + t*SP*4 + (REV ? S-1-r : r)*4
  const int st_lane = (REV ? (S - 1 - lane * K) : lane * K) * 4;
  const int st_row = (REV ? -SP : SP) * 4;
  long long st = static_cast<long long>(t0) * SP * 4 + st_lane;       // byte offset of (row, state j = 0)
  auto store_row = [&](const double (&v)[K]) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (valid[j]) {
        const int o = REV ? -4 * j : 4 * j;
        if (SM) sts32(ab_s + static_cast<uint32_t>(st) + o, pack_pd(v[j]));
        else ab_g[(st + o) >> 2] = pack_pd(v[j]);
      }
    }
  };
  // REV (beta) rows are stored WITHOUT the row's own emission probability: beta'_t(s) = beta_t(s) / p_t(l_s) is exactly
  // the sum the recursion forms before its multiply, and alpha_t(s) * beta'_t(s) is the posterior numerator - the
  // collect pass then needs no division (and no probability lookup) per state.  The boundary row's sums are 1.
  if (REV) {
    double one[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { const int r = lane * K + j; one[j] = (r == 0 || (r == 1 && S > 1)) ? 1.0 : 0.0; }
    store_row(one);
  } else {
    store_row(a);
  }
  // progress: one single-use mbarrier per 4 recursion steps (prog_s = shared address of this direction's array; the
  // two words in front of it hold the named-barrier thread counts).  Called after the rows of step i are stored, for
  // i % 4 == 0 (the renormalisation steps) and for the last step: the row of step t is covered by barrier (t+3)/4.
  // The arrive (release.cta) after a __syncwarp makes the rows and offsets visible to the collecting warps, which
  // sleep on named barriers until then (no polling while the recursions run alone): barrier 1 opens when both
  // directions have passed the middle row, barrier 2 when both are done.
  auto publish = [&](int i) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(prog_s + ((i + 3) >> 2) * 8) : "memory");
    const int lo = (Tb - 1) >> 1;
    const int mid = min(((REV ? Tb - 1 - lo : lo) + 3) & ~3, Tb - 1);
    if (i == mid) asm volatile("barrier.arrive 1, %0;" ::"r"(lds32(prog_s - 8)) : "memory");
    if (i == Tb - 1) {
      const uint32_t n_end = lds32(prog_s - 4);
      if (n_end > 64u) asm volatile("barrier.arrive 2, %0;" ::"r"(n_end) : "memory");
    }
  };
  int c = 0;
  if (lane == 0) sts32(off_s + t0 * 4, 0u);
  publish(0);
  int sh = fast_row_shift<K>(a);
  prow += row_step;
  if (Tb > 1) {
#pragma unroll
    for (int j = 0; j < K; ++j) q[j] = lds32(prow + lo4[j]);
  }
  for (int i = 1; i < Tb; ++i) {
    const int t = REV ? Tb - 1 - i : i;
    double pd[K];
#pragma unroll
    for (int j = 0; j < K; ++j) pd[j] = unpack_pd(q[j]);
    const int c_prev = c;                            // scale of the incoming row = scale of this row's sums
    if (sh != 0) {                                   // warp-uniform; exact power of two, off the dependency chain
      const double sc = pow2_neg(sh);
#pragma unroll
      for (int j = 0; j < K; ++j) pd[j] *= sc;
      c += sh;
    }
    if (lane == 0) sts32(off_s + t * 4, static_cast<uint32_t>(REV ? c_prev : c));
    double up1 = shfl_up_d(a[K - 1], 1);             // previous lane's last state
    if (lane == 0) up1 = 0.0;
    prow += row_step;
    if (i + 1 < Tb) {
#pragma unroll
      for (int j = 0; j < K; ++j) q[j] = lds32(prow + lo4[j]);
    }
    double n[K], su[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const double p1 = (j >= 1) ? a[j >= 1 ? j - 1 : 0] : up1;
      if (j & 1) {
        const double p2 = (j >= 2) ? a[j >= 2 ? j - 2 : 0] : up1;    // j == 1: two behind = previous lane's last state
        su[j] = fma(msk[j >> 1], p2, a[j] + p1);
      } else {
        su[j] = a[j] + p1;
      }
      n[j] = su[j] * pd[j];
    }
    st += st_row;
    if (REV) store_row(su); else store_row(n);
#pragma unroll
    for (int j = 0; j < K; ++j) a[j] = n[j];
    sh = 0;
    if ((i & (kFastRenorm - 1)) == 0) {
      sh = fast_row_shift<K>(n);
      publish(i);
    }
  }
  if (Tb > 1 && ((Tb - 1) & 3) != 0) publish(Tb - 1);
}

// packed word -> (mantissa float in [1, 2), exponent relative to 2^0); w != 0
__device__ __forceinline__ float pk_mant(uint32_t w) { return __uint_as_float(0x3F800000u | ((w & 0x3FFFFFu) << 1)); }
__device__ __forceinline__ int pk_exp(uint32_t w) { return static_cast<int>(w >> 22) - 1023; }
// packed probability -> fp32 (flushes below 2^-126: only used for the softmax term of the gradient)
__device__ __forceinline__ float pk_to_float(uint32_t w) {
  // exponents 897..1023 all carry bit 9, so (w << 1) keeps them modulo 512 and re-biasing is one subtraction;
  // anything at or below 2^-127 is clamped to the exponent field 0 (a denormal below 2^-126: nothing)
  return __uint_as_float((max(w, 896u << 22) << 1) - (384u << 23));
}

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

__device__ unsigned long long g_ctc_fallbacks = 0ull;
__device__ long long g_ctc_stamps[16];                      // HTRVT_CTC_DEBUG=1: clock64 at the phase boundaries of CTA 0
#define CTC_STAMP(i) do { if (P.dbg && b == 0 && (tid & 31) == 0) g_ctc_stamps[i] = clock64(); } while (0)      // sequences recomputed by the log-space path

#define CTC_DISPATCH_FAST(REV, SM, ...)                                  \
  switch (Kf) {                                                            \
    case 2: ctc_rec_fast<2, REV, SM>(__VA_ARGS__); break;                  \
    case 4: ctc_rec_fast<4, REV, SM>(__VA_ARGS__); break;                  \
    case 6: ctc_rec_fast<6, REV, SM>(__VA_ARGS__); break;                  \
    case 8: ctc_rec_fast<8, REV, SM>(__VA_ARGS__); break;                  \
    case 10: ctc_rec_fast<10, REV, SM>(__VA_ARGS__); break;                \
    case 12: ctc_rec_fast<12, REV, SM>(__VA_ARGS__); break;                \
    default: ctc_rec_fast<14, REV, SM>(__VA_ARGS__); break;                \
  }

__global__ void __launch_bounds__(kCtcThreads, 1) ctc_loss_grad_kernel(const CtcParams P) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x;
  if (P.only && !P.only[b]) return;                 // fix-up pass: the throughput kernel already did this sequence
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kCtcThreads / 32;
  const int T = P.T, C = P.C;
  const int ldp = (C + 1) | 1;                      // odd row stride with at least one pad column (column C)

  float* l2p = smem;                               // [T][ldp] log2-domain log-probabilities (fast path: packed doubles)
  float* post = l2p + T * ldp;                     // [NW][ldp] per-warp posterior row
  int* ext = reinterpret_cast<int*>(post + NW * ldp);   // [32*kmax] extended label sequence
  float* red = reinterpret_cast<float*>(ext + 32 * P.kmax);   // [40] reductions / broadcast
  double* offA = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(red + 40) + 7) & ~uintptr_t(7));         // [T] renormalisation offsets of alpha rows
  double* offB = offA + T;                                    // [T] ... of beta rows
  const int NB = ((T + 2) >> 2) + 2;                          // per direction: 1 count slot + progress barriers (4 steps each)
  uint64_t* bars = reinterpret_cast<uint64_t*>(offB + T);     // [2][NB]: {n_mid, n_end} | barriers
  float* AB = reinterpret_cast<float*>(bars + 2 * NB);        // alpha | beta when they fit in shared memory
  volatile int& s_bad = *reinterpret_cast<volatile int*>(red + 36);   // fast path: "recompute in log space" flag
  volatile int& s_badlabel = *reinterpret_cast<volatile int*>(red + 34);   // a target id outside [0, C)

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
  const float gs = P.grad_scale ? P.grad_scale[b] : P.grad_scale_const;

  if (tid == 0) {
    s_bad = 0;
    s_badlabel = 0;
    red[35] = __int_as_float(0);                                // row ticket of the collecting warps
  }
  for (int i = tid; i < 2 * NB; i += kCtcThreads) {
    if (i == 0 || i == NB) {
      // P.ovl: bit w set = warp w starts collecting at the middle row, clear = after the recursions (developer knob)
      const int n_early = P.grad ? __popc(static_cast<uint32_t>(P.ovl) & 0xFFFCu) : 0;
      reinterpret_cast<int2*>(bars)[i] = make_int2(32 * (2 + n_early), 32 * (2 + (P.grad ? 14 - n_early : 0)));
    } else {
      mbar_init(bars + i, 1);
    }
  }
  if (tid == 0) CTC_STAMP(0);
  __syncthreads();                                              // flags and barriers initialised before anyone uses them
  if (fits) {
    // labels index the staged probability rows and the per-class shared-memory atomics: an id outside [0, C) is clamped
    // (memory safety) and the sequence's nll comes back NaN - loud, like torch's device-side assert for such targets
    for (int s = tid; s < SP; s += kCtcThreads) {
      int lab = (s < S && (s & 1)) ? P.targets[toff + (s >> 1)] : 0;
      if (lab < 0 || lab >= C) { lab = min(max(lab, 0), C - 1); s_badlabel = 1; }
      ext[s] = lab;
    }
  }
  const bool run = fits && Tb > 0;

  // =============================== fast path (linear domain, FP64 pipe) ===========================
  const int Kf = ((S + 31) / 32 + 1) & ~1;             // even number of states per lane
  if (run && Kf <= 14 && !P.force_slow) {
    const uint32_t pk_s = smem_u32(l2p);
    const uint32_t ab_s = smem_u32(AB);
    uint32_t* Au = reinterpret_cast<uint32_t*>(A);
    uint32_t* Bu = reinterpret_cast<uint32_t*>(Bt);
    int* ioffA = reinterpret_cast<int*>(offA);
    int* ioffB = reinterpret_cast<int*>(offB);
    // ---- phase 0a: raw rows -> smem, every load of the CTA in flight at once --------------------
    if (((C & 3) == 0) && ((P.x_st & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0)) {
      const int C4 = C >> 2, n4 = Tb * C4;
#pragma unroll 4
      for (int i = tid; i < n4; i += kCtcThreads) {
        const int t = i / C4, c = (i - t * C4) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(xb + static_cast<long long>(t) * P.x_st + c));
        const uint32_t a4 = pk_s + (t * ldp + c) * 4;
        sts32(a4, __float_as_uint(v.x)); sts32(a4 + 4, __float_as_uint(v.y));
        sts32(a4 + 8, __float_as_uint(v.z)); sts32(a4 + 12, __float_as_uint(v.w));
      }
    } else {
      const int n = Tb * C;
#pragma unroll 8
      for (int i = tid; i < n; i += kCtcThreads) {
        const int t = i / C, c = i - t * C;
        sts32(pk_s + (t * ldp + c) * 4, __float_as_uint(__ldg(xb + static_cast<long long>(t) * P.x_st + c)));
      }
    }
    __syncthreads();
    if (tid == 0) CTC_STAMP(1);
    // ---- phase 0b: 4 threads per row: softmax -> packed double ---------------------------------------
    {
      const int part = tid & 3;
      bool oor = false;
      for (int t0 = 0; t0 < Tb; t0 += kCtcThreads / 4) {       // uniform trip count: the shuffles need whole warps
        const bool act = t0 + (tid >> 2) < Tb;
        const int t = act ? t0 + (tid >> 2) : Tb - 1;
        const uint32_t rowa = pk_s + t * ldp * 4;
        float lse2 = 0.f;
        if (!P.is_logprob) {
          float mx = -INFINITY;
#pragma unroll 5
          for (int c = part; c < C; c += 4) mx = fmaxf(mx, __uint_as_float(lds32(rowa + c * 4)));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float sum = 0.f;
          const float mxl = mx * kLog2e;
#pragma unroll 5
          for (int c = part; c < C; c += 4) sum += ex2f(fmaf(__uint_as_float(lds32(rowa + c * 4)), kLog2e, -mxl));
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          lse2 = mx * kLog2e + lg2f(sum);
        }
        if (!act) continue;                                    // (a padding thread may have read a row mid-rewrite)
#pragma unroll 5
        for (int c = part; c < C; c += 4) {
          const float l2 = __uint_as_float(lds32(rowa + c * 4)) * kLog2e - lse2;   // log2 p  (<= 0 up to rounding)
          uint32_t w = 0u;
          if (l2 > -1000.f) {
            const float fl = floorf(l2);
            const float m = fmaxf(ex2f(l2 - fl), 1.0f);      // [1, 2]
            const int X = 1023 + static_cast<int>(fl);
            w = (static_cast<uint32_t>(X) << 22) + ((__float_as_uint(m) - 0x3F800000u + 1u) >> 1);   // a carry bumps X
            if (X > 1023) oor = true;                        // p > 1: only a malformed "log-prob" input gets here
          } else if (l2 > -INFINITY || l2 != l2) {
            oor = true;                                      // finite but below the packed range, or NaN
          }
          sts32(rowa + c * 4, w);
        }
        if (part == 0) sts32(rowa + C * 4, 0u);              // the zero column out-of-range states read
      }
      if (oor) s_bad = 1;
    }
    __syncthreads();
    if (tid == 0) CTC_STAMP(2);
    if (!s_bad) {
      // ---- phase 1: alpha on warp 0, beta on warp 1, concurrently; every 4 steps they publish their progress ----
      // ---- phase 2 (overlapped): the other warps collect posterior rows as soon as BOTH directions have produced
      // them - from the middle of the sequence outwards while the recursions run their second half.  Rows are handed
      // out in completion order by an smem ticket; the likelihood the posteriors are normalised with comes from the
      // middle row (Z = sum_s alpha_m(s) beta'_m(s)), every warp computing the same sum for itself.
      const uint32_t progA_s = smem_u32(bars + 1), progB_s = smem_u32(bars + NB + 1);
      if (warp == 0) {
        if (P.scratch_in_smem) { CTC_DISPATCH_FAST(false, true, pk_s, ldp, C, ext, S, Tb, ab_s, Au, SP, smem_u32(ioffA), progA_s) }
        else { CTC_DISPATCH_FAST(false, false, pk_s, ldp, C, ext, S, Tb, 0u, Au, SP, smem_u32(ioffA), progA_s) }
        CTC_STAMP(3);
      } else if (warp == 1) {
        if (P.scratch_in_smem) { CTC_DISPATCH_FAST(true, true, pk_s, ldp, C, ext, S, Tb, ab_s + T * SP * 4, Bu, SP, smem_u32(ioffB), progB_s) }
        else { CTC_DISPATCH_FAST(true, false, pk_s, ldp, C, ext, S, Tb, 0u, Bu, SP, smem_u32(ioffB), progB_s) }
        CTC_STAMP(4);
      }
      if (gb) {
        auto wait_row = [&](int t) {             // alpha stores row t at step t, beta' at step Tb-1-t
          uint64_t* ba = bars + 1 + ((t + 3) >> 2);
          uint64_t* bb = bars + NB + 1 + ((Tb - 1 - t + 3) >> 2);
          while (!mbar_try_wait(ba, 0)) __nanosleep(100);
          while (!mbar_try_wait(bb, 0)) __nanosleep(100);
        };
        const int lo = (Tb - 1) >> 1, hi = Tb >> 1;           // first row(s) both directions reach
        if (warp >= 2) {
          const int2 cnt = reinterpret_cast<const int2*>(bars)[0];
          if ((static_cast<uint32_t>(P.ovl) >> warp) & 1u) asm volatile("barrier.sync 1, %0;" ::"r"(cnt.x) : "memory");
          else asm volatile("barrier.sync 2, %0;" ::"r"(cnt.y) : "memory");
        }
        const uint32_t pw_s = smem_u32(post + warp * ldp);
        const int NP = L + 1;                                 // pairs (2i, 2i+1); the last one has no label state
        constexpr int kPU = 3;                                // pairs per lane handled with labels in registers
        uint32_t lab4[kPU];
#pragma unroll
        for (int i = 0; i < kPU; ++i) lab4[i] = (lane + 32 * i < L) ? static_cast<uint32_t>(ext[2 * (lane + 32 * i) + 1]) * 4u : 0u;
        for (int c = lane; c < C; c += 32) sts32(pw_s + c * 4, 0u);
        __syncwarp();
        auto load_pair = [&](int t, int pi, uint2& wa, uint2& wb) {
          if (P.scratch_in_smem) {
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wa.x), "=r"(wa.y) : "r"(ab_s + (t * SP) * 4 + pi * 8));
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wb.x), "=r"(wb.y) : "r"(ab_s + ((T + t) * SP) * 4 + pi * 8));
          } else {
            wa = reinterpret_cast<const uint2*>(Au + t * SP)[pi];
            wb = reinterpret_cast<const uint2*>(Bu + t * SP)[pi];
          }
        };
        // ---- likelihood from the middle row ------------------------------------------------------
        wait_row(lo);
        double zsum = 0.0;
        for (int pi = lane; pi < NP; pi += 32) {
          uint2 wa, wb;
          load_pair(lo, pi, wa, wb);
          zsum += unpack_pd(wa.x) * unpack_pd(wb.x);
          if (pi < L) zsum += unpack_pd(wa.y) * unpack_pd(wb.y);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1)
          zsum += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(zsum), d),
                                   __shfl_xor_sync(0xffffffffu, __double2loint(zsum), d));
        const bool feas_m = zsum > 0.0;
        const int offm = ioffA[lo] + ioffB[lo];
        if (warp == 2 && lane == 0) {
          red[37] = __int_as_float(__double2hiint(zsum));
          red[38] = __int_as_float(__double2loint(zsum));
          red[39] = __int_as_float(offm);
        }
        // gamma_t(s) * 2^30 = (alpha_t(s) * c_t) * beta'_t(s) with ONE double constant per row,
        // c_t = 2^(30 + offA[t] + offB[t] - offA[m] - offB[m]) / Z_m: two DMULs per state on the FP64 pipe keep the
        // whole exponent range of the packed rows (alpha * c_t overflows only where beta' sits at the flush boundary),
        // and adding 2^52 leaves round(gamma * 2^30) in the low word - no conversion, no exponent arithmetic.
        // A lane takes (blank, label) state PAIRS: one 64-bit load per row, pair and direction.
        const double rza = feas_m ? 1.0 / zsum : 0.0;
        int* ticket = reinterpret_cast<int*>(red + 35);
        for (;;) {
          int k = 0;
          if (lane == 0) k = atomicAdd(ticket, 1);
          k = __shfl_sync(0xffffffffu, k, 0);
          if (k >= T) break;
          int t = k;
          if (k < Tb) {                                       // completion order: middle row(s) first, then outwards
            const int kq = k + (Tb & 1), j = kq >> 1;
            t = (kq & 1) ? hi + j : lo - j;
          }
          float* gr = gb + static_cast<long long>(t) * P.g_st;
          if (t >= Tb || !feas_m) {
            for (int c = lane; c < C; c += 32) gr[c] = 0.f;
            continue;
          }
          wait_row(t);
          const uint32_t pr_s = pk_s + t * ldp * 4;
          int kk = ioffA[t] + ioffB[t] - offm + 30;
          if (kk > 1000 || kk < -1000) { s_bad = 1; kk = 0; }  // cannot happen with consistent rows
          const double crow = rza * __hiloint2double((1023 + kk) << 20, 0);
          uint32_t blank = 0u, tot = 0u;
          auto fix30 = [&](uint32_t wa, uint32_t wb) -> uint32_t {
            const double q = (unpack_pd(wa) * crow) * unpack_pd(wb) + 4503599627370496.0;   // + 2^52
            return static_cast<uint32_t>(__double2loint(q));
          };
          auto add_pair = [&](bool okb, bool okl, uint32_t c4, uint2 wa, uint2 wb) {
            const uint32_t u0 = okb ? fix30(wa.x, wb.x) : 0u;
            const uint32_t u1 = okl ? fix30(wa.y, wb.y) : 0u;
            blank += u0;
            tot += u0 + u1;
            if (u1) reds_add_u32(pw_s + c4, u1);
          };
          {
            uint2 wa[kPU], wb[kPU];
#pragma unroll
            for (int i = 0; i < kPU; ++i) load_pair(t, min(lane + 32 * i, NP - 1), wa[i], wb[i]);   // all loads first
#pragma unroll
            for (int i = 0; i < kPU; ++i) add_pair(lane + 32 * i < NP, lane + 32 * i < L, lab4[i], wa[i], wb[i]);
          }
          for (int pi = lane + 32 * kPU; pi < NP; pi += 32) {
            uint2 wa, wb;
            load_pair(t, pi, wa, wb);
            add_pair(true, pi < L, pi < L ? static_cast<uint32_t>(ext[2 * pi + 1]) * 4u : 0u, wa, wb);
          }
          blank = __reduce_add_sync(0xffffffffu, blank);
          tot = __reduce_add_sync(0xffffffffu, tot);
          const int dev1 = static_cast<int>(tot) - (1 << 30);
          if (dev1 > 21475 || dev1 < -21475) s_bad = 1;    // the row's posteriors do not sum to 1 (2e-5): mass was flushed
          __syncwarp();
          for (int c = lane; c < C; c += 32) {
            const uint32_t pc = (c == 0) ? blank : lds32(pw_s + c * 4);
            sts32(pw_s + c * 4, 0u);                         // ready for this warp's next row
            gr[c] = (pk_to_float(lds32(pr_s + c * 4)) - static_cast<float>(pc) * 9.313225746154785e-10f) * gs;
          }
          __syncwarp();
        }
      }
      if (warp == 5) CTC_STAMP(5);
      if (!P.scratch_in_smem) __threadfence_block();
      __syncthreads();
      // ---- likelihood from both ends; all three (alpha end, beta start, middle row) must agree -------------
      const double za = unpack_pd(Au[(Tb - 1) * SP + S - 1]) + (S > 1 ? unpack_pd(Au[(Tb - 1) * SP + S - 2]) : 0.0);
      // beta rows hold beta' = beta / p: put row 0's emission probabilities back for the two start states
      const double zb = unpack_pd(Bu[0]) * unpack_pd(lds32(pk_s)) +
                        (S > 1 ? unpack_pd(Bu[1]) * unpack_pd(lds32(pk_s + ext[1] * 4)) : 0.0);
      const double zm = gb ? __hiloint2double(__float_as_int(red[37]), __float_as_int(red[38])) : za;
      const bool feasible = za > 0.0;
      bool bad = (za > 0.0) != (zb > 0.0) || (za > 0.0) != (zm > 0.0);
      double ll2 = 0.0;
      if (feasible && !bad) {
        // log2 of a packed value: exponent + lg2 of the 22-bit mantissa (abs err ~2e-7 of a bit)
        const uint32_t wa2 = pack_pd(za), wb2 = pack_pd(zb);
        ll2 = static_cast<double>(pk_exp(wa2) + ioffA[Tb - 1]) + static_cast<double>(lg2f(pk_mant(wa2)));
        const double ll2b = static_cast<double>(pk_exp(wb2) + ioffB[0]) + static_cast<double>(lg2f(pk_mant(wb2)));
        bad = fabs(ll2 - ll2b) > 3.0e-5;
        if (gb) {
          const uint32_t wm2 = pack_pd(zm);
          const double ll2m = static_cast<double>(pk_exp(wm2) + __float_as_int(red[39])) + static_cast<double>(lg2f(pk_mant(wm2)));
          bad = bad || fabs(ll2 - ll2m) > 3.0e-5;
        }
      }
      if (bad) s_bad = 1;
      __syncthreads();
      if (tid == 0) CTC_STAMP(6);
      if (!s_bad) {
        if (tid == 0) P.nll[b] = s_badlabel ? __int_as_float(0x7fc00000)
                                 : feasible ? static_cast<float>(-ll2 * 0.6931471805599453) : 0.f;   // zero_infinity=True
        return;
      }
    }
    if (tid == 0) atomicAdd(&g_ctc_fallbacks, 1ull);
    __syncthreads();
  }

  // =============================== log-space path ==================================================
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
    else if (K > P.kmax || s_badlabel) out = __int_as_float(0x7fc00000);   // provisioning error / bad target id: be loud
    P.nll[b] = out;
  }
  if (!gb) return;

  // ---- phase 2: posterior collect + gradient rows ----------------------------------------------
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
  const int ldp = (C + 1) | 1;
  size_t words = static_cast<size_t>(T) * ldp + (kCtcThreads / 32) * ldp + 32 * kmax + 40 + 4 * T + 2 +
                 4 * (((T + 2) >> 2) + 2);                  // + the progress barriers of the fast path
  if (scratch_in_smem) words += static_cast<size_t>(2) * T * 32 * kmax;
  return words * 4;
}

}  // namespace htrvt

using namespace htrvt;

namespace {
// Large batches go to the lane-group throughput kernel (ctc_grp.cu: G lanes per sequence, fp32 linear domain; labels up to
// 256).  HTRVT_CTC_TPUT / htrvt_ctc_set_mode: -1 automatic, 0 CTA-per-sequence kernel only, 1 lane-group kernel at any B.
int g_tput_mode = -2;                                 // -2: not set (read HTRVT_CTC_TPUT)
int tput_mode() {
  if (g_tput_mode == -2) {
    const char* env = getenv("HTRVT_CTC_TPUT");
    g_tput_mode = env ? atoi(env) : -1;
    if (g_tput_mode < -1 || g_tput_mode > 1) g_tput_mode = -1;
  }
  return g_tput_mode;
}
// measured on B200 (tools/ctc_grp_probe.py, T = 128, C = 80): the CTA-per-sequence kernel takes 23 us per wave of 148
// sequences (169 us at B = 1024, 334 us at 2048), the lane-group kernel ~170 us for anything up to one resident wave
// (B <= ~2400; 160 us with 8 states per lane) and 270 us at B = 4096: break-even at B ~ 1000
constexpr int kGrpMinBatch = 1024;
bool use_grp(int B, int T, int C, int lmax) {
  const int mode = tput_mode();
  if (mode == 0 || !ctc_grp_supported(B, T, C, lmax)) return false;
  return mode == 1 || B >= kGrpMinBatch;
}
size_t tput_flag_bytes(int B) { return (static_cast<size_t>(2 * B) * sizeof(int) + 255) & ~size_t(255); }   // flags + offsets
size_t fixup_bytes(int B, int T, int kmax) { return (static_cast<size_t>(B) * 2 * T * 32 * kmax * sizeof(float) + 255) & ~size_t(255); }
}  // namespace

// kernel selection: -1 automatic (lane-group throughput kernel for B >= 1024), 0 CTA-per-sequence kernel only,
// 1 lane-group kernel at any batch size (tests).  Returns the previous mode.
extern "C" int htrvt_ctc_set_mode(int mode) {
  const int prev = g_tput_mode == -2 ? -1 : g_tput_mode;
  g_tput_mode = mode < -1 || mode > 1 ? -1 : mode;
  return prev;
}

extern "C" size_t htrvt_ctc_workspace_bytes(int B, int T, int C, int max_target_len) {
  int lmax = max_target_len < 0 || max_target_len > T ? T : max_target_len;
  const int kmax = ctc_round_k((2 * lmax + 1 + 31) / 32);
  if (use_grp(B, T, C, lmax))       // flags + the fix-up pass's rows (touched only for flagged sequences) + the alpha slots
    return tput_flag_bytes(B) + fixup_bytes(B, T, kmax) + ctc_grp_scratch_bytes(B, T, C, lmax, ctc_num_sms());
  if (ctc_smem_bytes(T, C, kmax, true) <= 227 * 1024) return 0;
  return static_cast<size_t>(B) * 2 * T * 32 * kmax * sizeof(float);
}

// number of sequences the fast (linear fp64) path handed to the log-space path since the library was loaded
extern "C" int htrvt_ctc_debug_stamps(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_ctc_stamps, sizeof(long long) * 16) == cudaSuccess ? HTRVT_OK : HTRVT_ERR_LAUNCH;
}
extern "C" long long htrvt_ctc_fallback_count(void) {
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, g_ctc_fallbacks, sizeof(v)) != cudaSuccess) return -1;
  return static_cast<long long>(v);
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
  static const int force_slow = (getenv("HTRVT_CTC_SLOW") && atoi(getenv("HTRVT_CTC_SLOW")) != 0) ? 1 : 0;
  P.force_slow = force_slow;
  static const int dbg = getenv("HTRVT_CTC_DEBUG") ? 1 : 0;
  P.dbg = dbg;
  static const int ovl = getenv("HTRVT_CTC_OVL") ? static_cast<int>(strtol(getenv("HTRVT_CTC_OVL"), nullptr, 0)) : 0xFFFC;
  P.ovl = ovl;   // developer knob: run the log-space path only
  P.only = nullptr;
  if (!force_slow && use_grp(B, T, C, lmax)) {
    // throughput path: lane-group kernel, then the CTA-per-sequence kernel as a fix-up pass that exits at once for every
    // sequence the first kernel completed (flags[b] == 0)
    const size_t fb = tput_flag_bytes(B), fx = fixup_bytes(B, T, P.kmax);
    const size_t need = fb + fx + ctc_grp_scratch_bytes(B, T, C, lmax, ctc_num_sms());
    if (!workspace || workspace_bytes < need) return HTRVT_ERR_WORKSPACE;
    int* flags = static_cast<int*>(workspace);
    P.scratch = nullptr; P.scratch_in_smem = 0;
    if (P.tgt_stride <= 0) {
      int r0 = ctc_offsets_launch(P.target_lengths, B, flags + B, stream);
      if (r0) return r0;
    }
    int r = ctc_grp_launch(P, lmax, flags, flags + B, static_cast<char*>(workspace) + fb + fx, ctc_num_sms(), stream);
    if (r) return r;
    P.only = flags;
    P.scratch = reinterpret_cast<float*>(static_cast<char*>(workspace) + fb);
    const size_t smem_fix = ctc_smem_bytes(T, C, P.kmax, false);
    if (smem_fix > 227 * 1024) return HTRVT_ERR_SHAPE;
    if (!HTRVT_ENSURE_SMEM(ctc_loss_grad_kernel, 227 * 1024)) return HTRVT_ERR_LAUNCH;
    ctc_loss_grad_kernel<<<B, kCtcThreads, smem_fix, stream>>>(P);
    HTRVT_LAUNCH_CHECK();
    return HTRVT_OK;
  }
  P.scratch_in_smem = ctc_smem_bytes(T, C, P.kmax, true) <= 227 * 1024 ? 1 : 0;
  P.scratch = static_cast<float*>(workspace);
  const size_t smem = ctc_smem_bytes(T, C, P.kmax, P.scratch_in_smem);
  if (smem > 227 * 1024) return HTRVT_ERR_SHAPE;     // T*C slab itself does not fit
  if (!P.scratch_in_smem) {
    const size_t need = static_cast<size_t>(B) * 2 * T * 32 * P.kmax * sizeof(float);
    if (!workspace || workspace_bytes < need) return HTRVT_ERR_WORKSPACE;
  }
  if (!HTRVT_ENSURE_SMEM(ctc_loss_grad_kernel, 227 * 1024)) return HTRVT_ERR_LAUNCH;
  ctc_loss_grad_kernel<<<B, kCtcThreads, smem, stream>>>(P);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
