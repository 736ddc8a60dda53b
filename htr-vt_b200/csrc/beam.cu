// K-best CTC alignment paths for sm_100a: the per-frame beam of the reference's LM-rescoring evaluation
// (model_window/test_with_kenlm.py:25-59, `simple_ctc_beam_search_with_lm`) as one kernel, one warp per line.
//
// What the reference does per line, on the host, with ~T * K^2 Python list operations and one D2H copy per frame:
//   beams = [([], 0.0)]
//   for t: top_c = the K most probable classes of frame t (np.argsort(probs)[-K:][::-1]);
//          every beam is extended by every class of top_c (score += log_prob), the K best extensions survive
//          (Python's stable sort, descending: among equal scores the earlier (beam, class) pair wins)
//   every surviving PATH is collapsed (drop blanks and repeats), decoded and handed to the language model.
// Scores are sums of per-frame terms, so the K survivors of frame t are exactly the K best paths of length t+1.
//
// Here: the frame's top-K classes by K rounds of a warp arg-max over register-resident log-probs, the <= K*K
// candidate scores spread over the lanes (float64 accumulation, as numpy 1.24 - the reference's pinned version -
// promotes `0.0 + np.float32`), K rounds of a warp arg-max with the reference's tie order, back-pointers in shared
// memory, then K lanes backtrack and collapse their paths in place.  Output: ids[B,K,T] (collapsed, zero padded),
// lens[B,K], scores[B,K] in the reference's beam order; id -> char and the LM stay on the host.
// Tie rule for equal class log-probs inside a frame: the higher class index ranks first (numpy's stable order
// reversed; the reference's default introsort leaves it unspecified).
#include "common.cuh"

namespace htrvt {

constexpr int kBeamMaxK = 8;        // beams (K*K <= 64 candidates = 2 per lane)
constexpr int kBeamCpl = 8;         // classes per lane held in registers: C <= 256

struct Cand {
  double s;
  int n;                            // candidate number beam * KC + class rank; smaller wins ties
};
__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
  return a.s > b.s || (a.s == b.s && a.n < b.n);
}

__global__ void __launch_bounds__(128) ctc_kbest_kernel(const float* __restrict__ x, long long sb, long long st,
                                                        const int* __restrict__ lengths, int B, int T, int C, int K,
                                                        int* __restrict__ ids, int* __restrict__ lens,
                                                        double* __restrict__ scores) {
  extern __shared__ uint32_t bp_all[];                        // [warps][T][kBeamMaxK]: parent << 16 | class
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  uint32_t* bp = bp_all + static_cast<size_t>(warp) * T * kBeamMaxK;
  int Tb = lengths ? lengths[b] : T;
  Tb = min(max(Tb, 0), T);
  const float* xb = x + static_cast<long long>(b) * sb;
  const int KC = min(K, C);                                    // classes taken per frame
  double bs[kBeamMaxK];                                        // beam scores, replicated in every lane
#pragma unroll
  for (int i = 0; i < kBeamMaxK; ++i) bs[i] = 0.0;
  int nb = 1;                                                  // live beams
  float v[kBeamCpl], vn[kBeamCpl];                            // this frame's log-probs, the next frame's (prefetch)
#pragma unroll
  for (int i = 0; i < kBeamCpl; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < C && Tb > 0) ? __ldg(xb + c) : -INFINITY;
    vn[i] = -INFINITY;
  }
  for (int t = 0; t < Tb; ++t) {
    if (t + 1 < Tb) {
      const float* xr = xb + static_cast<long long>(t + 1) * st;
#pragma unroll
      for (int i = 0; i < kBeamCpl; ++i) {
        const int c = lane + 32 * i;
        if (c < C) vn[i] = __ldg(xr + c);
      }
    }
    // ---- the frame's KC best classes (value desc, index desc) ----------------------------------
    unsigned taken = 0u;
    float topv[kBeamMaxK];
    int topc[kBeamMaxK];
#pragma unroll
    for (int j = 0; j < kBeamMaxK; ++j) { topv[j] = -INFINITY; topc[j] = 0; }
#pragma unroll
    for (int j = 0; j < kBeamMaxK; ++j) {
      if (j < KC) {
        float bv = -INFINITY;
        int bc = -1;
#pragma unroll
        for (int i = 0; i < kBeamCpl; ++i) {
          const int c = lane + 32 * i;
          if (c < C && !((taken >> i) & 1u) && (bc < 0 || v[i] > bv || (v[i] == bv && c > bc))) { bv = v[i]; bc = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
          if (oc >= 0 && (bc < 0 || ov > bv || (ov == bv && oc > bc))) { bv = ov; bc = oc; }
        }
        topv[j] = bv; topc[j] = bc;
        if (bc >= 0 && (bc & 31) == lane) taken |= 1u << (bc >> 5);
      }
    }
    // ---- candidates n = beam * KC + rank, two per lane; the K best survive ----------------------
    const int ncand = nb * KC;
    Cand mine[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n = lane + 32 * q;
      mine[q].n = n;
      mine[q].s = -INFINITY;
      if (n < ncand) {
        const int i = n / KC, j = n - i * KC;
        double si = 0.0;
        float pj = 0.f;
#pragma unroll
        for (int u = 0; u < kBeamMaxK; ++u) {                 // register arrays: select, do not index
          if (u == i) si = bs[u];
          if (u == j) pj = topv[u];
        }
        mine[q].s = si + static_cast<double>(pj);
      }
    }
    const int nnew = min(K, ncand);
    double nbs[kBeamMaxK];
#pragma unroll
    for (int r = 0; r < kBeamMaxK; ++r) nbs[r] = 0.0;
    unsigned used = 0u;
#pragma unroll
    for (int r = 0; r < kBeamMaxK; ++r) {
      if (r < nnew) {
        Cand best{-INFINITY, 0x7fffffff};
        bool have = false;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (mine[q].n < ncand && !((used >> q) & 1u) && (!have || cand_better(mine[q], best))) { best = mine[q]; have = true; }
        }
        if (!have) best.n = 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          Cand other;
          other.s = __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(best.s), o),
                                     __shfl_xor_sync(0xffffffffu, __double2loint(best.s), o));
          other.n = __shfl_xor_sync(0xffffffffu, best.n, o);
          if (other.n != 0x7fffffff && (best.n == 0x7fffffff || cand_better(other, best))) best = other;
        }
        // every lane now knows the winner
        if ((best.n & 31) == lane) used |= 1u << (best.n >> 5);
        nbs[r] = best.s;
        const int i = best.n / KC, j = best.n - i * KC;
        int cj = 0;
#pragma unroll
        for (int u = 0; u < kBeamMaxK; ++u)
          if (u == j) cj = topc[u];
        if (lane == 0) bp[t * kBeamMaxK + r] = (static_cast<uint32_t>(i) << 16) | static_cast<uint32_t>(cj);
      }
    }
#pragma unroll
    for (int r = 0; r < kBeamMaxK; ++r) bs[r] = nbs[r];
#pragma unroll
    for (int i = 0; i < kBeamCpl; ++i) v[i] = vn[i];
    nb = nnew;
  }
  __syncwarp();
  // ---- backtrack + collapse: lane r owns beam r ---------------------------------------------------
  if (lane < K) {
    int* out = ids + (static_cast<long long>(b) * K + lane) * T;
    int n = 0;
    double sc = 0.0;
    const bool live = Tb > 0 ? lane < nb : lane == 0;          // T = 0: the single empty beam
    if (live && Tb > 0) {
      int cur = lane;
      for (int t = Tb - 1; t >= 0; --t) {
        const uint32_t e = bp[t * kBeamMaxK + cur];
        out[t] = static_cast<int>(e & 0xFFFFu);
        cur = static_cast<int>(e >> 16);
      }
      int prev = -1;                                           // test_with_kenlm.py:46-51
      for (int t = 0; t < Tb; ++t) {
        const int id = out[t];
        if (id != 0 && id != prev) out[n++] = id;
        prev = id;
      }
#pragma unroll
      for (int u = 0; u < kBeamMaxK; ++u)
        if (u == lane) sc = bs[u];
    }
    for (int t = n; t < T; ++t) out[t] = 0;
    lens[b * K + lane] = live ? n : -1;                        // -1: this beam does not exist (fewer than K paths)
    scores[b * K + lane] = live ? sc : -INFINITY;
  }
}

}  // namespace htrvt

using namespace htrvt;

// ids int32 [B, K, T], lens int32 [B, K] (-1 = no such beam), scores float64 [B, K]; log_probs [B?, T?, C] through
// element strides (class axis contiguous), so the reference's [T, B, C] layout needs no copy.
extern "C" int htrvt_ctc_kbest_paths(const float* log_probs, long long stride_b, long long stride_t, const int* lengths,
                                     int B, int T, int C, int K, int* ids, int* lens, double* scores,
                                     cudaStream_t stream) {
  if (B <= 0 || T <= 0 || C <= 0 || !log_probs || !ids || !lens || !scores) return HTRVT_ERR_SHAPE;
  if (K < 1 || K > kBeamMaxK || C > 32 * kBeamCpl || C > 65535 || T > 4096) return HTRVT_ERR_SHAPE;
  const int warps = 4;
  const size_t smem = static_cast<size_t>(warps) * T * kBeamMaxK * sizeof(uint32_t);
  if (smem > 227 * 1024) return HTRVT_ERR_SHAPE;
  if (smem > 48 * 1024 && !HTRVT_ENSURE_SMEM(ctc_kbest_kernel, 227 * 1024)) return HTRVT_ERR_LAUNCH;
  ctc_kbest_kernel<<<(B + warps - 1) / warps, warps * 32, smem, stream>>>(log_probs, stride_b, stride_t, lengths, B, T, C,
                                                                        K, ids, lens, scores);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
