// CTC prefix beam search for sm_100a, one warp per line (the search that replaces the toy
// per-frame beam of model_window/test_with_kenlm.py:25-59 in the LM-rescoring evaluation).
//
// The reference's `simple_ctc_beam_search_with_lm` ranks ALIGNMENT PATHS (ctc_kbest_kernel, beam.cu, restates it bit
// for bit); this kernel ranks LABELLINGS: every beam entry is a collapsed prefix l with the log-probability mass of
// all alignments of the frames seen so far that collapse to l, split into "ends in blank" (pb) and "ends in its last
// label" (pnb).  Per frame, with lp = the frame's log-probs and tot = lae(pb, pnb)  (lae = log-add-exp):
//   stay   l      : pb' = tot + lp[0];  pnb' = pnb + lp[last(l)]
//   extend l + c  : pnb' = (c == last(l) ? pb : tot) + lp[c];  pb' = -inf          for every label c >= 1
//   if l + c is itself an entry of the beam, its mass is ADDED to that entry's pnb' (one prefix, one entry)
//   the K entries with the largest lae(pb', pnb') survive; ties: stay entries (by rank) before extensions (by
//   parent rank, then label).
// With a beam wide enough to hold every prefix this is the exact labelling posterior (tests pin it to a brute-force
// enumeration of all alignments); all classes are extended, nothing is pruned by a per-frame class cut-off.
//
// Layout: the frame's log-probs live in registers (8 classes per lane) and in a per-warp smem row (for the stay
// terms); the two beam generations, the stay candidates and the winners of the selection rounds in per-warp smem;
// prefixes are identified by a 64-bit hash chain (the merge test is hash(parent) == hash of an entry one label
// shorter) and spelled out at the end from a per-line node array (parent node << 16 | label) in smem.  Scores are
// float64 in log space.  Selection = K rounds of a warp arg-max over the nb * (C - 1) + nb candidates in the strict
// total order (score desc, candidate number asc); a round looks for the best candidate that comes AFTER the previous
// winner in that order, so nothing has to be marked as taken.
#include "common.cuh"
#include <climits>

namespace htrvt {

constexpr int kPbMaxK = 16;         // beam entries
constexpr int kPbCpl = 8;           // classes per lane held in registers: C <= 256

struct PbGen {                      // one beam generation of one line
  double pb[kPbMaxK], pnb[kPbMaxK], tot[kPbMaxK];
  unsigned long long hash[kPbMaxK], phash[kPbMaxK];
  int last[kPbMaxK], len[kPbMaxK], node[kPbMaxK];
};
struct PbWarp {
  PbGen gen[2];
  double spb[kPbMaxK], spnb[kPbMaxK], stot[kPbMaxK];        // the stay candidate of every entry
  double win_s[kPbMaxK];
  int win_n[kPbMaxK];
  int mrg_i[kPbMaxK], mrg_c[kPbMaxK];                       // entry j absorbed the extension (mrg_i, mrg_c), or -1
  float lp[32 * kPbCpl];
};

__device__ __forceinline__ double lae(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  return fmax(a, b) + log1p(exp(-fabs(a - b)));
}
__device__ __forceinline__ unsigned long long pb_hash(unsigned long long h, int c) {
  h = (h ^ (static_cast<unsigned long long>(c) + 0x9e3779b97f4a7c15ull)) * 0xff51afd7ed558ccdull;
  return h ^ (h >> 32);
}
// order-preserving integer image of a double (-0.0 folded onto +0.0), its inverse, and the candidate order
// (score descending, candidate number ascending)
__device__ __forceinline__ long long pb_key(double s) {
  long long b = __double_as_longlong(s);
  if ((b << 1) == 0) b = 0;
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double pb_unkey(long long k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}
__device__ __forceinline__ bool pb_before(long long k, int n, long long bk, int bn) {  // (k, n) ranks before (bk, bn)
  return k > bk || (k == bk && n < bn);
}

template <int CPL>                  // class slots per lane: C <= 32 * CPL
__global__ void __launch_bounds__(128) ctc_prefix_beam_kernel(const float* __restrict__ x, long long sb, long long st,
                                                              const int* __restrict__ lengths, int B, int T, int C,
                                                              int K, int* __restrict__ ids, int* __restrict__ lens,
                                                              double* __restrict__ scores) {
  extern __shared__ __align__(16) uint8_t pb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int b = blockIdx.x * warps + warp;
  if (b >= B) return;
  PbWarp* W = reinterpret_cast<PbWarp*>(pb_smem) + warp;
  uint32_t* nodes = reinterpret_cast<uint32_t*>(pb_smem + sizeof(PbWarp) * warps) +
                    static_cast<size_t>(warp) * (static_cast<size_t>(T) * K + 1);
  int Tb = lengths ? lengths[b] : T;
  Tb = min(max(Tb, 0), T);
  const float* xb = x + static_cast<long long>(b) * sb;

  int cur = 0, nb = 1, nnodes = 1;
  if (lane == 0) {
    PbGen& g = W->gen[0];
    g.pb[0] = 0.0; g.pnb[0] = -INFINITY; g.tot[0] = 0.0;
    g.hash[0] = 0x243f6a8885a308d3ull; g.phash[0] = 0ull;
    g.last[0] = -1; g.len[0] = 0; g.node[0] = 0;
    nodes[0] = 0u;
  }
  float v[CPL], vn[CPL];
#pragma unroll
  for (int u = 0; u < CPL; ++u) {
    const int c = lane + 32 * u;
    v[u] = (c < C && Tb > 0) ? __ldg(xb + c) : -INFINITY;
    vn[u] = -INFINITY;
  }
  __syncwarp();

  for (int t = 0; t < Tb; ++t) {
    if (t + 1 < Tb) {
      const float* xr = xb + static_cast<long long>(t + 1) * st;
#pragma unroll
      for (int u = 0; u < CPL; ++u) {
        const int c = lane + 32 * u;
        if (c < C) vn[u] = __ldg(xr + c);
      }
    }
#pragma unroll
    for (int u = 0; u < CPL; ++u) W->lp[lane + 32 * u] = v[u];
    __syncwarp();
    const PbGen& g = W->gen[cur];
    PbGen& gn = W->gen[cur ^ 1];

    // ---- stay candidates; an extension that spells an existing entry is folded into that entry ----------
    if (lane < nb) {
      const int j = lane;
      const double totj = g.tot[j];
      const double npb = totj + static_cast<double>(W->lp[0]);
      double npnb = g.len[j] > 0 ? g.pnb[j] + static_cast<double>(W->lp[g.last[j]]) : -INFINITY;
      int mi = -1;
      if (g.len[j] > 0) {
        const unsigned long long ph = g.phash[j];
        for (int i = 0; i < nb; ++i)
          if (g.hash[i] == ph && g.len[i] == g.len[j] - 1) mi = i;
      }
      if (mi >= 0) {
        const int c = g.last[j];
        const double base = (g.len[mi] > 0 && g.last[mi] == c) ? g.pb[mi] : g.tot[mi];
        npnb = lae(npnb, base + static_cast<double>(W->lp[c]));
      }
      W->mrg_i[j] = mi;
      W->mrg_c[j] = g.last[j];
      W->spb[j] = npb;
      W->spnb[j] = npnb;
      W->stot[j] = lae(npb, npnb);
    }
    __syncwarp();
    unsigned excl[CPL];                                   // bit i of excl[u]: extension (i, lane + 32 u) was folded
#pragma unroll
    for (int u = 0; u < CPL; ++u) {                       // the blank and classes beyond C are never extensions
      const int c = lane + 32 * u;
      excl[u] = (c >= 1 && c < C) ? 0u : 0xffffffffu;
    }
    for (int j = 0; j < nb; ++j) {
      const int mi = W->mrg_i[j], mc = W->mrg_c[j];
      if (mi >= 0 && (mc & 31) == lane) {
#pragma unroll
        for (int u = 0; u < CPL; ++u)
          if (u == (mc >> 5)) excl[u] |= 1u << mi;
      }
    }

    // ---- K selection rounds ---------------------------------------------------------------------------------
    // Scores are compared as order-preserving 64-bit integer images of the doubles; the sentinel (LLONG_MIN, INT_MAX)
    // loses against every real candidate.  Every lane scans ITS candidates (its class slots x all entries, + its
    // stay candidate) ONCE per frame for its local best; a round is then a warp arg-max over the 32 local bests, and
    // only the winner's lane needs a new local best (the best of its candidates that comes after the winner in the
    // strict order) - found by all 32 lanes together, <= ceil((nb * CPL + 1) / 32) candidates each, from the smem
    // copy of the frame's log-probs.  K scans of nb * CPL candidates per lane become one.
    long long lk = LLONG_MIN;                                // this lane's best remaining candidate
    int ln = 0x7fffffff;
    {
      long long bku[CPL];
      int bnu[CPL];
#pragma unroll
      for (int u = 0; u < CPL; ++u) { bku[u] = LLONG_MIN; bnu[u] = 0x7fffffff; }
      if (lane < nb) { bku[0] = pb_key(W->stot[lane]); bnu[0] = lane; }
#pragma unroll 4
      for (int i = 0; i < nb; ++i) {
        const double ti = g.tot[i], pbi = g.pb[i];
        const int li = g.len[i] > 0 ? g.last[i] : -1;
        const int n0 = K + i * C;
#pragma unroll
        for (int u = 0; u < CPL; ++u) {
          const int c = lane + 32 * u;
          if (!((excl[u] >> i) & 1u)) {
            const long long k = pb_key((c == li ? pbi : ti) + static_cast<double>(v[u]));
            if (pb_before(k, n0 + c, bku[u], bnu[u])) { bku[u] = k; bnu[u] = n0 + c; }
          }
        }
      }
      lk = bku[0]; ln = bnu[0];
#pragma unroll
      for (int u = 1; u < CPL; ++u)
        if (pb_before(bku[u], bnu[u], lk, ln)) { lk = bku[u]; ln = bnu[u]; }
    }
    int nnew = 0;
    for (int r = 0; r < K; ++r) {
      long long bk = lk;
      int bn = ln;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int on = __shfl_xor_sync(0xffffffffu, bn, o);
        if (pb_before(ok, on, bk, bn)) { bk = ok; bn = on; }
      }
      if (bn == 0x7fffffff) break;                            // fewer than K candidates exist
      if (lane == 0) { W->win_s[r] = pb_unkey(bk); W->win_n[r] = bn; }
      nnew = r + 1;
      if (r + 1 == K) break;
      // the owner lane of the winner and its class slots' fold masks
      const int own = bn < K ? bn : (((bn - K) % C) & 31);
      unsigned oex[CPL];
#pragma unroll
      for (int u = 0; u < CPL; ++u) oex[u] = __shfl_sync(0xffffffffu, excl[u], own);
      long long ck = LLONG_MIN;
      int cn = 0x7fffffff;
      const int M = nb * CPL;
      for (int idx = lane; idx <= M; idx += 32) {
        long long k;
        int n;
        bool ok;
        if (idx == M) {                                       // the owner's stay candidate
          ok = own < nb;
          k = ok ? pb_key(W->stot[own]) : LLONG_MIN;
          n = own;
        } else {
          const int i = idx / CPL, u = idx - i * CPL;
          const int c = own + 32 * u;
          unsigned ex = 0u;
#pragma unroll
          for (int q = 0; q < CPL; ++q)
            if (q == u) ex = oex[q];
          ok = !((ex >> i) & 1u);                             // (invalid classes carry an all-ones mask)
          const int li = g.len[i] > 0 ? g.last[i] : -1;
          k = ok ? pb_key((c == li ? g.pb[i] : g.tot[i]) + static_cast<double>(W->lp[c])) : LLONG_MIN;
          n = K + i * C + c;
        }
        if (ok && pb_before(bk, bn, k, n) && pb_before(k, n, ck, cn)) { ck = k; cn = n; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const long long ok2 = __shfl_xor_sync(0xffffffffu, ck, o);
        const int on = __shfl_xor_sync(0xffffffffu, cn, o);
        if (pb_before(ok2, on, ck, cn)) { ck = ok2; cn = on; }
      }
      if (lane == own) { lk = ck; ln = cn; }
    }
    __syncwarp();

    // ---- the next generation: lane r builds entry r -------------------------------------------------------
    int n = 0, pi = 0, c = 0;
    bool ext = false;
    if (lane < nnew) {
      n = W->win_n[lane];
      ext = n >= K;
      if (ext) { pi = (n - K) / C; c = (n - K) - pi * C; }
    }
    const unsigned em = __ballot_sync(0xffffffffu, ext);
    if (lane < nnew) {
      if (!ext) {
        gn.pb[lane] = W->spb[n]; gn.pnb[lane] = W->spnb[n]; gn.tot[lane] = W->win_s[lane];
        gn.hash[lane] = g.hash[n]; gn.phash[lane] = g.phash[n];
        gn.last[lane] = g.last[n]; gn.len[lane] = g.len[n]; gn.node[lane] = g.node[n];
      } else {
        const int nd = nnodes + __popc(em & ((1u << lane) - 1u));
        nodes[nd] = (static_cast<uint32_t>(g.node[pi]) << 16) | static_cast<uint32_t>(c);
        gn.pb[lane] = -INFINITY; gn.pnb[lane] = W->win_s[lane]; gn.tot[lane] = W->win_s[lane];
        gn.hash[lane] = pb_hash(g.hash[pi], c); gn.phash[lane] = g.hash[pi];
        gn.last[lane] = c; gn.len[lane] = g.len[pi] + 1; gn.node[lane] = nd;
      }
    }
    nnodes += __popc(em);
    nb = nnew;
    cur ^= 1;
#pragma unroll
    for (int u = 0; u < CPL; ++u) v[u] = vn[u];
    __syncwarp();
  }

  // ---- spell the surviving prefixes --------------------------------------------------------------------------
  if (lane < K) {
    const PbGen& g = W->gen[cur];
    int* out = ids + (static_cast<long long>(b) * K + lane) * T;
    const bool live = lane < nb;
    int n = 0;
    if (live) {
      n = g.len[lane];
      int nd = g.node[lane];
      for (int p = n - 1; p >= 0; --p) {
        const uint32_t e = nodes[nd];
        out[p] = static_cast<int>(e & 0xFFFFu);
        nd = static_cast<int>(e >> 16);
      }
    }
    for (int p = n; p < T; ++p) out[p] = 0;
    lens[b * K + lane] = live ? n : -1;
    scores[b * K + lane] = live ? g.tot[lane] : -INFINITY;
  }
}

}  // namespace htrvt

using namespace htrvt;

// ids int32 [B, K, T] (labels of the K most probable prefixes, zero padded), lens int32 [B, K] (-1: fewer than K
// prefixes exist), scores float64 [B, K] = log P(labelling's alignments over the line's frames), best first.
// log_probs fp32 through element strides (class axis contiguous).  K <= 16, C <= 256, T * K < 65535.
extern "C" int htrvt_ctc_prefix_beam(const float* log_probs, long long stride_b, long long stride_t,
                                     const int* lengths, int B, int T, int C, int K, int* ids, int* lens,
                                     double* scores, cudaStream_t stream) {
  if (B <= 0 || T <= 0 || C <= 0 || !log_probs || !ids || !lens || !scores) return HTRVT_ERR_SHAPE;
  if (K < 1 || K > kPbMaxK || C > 32 * kPbCpl || static_cast<long long>(T) * K >= 65535) return HTRVT_ERR_SHAPE;
  int warps = 4;
  auto need = [&](int w) { return static_cast<size_t>(w) * (sizeof(PbWarp) + (static_cast<size_t>(T) * K + 1) * 4); };
  while (warps > 1 && need(warps) > 200 * 1024) warps >>= 1;
  const size_t smem = need(warps);
  if (smem > 227 * 1024) return HTRVT_ERR_SHAPE;
  const dim3 grid((B + warps - 1) / warps), block(warps * 32);
#define PB_LAUNCH(CPL)                                                                                         \
  do {                                                                                                         \
    if (smem > 48 * 1024 && !HTRVT_ENSURE_SMEM(ctc_prefix_beam_kernel<CPL>, 227 * 1024)) return HTRVT_ERR_LAUNCH; \
    ctc_prefix_beam_kernel<CPL><<<grid, block, smem, stream>>>(log_probs, stride_b, stride_t, lengths, B, T, C, K, \
                                                               ids, lens, scores);                             \
  } while (0)
  if (C <= 32) PB_LAUNCH(1);
  else if (C <= 64) PB_LAUNCH(2);
  else if (C <= 96) PB_LAUNCH(3);
  else if (C <= 128) PB_LAUNCH(4);
  else PB_LAUNCH(8);
#undef PB_LAUNCH
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
