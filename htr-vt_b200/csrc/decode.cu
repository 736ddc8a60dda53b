// Greedy CTC decoding for sm_100a: per-frame argmax + collapse (drop blanks, repeats and
// out-of-alphabet ids) + compaction in ONE kernel, one CTA per line; a single D2H copy of
// ids[B,T] / lens[B] then replaces the reference's per-element device syncs.
//
// Replaces model_v1/valid.py:40-42 (`preds.max(2)`, transpose, view) and the id filtering of
// CTCLabelConverter.decode (model_v1/utils/utils.py:72-86); id -> char stays on the host.
// Semantics: argmax = lowest index among ties, a NaN beats any number (first NaN);
// keep id iff id != 0 and id != previous RAW frame id and id < n_character.
#include "common.cuh"

namespace htrvt {

constexpr int kDecThreads = 128;

struct Best {
  float v;
  int i;
};
__device__ __forceinline__ bool better(const Best& a, const Best& b) {   // is a strictly preferred over b
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an != bn) return an;
  if (an) return a.i < b.i;
  return a.v > b.v || (a.v == b.v && a.i < b.i);
}

// Shared tail: collapse + compact raw[0..Tb) (shared memory) into ids_out, return count (warp 0 only).
__device__ __forceinline__ int collapse_compact(const int* raw, int Tb, int n_character, int* ids_out) {
  const int lane = threadIdx.x & 31;
  int count = 0;
  for (int base = 0; base < Tb; base += 32) {
    const int t = base + lane;
    bool keep = false;
    int id = 0;
    if (t < Tb) {
      id = raw[t];
      keep = id != 0 && !(t > 0 && raw[t - 1] == id) && id < n_character && id >= 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) ids_out[count + __popc(m & ((1u << lane) - 1u))] = id;
    count += __popc(m);
  }
  return count;
}

__global__ void __launch_bounds__(kDecThreads) greedy_decode_kernel(
    const float* __restrict__ x, long long sb, long long st, int B, int T, int C, const int* __restrict__ lengths,
    int n_character, int* __restrict__ ids, int* __restrict__ lens, int* __restrict__ raw_out) {
  extern __shared__ int raw[];                     // [T]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int Tb = lengths ? lengths[b] : T;
  Tb = min(max(Tb, 0), T);
  const float* xb = x + static_cast<long long>(b) * sb;
  for (int t = warp; t < Tb; t += kDecThreads / 32) {
    const float* xr = xb + static_cast<long long>(t) * st;
    Best best{-INFINITY, 0x7fffffff};
    for (int c = lane; c < C; c += 32) {
      Best cand{xr[c], c};
      if (best.i == 0x7fffffff || better(cand, best)) best = cand;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Best other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      if (other.i != 0x7fffffff && (best.i == 0x7fffffff || better(other, best))) best = other;
    }
    if (lane == 0) {
      raw[t] = best.i;
      if (raw_out) raw_out[static_cast<long long>(b) * T + t] = best.i;
    }
  }
  __syncthreads();
  if (warp == 0) {
    int* out = ids + static_cast<long long>(b) * T;
    const int n = collapse_compact(raw, Tb, n_character, out);
    for (int t = n + lane; t < T; t += 32) out[t] = 0;
    if (lane == 0) lens[b] = n;
  }
}

// Collapse an already-argmaxed index stream (the reference decode() input: [sum T_b] sample-major).
template <typename IdxT>
__global__ void __launch_bounds__(kDecThreads) collapse_kernel(const IdxT* __restrict__ index,
                                                               const int* __restrict__ lengths, int B, int Tmax,
                                                               int n_character, int* __restrict__ ids,
                                                               int* __restrict__ lens) {
  extern __shared__ int raw[];                     // [Tmax]
  __shared__ float red[40];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float part = 0.f;
  for (int i = threadIdx.x; i < b; i += kDecThreads) part += static_cast<float>(lengths[i]);
  const long long off = static_cast<long long>(block_sum(part, red) + 0.5f);
  const int Tb = min(max(lengths[b], 0), Tmax);
  for (int t = threadIdx.x; t < Tb; t += kDecThreads) {
    const long long v = static_cast<long long>(index[off + t]);
    raw[t] = v < 0 || v > 0x7ffffffe ? 0x7ffffffe : static_cast<int>(v);
  }
  __syncthreads();
  if (warp == 0) {
    int* out = ids + static_cast<long long>(b) * Tmax;
    const int n = collapse_compact(raw, Tb, n_character, out);
    for (int t = n + lane; t < Tmax; t += 32) out[t] = 0;
    if (lane == 0) lens[b] = n;
  }
}

}  // namespace htrvt

using namespace htrvt;

extern "C" int htrvt_greedy_decode(const float* logits, long long stride_b, long long stride_t, int B, int T, int C,
                                   const int* lengths, int n_character, int* ids, int* lens, int* raw_index,
                                   cudaStream_t stream) {
  if (B <= 0 || T <= 0 || C <= 0 || !logits || !ids || !lens) return HTRVT_ERR_SHAPE;
  if (static_cast<size_t>(T) * 4 > 40 * 1024) return HTRVT_ERR_SHAPE;
  greedy_decode_kernel<<<B, kDecThreads, T * sizeof(int), stream>>>(logits, stride_b, stride_t, B, T, C, lengths,
                                                                    n_character, ids, lens, raw_index);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_ctc_collapse(const void* index, int index_is_int64, const int* lengths, int B, int Tmax,
                                  int n_character, int* ids, int* lens, cudaStream_t stream) {
  if (B <= 0 || Tmax <= 0 || !index || !lengths || !ids || !lens) return HTRVT_ERR_SHAPE;
  if (static_cast<size_t>(Tmax) * 4 > 40 * 1024) return HTRVT_ERR_SHAPE;
  if (index_is_int64)
    collapse_kernel<long long><<<B, kDecThreads, Tmax * sizeof(int), stream>>>(
        static_cast<const long long*>(index), lengths, B, Tmax, n_character, ids, lens);
  else
    collapse_kernel<int><<<B, kDecThreads, Tmax * sizeof(int), stream>>>(static_cast<const int*>(index), lengths, B,
                                                                         Tmax, n_character, ids, lens);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_version() { return 100; }   // 0.1.0

unsigned long long htrvt_launch_counter = 0;
extern "C" unsigned long long htrvt_launch_count() { return htrvt_launch_counter; }
