// Greedy CTC decoding for sm_100a: per-frame argmax + collapse (drop blanks, repeats and
// out-of-alphabet ids) + compaction in ONE kernel, one CTA per line (the line's logits staged in shared memory by
// 16-byte asynchronous copies); a single D2H copy of ids[B,T] / lens[B] then replaces the reference's per-element
// device syncs.
//
// Replaces model_v1/valid.py:40-42 (`preds.max(2)`, transpose, view) and the id filtering of
// CTCLabelConverter.decode (model_v1/utils/utils.py:72-86); id -> char stays on the host.
// Semantics: argmax = lowest index among ties, a NaN beats any number (first NaN);
// keep id iff id != 0 and id != previous RAW frame id and id < n_character.
#include "common.cuh"

namespace htrvt {

constexpr int kDecThreads = 128;

// Order-preserving integer image of a logit for the arg-max: a > b as floats <=> okey(a) > okey(b), -0 and +0 share one
// key (equal floats tie, and ties keep the lowest index), every NaN maps to the largest key (a NaN beats any number;
// the first NaN wins).  0 is not the key of any float.
__device__ __forceinline__ unsigned okey(float v) {
  const unsigned u = __float_as_uint(__fadd_rn(v, 0.0f));                    // -0 -> +0
  const unsigned k = u ^ (static_cast<unsigned>(static_cast<int>(u) >> 31) | 0x80000000u);
  return (u & 0x7fffffffu) > 0x7f800000u ? 0xffffffffu : k;
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

// ---------------------------------------------------------------------------------------------
// arg-max + collapse, one CTA (256 threads) per line.  The kernel is pure streaming (B*T*C*4 bytes in, B*T*4 out), so
// the whole line is brought into shared memory with 16-byte cp.async copies that are ALL in flight at once (rows in
// chunks when T*C*4 exceeds the staging budget), issued as four commit groups that are scanned as they land: eight
// lanes per frame read the staged row with 16-byte loads (strict comparisons keep the lowest index, a NaN wins), and
// warp 0 compacts.  512 x 128 x 80: 10.4 us per launch back to back (same-box A/B against the two-threads-per-frame scan
// with its 8-way bank conflicts: 10.6 us) - the scan is NOT what bounds this kernel; launch + the DRAM ramp of a
// 21 MB single-wave grid is (tools/decode_probe.py).
// ---------------------------------------------------------------------------------------------
constexpr int kGdThreads = 256;
constexpr int kGdStageBytes = 96 * 1024;

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kGdThreads) greedy_decode_kernel(
    const float* __restrict__ x, long long sb, long long st, int B, int T, int C, const int* __restrict__ lengths,
    int n_character, int* __restrict__ ids, int* __restrict__ lens, int* __restrict__ raw_out, int rows_per_chunk,
    int vec_ok) {
  extern __shared__ __align__(16) unsigned char gd_smem[];
  int* raw = reinterpret_cast<int*>(gd_smem);                               // [T] arg-max ids
  const int ldc = (C + 3) & ~3;                                             // staged row stride (floats), 16-byte rows
  float* stage = reinterpret_cast<float*>(gd_smem + ((static_cast<size_t>(T) * 4 + 15) & ~size_t(15)));
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int Tb = lengths ? lengths[b] : T;
  Tb = min(max(Tb, 0), T);
  const float* xb = x + static_cast<long long>(b) * sb;
  const int nv_row = ldc >> 2, l8 = lane & 7, fr = lane >> 3;
  for (int t0 = 0; t0 < Tb; t0 += rows_per_chunk) {
    const int nr = min(rows_per_chunk, Tb - t0);
    if (t0 > 0) __syncthreads();                                            // previous chunk fully scanned
    // the chunk's rows go out as four cp.async groups (multiples of 32 rows; all in flight at once): group g is scanned
    // as soon as it has landed, under the copies of the groups behind it
    const int rg = (((nr + 3) >> 2) + 31) & ~31;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int r_lo = min(g * rg, nr), r_hi = min(r_lo + rg, nr);
      if (vec_ok) {
        const int v_per_row = C >> 2;
        for (int i = r_lo * v_per_row + threadIdx.x; i < r_hi * v_per_row; i += kGdThreads) {
          const int r = i / v_per_row, v = i - r * v_per_row;
          cp_async_16(smem_u32(stage + r * ldc + 4 * v), xb + static_cast<long long>(t0 + r) * st + 4 * v);
        }
      } else {
        for (int i = r_lo * C + threadIdx.x; i < r_hi * C; i += kGdThreads) {
          const int r = i / C, c = i - r * C;
          cp_async_4(smem_u32(stage + r * ldc + c), xb + static_cast<long long>(t0 + r) * st + c);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // eight lanes per frame, four frames per warp and pass (32 rows per pass): lane l of a frame reads the 16-byte groups
    // l, l + 8, ... of the staged row, so every quarter warp of an LDS.128 covers 128 contiguous bytes of one row (no
    // bank conflicts).  A lane walks its classes in ascending order on order-preserving integer keys (okey), so a strict
    // comparison keeps the lowest index; three shuffle rounds merge the eight partial results by (key desc, index asc)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g == 0) asm volatile("cp.async.wait_group 3;" ::: "memory");
      if (g == 1) asm volatile("cp.async.wait_group 2;" ::: "memory");
      if (g == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
      if (g == 3) asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const int r_lo = min(g * rg, nr), r_hi = min(r_lo + rg, nr);
      for (int base = r_lo; base < r_hi; base += 4 * (kGdThreads / 32)) {   // uniform trip count: the shuffles need whole warps
        const int r = base + 4 * warp + fr;
        const bool act = r < r_hi;
        unsigned bk = 0u;                                                   // below every key of a real element
        int bi = 0x7fffffff;
        if (act) {
          const float4* row4 = reinterpret_cast<const float4*>(stage + r * ldc);
          for (int v = l8; v < nv_row; v += 8) {
            const float4 q = row4[v];
            const int c = 4 * v;                                            // c < C always; the row's last group may be padded
            const unsigned k0 = okey(q.x), k1 = okey(q.y), k2 = okey(q.z), k3 = okey(q.w);
            if (k0 > bk) { bk = k0; bi = c; }
            if (c + 1 < C && k1 > bk) { bk = k1; bi = c + 1; }
            if (c + 2 < C && k2 > bk) { bk = k2; bi = c + 2; }
            if (c + 3 < C && k3 > bk) { bk = k3; bi = c + 3; }
          }
        }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
          const unsigned ok = __shfl_xor_sync(0xffffffffu, bk, d);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
          if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
        }
        if (act && l8 == 0) {
          raw[t0 + r] = bi;
          if (raw_out) raw_out[static_cast<long long>(b) * T + t0 + r] = bi;
        }
      }
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
                                                               const int* __restrict__ lengths,
                                                               const long long* __restrict__ offsets, int B, int Tmax,
                                                               int n_character, int* __restrict__ ids,
                                                               int* __restrict__ lens) {
  extern __shared__ int raw[];                     // [Tmax]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long off = offsets[b];                // exclusive prefix sum of the lengths (int64, one scan by the caller)
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
  const int ldc = (C + 3) & ~3;
  if (static_cast<size_t>(ldc) * 4 > static_cast<size_t>(kGdStageBytes)) return HTRVT_ERR_SHAPE;
  int rows = static_cast<int>(kGdStageBytes / (static_cast<size_t>(ldc) * 4));
  if (rows > T) rows = T;
  // 16-byte copies need every row start on a 16-byte boundary
  const int vec_ok = ((C & 3) == 0) && ((stride_b & 3) == 0) && ((stride_t & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const size_t smem = ((static_cast<size_t>(T) * 4 + 15) & ~size_t(15)) + static_cast<size_t>(rows) * ldc * 4;
  if (smem > 48 * 1024 && !HTRVT_ENSURE_SMEM(greedy_decode_kernel, 160 * 1024)) return HTRVT_ERR_LAUNCH;
  greedy_decode_kernel<<<B, kGdThreads, smem, stream>>>(logits, stride_b, stride_t, B, T, C, lengths, n_character, ids,
                                                        lens, raw_index, rows, vec_ok);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// offsets: int64 [B] device array, exclusive prefix sum of `lengths` (start of each line in the index stream)
extern "C" int htrvt_ctc_collapse(const void* index, int index_is_int64, const int* lengths, const long long* offsets,
                                  int B, int Tmax, int n_character, int* ids, int* lens, cudaStream_t stream) {
  if (B <= 0 || Tmax <= 0 || !index || !lengths || !offsets || !ids || !lens) return HTRVT_ERR_SHAPE;
  if (static_cast<size_t>(Tmax) * 4 > 40 * 1024) return HTRVT_ERR_SHAPE;
  if (index_is_int64)
    collapse_kernel<long long><<<B, kDecThreads, Tmax * sizeof(int), stream>>>(
        static_cast<const long long*>(index), lengths, offsets, B, Tmax, n_character, ids, lens);
  else
    collapse_kernel<int><<<B, kDecThreads, Tmax * sizeof(int), stream>>>(static_cast<const int*>(index), lengths,
                                                                         offsets, B, Tmax, n_character, ids, lens);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_version() { return 100; }   // 0.1.0

unsigned long long htrvt_launch_counter = 0;
extern "C" unsigned long long htrvt_launch_count() { return htrvt_launch_counter; }
