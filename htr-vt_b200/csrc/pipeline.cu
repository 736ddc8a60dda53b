// The two callers either side of the encoder that SURVEY.md 8(f) ranks next (rows 2 and 3):
//
//  * line_prep_u8: the device half of the input pipeline.  The reference's loader hands the model u8 / 255 images
//    right-padded with 1.0 (model_v1/data/dataset.py:13-45 SameTrCollate: `np.uint8(img * 255)` ... `/ 255.`;
//    :104-135 get_images: pad with constant 1.0) and the model's first op is the whole-sample LayerNorm
//    (model_v1/model/HTR_VT.py:134-136, 224).  Here the uint8 line is what crosses PCIe (4x fewer bytes) and ONE
//    kernel does u8 -> /255 -> pad-to-W with 1.0 -> LayerNorm; the statistics come from exact integer sums.
//
//  * edit_distance: Levenshtein distances of B (prediction, ground truth) id sequences, one warp per pair - the
//    `editdistance.eval` calls of the validation loop (model_v1/valid.py:49-75; restated in model_v1/test.py:114-133)
//    without the device -> host round trip of the decoded ids: CER = sum(dist) / sum(len(gt)).
#include "common.cuh"

namespace htrvt {

// ------------------------------------------------------------------------------------------------
// u8 line -> normalised fp32 sample.  One CTA per sample, N = H * W (W % 4 == 0), row stride ld (bytes).
// widths (optional): columns >= widths[b] are padding and read as 255 (= 1.0) whatever the buffer holds.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) line_prep_u8_kernel(const uint8_t* __restrict__ img, long long sample_stride,
                                                            int ld, const int* __restrict__ widths, int H, int W,
                                                            float* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, float eps) {
  __shared__ unsigned long long red_s[32], red_q[32];
  __shared__ float stat[2];
  const int b = blockIdx.x;
  const uint8_t* xb = img + static_cast<long long>(b) * sample_stride;
  float* yb = y + static_cast<long long>(b) * H * W;
  const int wv = widths ? min(max(widths[b], 0), W) : W;
  const int W4 = W >> 2, n4 = H * W4;
  auto load4 = [&](int i) -> uchar4 {
    const int r = i / W4, c = (i - r * W4) * 4;
    uchar4 v = *reinterpret_cast<const uchar4*>(xb + static_cast<long long>(r) * ld + c);
    if (c + 3 >= wv) {
      if (c >= wv) v.x = 255;
      if (c + 1 >= wv) v.y = 255;
      if (c + 2 >= wv) v.z = 255;
      v.w = 255;
    }
    return v;
  };
  unsigned int s = 0;
  unsigned long long q = 0;
  for (int i = threadIdx.x; i < n4; i += 1024) {
    const uchar4 v = load4(i);
    s += v.x + v.y + v.z + v.w;
    q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  unsigned long long s64 = s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s64 += __shfl_xor_sync(0xffffffffu, s64, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red_s[wid] = s64; red_q[wid] = q; }
  __syncthreads();
  if (wid == 0) {
    s64 = red_s[lane]; q = red_q[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s64 += __shfl_xor_sync(0xffffffffu, s64, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      // exact integer moments of the u8 image; mean / var of u8 / 255 in double
      const double N = static_cast<double>(H) * W;
      const double mean = static_cast<double>(s64) / (255.0 * N);
      const double var = fmax(static_cast<double>(q) / (65025.0 * N) - mean * mean, 0.0);
      const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
      stat[0] = static_cast<float>(mean); stat[1] = rstd;
      if (mean_out) mean_out[b] = stat[0];
      if (rstd_out) rstd_out[b] = rstd;
    }
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (int i = threadIdx.x; i < n4; i += 1024) {
    const uchar4 v = load4(i);
    const int r = i / W4, c = (i - r * W4) * 4;
    // u8 / 255 rounded to fp32 exactly as `tensor.float() / 255.` does, then normalised
    const float4 o = make_float4((__fdiv_rn(static_cast<float>(v.x), 255.0f) - mean) * rstd,
                                 (__fdiv_rn(static_cast<float>(v.y), 255.0f) - mean) * rstd,
                                 (__fdiv_rn(static_cast<float>(v.z), 255.0f) - mean) * rstd,
                                 (__fdiv_rn(static_cast<float>(v.w), 255.0f) - mean) * rstd);
    *reinterpret_cast<float4*>(yb + static_cast<long long>(r) * W + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// Levenshtein distance (unit costs), one warp per pair.  The DP row lives in shared memory; a row is
// processed in 32-column chunks: t[j] = min(prev[j] + 1, prev[j-1] + (a_i != b_j)) is elementwise, and the
// remaining dependency cur[j] = min(t[j], cur[j-1] + 1) is a min-plus prefix scan (5 shuffles on t[j] - j).
// ------------------------------------------------------------------------------------------------
constexpr int kEdWarps = 4;

__global__ void __launch_bounds__(kEdWarps * 32) edit_distance_kernel(
    const int* __restrict__ a, const int* __restrict__ a_off, int a_stride, const int* __restrict__ a_len,
    const int* __restrict__ b, const int* __restrict__ b_off, int b_stride, const int* __restrict__ b_len, int n,
    int max_b, int* __restrict__ out) {
  extern __shared__ int ed_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * kEdWarps + warp;
  if (p >= n) return;
  int* prev = ed_smem + warp * 2 * (max_b + 1);          // prev[0..lb], then the b ids
  int* bs = prev + max_b + 1;
  const int la = max(a_len[p], 0), lb = min(max(b_len[p], 0), max_b);
  const int* ap = a + (a_off ? static_cast<long long>(a_off[p]) : static_cast<long long>(p) * a_stride);
  const int* bp = b + (b_off ? static_cast<long long>(b_off[p]) : static_cast<long long>(p) * b_stride);
  for (int j = lane; j <= lb; j += 32) prev[j] = j;
  for (int j = lane; j < lb; j += 32) bs[j] = bp[j];
  __syncwarp();
  for (int i = 1; i <= la; ++i) {
    const int ai = ap[i - 1];
    int carry = i;                                         // cur[0]
    int diag_in = i - 1;                                   // prev[0] of this row (= i - 1), the first diagonal
    for (int j0 = 1; j0 <= lb; j0 += 32) {
      const int j = j0 + lane;
      const bool ok = j <= lb;
      const int pj = ok ? prev[j] : 0x3fffffff;
      int pjm1 = __shfl_up_sync(0xffffffffu, pj, 1);
      if (lane == 0) pjm1 = diag_in;
      diag_in = __shfl_sync(0xffffffffu, pj, 31);          // prev[j0 + 31] feeds the next chunk's first diagonal
      const int cost = (ok && bs[j - 1] == ai) ? 0 : 1;
      int t = ok ? min(pj + 1, pjm1 + cost) : 0x3fffffff;
      // cur[j] = min_k<=j (t[k] + j - k), also against the carry-in cur[j0 - 1] + (j - j0 + 1)
      int v = t - j;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = min(v, u);
      }
      v = min(v, carry - (j0 - 1));
      const int cur = v + j;
      __syncwarp();
      if (ok) prev[j] = cur;
      carry = __shfl_sync(0xffffffffu, cur, 31);
    }
    if (lane == 0) prev[0] = i;
    __syncwarp();
  }
  __syncwarp();
  if (lane == 0) out[p] = (la == 0) ? lb : prev[lb];
}

}  // namespace htrvt

using namespace htrvt;

// img: uint8 [B, H, ld] (ld >= W bytes per row, sample stride in bytes); y: fp32 [B, H, W] normalised with eps.
extern "C" int htrvt_line_prep_u8(const void* img, long long sample_stride, int ld, const int* widths, int B, int H,
                                  int W, float* y, float* mean, float* rstd, float eps, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || (W & 3) || (ld & 3) || ld < W || !img || !y) return HTRVT_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(img) & 3) || (sample_stride & 3) || (reinterpret_cast<uintptr_t>(y) & 15))
    return HTRVT_ERR_ALIGN;
  if (static_cast<long long>(H) * W > (1ll << 24)) return HTRVT_ERR_SHAPE;     // 32-bit sum of u8 values per thread
  line_prep_u8_kernel<<<B, 1024, 0, stream>>>(static_cast<const uint8_t*>(img), sample_stride, ld, widths, H, W, y,
                                              mean, rstd, eps);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// out[p] = Levenshtein(a_p, b_p), p < n.  Sequences are int32 ids; sequence p starts at x_off[p] (if x_off) or at
// p * x_stride.  max_b_len: upper bound of b_len (host side), <= 4095.
extern "C" int htrvt_edit_distance(const int* a, const int* a_off, int a_stride, const int* a_len, const int* b,
                                   const int* b_off, int b_stride, const int* b_len, int n, int max_b_len, int* out,
                                   cudaStream_t stream) {
  if (n <= 0) return HTRVT_OK;
  if (!a_len || !b_len || !out || max_b_len < 0 || max_b_len > 4095) return HTRVT_ERR_SHAPE;
  const int smem = kEdWarps * 2 * (max_b_len + 1) * static_cast<int>(sizeof(int));
  if (smem > 48 * 1024 && !HTRVT_ENSURE_SMEM(edit_distance_kernel, smem)) return HTRVT_ERR_LAUNCH;
  edit_distance_kernel<<<(n + kEdWarps - 1) / kEdWarps, kEdWarps * 32, smem, stream>>>(a, a_off, a_stride, a_len, b,
                                                                                      b_off, b_stride, b_len, n,
                                                                                      max_b_len, out);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
