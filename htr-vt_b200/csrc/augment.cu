// Training-time augmentation of a uint8 batch of line images for sm_100a: the pixel work of
// the reference's SameTrCollate (model_v1/data/dataset.py:13-45) in ONE launch, one CTA per line image, the image
// resident in shared memory from the first stage to the last:
//   1. RandomTransform (model_v1/data/transform.py:164-230): projective warp to the bounding box of the warped corners
//      (skimage.transform.warp: bilinear on floor / ceil neighbours, 255 outside) followed by skimage.transform.resize
//      back to H x W (anti-aliasing Gaussian down the rows when they shrink by more than 1.25, scipy.ndimage.zoom
//      order 1 / mirror / grid_mode) and truncation to uint8.  The intermediate image (up to ~81 x 529 float64) is
//      never materialised: every output pixel evaluates its 4 zoom taps (x 3 filter rows) on the fly from the uint8
//      source in shared memory, in float64 and in the operation order of the CPU code (explicit _rn intrinsics:
//      no FMA contraction), because the final truncation makes 254.99999999999997 and 255 different pixels.
//   2. cv2.erode / cv2.dilate with an all-ones k_rows x k_cols rectangle, anchor at size / 2, `iterations` folded into
//      a grown rectangle, pixels outside the image never win (transform.py:11-33).
//   3. ColorJitter on a grey image: brightness = PIL blend with black, contrast = PIL blend with the rounded image
//      mean (exact integer sum by a block reduction), float32 with PIL's truncate / clip rule; saturation and hue
//      are identities on mode 'L' and never reach the device.
// The random decisions are drawn on the host in the reference's order (htr-vt_b200/augment.py) and arrive as one
// 128-byte record per image.
#include "common.cuh"

namespace htrvt {

struct AugLine {                 // mirrors augment.py::_REC
  double m[9];                   // inverse projective map: output (col, row, 1) of the warped image -> input position
  double w0, w1;                 // anti-aliasing weights (centre, neighbour) when gauss != 0
  int warp;                      // 0: stage 1 off
  int rows, cols;                // shape of the intermediate warped image
  int gauss;                     // 1: three-tap Gaussian down the rows of the warped image
  int jit_n;                     // 0..2 jitter ops
  int jit_op[2];                 // 0 brightness, 1 contrast
  float jit_f[2];
  int pad;
};
static_assert(sizeof(AugLine) == 128, "record layout");

constexpr int kAugThreads = 512;

__device__ __forceinline__ int mirror_idx(int i, int n) {          // scipy 'mirror': reflect about the edge pixel centres
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - i;
}

// one pixel of the warped image: skimage `_warp_fast` + `bilinear_interpolation`, mode constant, cval 255
__device__ __forceinline__ double warped_px(const uint8_t* __restrict__ src, int H, int W, const double* m, int r,
                                            int c) {
  const double x = static_cast<double>(c), y = static_cast<double>(r);
  const double xx = __dadd_rn(__dadd_rn(__dmul_rn(m[0], x), __dmul_rn(m[1], y)), m[2]);
  const double yy = __dadd_rn(__dadd_rn(__dmul_rn(m[3], x), __dmul_rn(m[4], y)), m[5]);
  const double zz = __dadd_rn(__dadd_rn(__dmul_rn(m[6], x), __dmul_rn(m[7], y)), m[8]);
  const double cc = __ddiv_rn(xx, zz), rr = __ddiv_rn(yy, zz);
  const double minr = floor(rr), minc = floor(cc), maxr = ceil(rr), maxc = ceil(cc);
  const double dr = __dsub_rn(rr, minr), dc = __dsub_rn(cc, minc);
  auto px = [&](double pr, double pc) -> double {
    if (!(pr >= 0.0 && pr < static_cast<double>(H) && pc >= 0.0 && pc < static_cast<double>(W))) return 255.0;
    return static_cast<double>(src[static_cast<int>(pr) * W + static_cast<int>(pc)]);
  };
  const double omc = __dsub_rn(1.0, dc), omr = __dsub_rn(1.0, dr);
  const double top = __dadd_rn(__dmul_rn(omc, px(minr, minc)), __dmul_rn(dc, px(minr, maxc)));
  const double bot = __dadd_rn(__dmul_rn(omc, px(maxr, minc)), __dmul_rn(dc, px(maxr, maxc)));
  return __dadd_rn(__dmul_rn(omr, top), __dmul_rn(dr, bot));
}

// PIL Image.blend(degenerate, px, f) on uint8, float32 arithmetic (libImaging/Blend.c)
__device__ __forceinline__ uint8_t pil_blend(int deg, int px, float f, bool inside) {
  const float t = __fadd_rn(static_cast<float>(deg), __fmul_rn(f, static_cast<float>(px - deg)));
  if (inside) return static_cast<uint8_t>(static_cast<int>(t));
  if (t <= 0.f) return 0;
  if (t >= 255.f) return 255;
  return static_cast<uint8_t>(static_cast<int>(t));
}

__global__ void __launch_bounds__(kAugThreads) augment_lines_kernel(const uint8_t* __restrict__ in, long long in_sb,
                                                                    uint8_t* __restrict__ out,
                                                                    const AugLine* __restrict__ recs, int H, int W,
                                                                    int morph, int k_rows, int k_cols, int iters) {
  extern __shared__ __align__(16) uint8_t aug_smem[];
  __shared__ AugLine R;
  __shared__ unsigned long long red[kAugThreads / 32];
  __shared__ int mean_sh;
  __shared__ double rng_lo[kAugThreads / 32], rng_hi[kAugThreads / 32];
  const int n = H * W;
  uint8_t* A = aug_smem;
  uint8_t* Bf = aug_smem + ((n + 15) & ~15);
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < static_cast<int>(sizeof(AugLine) / 4))
    reinterpret_cast<uint32_t*>(&R)[tid] = reinterpret_cast<const uint32_t*>(recs + b)[tid];
  const uint8_t* src = in + static_cast<long long>(b) * in_sb;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n & 15) == 0) {
    for (int i = tid; i < n / 16; i += kAugThreads)
      reinterpret_cast<uint4*>(A)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
  } else {
    for (int i = tid; i < n; i += kAugThreads) A[i] = src[i];
  }
  __syncthreads();
  uint8_t* cur = A;
  uint8_t* nxt = Bf;

  // ---- 1. projective warp + resize -------------------------------------------------------------------------------
  if (R.warp) {
    const int ih = R.rows, iw = R.cols;
    // resize clips its result to the value range of ITS input (the warped image, before the filter): a pass over the
    // warped pixels for their minimum / maximum
    double lo = 255.0, hi = 0.0;
    for (int i = tid; i < ih * iw; i += kAugThreads) {
      const int r = i / iw;
      const double v = warped_px(cur, H, W, R.m, r, i - r * iw);
      lo = fmin(lo, v); hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((tid & 31) == 0) { rng_lo[tid >> 5] = lo; rng_hi[tid >> 5] = hi; }
    __syncthreads();
    lo = rng_lo[0]; hi = rng_hi[0];
    for (int w = 1; w < kAugThreads / 32; ++w) { lo = fmin(lo, rng_lo[w]); hi = fmax(hi, rng_hi[w]); }
    const double zr = __ddiv_rn(static_cast<double>(ih), static_cast<double>(H));
    const double zc = __ddiv_rn(static_cast<double>(iw), static_cast<double>(W));
    for (int i = tid; i < n; i += kAugThreads) {
      const int oy = i / W, ox = i - oy * W;
      double cr = __dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(oy), 0.5), zr), 0.5);
      double cc = __dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(ox), 0.5), zc), 0.5);
      if (ih == 1) cr = 0.0;                                    // a one-pixel axis maps every coordinate onto pixel 0
      if (iw == 1) cc = 0.0;
      const double fr = floor(cr), fc = floor(cc);
      const double yr = __dsub_rn(cr, fr), yc = __dsub_rn(cc, fc);
      const int r0 = static_cast<int>(fr), c0 = static_cast<int>(fc);
      // taps (index, weight) per axis in scipy's visiting order: floor, floor + 1 - except on a growing axis' first
      // coordinate (floor = -1), where the in-image tap comes before the mirrored one (last bit of the 4-term sum)
      int tr[2] = {mirror_idx(r0, ih), mirror_idx(r0 + 1, ih)}, tc[2] = {mirror_idx(c0, iw), mirror_idx(c0 + 1, iw)};
      double wr[2] = {__dsub_rn(1.0, yr), yr}, wc[2] = {__dsub_rn(1.0, yc), yc};
      if (r0 < 0) { int s = tr[0]; tr[0] = tr[1]; tr[1] = s; double q = wr[0]; wr[0] = wr[1]; wr[1] = q; }
      if (c0 < 0) { int s = tc[0]; tc[0] = tc[1]; tc[1] = s; double q = wc[0]; wc[0] = wc[1]; wc[1] = q; }
      double t = 0.0;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int rr = tr[a];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc2 = tc[e];
          double v;
          if (R.gauss) {        // scipy correlate1d, symmetric kernel: centre first, then (left + right) * w
            const double vc = warped_px(cur, H, W, R.m, rr, cc2);
            const double vl = warped_px(cur, H, W, R.m, mirror_idx(rr - 1, ih), cc2);
            const double vh = warped_px(cur, H, W, R.m, mirror_idx(rr + 1, ih), cc2);
            v = __dadd_rn(__dmul_rn(vc, R.w0), __dmul_rn(__dadd_rn(vl, vh), R.w1));
          } else {
            v = warped_px(cur, H, W, R.m, rr, cc2);
          }
          t = __dadd_rn(t, __dmul_rn(__dmul_rn(v, wr[a]), wc[e]));
        }
      }
      t = fmin(fmax(t, lo), hi);
      nxt[i] = static_cast<uint8_t>(static_cast<int>(t));
    }
    __syncthreads();
    uint8_t* s = cur; cur = nxt; nxt = s;
  }

  // ---- 2. erosion / dilation ---------------------------------------------------------------------------------------
  if (morph) {
    const int ay = k_rows / 2, ax = k_cols / 2;
    const int lo_y = -ay * iters, hi_y = lo_y + (k_rows - 1) * iters;
    const int lo_x = -ax * iters, hi_x = lo_x + (k_cols - 1) * iters;
    const bool erode = morph == 1;
    for (int i = tid; i < n; i += kAugThreads) {
      const int y = i / W, x = i - y * W;
      int acc = erode ? 255 : 0;
      for (int dy = lo_y; dy <= hi_y; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = lo_x; dx <= hi_x; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          const int v = cur[yy * W + xx];
          acc = erode ? min(acc, v) : max(acc, v);
        }
      }
      nxt[i] = static_cast<uint8_t>(acc);
    }
    __syncthreads();
    uint8_t* s = cur; cur = nxt; nxt = s;
  }

  // ---- 3. brightness / contrast ------------------------------------------------------------------------------------
  for (int j = 0; j < R.jit_n; ++j) {
    const float f = R.jit_f[j];
    const bool inside = f >= 0.f && f <= 1.f;
    int deg = 0;
    if (R.jit_op[j] == 1) {                                    // contrast: int(mean + 0.5) of the CURRENT image
      unsigned long long s = 0;
      for (int i = tid; i < n; i += kAugThreads) s += cur[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) red[tid >> 5] = s;
      __syncthreads();
      if (tid == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < kAugThreads / 32; ++w) tot += red[w];
        mean_sh = static_cast<int>(__dadd_rn(__ddiv_rn(static_cast<double>(tot), static_cast<double>(n)), 0.5));
      }
      __syncthreads();
      deg = mean_sh;
    }
    for (int i = tid; i < n; i += kAugThreads) cur[i] = pil_blend(deg, cur[i], f, inside);
    __syncthreads();
  }

  uint8_t* dst = out + static_cast<long long>(b) * n;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (n & 15) == 0) {
    for (int i = tid; i < n / 16; i += kAugThreads) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(cur)[i];
  } else {
    for (int i = tid; i < n; i += kAugThreads) dst[i] = cur[i];
  }
}

}  // namespace htrvt

using namespace htrvt;

// in uint8 [B, H, W] (image stride in bytes, rows contiguous), out uint8 [B, H, W] contiguous, recs: B 128-byte
// records on the device (layout: struct AugLine above = htr-vt_b200/augment.py::_REC).  morph: 0 none, 1 erode,
// 2 dilate with a k_rows x k_cols all-ones rectangle, `iterations` times - the same for every image of the batch, as in
// the reference.  2 * H * W bytes of shared memory: H * W <= 113 KB (64 x 1024 lines: 128 KB).
extern "C" int htrvt_augment_lines(const void* in, long long stride_b, void* out, const void* recs, int B, int H, int W,
                                   int morph, int k_rows, int k_cols, int iterations, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || !in || !out || !recs) return HTRVT_ERR_SHAPE;
  if (morph < 0 || morph > 2 || k_rows < 1 || k_cols < 1 || iterations < 1 || k_rows > 15 || k_cols > 15 ||
      iterations > 8)
    return HTRVT_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(recs) & 7) return HTRVT_ERR_ALIGN;
  const size_t n = static_cast<size_t>(H) * W;
  const size_t smem = 2 * ((n + 15) & ~static_cast<size_t>(15));
  if (smem > 220 * 1024) return HTRVT_ERR_SHAPE;
  if (smem > 48 * 1024 && !HTRVT_ENSURE_SMEM(augment_lines_kernel, 220 * 1024)) return HTRVT_ERR_LAUNCH;
  augment_lines_kernel<<<B, kAugThreads, smem, stream>>>(static_cast<const uint8_t*>(in), stride_b,
                                                       static_cast<uint8_t*>(out), static_cast<const AugLine*>(recs), H,
                                                       W, morph, k_rows, k_cols, iterations);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
