// Multi-tensor optimizer kernels for the training loop around the hot path:
// SAM (model_v1/utils/sam.py:15-59: global gradient norm, climb w + rho g / |g|, restore) around AdamW
// (torch.optim.AdamW as configured at model_v1/train.py:93) and the EMA of the state_dict
// (model_v1/utils/utils.py:158-173).  The reference issues ~10^3 tiny launches per iteration (a norm, clone and
// add_ per parameter, a lerp per state_dict entry); here every pass over the 53 M parameters is one launch per
// <= 192 tensors: pointer tables travel as kernel parameters, a block finds its (tensor, chunk) by binary search over
// the chunk prefix sums.  Pure HBM-bound streaming: 4 B/param (norm), 16 B/param (first step), 36 B/param (AdamW).
#include "common.cuh"

namespace htrvt {

constexpr int kMtMax = 192;              // 10 KB of kernel parameters (CUDA >= 12.1 takes up to 32 KB)
constexpr int kMtChunk = 16384;           // elements per block

struct MtTable {
  void* a[kMtMax];                        // meaning depends on the kernel
  void* b[kMtMax];
  void* c[kMtMax];
  void* d[kMtMax];
  void* e[kMtMax];
  long long numel[kMtMax];
  int chunk_start[kMtMax + 1];            // prefix sums of ceil(numel / kMtChunk)
  int n;
};

__device__ __forceinline__ void mt_locate(const MtTable& T, int& t, long long& lo, long long& hi) {
  const int id = blockIdx.x;
  int l = 0, r = T.n;                     // chunk_start[l] <= id < chunk_start[r]
  while (r - l > 1) {
    const int m = (l + r) >> 1;
    if (T.chunk_start[m] <= id) l = m; else r = m;
  }
  t = l;
  lo = static_cast<long long>(id - T.chunk_start[l]) * kMtChunk;
  hi = lo + kMtChunk < T.numel[l] ? lo + kMtChunk : T.numel[l];
}

// Streaming helper: fn4(i, n) is called for 4-element groups while every stream is 16-byte aligned (n = 4: the caller
// uses float4 accesses), and for single elements otherwise / in the tail (n = 1).  Unrolled x4: each thread keeps four
// independent 16-byte loads per stream in flight (these kernels are pure HBM streaming).
template <typename F4, typename F1>
__device__ __forceinline__ void mt_stream(long long lo, long long hi, bool aligned, F4 f4, F1 f1) {
  long long i = lo;
  if (aligned) {
    const long long n4 = (hi - lo) >> 2;
#pragma unroll 4
    for (long long k = threadIdx.x; k < n4; k += 256) f4(lo + 4 * k);
    i = lo + 4 * n4;
  }
  for (i += threadIdx.x; i < hi; i += 256) f1(i);
}
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// out (double, +=) = sum over tensors of |g|^2  (adaptive: |abs(p) g|^2)      a = g, b = p
__global__ void __launch_bounds__(256) mt_sqnorm_kernel(const __grid_constant__ MtTable T, int adaptive,
                                                        double* __restrict__ out) {
  __shared__ float red[40];
  int t; long long lo, hi;
  mt_locate(T, t, lo, hi);
  const float* g = static_cast<const float*>(T.a[t]);
  const float* p = static_cast<const float*>(T.b[t]);
  float s = 0.f;
  mt_stream(lo, hi, al16(g) && (!adaptive || al16(p)),
            [&](long long i) {
              float4 v = ld4(g + i);
              if (adaptive) { const float4 w = ld4(p + i); v.x *= fabsf(w.x); v.y *= fabsf(w.y); v.z *= fabsf(w.z); v.w *= fabsf(w.w); }
              s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
            },
            [&](long long i) {
              float v = g[i];
              if (adaptive) v *= fabsf(p[i]);
              s = fmaf(v, v, s);
            });
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, static_cast<double>(s));
}

// SAM first step: old = p; p += (adaptive ? p^2 : 1) * g * rho / (sqrt(norm2) + 1e-12)      a = p, b = g, c = old
__global__ void __launch_bounds__(256) mt_sam_first_kernel(const __grid_constant__ MtTable T,
                                                           const double* __restrict__ norm2, float rho, int adaptive) {
  int t; long long lo, hi;
  mt_locate(T, t, lo, hi);
  float* p = static_cast<float*>(T.a[t]);
  const float* g = static_cast<const float*>(T.b[t]);
  float* old = static_cast<float*>(T.c[t]);
  const float scale = static_cast<float>(static_cast<double>(rho) / (sqrt(*norm2) + 1e-12));
  auto one = [&](float w, float gi) { return w + (adaptive ? w * w : 1.0f) * gi * scale; };
  mt_stream(lo, hi, al16(p) && al16(g) && al16(old),
            [&](long long i) {
              const float4 w = ld4(p + i), gi = ld4(g + i);
              st4(old + i, w);
              st4(p + i, make_float4(one(w.x, gi.x), one(w.y, gi.y), one(w.z, gi.z), one(w.w, gi.w)));
            },
            [&](long long i) {
              const float w = p[i];
              old[i] = w;
              p[i] = one(w, g[i]);
            });
}

// SAM second step + AdamW: w = old (if given) ; torch.optim.AdamW single-tensor update     a = p, b = g, c = m, d = v, e = old
__global__ void __launch_bounds__(256) mt_adamw_kernel(const __grid_constant__ MtTable T, float lr, float beta1,
                                                       float beta2, float eps, float wd, float bc1, float rsqrt_bc2) {
  int t; long long lo, hi;
  mt_locate(T, t, lo, hi);
  float* p = static_cast<float*>(T.a[t]);
  const float* g = static_cast<const float*>(T.b[t]);
  float* m = static_cast<float*>(T.c[t]);
  float* v = static_cast<float*>(T.d[t]);
  const float* old = static_cast<const float*>(T.e[t]);
  const float step_size = lr / bc1;
  auto one = [&](float w, float gi, float& mi, float& vi) {
    w *= 1.0f - lr * wd;                                  // decoupled weight decay
    mi = mi + (gi - mi) * (1.0f - beta1);                 // exp_avg.lerp_(g, 1 - beta1)
    vi = vi * beta2 + gi * gi * (1.0f - beta2);
    return w - step_size * (mi / (sqrtf(vi) * rsqrt_bc2 + eps));
  };
  const float* wsrc = old ? old : p;
  mt_stream(lo, hi, al16(p) && al16(g) && al16(m) && al16(v) && al16(wsrc),
            [&](long long i) {
              const float4 w = ld4(wsrc + i), gi = ld4(g + i);
              float4 mi = ld4(m + i), vi = ld4(v + i);
              const float4 o = make_float4(one(w.x, gi.x, mi.x, vi.x), one(w.y, gi.y, mi.y, vi.y),
                                           one(w.z, gi.z, mi.z, vi.z), one(w.w, gi.w, mi.w, vi.w));
              st4(m + i, mi); st4(v + i, vi); st4(p + i, o);
            },
            [&](long long i) {
              float mi = m[i], vi = v[i];
              const float o = one(wsrc[i], g[i], mi, vi);
              m[i] = mi; v[i] = vi; p[i] = o;
            });
}

// EMA: ema = ema * decay + (1 - decay) * src          a = ema, b = src
__global__ void __launch_bounds__(256) mt_ema_kernel(const __grid_constant__ MtTable T, float decay) {
  int t; long long lo, hi;
  mt_locate(T, t, lo, hi);
  float* e = static_cast<float*>(T.a[t]);
  const float* s = static_cast<const float*>(T.b[t]);
  const float om = 1.0f - decay;
  mt_stream(lo, hi, al16(e) && al16(s),
            [&](long long i) {
              const float4 a = ld4(e + i), b = ld4(s + i);
              st4(e + i, make_float4(a.x * decay + om * b.x, a.y * decay + om * b.y, a.z * decay + om * b.z,
                                     a.w * decay + om * b.w));
            },
            [&](long long i) { e[i] = e[i] * decay + om * s[i]; });
}

}  // namespace htrvt

using namespace htrvt;

namespace {
// fills T from host pointer arrays [base, base + cnt); returns the number of chunks (= blocks)
int mt_fill(MtTable& T, int base, int cnt, void* const* a, void* const* b, void* const* c, void* const* d,
            void* const* e, const long long* numel) {
  T.n = cnt;
  int chunks = 0;
  for (int i = 0; i < cnt; ++i) {
    T.a[i] = a ? a[base + i] : nullptr;
    T.b[i] = b ? b[base + i] : nullptr;
    T.c[i] = c ? c[base + i] : nullptr;
    T.d[i] = d ? d[base + i] : nullptr;
    T.e[i] = e ? e[base + i] : nullptr;
    T.numel[i] = numel[base + i];
    T.chunk_start[i] = chunks;
    chunks += static_cast<int>((numel[base + i] + kMtChunk - 1) / kMtChunk);
  }
  T.chunk_start[cnt] = chunks;
  return chunks;
}
}  // namespace

// norm2 (device double, caller zeroes it) += sum |g|^2; p may be NULL unless adaptive
extern "C" int htrvt_mt_sqnorm(int n, void* const* g, void* const* p, const long long* numel, int adaptive,
                               double* norm2, cudaStream_t stream) {
  if (n < 0 || !norm2 || (adaptive && !p)) return HTRVT_ERR_SHAPE;
  for (int base = 0; base < n; base += kMtMax) {
    MtTable T = {};
    const int chunks = mt_fill(T, base, n - base < kMtMax ? n - base : kMtMax, g, p, nullptr, nullptr, nullptr, numel);
    if (!chunks) continue;
    mt_sqnorm_kernel<<<chunks, 256, 0, stream>>>(T, adaptive, norm2);
    HTRVT_LAUNCH_CHECK();
  }
  return HTRVT_OK;
}

extern "C" int htrvt_mt_sam_first(int n, void* const* p, void* const* g, void* const* old_p, const long long* numel,
                                  const double* norm2, float rho, int adaptive, cudaStream_t stream) {
  if (n < 0 || !norm2) return HTRVT_ERR_SHAPE;
  for (int base = 0; base < n; base += kMtMax) {
    MtTable T = {};
    const int chunks = mt_fill(T, base, n - base < kMtMax ? n - base : kMtMax, p, g, old_p, nullptr, nullptr, numel);
    if (!chunks) continue;
    mt_sam_first_kernel<<<chunks, 256, 0, stream>>>(T, norm2, rho, adaptive);
    HTRVT_LAUNCH_CHECK();
  }
  return HTRVT_OK;
}

// old_p: NULL (plain AdamW) or the saved weights to restore before the update (SAM second step)
extern "C" int htrvt_mt_adamw(int n, void* const* p, void* const* g, void* const* exp_avg, void* const* exp_avg_sq,
                              void* const* old_p, const long long* numel, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int step, cudaStream_t stream) {
  if (n < 0 || step < 1) return HTRVT_ERR_SHAPE;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  for (int base = 0; base < n; base += kMtMax) {
    MtTable T = {};
    const int chunks = mt_fill(T, base, n - base < kMtMax ? n - base : kMtMax, p, g, exp_avg, exp_avg_sq, old_p, numel);
    if (!chunks) continue;
    mt_adamw_kernel<<<chunks, 256, 0, stream>>>(T, lr, beta1, beta2, eps, weight_decay, bc1, 1.0f / sqrtf(bc2));
    HTRVT_LAUNCH_CHECK();
  }
  return HTRVT_OK;
}

extern "C" int htrvt_mt_ema(int n, void* const* ema, void* const* src, const long long* numel, float decay,
                            cudaStream_t stream) {
  if (n < 0) return HTRVT_ERR_SHAPE;
  for (int base = 0; base < n; base += kMtMax) {
    MtTable T = {};
    const int chunks = mt_fill(T, base, n - base < kMtMax ? n - base : kMtMax, ema, src, nullptr, nullptr, nullptr, numel);
    if (!chunks) continue;
    mt_ema_kernel<<<chunks, 256, 0, stream>>>(T, decay);
    HTRVT_LAUNCH_CHECK();
  }
  return HTRVT_OK;
}
