// HBM-bound kernels of the conv stem (sm_100a), NHWC bf16 activations:
//   first conv (1 input channel, K = 9: direct, memory bound) + batch statistics,
//   BatchNorm finalise (batch or running statistics, running-stat update), BN-apply + ReLU (+ residual),
//   3x3 / stride (2,1) max-pool fused with BN + ReLU (+ arg-max byte for the backward),
//   BatchNorm backward (two-stage reduction + apply, ReLU mask fused), max-pool backward,
//   first-conv weight gradient.
// Replaces the ATen / cuDNN dispatches behind ResNet18 / BasicBlock (model_v1/model/resnet18.py:10-39,
// 42-84): BatchNorm2d(eps=1e-5, momentum 0.1), ReLU, MaxPool2d(3, (2,1), 1), residual adds.
#include "common.cuh"

namespace htrvt {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
}
// 16-byte read of data that is touched once: keep it out of L1
__device__ __forceinline__ uint4 ld_stream(const __nv_bfloat16* p) {
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}
// F16 = true: the FORWARD stem tensors (activations / raw conv outputs) are IEEE fp16; gradients are always bf16.
// The 16-bit element type is only a bit pattern to these kernels, so pointers stay __nv_bfloat16* ("16-bit word").
template <bool F16>
__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  if (!F16) { unpack8(u, f); return; }
  float2 t;
  t = unpack_f16(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_f16(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_f16(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_f16(u.w); f[6] = t.x; f[7] = t.y;
}
template <bool F16>
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  if (!F16) return pack8(f);
  uint4 u;
  u.x = pack_f16(f[0], f[1]); u.y = pack_f16(f[2], f[3]);
  u.z = pack_f16(f[4], f[5]); u.w = pack_f16(f[6], f[7]);
  return u;
}
template <bool F16>
__device__ __forceinline__ float round16(float v) {      // the value a 16-bit store of v would hold
  return F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v));
}

// ------------------------------------------------------------------------------------------------
// conv1: x [B,H,W] bf16 (one channel) -> raw [B,H/2,W,C] bf16, stride (2,1), pad 1, + channel statistics
// grid (W/128 tiles, H/2, B); blockDim = C threads: thread -> channel pair (t % (C/2)), pixel half (t / (C/2))
// ------------------------------------------------------------------------------------------------
__global__ void conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 __nv_bfloat16* __restrict__ raw, float* __restrict__ partial, int B, int H, int W,
                                 int C) {
  __shared__ float in[3][132];
  extern __shared__ float red[];                     // [2][C] statistics of the second pixel half
  const int w0 = blockIdx.x * 128, ho = blockIdx.y, Ho = H / 2;
  const int half_c = C / 2;
  const int cp = threadIdx.x % half_c, ph = threadIdx.x / half_c;
  float wa[9], wb[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { wa[k] = w[(2 * cp) * 9 + k]; wb[k] = w[(2 * cp + 1) * 9 + k]; }
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (int n = blockIdx.z; n < B; n += gridDim.z) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 130; i += blockDim.x) {
      const int r = i / 130, c = i - r * 130;
      const int hh = 2 * ho + r - 1, ww = w0 + c - 1;
      in[r][c] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(static_cast<long long>(n) * H + hh) * W + ww] : 0.f;
    }
    __syncthreads();
    __nv_bfloat16* orow = raw + ((static_cast<long long>(n) * Ho + ho) * W + w0) * C + 2 * cp;
    for (int p = ph * 64; p < ph * 64 + 64; ++p) {
      if (w0 + p >= W) break;
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float v = in[kh][p + kw];
          a = fmaf(v, wa[kh * 3 + kw], a);
          b = fmaf(v, wb[kh * 3 + kw], b);
        }
      const uint32_t u = pack_bf16(a, b);
      *reinterpret_cast<uint32_t*>(orow + static_cast<long long>(p) * C) = u;
      const float2 r = unpack_bf16(u);                 // statistics of what is stored
      s0 += r.x; s1 += r.y; q0 += r.x * r.x; q1 += r.y * r.y;
    }
  }
  if (!partial) return;
  if (ph == 1) {
    red[2 * cp] = s0; red[2 * cp + 1] = s1; red[C + 2 * cp] = q0; red[C + 2 * cp + 1] = q1;
  }
  __syncthreads();
  if (ph == 0) {
    const long long cta = (static_cast<long long>(blockIdx.z) * gridDim.y + ho) * gridDim.x + blockIdx.x;
    float* dst = partial + cta * 2 * C;
    dst[2 * cp] = s0 + red[2 * cp];
    dst[2 * cp + 1] = s1 + red[2 * cp + 1];
    dst[C + 2 * cp] = q0 + red[C + 2 * cp];
    dst[C + 2 * cp + 1] = q1 + red[C + 2 * cp + 1];
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm finalise: partial [R][2][C] (sum, sum of squares) -> mean, rstd, scale, shift (+ running stats)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int R, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches_tracked, float momentum, float eps,
                                   int training, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                   float* __restrict__ scale_out, float* __restrict__ shift_out, int C) {
  // block = 8 channels x 128 row lanes (C / 8 blocks: 24 at C = 192 instead of the 6 a 32-channel block gives - the
  // pass reads up to 6 MB of partial rows and was limited by the number of SMs it reached); a warp reads four rows x
  // 32-byte segments; four independent accumulators per lane keep the loads in flight
  __shared__ double sh[2][32][8];
  const int c8 = threadIdx.x & 7, lr = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + c8;
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
  double s = 0.0, q = 0.0;
  if (training && c < C) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
    int r = lr;
    for (; r + 384 < R; r += 512) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s4[u] += partial[(static_cast<long long>(r + 128 * u) * 2) * C + c];
        q4[u] += partial[(static_cast<long long>(r + 128 * u) * 2 + 1) * C + c];
      }
    }
    for (; r < R; r += 128) {
      s4[0] += partial[(static_cast<long long>(r) * 2) * C + c];
      q4[0] += partial[(static_cast<long long>(r) * 2 + 1) * C + c];
    }
    s = (static_cast<double>(s4[0]) + s4[1]) + (static_cast<double>(s4[2]) + s4[3]);
    q = (static_cast<double>(q4[0]) + q4[1]) + (static_cast<double>(q4[2]) + q4[3]);
  }
  // lanes with the same channel inside a warp differ in lane bits 3 and 4
  s += __shfl_xor_sync(0xffffffffu, s, 8);  q += __shfl_xor_sync(0xffffffffu, q, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16); q += __shfl_xor_sync(0xffffffffu, q, 16);
  if ((threadIdx.x & 31) < 8) { sh[0][threadIdx.x >> 5][c8] = s; sh[1][threadIdx.x >> 5][c8] = q; }
  __syncthreads();
  if (threadIdx.x >= 8 || c >= C) return;
  float mean, var;
  if (training) {
    s = 0.0; q = 0.0;
    for (int k = 0; k < 32; ++k) { s += sh[0][k][c8]; q += sh[1][k][c8]; }
    const double m = s / count;
    double v = q / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = static_cast<float>(m);
    var = static_cast<float>(v);
    if (running_mean) {
      const double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  scale_out[c] = sc;
  shift_out[c] = beta[c] - mean * sc;
}

// ------------------------------------------------------------------------------------------------
// y = [relu](raw*scale + shift [+ res | + raw2*scale2 + shift2])       8 channels per thread
// ------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ raw, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res,
                                  const __nv_bfloat16* __restrict__ raw2, const float* __restrict__ scale2,
                                  const float* __restrict__ shift2, __nv_bfloat16* __restrict__ y,
                                  __nv_bfloat16* __restrict__ y_bf, uint8_t* __restrict__ mask, long long n8, int C,
                                  int relu) {
  const int G = C / 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % G) * 8;
    float v[8], s[8], b[8];
    unpack8f<F16>(*reinterpret_cast<const uint4*>(raw + i * 8), v);
    *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(scale + c);
    *reinterpret_cast<float4*>(s + 4) = *reinterpret_cast<const float4*>(scale + c + 4);
    *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(shift + c);
    *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(shift + c + 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], s[k], b[k]);
    if (res) {
      float r[8];
      unpack8f<F16>(*reinterpret_cast<const uint4*>(res + i * 8), r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += r[k];
    } else if (raw2) {
      float r[8];
      unpack8f<F16>(*reinterpret_cast<const uint4*>(raw2 + i * 8), r);
      *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(scale2 + c);
      *reinterpret_cast<float4*>(s + 4) = *reinterpret_cast<const float4*>(scale2 + c + 4);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(shift2 + c);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(shift2 + c + 4);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += fmaf(r[k], s[k], b[k]);
    }
    if (relu) {
      unsigned m = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        m |= (v[k] > 0.f ? 1u : 0u) << k;
        v[k] = fmaxf(v[k], 0.f);
      }
      if (mask) mask[i] = static_cast<uint8_t>(m);       // ReLU mask bits for the backward (1 byte per 8 channels)
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8f<F16>(v);
    if (y_bf) *reinterpret_cast<uint4*>(y_bf + i * 8) = pack8(v);   // bf16 copy: operand of the next conv's weight gradient
  }
}

// ------------------------------------------------------------------------------------------------
// max-pool 3x3, stride (2,1), pad 1 over a = relu(raw*scale+shift) (scale == null: a = raw as is)
// out [B,Ho,W,C]; idx (optional) = kh*3+kw of the first maximum (torch's arg-max rule)
// ------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void pool_fwd_kernel(const __nv_bfloat16* __restrict__ raw, const float* __restrict__ scale,
                                const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                                uint8_t* __restrict__ idx, int B, int H, int W, int C) {
  const int G = C / 8, Ho = (H - 1) / 2 + 1;
  const long long n8 = static_cast<long long>(B) * Ho * W * G;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % G);
    long long r = i / G;
    const int wo = static_cast<int>(r % W); r /= W;
    const int ho = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    float s[8], b[8];
    if (scale) {
      *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(scale + g * 8);
      *reinterpret_cast<float4*>(s + 4) = *reinterpret_cast<const float4*>(scale + g * 8 + 4);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(shift + g * 8);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(shift + g * 8 + 4);
    }
    float best[8];
    int bi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bi[k] = 0; }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hh = 2 * ho + kh - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ww = wo + kw - 1;
        if (ww < 0 || ww >= W) continue;
        float v[8];
        unpack8f<F16>(*reinterpret_cast<const uint4*>(raw + ((static_cast<long long>(n) * H + hh) * W + ww) * C + g * 8), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (scale) v[k] = round16<F16>(fmaxf(fmaf(v[k], s[k], b[k]), 0.f));
          if (v[k] > best[k]) { best[k] = v[k]; bi[k] = kh * 3 + kw; }
        }
      }
    }
    *reinterpret_cast<uint4*>(out + i * 8) = pack8f<F16>(best);
    if (idx) {
      uint2 u;
      u.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      u.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + i * 8) = u;
    }
  }
}

// gin[n,h,w,c] = sum over windows (ho,wo) containing (h,w) of gout[ho,wo] * [idx[ho,wo] == position code]
// optional relu mask: multiply by (relu(raw*scale+shift) > 0); optional fp32 gout (token gradient).
template <typename GT, bool F16>
__global__ void pool_bwd_kernel(const GT* __restrict__ gout, const uint8_t* __restrict__ idx,
                                const __nv_bfloat16* __restrict__ raw, const float* __restrict__ scale,
                                const float* __restrict__ shift, __nv_bfloat16* __restrict__ gin, int B, int H,
                                int W, int C) {
  const int G = C / 8, Ho = (H - 1) / 2 + 1;
  const long long n8 = static_cast<long long>(B) * H * W * G;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % G);
    long long r = i / G;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // windows: ho with 2ho-1 <= h <= 2ho+1
    const int ho_lo = h / 2, ho_hi = (h + 1) / 2;       // equal when h is even
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      if (ho >= Ho) continue;
      const int kh = h - (2 * ho - 1);
      for (int wo = w - 1; wo <= w + 1; ++wo) {
        if (wo < 0 || wo >= W) continue;
        const int kw = w - (wo - 1);
        const int code = kh * 3 + kw;
        const long long o = ((static_cast<long long>(n) * Ho + ho) * W + wo) * C + g * 8;
        const uint2 iv = *reinterpret_cast<const uint2*>(idx + o);
        float gv[8];
        if (sizeof(GT) == 4) {
          *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(gout) + o);
          *reinterpret_cast<float4*>(gv + 4) = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(gout) + o + 4);
        } else {
          unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(gout) + o), gv);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int id = ((k < 4 ? iv.x : iv.y) >> (8 * (k & 3))) & 0xff;
          if (id == code) acc[k] += gv[k];
        }
      }
    }
    if (scale) {
      float v[8];
      unpack8f<F16>(*reinterpret_cast<const uint4*>(raw + i * 8), v);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (!(fmaf(v[k], scale[g * 8 + k], shift[g * 8 + k]) > 0.f)) acc[k] = 0.f;
    }
    *reinterpret_cast<uint4*>(gin + i * 8) = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm backward, stage 1: per-CTA partial sums of g' = g * [y > 0] and g' * xhat (one or two BNs)
// partial [cta][3][C] : sum g', sum g' xhat_a, sum g' xhat_b
// ------------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256, 2)             // <= 128 registers: two CTAs per SM keep enough loads in flight
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ g, const uint8_t* __restrict__ mask,
                                     const __nv_bfloat16* __restrict__ raw_a, const float* __restrict__ mean_a,
                                     const float* __restrict__ rstd_a, const __nv_bfloat16* __restrict__ raw_b,
                                     const float* __restrict__ mean_b, const float* __restrict__ rstd_b,
                                     float* __restrict__ partial, long long P, int C, int rows_per_cta) {
  extern __shared__ float sm[];                         // [R][3][C]
  const int G = C / 8, R = blockDim.x / G;
  const int grp = threadIdx.x % G, lane_r = threadIdx.x / G;
  float s0[8], s1[8], s2[8], ma[8], mb[8];               // s1 / s2 accumulate g' (x - mean); rstd multiplies once at the end
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s0[k] = s1[k] = s2[k] = 0.f;
    ma[k] = mean_a[grp * 8 + k];
    mb[k] = raw_b ? mean_b[grp * 8 + k] : 0.f;
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < P ? r0 + rows_per_cta : P;
  if (lane_r < R) {
    for (long long r = r0 + lane_r; r < r1; r += 4 * R) {
      uint4 gq[4], xq[4], bq[4];
      unsigned mk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                       // issue every load of four rows before using any
        const long long rr = r + static_cast<long long>(u) * R;
        const bool ok = rr < r1;
        const long long o = rr * C + grp * 8;
        gq[u] = ok ? ld_stream(g + o) : make_uint4(0u, 0u, 0u, 0u);
        xq[u] = ok ? ld_stream(raw_a + o) : make_uint4(0u, 0u, 0u, 0u);
        bq[u] = (ok && raw_b) ? ld_stream(raw_b + o) : make_uint4(0u, 0u, 0u, 0u);
        mk[u] = (ok && mask) ? mask[rr * G + grp] : 0xffu;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float gv[8], xa[8];
        unpack8(gq[u], gv);
        unpack8f<F16>(xq[u], xa);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (!((mk[u] >> k) & 1u)) gv[k] = 0.f;
          s0[k] += gv[k];
          s1[k] = fmaf(gv[k], xa[k] - ma[k], s1[k]);
        }
        if (raw_b) {
          float xb[8];
          unpack8f<F16>(bq[u], xb);
#pragma unroll
          for (int k = 0; k < 8; ++k) s2[k] = fmaf(gv[k], xb[k] - mb[k], s2[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sm[(lane_r * 3 + 0) * C + grp * 8 + k] = s0[k];
      sm[(lane_r * 3 + 1) * C + grp * 8 + k] = s1[k] * rstd_a[grp * 8 + k];
      sm[(lane_r * 3 + 2) * C + grp * 8 + k] = raw_b ? s2[k] * rstd_b[grp * 8 + k] : 0.f;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
    float s = 0.f;
    for (int rr = 0; rr < R; ++rr) s += sm[rr * 3 * C + i];
    atomicAdd(partial + i, s);            // one zeroed [3][C] row: a few hundred fp32 atomics per column
  }
}

// stage 2 (inside the apply kernel): dgamma += sum g' xhat, dbeta += sum g'; per-channel affine coefficients of stage 3:
//   d_raw = gamma*rstd*(g' - k1 - xhat*k2) = A*g' + Bc*raw + Cc,  k1 = sum g'/N, k2 = sum g' xhat / N,
//   A = gamma*rstd, Bc = -A*rstd*k2, Cc = -A*k1 + A*rstd*k2*mean.
// stage 3: d_raw = A*g' + Bc*raw + Cc for one or two BNs; optional gz = g' (identity-residual gradient)
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  *reinterpret_cast<float4*>(f) = *reinterpret_cast<const float4*>(p);
  *reinterpret_cast<float4*>(f + 4) = *reinterpret_cast<const float4*>(p + 4);
}
// The per-channel coefficients (stage 2 above) are recomputed by every CTA from the [3][C] sums into shared memory -
// a few flops per channel - so no finalise launch sits between the reduction and this pass; CTA 0 also accumulates
// dgamma / dbeta.
struct BnBwdSide {
  const float* gamma; const float* mean; const float* rstd; float* dgamma; float* dbeta;
};
template <bool F16>
__global__ void __launch_bounds__(256, 4) bn_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ g, const uint8_t* __restrict__ mask, const __nv_bfloat16* __restrict__ raw_a,
    const float* __restrict__ sums, double inv_count, BnBwdSide sa, __nv_bfloat16* __restrict__ d_a,
    const __nv_bfloat16* __restrict__ raw_b, BnBwdSide sb, __nv_bfloat16* __restrict__ d_b,
    __nv_bfloat16* __restrict__ gz, long long n8, int C) {
  extern __shared__ float coef_sm[];                    // [2][3][C]: A, Bc, Cc of BN a and BN b
  float* coef_a = coef_sm;
  float* coef_b = coef_sm + 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s0 = sums[c], s1 = sums[C + c], s2 = sums[2 * C + c];
    const float k1 = static_cast<float>(static_cast<double>(s0) * inv_count);      // (1 / N from the host: no fp64 divide per channel per CTA)
    {
      const float k2 = static_cast<float>(static_cast<double>(s1) * inv_count);
      const float A = sa.gamma[c] * sa.rstd[c];
      coef_a[c] = A;
      coef_a[C + c] = -A * sa.rstd[c] * k2;
      coef_a[2 * C + c] = -A * k1 + A * sa.rstd[c] * k2 * sa.mean[c];
      if (blockIdx.x == 0 && sa.dgamma) { sa.dgamma[c] += s1; sa.dbeta[c] += s0; }
    }
    if (raw_b) {
      const float k2 = static_cast<float>(static_cast<double>(s2) * inv_count);
      const float A = sb.gamma[c] * sb.rstd[c];
      coef_b[c] = A;
      coef_b[C + c] = -A * sb.rstd[c] * k2;
      coef_b[2 * C + c] = -A * k1 + A * sb.rstd[c] * k2 * sb.mean[c];
      if (blockIdx.x == 0 && sb.dgamma) { sb.dgamma[c] += s2; sb.dbeta[c] += s0; }
    }
  }
  __syncthreads();
  const int G = C / 8;
  // a CTA walks 2 x 256 ADJACENT groups per iteration (8 KB per stream: same DRAM pages), then jumps by the grid
  const long long stride = static_cast<long long>(blockDim.x);
  const long long jump = 2 * static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = 2 * blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < n8; i0 += jump) {
    uint4 gq[2], xq[2], bq[2];
    unsigned mk[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {                         // both groups' loads are issued before any use
      const long long i = i0 + u * stride;
      const bool ok = i < n8;
      gq[u] = ok ? ld_stream(g + i * 8) : make_uint4(0u, 0u, 0u, 0u);
      xq[u] = ok ? ld_stream(raw_a + i * 8) : make_uint4(0u, 0u, 0u, 0u);
      bq[u] = (ok && raw_b) ? ld_stream(raw_b + i * 8) : make_uint4(0u, 0u, 0u, 0u);
      mk[u] = (ok && mask) ? mask[i] : 0xffu;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n8) break;
      const int c = static_cast<int>(i % G) * 8;
      float gv[8], xa[8], o[8], A[8], Bc[8], Cc[8];
      unpack8(gq[u], gv);
      unpack8f<F16>(xq[u], xa);
#pragma unroll
      for (int k = 0; k < 8; ++k) if (!((mk[u] >> k) & 1u)) gv[k] = 0.f;
      if (gz) *reinterpret_cast<uint4*>(gz + i * 8) = pack8(gv);
      load8(coef_a + c, A); load8(coef_a + C + c, Bc); load8(coef_a + 2 * C + c, Cc);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], gv[k], fmaf(Bc[k], xa[k], Cc[k]));
      *reinterpret_cast<uint4*>(d_a + i * 8) = pack8(o);
      if (raw_b) {
        float xb[8];
        unpack8f<F16>(bq[u], xb);
        load8(coef_b + c, A); load8(coef_b + C + c, Bc); load8(coef_b + 2 * C + c, Cc);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], gv[k], fmaf(Bc[k], xb[k], Cc[k]));
        *reinterpret_cast<uint4*>(d_b + i * 8) = pack8(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// conv1 weight gradient: dw[c][kh][kw] = sum dy[n,ho,wo,c] * x[n, 2ho+kh-1, wo+kw-1]; partial [cta][9][C]
// persistent CTAs over (n, ho, w-tile) work items; thread -> channel pair / pixel half as in conv1_fwd
// ------------------------------------------------------------------------------------------------
__global__ void conv1_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                   float* __restrict__ partial, int B, int H, int W, int C) {
  __shared__ float in[3][132];
  extern __shared__ float red[];                      // [18][C/2] for the second pixel half
  const int Ho = H / 2, tiles_w = (W + 127) / 128;
  const long long items = static_cast<long long>(B) * Ho * tiles_w;
  const int half_c = C / 2;
  const int cp = threadIdx.x % half_c, ph = threadIdx.x / half_c;
  float ga[9], gb[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { ga[k] = 0.f; gb[k] = 0.f; }
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int tw = static_cast<int>(it % tiles_w);
    const int ho = static_cast<int>((it / tiles_w) % Ho);
    const int n = static_cast<int>(it / (static_cast<long long>(tiles_w) * Ho));
    const int w0 = tw * 128;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 130; i += blockDim.x) {
      const int r = i / 130, c = i - r * 130;
      const int hh = 2 * ho + r - 1, ww = w0 + c - 1;
      in[r][c] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(static_cast<long long>(n) * H + hh) * W + ww] : 0.f;
    }
    __syncthreads();
    const __nv_bfloat16* drow = dy + ((static_cast<long long>(n) * Ho + ho) * W + w0) * C + 2 * cp;
    for (int p0 = ph * 64; p0 < ph * 64 + 64; p0 += 8) {
      uint32_t dv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        dv[j] = (w0 + p0 + j < W) ? *reinterpret_cast<const uint32_t*>(drow + static_cast<long long>(p0 + j) * C) : 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 d = unpack_bf16(dv[j]);
        const int p = p0 + j;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float v = in[kh][p + kw];
            ga[kh * 3 + kw] = fmaf(d.x, v, ga[kh * 3 + kw]);
            gb[kh * 3 + kw] = fmaf(d.y, v, gb[kh * 3 + kw]);
          }
      }
    }
  }
  __syncthreads();
  if (ph == 1) {
#pragma unroll
    for (int k = 0; k < 9; ++k) { red[(2 * k) * half_c + cp] = ga[k]; red[(2 * k + 1) * half_c + cp] = gb[k]; }
  }
  __syncthreads();
  if (ph == 0) {
    float* dst = partial + static_cast<long long>(blockIdx.x) * 9 * C;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      dst[(2 * cp) * 9 + k] = ga[k] + red[(2 * k) * half_c + cp];
      dst[(2 * cp + 1) * 9 + k] = gb[k] + red[(2 * k + 1) * half_c + cp];
    }
  }
}

__global__ void colsum_finalize2_kernel(const float* __restrict__ partial, int R, long long stride, int n,
                                        float* __restrict__ out, int accumulate) {
  __shared__ double sh[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), lr = threadIdx.x >> 5;
  double s = 0.0;
  if (c < n)
    for (int r = lr; r < R; r += 8) s += partial[r * stride + c];
  sh[lr][threadIdx.x & 31] = s;
  __syncthreads();
  if (lr != 0 || c >= n) return;
  s = 0.0;
  for (int k = 0; k < 8; ++k) s += sh[k][threadIdx.x];
  out[c] = accumulate ? out[c] + static_cast<float>(s) : static_cast<float>(s);
}

}  // namespace htrvt

using namespace htrvt;

static inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  return static_cast<int>(g < 1 ? 1 : (g > 148LL * 16 ? 148LL * 16 : g));
}

extern "C" int htrvt_conv1_fwd_zdim(int B) { return B < 8 ? B : 8; }

// partial: fp32 [htrvt_conv1_fwd_zdim(B) * (H/2) * ceil(W/128)][2][C]  (null: no statistics, eval mode)
extern "C" int htrvt_conv1_fwd(const float* x, const float* w, void* raw_bf16, float* partial, int B, int H,
                               int W, int C, cudaStream_t stream) {
  if (B <= 0 || (H & 1) || W <= 0 || (C & 1) || C > 1024 || C < 2) return HTRVT_ERR_SHAPE;
  dim3 grid((W + 127) / 128, H / 2, htrvt_conv1_fwd_zdim(B));
  conv1_fwd_kernel<<<grid, C, 2 * C * sizeof(float), stream>>>(x, w, static_cast<__nv_bfloat16*>(raw_bf16), partial,
                                                               B, H, W, C);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_bn_finalize(const float* partial, int R, double count, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, long long* num_batches_tracked,
                                 float momentum, float eps, int training, float* mean, float* rstd, float* scale,
                                 float* shift, int C, cudaStream_t stream) {
  if (C <= 0 || (training && (!partial || R <= 0))) return HTRVT_ERR_SHAPE;
  bn_finalize_kernel<<<(C + 7) / 8, 1024, 0, stream>>>(partial, R, count, gamma, beta, running_mean, running_var,
                                                          num_batches_tracked, momentum, eps, training, mean, rstd,
                                                          scale, shift, C);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_bn_act_fwd(const void* raw, const float* scale, const float* shift, const void* res,
                                const void* raw2, const float* scale2, const float* shift2, void* y, void* y_bf16,
                                void* mask, long long P, int C, int relu, int f16, cudaStream_t stream) {
  if (P <= 0 || (C & 7)) return HTRVT_ERR_SHAPE;
  const long long n8 = P * C / 8;
  auto kern = f16 ? bn_act_fwd_kernel<true> : bn_act_fwd_kernel<false>;     // f16: raw / res / raw2 / y are fp16
  kern<<<grid_for(n8, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(raw), scale, shift, static_cast<const __nv_bfloat16*>(res),
      static_cast<const __nv_bfloat16*>(raw2), scale2, shift2, static_cast<__nv_bfloat16*>(y),
      static_cast<__nv_bfloat16*>(y_bf16), static_cast<uint8_t*>(mask), n8, C, relu);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_pool_fwd(const void* raw, const float* scale, const float* shift, void* out, void* idx, int B,
                              int H, int W, int C, int f16, cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || (C & 7)) return HTRVT_ERR_SHAPE;
  const int Ho = (H - 1) / 2 + 1;
  const long long n8 = static_cast<long long>(B) * Ho * W * C / 8;
  auto kern = f16 ? pool_fwd_kernel<true> : pool_fwd_kernel<false>;         // f16: raw and out are fp16
  kern<<<grid_for(n8, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(raw), scale, shift,
                                              static_cast<__nv_bfloat16*>(out), static_cast<uint8_t*>(idx), B, H, W, C);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_pool_bwd(const void* gout, int gout_is_f32, const void* idx, const void* raw, const float* scale,
                              const float* shift, void* gin, int B, int H, int W, int C, int raw_f16,
                              cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || (C & 7) || !idx) return HTRVT_ERR_SHAPE;
  const long long n8 = static_cast<long long>(B) * H * W * C / 8;
  const uint8_t* ix = static_cast<const uint8_t*>(idx);
  const __nv_bfloat16* rw = static_cast<const __nv_bfloat16*>(raw);
  __nv_bfloat16* gi = static_cast<__nv_bfloat16*>(gin);
  const int grid = grid_for(n8, 256);
  if (gout_is_f32) {
    const float* go = static_cast<const float*>(gout);
    if (raw_f16) pool_bwd_kernel<float, true><<<grid, 256, 0, stream>>>(go, ix, rw, scale, shift, gi, B, H, W, C);
    else pool_bwd_kernel<float, false><<<grid, 256, 0, stream>>>(go, ix, rw, scale, shift, gi, B, H, W, C);
  } else {
    const __nv_bfloat16* go = static_cast<const __nv_bfloat16*>(gout);
    if (raw_f16) pool_bwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(go, ix, rw, scale, shift, gi, B, H, W, C);
    else pool_bwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(go, ix, rw, scale, shift, gi, B, H, W, C);
  }
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// the apply pass: every CTA first derives the per-channel coefficients (C channels), so one resident wave of CTAs
// (4 per SM) strides over the tensor instead of up to 16 waves repeating that prologue
static inline int apply_grid(long long n8) {
  const long long g = (n8 + 511) / 512;
  return static_cast<int>(g < 1 ? 1 : (g > 148LL * 4 ? 148LL * 4 : g));
}

extern "C" int htrvt_bn_bwd_ctas(long long P) {
  long long c = (P + 63) / 64;
  return static_cast<int>(c > 592 ? 592 : (c < 1 ? 1 : c));
}

// BatchNorm backward for the BN (a) that produced `raw_a` (and optionally a second BN (b) fed by the same
// upstream gradient: the downsample branch).  g: gradient w.r.t. the post-activation output; `mask` (optional) holds
// the ReLU mask bits written by htrvt_bn_act_fwd (1 byte per 8 channels).  Writes d_a (/d_b) = gradient w.r.t. the raw conv outputs, accumulates
// dgamma / dbeta, optionally writes gz = masked g (identity-residual gradient).
// partial: fp32 [3][C] column sums, ZEROED by the caller (two launches: reduce -> apply; the apply pass derives the
// per-channel coefficients itself); coef: unused, kept for ABI stability.
extern "C" int htrvt_bn_bwd(const void* g, const void* mask, const void* raw_a, const float* mean_a,
                            const float* rstd_a, const float* gamma_a, float* dgamma_a, float* dbeta_a, void* d_a,
                            const void* raw_b, const float* mean_b, const float* rstd_b, const float* gamma_b,
                            float* dgamma_b, float* dbeta_b, void* d_b, void* gz, long long P, int C,
                            float* partial, float* coef, int raw_f16, cudaStream_t stream) {
  if (P <= 0 || (C & 7) || C > 2048) return HTRVT_ERR_SHAPE;
  const int G = C / 8;
  if (G > 256) return HTRVT_ERR_SHAPE;
  const int R = 256 / G;
  const int threads = G * R;
  const int ctas = htrvt_bn_bwd_ctas(P);
  const int rows = static_cast<int>((P + ctas - 1) / ctas);
  const size_t smem = static_cast<size_t>(R) * 3 * C * sizeof(float);
  if (!(raw_f16 ? HTRVT_ENSURE_SMEM(bn_bwd_reduce_kernel<true>, 96 * 1024)
                : HTRVT_ENSURE_SMEM(bn_bwd_reduce_kernel<false>, 96 * 1024)))
    return HTRVT_ERR_LAUNCH;
  auto k_reduce = raw_f16 ? bn_bwd_reduce_kernel<true> : bn_bwd_reduce_kernel<false>;
  auto k_apply = raw_f16 ? bn_bwd_apply_kernel<true> : bn_bwd_apply_kernel<false>;
  if (smem > 96 * 1024) return HTRVT_ERR_SHAPE;
  // partial: [3][C] sums, ZERO on entry (the caller hands out slices of one buffer it cleared once per backward)
  k_reduce<<<ctas, threads, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(g), static_cast<const uint8_t*>(mask),
      static_cast<const __nv_bfloat16*>(raw_a), mean_a, rstd_a, static_cast<const __nv_bfloat16*>(raw_b), mean_b,
      rstd_b, partial, P, C, rows);
  HTRVT_LAUNCH_CHECK();
  (void)coef;
  const long long n8 = P * C / 8;
  const BnBwdSide sa = {gamma_a, mean_a, rstd_a, dgamma_a, dbeta_a};
  const BnBwdSide sb = {gamma_b, mean_b, rstd_b, dgamma_b, dbeta_b};
  k_apply<<<apply_grid(n8), 256, static_cast<size_t>(6) * C * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(g), static_cast<const uint8_t*>(mask),
      static_cast<const __nv_bfloat16*>(raw_a), partial, 1.0 / static_cast<double>(P), sa, static_cast<__nv_bfloat16*>(d_a),
      static_cast<const __nv_bfloat16*>(raw_b), sb, static_cast<__nv_bfloat16*>(d_b),
      static_cast<__nv_bfloat16*>(gz), n8, C);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// Second half of htrvt_bn_bwd alone: the [3][C] sums were produced elsewhere (htrvt_conv_dgrad_bn's epilogue) and g is
// already masked (g' = g * [y > 0]).  d_a = A g' + B raw + C, dgamma += sum g' xhat, dbeta += sum g'.
extern "C" int htrvt_bn_bwd_apply(const void* g_masked, const void* raw_a, const float* mean_a, const float* rstd_a,
                                  const float* gamma_a, float* dgamma_a, float* dbeta_a, void* d_a, long long P, int C,
                                  const float* sums, int raw_f16, cudaStream_t stream) {
  if (P <= 0 || (C & 7) || C > 2048 || !sums) return HTRVT_ERR_SHAPE;
  auto k_apply = raw_f16 ? bn_bwd_apply_kernel<true> : bn_bwd_apply_kernel<false>;
  const long long n8 = P * C / 8;
  const BnBwdSide sa = {gamma_a, mean_a, rstd_a, dgamma_a, dbeta_a};
  const BnBwdSide sb = {nullptr, nullptr, nullptr, nullptr, nullptr};
  k_apply<<<apply_grid(n8), 256, static_cast<size_t>(6) * C * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(g_masked), nullptr, static_cast<const __nv_bfloat16*>(raw_a), sums,
      1.0 / static_cast<double>(P), sa, static_cast<__nv_bfloat16*>(d_a), nullptr, sb, nullptr, nullptr, n8, C);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

extern "C" int htrvt_conv1_wgrad_ctas() { return 148 * 4; }

// grad [C][1][3][3] fp32 (+=); partial: fp32 [htrvt_conv1_wgrad_ctas()][9*C]
extern "C" int htrvt_conv1_wgrad(const void* dy_bf16, const float* x, float* grad, int accumulate,
                                 float* partial, int B, int H, int W, int C, cudaStream_t stream) {
  if (B <= 0 || (H & 1) || W <= 0 || (C & 1) || C > 1024 || C < 2) return HTRVT_ERR_SHAPE;
  const int ctas = htrvt_conv1_wgrad_ctas();
  conv1_wgrad_kernel<<<ctas, C, 18 * (C / 2) * sizeof(float), stream>>>(
      static_cast<const __nv_bfloat16*>(dy_bf16), x, partial, B, H, W, C);
  HTRVT_LAUNCH_CHECK();
  colsum_finalize2_kernel<<<(9 * C + 31) / 32, 256, 0, stream>>>(partial, ctas, 9LL * C, 9 * C, grad, accumulate);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
