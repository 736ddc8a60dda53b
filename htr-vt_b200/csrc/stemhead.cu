// Fused stem head (sm_100a): conv1 (1 -> C channels, 3x3, stride (2,1), pad 1, no bias) -> BatchNorm ->
// ReLU -> MaxPool(3, stride (2,1), pad 1), forward and backward, WITHOUT materialising the conv output.
// Replaces model_v1/model/resnet18.py:48-51 (construction) / :74-77 (forward) and their autograd backward.
//
// The conv has K = 9 on a single input channel, so its [B, H/2, W, C] output (805 MB in bf16 at B = 128)
// is 12x larger than everything needed to recompute it (the fp32 image, 16 MB).  Instead of storing it:
//   * batch statistics come from the 9-tap patch moments of the image: with v(p) the 3x3 input patch of
//     output pixel p, S = sum_p v and R = sum_p v v^T give  sum_p raw_c = w_c . S,  sum_p raw_c^2 = w_c^T R w_c
//     exactly (54 numbers for the whole batch, one cheap pass over the image);
//   * the forward recomputes the conv in fp32 per pooled output (sliding 3-column window in registers) and
//     writes only the pooled activation + a 4-bit arg-max code (4 kh + kw; 15 = ReLU inactive);
//   * the backward needs no per-pixel gradient tensor at all: with g the gradient of the pooled output,
//       dbeta_c = sum_o g,  dgamma_c = sum_o g xhat(argmax o),  G3[c][t] = sum_o g v_t(argmax o)
//     are accumulated in ONE pass over g, and the conv weight gradient follows in closed form from the
//     BatchNorm backward's affine structure  d_raw = A g' + Bc raw + Cc:
//       dW[c][t] = A_c G3[c][t] + Bc_c (w_c^T R)[t] + Cc_c S[t].
// HBM traffic per training step at B = 128: ~0.5 GB (pooled activation, codes, pooled gradient) instead of ~8 GB.
#include "common.cuh"

namespace htrvt {

constexpr int kMom = 54;               // 9 first moments + 45 upper-triangular second moments
constexpr int kHeadInW = 68;           // smem row pitch of the staged image tile: 68 columns used; 68 = 4 mod 32 keeps the
                                       // backward's per-lane window reads (rows 2 kh + i, columns kw + j, kh / kw from the arg-max
                                       // code) on distinct banks - a pitch of 72 put kh = 0 and kh = 2 on the same bank
static_assert((7 * kHeadInW * 4) % 16 == 0, "stage alignment");

// ------------------------------------------------------------------------------------------------
// patch moments: x fp32 [B,H,W] -> partial [gridDim.x][54]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1_moments_kernel(const float* __restrict__ x, float* __restrict__ partial,
                                                            int B, int H, int W) {
  __shared__ float red[8][kMom];
  const int Hp = H / 2;
  const long long total = static_cast<long long>(B) * Hp * W;
  float acc[kMom];
#pragma unroll
  for (int k = 0; k < kMom; ++k) acc[k] = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const long long r = i / W;
    const int hp = static_cast<int>(r % Hp);
    const long long n = r / Hp;
    float v[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hh = 2 * hp + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ww = w + kw - 1;
        v[kh * 3 + kw] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (n * H + hh) * W + ww) : 0.f;
      }
    }
    int k = 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      acc[t] += v[t];
#pragma unroll
      for (int u = t; u < 9; ++u) { acc[k] = fmaf(v[t], v[u], acc[k]); ++k; }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kMom; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[wid][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < kMom) {
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s += red[j][threadIdx.x];
    partial[static_cast<long long>(blockIdx.x) * kMom + threadIdx.x] = s;
  }
}

// moments [54] (fp32) and the per-channel (sum, sum of squares) of the conv output, stats [2][C]
__global__ void conv1_moments_finalize_kernel(const float* __restrict__ partial, int R, const float* __restrict__ w,
                                              int C, float* __restrict__ moments, float* __restrict__ stats) {
  __shared__ double m[kMom];
  for (int t = threadIdx.x; t < kMom; t += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < R; ++r) s += partial[static_cast<long long>(r) * kMom + t];
    m[t] = s;
    moments[t] = static_cast<float>(s);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double wv[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[t] = w[c * 9 + t];
    double s = 0.0, q = 0.0;
    int k = 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      s += wv[t] * m[t];
#pragma unroll
      for (int u = t; u < 9; ++u) { q += (t == u ? 1.0 : 2.0) * wv[t] * wv[u] * m[k]; ++k; }
    }
    stats[c] = static_cast<float>(s);
    stats[C + c] = static_cast<float>(q < 0.0 ? 0.0 : q);
  }
}

// ------------------------------------------------------------------------------------------------
// forward: x fp32 [B,H,W] -> out bf16 [B,Ho,W,C] (Ho = (H/2 - 1)/2 + 1), code nibbles [B,Ho,W,C/2]
// grid (ceil(W/64), ceil(Ho/2), B); blockDim = C: thread -> channel pair (tid % (C/2)), column half (tid / (C/2))
// ------------------------------------------------------------------------------------------------
template <bool CODE, int FMT>       // CODE = false (inference): no arg-max bookkeeping; FMT: output 0 bf16, 1 fp16, 2 fp32
__global__ void __launch_bounds__(256) stem_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             __nv_bfloat16* __restrict__ out,
                                                             __nv_bfloat16* __restrict__ out_bf,
                                                             uint8_t* __restrict__ code, int B, int H, int W, int C) {
  __shared__ float in[11][kHeadInW];
  const int Hp = H / 2, Ho = (Hp - 1) / 2 + 1;
  const int w0 = blockIdx.x * 64, rp = blockIdx.y, n = blockIdx.z;
  const int half_c = C / 2;
  const int cp = threadIdx.x % half_c, half = threadIdx.x / half_c;
  for (int i = threadIdx.x; i < 11 * kHeadInW; i += blockDim.x) {
    const int r = i / kHeadInW, c = i - r * kHeadInW;
    const int hh = 8 * rp - 3 + r, ww = w0 - 2 + c;
    in[r][c] = (c < 68 && hh >= 0 && hh < H && ww >= 0 && ww < W)
                   ? __ldg(x + (static_cast<long long>(n) * H + hh) * W + ww) : 0.f;
  }
  float wa[9], wb[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { wa[k] = __ldg(w + (2 * cp) * 9 + k); wb[k] = __ldg(w + (2 * cp + 1) * 9 + k); }
  const float sc0 = scale[2 * cp], sc1 = scale[2 * cp + 1], sh0 = shift[2 * cp], sh1 = shift[2 * cp + 1];
  __syncthreads();

  const int cb = 32 * half;
  float win[11][3];                 // ring of three input columns
  float cva[2][3], cvb[2][3];       // per pooled row q: column maxima of the last three pre-pool columns
  int cka[2][3], ckb[2][3];         // ... and the kh of each
#pragma unroll
  for (int r = 0; r < 11; ++r) { win[r][0] = in[r][cb]; win[r][1] = in[r][cb + 1]; win[r][2] = 0.f; }
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int s = 0; s < 3; ++s) { cva[q][s] = cvb[q][s] = -INFINITY; cka[q][s] = ckb[q][s] = 0; }

#pragma unroll 1
  for (int j0 = 0; j0 < 36; j0 += 3) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int j = j0 + s;
#pragma unroll
      for (int r = 0; r < 11; ++r) win[r][(s + 2) % 3] = in[r][cb + j + 2];
      const int jj = w0 + cb - 1 + j;                     // pre-pool column
      const bool col_ok = (jj >= 0) && (jj < W);
      float ya[5], yb[5];
#pragma unroll
      for (int r5 = 0; r5 < 5; ++r5) {
        float za = 0.f, zb = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float v = win[2 * r5 + kh][(s + kw) % 3];
            za = fmaf(v, wa[kh * 3 + kw], za);
            zb = fmaf(v, wb[kh * 3 + kw], zb);
          }
        const int hpre = 4 * rp - 1 + r5;
        const bool ok = col_ok && hpre >= 0 && hpre < Hp;
        ya[r5] = ok ? fmaxf(fmaf(za, sc0, sh0), 0.f) : -INFINITY;
        yb[r5] = ok ? fmaxf(fmaf(zb, sc1, sh1), 0.f) : -INFINITY;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float va = ya[2 * q], vb = yb[2 * q];
        int ka = 0, kb = 0;
        if (CODE) {
          if (ya[2 * q + 1] > va) { va = ya[2 * q + 1]; ka = 1; }
          if (ya[2 * q + 2] > va) { va = ya[2 * q + 2]; ka = 2; }
          if (yb[2 * q + 1] > vb) { vb = yb[2 * q + 1]; kb = 1; }
          if (yb[2 * q + 2] > vb) { vb = yb[2 * q + 2]; kb = 2; }
        } else {
          va = fmaxf(va, fmaxf(ya[2 * q + 1], ya[2 * q + 2]));
          vb = fmaxf(vb, fmaxf(yb[2 * q + 1], yb[2 * q + 2]));
        }
        cva[q][s] = va; cka[q][s] = ka; cvb[q][s] = vb; ckb[q][s] = kb;
      }
      const int wo = w0 + cb + j - 2;
      if (j >= 2 && j < 34 && wo < W) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int ho = 2 * rp + q;
          if (ho >= Ho) continue;
          // first maximum in (kh, kw) row-major order = torch's max_pool2d arg-max rule
          float ba = cva[q][(s + 1) % 3], bb = cvb[q][(s + 1) % 3];
          int kha = cka[q][(s + 1) % 3], khb = ckb[q][(s + 1) % 3], kwa = 0, kwb = 0;
          if (CODE) {
#pragma unroll
            for (int kw = 1; kw < 3; ++kw) {
              const int sl = (s + 1 + kw) % 3;
              if (cva[q][sl] > ba || (cva[q][sl] == ba && cka[q][sl] < kha)) { ba = cva[q][sl]; kha = cka[q][sl]; kwa = kw; }
              if (cvb[q][sl] > bb || (cvb[q][sl] == bb && ckb[q][sl] < khb)) { bb = cvb[q][sl]; khb = ckb[q][sl]; kwb = kw; }
            }
          } else {
            ba = fmaxf(ba, fmaxf(cva[q][(s + 2) % 3], cva[q][s % 3]));
            bb = fmaxf(bb, fmaxf(cvb[q][(s + 2) % 3], cvb[q][s % 3]));
          }
          const long long o = ((static_cast<long long>(n) * Ho + ho) * W + wo);
          if (FMT == 2) *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + o * C + 2 * cp) = make_float2(ba, bb);
          else *reinterpret_cast<uint32_t*>(out + o * C + 2 * cp) = FMT == 1 ? pack_f16(ba, bb) : pack_bf16(ba, bb);
          if (out_bf) *reinterpret_cast<uint32_t*>(out_bf + o * C + 2 * cp) = pack_bf16(ba, bb);
          if (CODE && code) {
            const unsigned na = ba > 0.f ? static_cast<unsigned>(kha * 4 + kwa) : 15u;     // code = 4 kh + kw, 15 = no gradient
            const unsigned nb = bb > 0.f ? static_cast<unsigned>(khb * 4 + kwb) : 15u;
            code[o * half_c + cp] = static_cast<uint8_t>(na | (nb << 4));
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward pass over the pooled gradient: partial [cta][11][C] = (G1, G2, G3[9]) with
//   G1 = sum g, G2 = sum g raw(argmax), G3[t] = sum g v_t(argmax)   (ReLU-inactive outputs contribute nothing)
// persistent CTAs over (n, ho, 64-column tile) items; g / code tiles are staged with cp.async (double buffered)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(256) stem_head_bwd_kernel(const __nv_bfloat16* __restrict__ g,
                                                             const uint8_t* __restrict__ code,
                                                             const float* __restrict__ x, const float* __restrict__ w,
                                                             float* __restrict__ partial, int B, int H, int W, int C) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int Hp = H / 2, Ho = (Hp - 1) / 2 + 1;
  const int tiles_w = (W + 63) / 64;
  const long long items = static_cast<long long>(B) * Ho * tiles_w;
  const int half_c = C / 2;
  const int cp = threadIdx.x % half_c, half = threadIdx.x / half_c;
  const int g_bytes = 64 * C * 2, c_bytes = 64 * half_c;
  const int in_bytes = 7 * kHeadInW * 4;
  const int stage_bytes = g_bytes + c_bytes + in_bytes;            // multiples of 16
  float wa[9], wb[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { wa[k] = __ldg(w + (2 * cp) * 9 + k); wb[k] = __ldg(w + (2 * cp + 1) * 9 + k); }
  float g1a = 0.f, g1b = 0.f, g2a = 0.f, g2b = 0.f, g3a[9], g3b[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { g3a[k] = 0.f; g3b[k] = 0.f; }

  auto issue = [&](long long it, int st) {
    const int tw = static_cast<int>(it % tiles_w);
    const int ho = static_cast<int>((it / tiles_w) % Ho);
    const long long n = it / (static_cast<long long>(tiles_w) * Ho);
    const int w0 = tw * 64;
    const int px = (W - w0) < 64 ? (W - w0) : 64;
    uint8_t* sg = dsm + st * stage_bytes;
    uint8_t* sc = sg + g_bytes;
    float* sin = reinterpret_cast<float*>(sc + c_bytes);
    const long long o = (n * Ho + ho) * W + w0;
    const uint8_t* gsrc = reinterpret_cast<const uint8_t*>(g + o * C);
    const uint8_t* csrc = code + o * half_c;
    for (int i = threadIdx.x; i < px * C * 2 / 16; i += blockDim.x) cp_async16(sg + i * 16, gsrc + i * 16);
    for (int i = threadIdx.x; i < px * half_c / 16; i += blockDim.x) cp_async16(sc + i * 16, csrc + i * 16);
    for (int i = threadIdx.x; i < 7 * kHeadInW; i += blockDim.x) {
      const int r = i / kHeadInW, c = i - r * kHeadInW;
      const int hh = 4 * ho - 3 + r, ww = w0 - 2 + c;
      sin[i] = (c < 68 && hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (n * H + hh) * W + ww) : 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int st = 0;
  if (blockIdx.x < items) issue(blockIdx.x, 0);
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const long long nxt = it + gridDim.x;
    if (nxt < items) {
      issue(nxt, st ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int w0 = static_cast<int>(it % tiles_w) * 64;
    const uint8_t* sg = dsm + st * stage_bytes;
    const uint8_t* sc = sg + g_bytes;
    const float* sin = reinterpret_cast<const float*>(sc + c_bytes);
    const int pend = (W - w0 - 32 * half) < 32 ? (W - w0 - 32 * half) : 32;
#pragma unroll 2
    for (int p = 0; p < pend; ++p) {
      const int pl = 32 * half + p;
      const float2 gv = unpack_bf16(*reinterpret_cast<const uint32_t*>(sg + (pl * half_c + cp) * 4));
      const unsigned nb = sc[pl * half_c + cp];
      {
        unsigned cd = nb & 15u;
        const float ga = cd < 11u ? gv.x : 0.f;
        cd = cd < 11u ? cd : 0u;
        const unsigned kh = cd >> 2, kw = cd & 3u;
        const float* base = sin + (2 * kh) * kHeadInW + pl + kw;
        float raw = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float v = base[i * kHeadInW + j];
            raw = fmaf(v, wa[i * 3 + j], raw);
            g3a[i * 3 + j] = fmaf(ga, v, g3a[i * 3 + j]);
          }
        g1a += ga;
        g2a = fmaf(ga, raw, g2a);
      }
      {
        unsigned cd = nb >> 4;
        const float gb = cd < 11u ? gv.y : 0.f;
        cd = cd < 11u ? cd : 0u;
        const unsigned kh = cd >> 2, kw = cd & 3u;
        const float* base = sin + (2 * kh) * kHeadInW + pl + kw;
        float raw = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float v = base[i * kHeadInW + j];
            raw = fmaf(v, wb[i * 3 + j], raw);
            g3b[i * 3 + j] = fmaf(gb, v, g3b[i * 3 + j]);
          }
        g1b += gb;
        g2b = fmaf(gb, raw, g2b);
      }
    }
    __syncthreads();
    st ^= 1;
  }
  // block reduction of the two column halves -> partial [cta][11][C]
  float* red = reinterpret_cast<float*>(dsm);                     // [11][C]
  __syncthreads();
  if (half == 1) {
    red[0 * C + 2 * cp] = g1a; red[0 * C + 2 * cp + 1] = g1b;
    red[1 * C + 2 * cp] = g2a; red[1 * C + 2 * cp + 1] = g2b;
#pragma unroll
    for (int k = 0; k < 9; ++k) { red[(2 + k) * C + 2 * cp] = g3a[k]; red[(2 + k) * C + 2 * cp + 1] = g3b[k]; }
  }
  __syncthreads();
  if (half == 0) {
    float* dst = partial + static_cast<long long>(blockIdx.x) * 11 * C;
    dst[0 * C + 2 * cp] = g1a + red[0 * C + 2 * cp]; dst[0 * C + 2 * cp + 1] = g1b + red[0 * C + 2 * cp + 1];
    dst[1 * C + 2 * cp] = g2a + red[1 * C + 2 * cp]; dst[1 * C + 2 * cp + 1] = g2b + red[1 * C + 2 * cp + 1];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      dst[(2 + k) * C + 2 * cp] = g3a[k] + red[(2 + k) * C + 2 * cp];
      dst[(2 + k) * C + 2 * cp + 1] = g3b[k] + red[(2 + k) * C + 2 * cp + 1];
    }
  }
}

// finalise: dbeta += G1, dgamma += rstd (G2 - mean G1), dW[c][t] += A G3 + Bc (w^T R)[t] + Cc S[t]
// block = 32 channels x 16 row lanes
__global__ void __launch_bounds__(512) stem_head_bwd_finalize_kernel(
    const float* __restrict__ partial, int R, double count, int C, const float* __restrict__ w,
    const float* __restrict__ moments, const float* __restrict__ gamma, const float* __restrict__ mean,
    const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw) {
  __shared__ double sh[16][11][32];
  const int cl = threadIdx.x & 31, lr = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double acc[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) acc[k] = 0.0;
  if (c < C) {
    for (int r = lr; r < R; r += 16) {
      const float* p = partial + static_cast<long long>(r) * 11 * C + c;
#pragma unroll
      for (int k = 0; k < 11; ++k) acc[k] += p[k * C];
    }
  }
#pragma unroll
  for (int k = 0; k < 11; ++k) sh[lr][k][cl] = acc[k];
  __syncthreads();
  if (lr != 0 || c >= C) return;
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    double s = 0.0;
    for (int j = 0; j < 16; ++j) s += sh[j][k][cl];
    acc[k] = s;
  }
  const double mu = mean[c], rs = rstd[c], gm = gamma[c];
  const double G1 = acc[0], G2x = rs * (acc[1] - mu * acc[0]);       // sum g', sum g' xhat
  dbeta[c] += static_cast<float>(G1);
  dgamma[c] += static_cast<float>(G2x);
  const double k1 = G1 / count, k2 = G2x / count;
  const double A = gm * rs, Bc = -A * rs * k2, Cc = -A * k1 + A * rs * k2 * mu;
  double wv[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wv[t] = w[c * 9 + t];
  // symmetric R from the packed upper triangle
  for (int t = 0; t < 9; ++t) {
    double wr = 0.0;
    for (int u = 0; u < 9; ++u) {
      const int a = t < u ? t : u, b = t < u ? u : t;
      const int k = 9 + a * 9 - a * (a - 1) / 2 + (b - a);
      wr += wv[u] * moments[k];
    }
    dw[c * 9 + t] += static_cast<float>(A * acc[2 + t] + Bc * wr + Cc * moments[t]);
  }
}

}  // namespace htrvt

using namespace htrvt;

namespace htrvt {
int stem_head_tc_supported(int B, int H, int W, int C, int out_fmt);
int stem_head_tc_launch(const float* x, const float* w, const float* scale, const float* shift, void* out, void* out_bf,
                        void* code, int B, int H, int W, int C, int out_fmt, cudaStream_t stream);
}  // namespace htrvt

extern "C" int htrvt_stem_head_moment_ctas() { return 148 * 2; }
extern "C" int htrvt_stem_head_bwd_ctas() { return 148 * 3; }

static int head_shape_ok(int B, int H, int W, int C) {
  return B > 0 && H >= 4 && !(H & 1) && W > 0 && !(W & 3) && C >= 64 && !(C & 63) && C <= 256;
}

// moments [54] + conv-output channel statistics stats [2][C] (sum, sum of squares over B*(H/2)*W pixels)
// partial: fp32 [htrvt_stem_head_moment_ctas()][54] scratch
extern "C" int htrvt_stem_head_moments(const float* x, const float* w, float* partial, float* moments, float* stats,
                                       int B, int H, int W, int C, cudaStream_t stream) {
  if (!head_shape_ok(B, H, W, C)) return HTRVT_ERR_SHAPE;
  const int ctas = htrvt_stem_head_moment_ctas();
  conv1_moments_kernel<<<ctas, 256, 0, stream>>>(x, partial, B, H, W);
  HTRVT_LAUNCH_CHECK();
  conv1_moments_finalize_kernel<<<1, 256, 0, stream>>>(partial, ctas, w, C, moments, stats);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// out [B,Ho,W,C]: out_fmt 0 = bf16, 1 = fp16 (the engine's forward stem format), 2 = fp32 (fp32-parity mode);
// out_bf16 (nullable): a bf16 copy of out (train mode: operand of layer1.0's weight-gradient GEMMs);
// code (nullable: eval mode) uint8 [B,Ho,W,C/2]
extern "C" int htrvt_stem_head_fwd(const float* x, const float* w, const float* scale, const float* shift, void* out,
                                   void* out_bf16, void* code, int B, int H, int W, int C, int out_fmt,
                                   cudaStream_t stream) {
  if (!head_shape_ok(B, H, W, C)) return HTRVT_ERR_SHAPE;
  const int Ho = (H / 2 - 1) / 2 + 1;
  dim3 grid((W + 63) / 64, (Ho + 1) / 2, B);
  if (out_fmt < 0 || out_fmt > 2) return HTRVT_ERR_SHAPE;
  // the tensor-pipe kernel (stemhead_tc.cu) where the shape allows; the FP32-pipe kernel below for the rest (widths
  // that are not a multiple of 64, fp32 output of the fp32-parity mode)
  if (htrvt::stem_head_tc_supported(B, H, W, C, out_fmt))
    return htrvt::stem_head_tc_launch(x, w, scale, shift, out, out_bf16, code, B, H, W, C, out_fmt, stream);
  auto kern = code ? (out_fmt == 1 ? stem_head_fwd_kernel<true, 1> : out_fmt == 2 ? stem_head_fwd_kernel<true, 2>
                                                                                  : stem_head_fwd_kernel<true, 0>)
                   : (out_fmt == 1 ? stem_head_fwd_kernel<false, 1> : out_fmt == 2 ? stem_head_fwd_kernel<false, 2>
                                                                                   : stem_head_fwd_kernel<false, 0>);
  kern<<<grid, C, 0, stream>>>(x, w, scale, shift, static_cast<__nv_bfloat16*>(out),
                               static_cast<__nv_bfloat16*>(out_bf16), static_cast<uint8_t*>(code), B, H, W, C);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// g bf16 [B,Ho,W,C] gradient of the pooled output; accumulates (+=) dgamma, dbeta [C] and dw [C][9].
// partial: fp32 [htrvt_stem_head_bwd_ctas()][11][C] scratch
extern "C" int htrvt_stem_head_bwd(const void* g, const void* code, const float* x, const float* w,
                                   const float* moments, const float* gamma, const float* mean, const float* rstd,
                                   float* dgamma, float* dbeta, float* dw, float* partial, int B, int H, int W, int C,
                                   cudaStream_t stream) {
  if (!head_shape_ok(B, H, W, C)) return HTRVT_ERR_SHAPE;
  const int ctas = htrvt_stem_head_bwd_ctas();
  const int stage = 64 * C * 2 + 64 * (C / 2) + 7 * kHeadInW * 4;
  int smem = 2 * stage;
  if (smem < 11 * C * 4) smem = 11 * C * 4;
  if (smem > 220 * 1024) return HTRVT_ERR_SHAPE;
  if (!HTRVT_ENSURE_SMEM(stem_head_bwd_kernel, smem)) return HTRVT_ERR_LAUNCH;
  stem_head_bwd_kernel<<<ctas, C, smem, stream>>>(static_cast<const __nv_bfloat16*>(g),
                                                  static_cast<const uint8_t*>(code), x, w, partial, B, H, W, C);
  HTRVT_LAUNCH_CHECK();
  const double count = static_cast<double>(B) * (H / 2) * W;
  stem_head_bwd_finalize_kernel<<<(C + 31) / 32, 512, 0, stream>>>(partial, ctas, count, C, w, moments, gamma, mean,
                                                                    rstd, dgamma, dbeta, dw);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
