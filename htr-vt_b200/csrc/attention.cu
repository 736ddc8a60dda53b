// Multi-head self-attention for sm_100a, one CTA per (batch, head), T <= 128 tokens, head_dim 128:
//   fwd: S = Q K^T (tcgen05, fp32 in TMEM) -> row softmax in registers -> P (bf16, swizzled smem) ->
//        O = P V (tcgen05, V consumed MN-major straight from its [T, hd] layout) -> bf16 [B, T, H*hd].
//   bwd: recompute S, P; dP = dO V^T; dS = P o (dP - rowsum(dO o O)) * scale;
//        dV = P^T dO, dK = dS^T Q, dQ = dS K - five tcgen05 contractions off the same SWIZZLE_128B
//        smem images (a [rows][cols] image is a K-major operand over cols or an MN-major one over rows).
// The score matrix never touches HBM.  Replaces Attention.forward (model_v1/model/HTR_VT.py:27-39):
// `q @ k^T * scale -> softmax -> @ v -> transpose/reshape` and its autograd backward.
#include "common.cuh"
#include <cuda.h>

namespace htrvt {

constexpr int kAttnThreads = 160;          // warps 0-3: one thread per query row; warp 4: TMA + MMA issue
constexpr int kHd = 128;
constexpr int kTq = 128;
constexpr int kImg = kTq * 128;            // bytes of one [128 rows][64 bf16] swizzled chunk image

// image = two 64-column chunks, each [128 rows][128 B], SWIZZLE_128B
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t img, int k16) {       // MN = rows, K = cols
  return umma_desc_sw128(img + (k16 >> 2) * kImg + (k16 & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t img, int k16) {      // MN = cols, K = rows
  return umma_desc_sw128(img + k16 * 2048, kImg, 1024);
}
// write 8 consecutive bf16 (cols 8g..8g+7 of chunk c) of row r into an image
__device__ __forceinline__ void img_store8(uint8_t* img, int r, int c, int g, uint4 v) {
  *reinterpret_cast<uint4*>(img + c * kImg + r * 128 + ((g ^ (r & 7)) << 4)) = v;
}

__device__ __forceinline__ void load_image(uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int col0, int row0,
                                           int z) {
  tma_load_3d(dst, map, bar, col0, row0, z);
  tma_load_3d(dst + kImg, map, bar, col0 + 64, row0, z);
}

struct AttnP {
  int B, H, T;
  float scale;
  __nv_bfloat16* out;          // fwd: [B, T, H*hd]
  float* lse;                  // [B, H, T] natural-log row log-sum-exp of scale*S
  const __nv_bfloat16* o;      // bwd: forward output [B, T, H*hd]
  const __nv_bfloat16* dout;   // bwd: [B, T, H*hd]
  __nv_bfloat16* dqkv;         // bwd: [B, T, 3, H, hd]  (token-major: the dY of the qkv projection)
};

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ AttnP P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                  // 32 KB   (P aliases Q after S is complete)
  uint8_t* sK = smem + 2 * kImg;       // 32 KB
  uint8_t* sV = smem + 4 * kImg;       // 32 KB
  uint8_t* sP = sQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kImg);
  uint64_t* bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_o = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / P.H, h = bh - b * P.H;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 0, 1);

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_load, 6 * kImg);
      const int D = P.H * kHd;
      load_image(sQ, &tmQKV, bar_load, h * kHd, 0, b);
      load_image(sK, &tmQKV, bar_load, D + h * kHd, 0, b);
      load_image(sV, &tmQKV, bar_load, 2 * D + h * kHd, 0, b);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      const uint32_t q = smem_u32(sQ), k = smem_u32(sK);
#pragma unroll
      for (int i = 0; i < 8; ++i) umma_bf16(tmem, desc_kmajor(q, i), desc_kmajor(k, i), idesc_s, i ? 1u : 0u);
      umma_commit(bar_s);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t pp = smem_u32(sP), v = smem_u32(sV);
#pragma unroll
      for (int i = 0; i < 8; ++i) umma_bf16(tmem + 128, desc_kmajor(pp, i), desc_mnmajor(v, i), idesc_o, i ? 1u : 0u);
      umma_commit(bar_o);
    }
  } else {
    const int r = warp * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float s[128];
#pragma unroll
    for (int c = 0; c < 128; c += 16) {
      uint32_t raw[16];
      tmem_ld16(trow + c, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) s[c + i] = __uint_as_float(raw[i]);
    }
    const float sl2 = P.scale * kLog2e;
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 128; ++j) {
      if (j >= P.T) s[j] = -INFINITY;
      m = fmaxf(m, s[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 128; ++j) {
      s[j] = ex2f((s[j] - m) * sl2);
      sum += s[j];
    }
    const float inv = __fdividef(1.0f, sum);
    // all rows have consumed S (and therefore Q, K) only after every thread passed its tmem loads;
    // P overwrites Q's image, so wait for the whole CTA's row threads first
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      uint4 u;
      u.x = pack_bf16(s[8 * g], s[8 * g + 1]); u.y = pack_bf16(s[8 * g + 2], s[8 * g + 3]);
      u.z = pack_bf16(s[8 * g + 4], s[8 * g + 5]); u.w = pack_bf16(s[8 * g + 6], s[8 * g + 7]);
      img_store8(sP, r, g >> 3, g & 7, u);
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bar_p);
    if (r < P.T && P.lse) P.lse[static_cast<long long>(bh) * P.T + r] = m * P.scale + __logf(sum);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    __nv_bfloat16* orow = P.out + (static_cast<long long>(b) * P.T + r) * (P.H * kHd) + h * kHd;
#pragma unroll
    for (int c = 0; c < 128; c += 16) {
      uint32_t raw[16];
      tmem_ld16(trow + 128 + c, raw);
      tmem_ld_wait();
      if (r < P.T) {
        uint4 u0, u1;
        u0.x = pack_bf16(__uint_as_float(raw[0]) * inv, __uint_as_float(raw[1]) * inv);
        u0.y = pack_bf16(__uint_as_float(raw[2]) * inv, __uint_as_float(raw[3]) * inv);
        u0.z = pack_bf16(__uint_as_float(raw[4]) * inv, __uint_as_float(raw[5]) * inv);
        u0.w = pack_bf16(__uint_as_float(raw[6]) * inv, __uint_as_float(raw[7]) * inv);
        u1.x = pack_bf16(__uint_as_float(raw[8]) * inv, __uint_as_float(raw[9]) * inv);
        u1.y = pack_bf16(__uint_as_float(raw[10]) * inv, __uint_as_float(raw[11]) * inv);
        u1.z = pack_bf16(__uint_as_float(raw[12]) * inv, __uint_as_float(raw[13]) * inv);
        u1.w = pack_bf16(__uint_as_float(raw[14]) * inv, __uint_as_float(raw[15]) * inv);
        *reinterpret_cast<uint4*>(orow + c) = u0;
        *reinterpret_cast<uint4*>(orow + c + 8) = u1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ AttnP P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 2 * kImg;
  uint8_t* sV = smem + 4 * kImg;
  uint8_t* sDO = smem + 6 * kImg;
  uint8_t* sP = smem + 8 * kImg;
  uint8_t* sDS = smem + 10 * kImg;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * kImg);
  uint64_t* bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_g = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / P.H, h = bh - b * P.H;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_g, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S [0,128)  dP [128,256)  dQ [256,384)  dV [384,512)  dK reuses [0,128)
  constexpr uint32_t id_kk = umma_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t id_km = umma_idesc_bf16(128, 128, 0, 1);
  constexpr uint32_t id_mm = umma_idesc_bf16(128, 128, 1, 1);

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_load, 8 * kImg);
      const int Dm = P.H * kHd;
      load_image(sQ, &tmQKV, bar_load, h * kHd, 0, b);
      load_image(sK, &tmQKV, bar_load, Dm + h * kHd, 0, b);
      load_image(sV, &tmQKV, bar_load, 2 * Dm + h * kHd, 0, b);
      load_image(sDO, &tmDO, bar_load, h * kHd, 0, b);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      const uint32_t q = smem_u32(sQ), k = smem_u32(sK), v = smem_u32(sV), d_o = smem_u32(sDO);
      const uint32_t pp = smem_u32(sP), ds = smem_u32(sDS);
#pragma unroll
      for (int i = 0; i < 8; ++i) umma_bf16(tmem, desc_kmajor(q, i), desc_kmajor(k, i), id_kk, i ? 1u : 0u);        // S
#pragma unroll
      for (int i = 0; i < 8; ++i) umma_bf16(tmem + 128, desc_kmajor(d_o, i), desc_kmajor(v, i), id_kk, i ? 1u : 0u);  // dP
      umma_commit(bar_s);
      mbar_wait(bar_p, 0);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < 8; ++i) umma_bf16(tmem + 384, desc_mnmajor(pp, i), desc_mnmajor(d_o, i), id_mm, i ? 1u : 0u);  // dV = P^T dO
#pragma unroll
      for (int i = 0; i < 8; ++i) umma_bf16(tmem, desc_mnmajor(ds, i), desc_mnmajor(q, i), id_mm, i ? 1u : 0u);       // dK = dS^T Q
#pragma unroll
      for (int i = 0; i < 8; ++i) umma_bf16(tmem + 256, desc_kmajor(ds, i), desc_mnmajor(k, i), id_km, i ? 1u : 0u);  // dQ = dS K
      umma_commit(bar_g);
    }
  } else {
    const int r = warp * 32 + lane;
    const bool rok = r < P.T;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const int D = P.H * kHd;
    // delta = rowsum(dO o O) from global (256 B per row each), while the MMAs run
    float delta = 0.f;
    if (rok) {
      const uint4* po = reinterpret_cast<const uint4*>(P.o + (static_cast<long long>(b) * P.T + r) * D + h * kHd);
      const uint4* pd = reinterpret_cast<const uint4*>(P.dout + (static_cast<long long>(b) * P.T + r) * D + h * kHd);
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const uint4 a = __ldg(po + i), c = __ldg(pd + i);
        float2 x, y;
        x = unpack_bf16(a.x); y = unpack_bf16(c.x); delta += x.x * y.x + x.y * y.y;
        x = unpack_bf16(a.y); y = unpack_bf16(c.y); delta += x.x * y.x + x.y * y.y;
        x = unpack_bf16(a.z); y = unpack_bf16(c.z); delta += x.x * y.x + x.y * y.y;
        x = unpack_bf16(a.w); y = unpack_bf16(c.w); delta += x.x * y.x + x.y * y.y;
      }
    }
    const float lse = rok ? P.lse[static_cast<long long>(bh) * P.T + r] : 0.f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    const float sl2 = P.scale * kLog2e, lse2 = lse * kLog2e;
#pragma unroll 1
    for (int c = 0; c < 128; c += 16) {
      uint32_t rs[16], rp[16];
      tmem_ld16(trow + c, rs);
      tmem_ld16(trow + 128 + c, rp);
      tmem_ld_wait();
      float p[16], ds[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const bool ok = rok && (c + i) < P.T;
        p[i] = ok ? ex2f(__uint_as_float(rs[i]) * sl2 - lse2) : 0.f;
        ds[i] = ok ? p[i] * (__uint_as_float(rp[i]) - delta) * P.scale : 0.f;
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint4 u, w;
        u.x = pack_bf16(p[8 * g], p[8 * g + 1]); u.y = pack_bf16(p[8 * g + 2], p[8 * g + 3]);
        u.z = pack_bf16(p[8 * g + 4], p[8 * g + 5]); u.w = pack_bf16(p[8 * g + 6], p[8 * g + 7]);
        w.x = pack_bf16(ds[8 * g], ds[8 * g + 1]); w.y = pack_bf16(ds[8 * g + 2], ds[8 * g + 3]);
        w.z = pack_bf16(ds[8 * g + 4], ds[8 * g + 5]); w.w = pack_bf16(ds[8 * g + 6], ds[8 * g + 7]);
        const int gg = (c >> 3) + g;
        img_store8(sP, r, gg >> 3, gg & 7, u);
        img_store8(sDS, r, gg >> 3, gg & 7, w);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_g, 0);
    tc_fence_after();
    // rows of dQ are queries, rows of dK / dV are keys: all indexed by token r
    __nv_bfloat16* base = P.dqkv + (static_cast<long long>(b) * P.T + r) * (3 * D) + h * kHd;
    const uint32_t cols[3] = {256u, 0u, 384u};       // dQ, dK, dV
#pragma unroll 1
    for (int w = 0; w < 3; ++w) {
      __nv_bfloat16* dst = base + w * D;
#pragma unroll 1
      for (int c = 0; c < 128; c += 16) {
        uint32_t raw[16];
        tmem_ld16(trow + cols[w] + c, raw);
        tmem_ld_wait();
        if (rok) {
          uint4 u0, u1;
          u0.x = pack_bf16(__uint_as_float(raw[0]), __uint_as_float(raw[1]));
          u0.y = pack_bf16(__uint_as_float(raw[2]), __uint_as_float(raw[3]));
          u0.z = pack_bf16(__uint_as_float(raw[4]), __uint_as_float(raw[5]));
          u0.w = pack_bf16(__uint_as_float(raw[6]), __uint_as_float(raw[7]));
          u1.x = pack_bf16(__uint_as_float(raw[8]), __uint_as_float(raw[9]));
          u1.y = pack_bf16(__uint_as_float(raw[10]), __uint_as_float(raw[11]));
          u1.z = pack_bf16(__uint_as_float(raw[12]), __uint_as_float(raw[13]));
          u1.w = pack_bf16(__uint_as_float(raw[14]), __uint_as_float(raw[15]));
          *reinterpret_cast<uint4*>(dst + c) = u0;
          *reinterpret_cast<uint4*>(dst + c + 8) = u1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

}  // namespace htrvt

using namespace htrvt;

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn attn_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 3-D map (cols, rows, z), box {64, 128, 1}
int map3(CUtensorMap* m, const void* ptr, long long cols, long long rows, long long z, long long row_stride,
         long long z_stride) {
  EncodeTiledFn enc = attn_encode();
  if (!enc) return HTRVT_ERR_DRIVER;
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return HTRVT_ERR_ALIGN;
  cuuint64_t gd[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(z)};
  cuuint64_t gs[2] = {static_cast<cuuint64_t>(row_stride) * 2, static_cast<cuuint64_t>(z_stride) * 2};
  cuuint32_t bx[3] = {64, 128, 1}, es[3] = {1, 1, 1};
  if ((gs[0] & 15) || (gs[1] & 15)) return HTRVT_ERR_ALIGN;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HTRVT_OK : HTRVT_ERR_DRIVER;
}
}  // namespace

// qkv: bf16 token-major [B][T][3*H*128] (the QKV projection output as is); out: bf16 [B][T][H*128];
// lse: fp32 [B][H][T] (may be null)
extern "C" int htrvt_attention_fwd(const void* qkv, int B, int H, int T, int hd, float scale, void* out, float* lse,
                                   cudaStream_t stream) {
  if (hd != kHd || T < 1 || T > kTq || B < 1 || H < 1) return HTRVT_ERR_SHAPE;
  CUtensorMap tm;
  int r = map3(&tm, qkv, 3LL * H * kHd, T, B, 3LL * H * kHd, static_cast<long long>(T) * 3 * H * kHd);
  if (r) return r;
  AttnP P = {};
  P.B = B; P.H = H; P.T = T; P.scale = scale; P.out = static_cast<__nv_bfloat16*>(out); P.lse = lse;
  const int smem = 6 * kImg + 1024 + 128;
  if (!HTRVT_ENSURE_SMEM(attn_fwd_kernel, smem)) return HTRVT_ERR_LAUNCH;
  attn_fwd_kernel<<<B * H, kAttnThreads, smem, stream>>>(tm, P);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// dqkv: bf16 [B][T][3][H][128] (token-major)
extern "C" int htrvt_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int H,
                                   int T, int hd, float scale, void* dqkv, cudaStream_t stream) {
  if (hd != kHd || T < 1 || T > kTq || B < 1 || H < 1 || !lse) return HTRVT_ERR_SHAPE;
  CUtensorMap tm, tdo;
  int r = map3(&tm, qkv, 3LL * H * kHd, T, B, 3LL * H * kHd, static_cast<long long>(T) * 3 * H * kHd);
  if (r) return r;
  r = map3(&tdo, dout, static_cast<long long>(H) * kHd, T, B, static_cast<long long>(H) * kHd,
           static_cast<long long>(T) * H * kHd);
  if (r) return r;
  AttnP P = {};
  P.B = B; P.H = H; P.T = T; P.scale = scale; P.lse = const_cast<float*>(lse);
  P.o = static_cast<const __nv_bfloat16*>(out); P.dout = static_cast<const __nv_bfloat16*>(dout);
  P.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  const int smem = 12 * kImg + 1024 + 128;
  if (!HTRVT_ENSURE_SMEM(attn_bwd_kernel, smem)) return HTRVT_ERR_LAUNCH;
  attn_bwd_kernel<<<B * H, kAttnThreads, smem, stream>>>(tm, tdo, P);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
