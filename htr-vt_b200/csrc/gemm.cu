// The sm_100a tap GEMM: persistent, warp-specialised tcgen05 kernel.
//   warp 0   : TMA producer (one elected lane) - 5-D tiled tensor maps, SWIZZLE_128B, OOB zero fill
//              implements conv padding, ragged tiles and K tails
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16, bf16 -> fp32 in TMEM)
//   warps 2-9: epilogue (tcgen05.ld 32x32b, two warps per TMEM lane quadrant), double-buffered
//              accumulator so the epilogue of tile i overlaps the main loop of tile i+1
// Operands can be K-major or MN-major (smem descriptor + instruction-descriptor major bits), which is
// what lets dgrad (W as [K,N]) and wgrad (dY^T X, both [pixels, channels]) run without transposes.
//
// Replaces the cuBLAS / cuDNN dispatches behind nn.Linear (model_v1/model/HTR_VT.py:22-24,29,37,170,
// timm Mlp fc1/fc2) and nn.Conv2d 3x3 / 1x1 (model_v1/model/resnet18.py:4-7,56-63) and their backward.
#include "common.cuh"
#include "gemm.cuh"

namespace htrvt {

constexpr int kGemmThreads = 320;
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kStatCols = 768;             // widest output a statistics epilogue supports (Cout of the stem)

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 128) ? 6 : 4;
  static_assert(true, "");
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStatsBytes = 8 * 2 * 16 * 4;                  // per-warp column partials
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 /*barriers*/ + 2 * BN * 4 * 4 /* stats: [4 quadrants][2][BN] */ +
                                    2 * kStatCols * 4 /* per-CTA column sums */ + 8 * 32 * 20 * 4 /* epilogue transpose staging */;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ float gelu_erf(float x) {
  // 0.5 x (1 + erf(x / sqrt2)); erf by Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below bf16 eps)
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = 1.0f - poly * ex2f(-z * z * kLog2e);
  const float erfv = copysignf(e, x);
  return 0.5f * x * (1.0f + erfv);
}

struct TileCoord {
  int m_tile, n_tile, tap, split;
  int n, h, w0;          // kind 0: output row (n,h) and first column
  int q_begin, q_end;    // kind 1: pixel-chunk range
};

__device__ __forceinline__ TileCoord decode_tile(const GemmP& P, int id) {
  TileCoord c;
  c.m_tile = id % P.tiles_m;
  int r = id / P.tiles_m;
  c.n_tile = r % P.tiles_n;
  r /= P.tiles_n;
  c.tap = 0; c.split = 0; c.n = 0; c.h = 0; c.w0 = 0; c.q_begin = 0; c.q_end = 0;
  if (P.kind == 0) {
    const int row = c.m_tile / P.tiles_per_row;
    c.w0 = (c.m_tile - row * P.tiles_per_row) * kBM;
    c.n = row / P.Ho;
    c.h = row - c.n * P.Ho;
  } else {
    c.tap = r % P.n_taps;
    c.split = r / P.n_taps;
    const long long Q = static_cast<long long>(P.NB) * P.Ho * P.k_chunks;
    c.q_begin = static_cast<int>(Q * c.split / P.splits);
    c.q_end = static_cast<int>(Q * (c.split + 1) / P.splits);
  }
  return c;
}

template <int BN, int KIND, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmP P) {
  using Cfg = GemmCfg<BN>;
  constexpr bool A_MN = (KIND == 1);
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];      // SWIZZLE_128B tiles need 1024-byte alignment
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  float* stat_smem = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256);   // [4][2][BN]
  float* stat_acc = stat_smem + 8 * BN;                                                    // [2][kStatCols]
  float* stage_f32 = stat_acc + 2 * kStatCols;                                              // [8 warps][32][20]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = P.tiles_m * P.tiles_n * (P.kind == 0 ? 1 : P.n_taps * P.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int id = blockIdx.x; id < total_tiles; id += gridDim.x) {
        const TileCoord tc = decode_tile(P, id);
        const int n0 = tc.n_tile * BN, m0 = tc.m_tile * kBM;
        const int kiters = (KIND == 0) ? P.n_taps * P.k_chunks : (tc.q_end - tc.q_begin);
        for (int k = 0; k < kiters; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
          if (KIND == 0) {
            const int tap = k / P.k_chunks, cc = k - tap * P.k_chunks;
            tma_load_5d(sa, &tmA, &full[stage], cc * kBK, P.tap.pw[tap], tc.w0 + P.tap.dw[tap],
                        tc.h * P.a_sh + P.tap.dh[tap], tc.n);
            if (!B_MN) {
              tma_load_5d(sb, &tmB, &full[stage], P.tap.widx[tap] * P.b_tap_stride + cc * kBK, n0, 0, 0, 0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_5d(sb + i * 8192, &tmB, &full[stage], P.tap.widx[tap] * P.b_tap_stride + n0 + 64 * i,
                            cc * kBK, 0, 0, 0);
            }
          } else {
            const int q = tc.q_begin + k;
            const int row = q / P.k_chunks, w0c = (q - row * P.k_chunks) * kBK;
            const int n = row / P.Ho, ho = row - n * P.Ho;
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i)
              tma_load_5d(sa + i * 8192, &tmA, &full[stage], m0 + 64 * i, 0, w0c, ho, n);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_5d(sb + i * 8192, &tmB, &full[stage], n0 + 64 * i, P.tap.pw[tc.tap],
                          w0c + P.tap.dw[tc.tap], ho * P.a_sh + P.tap.dh[tc.tap], n);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int id = blockIdx.x; id < total_tiles; id += gridDim.x) {
        const TileCoord tc = decode_tile(P, id);
        const int kiters = (KIND == 0) ? P.n_taps * P.k_chunks : (tc.q_end - tc.q_begin);
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int k = 0; k < kiters; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            const uint64_t da = A_MN ? umma_desc_sw128(sa + kk * 2048, 8192, 1024)
                                     : umma_desc_sw128(sa + kk * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc_sw128(sb + kk * 2048, 8192, 1024)
                                     : umma_desc_sw128(sb + kk * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);
        if (kiters == 0) { /* unreachable: host never creates empty tiles */ }
        as ^= 1; if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    // Per 16-column chunk: tcgen05.ld (thread = row) -> per-warp smem transpose -> phase 2 where a lane owns
    // (row = lane/4 + 8i, 4 consecutive columns), so every global access of the warp covers whole 32-byte
    // sectors of 8 rows (the thread-per-row pattern touched 32 separate cache lines per instruction and paced
    // the whole kernel).  The next chunk's TMEM load is in flight while the current one is processed.
    const int ew = warp - 2;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                // column half handled by this warp
    constexpr int kColsPerWarp = BN / 2;
    constexpr int kStgStride = 20;           // floats per staged row (16 + pad: conflict-free 128-bit stores)
    float* stg = stage_f32 + ew * (32 * kStgStride);
    const int r8 = lane >> 2, cs = lane & 3;
    int as = 0; uint32_t aphase = 0;
    if (P.flags & EPI_STATS) {
      for (int c = threadIdx.x - 64; c < 2 * kStatCols; c += 256) stat_acc[c] = 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    for (int id = blockIdx.x; id < total_tiles; id += gridDim.x) {
      const TileCoord tc = decode_tile(P, id);
      const int n0 = tc.n_tile * BN;
      // the four rows this lane stores in phase 2
      bool rok[4];
      long long roff[4];
      int qb[4], qt[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = quad * 32 + r8 + 8 * i;
        qb[i] = 0; qt[i] = 0;
        if (KIND == 0) {
          const int w = tc.w0 + r;
          rok[i] = w < P.Wo;
          roff[i] = tc.n * P.o_sn + tc.h * P.o_sh + w * P.o_sw + P.o_base;
          if (P.flags & EPI_QKV) {
            const int m = tc.m_tile * kBM + r;
            qb[i] = m / P.qkv_T; qt[i] = m - qb[i] * P.qkv_T;
          }
        } else {
          const int co = tc.m_tile * kBM + r;
          rok[i] = co < P.M_valid;
          roff[i] = tc.split * P.o_split + co * P.o_sw + tc.tap * P.o_tap + P.o_base;
        }
      }
      if ((P.flags & (EPI_RESID | EPI_ACCUM)) && !(P.flags & EPI_QKV) && id + static_cast<int>(gridDim.x) < total_tiles) {
        // warm L2 with the rows the NEXT tile's epilogue will read (each lane: one 128-byte line per row and step)
        const TileCoord nt = decode_tile(P, id + gridDim.x);
        const int esz = (P.flags & EPI_RESID) ? 4 : ((P.flags & EPI_BF16) ? 2 : 4);
        const char* basep = (P.flags & EPI_RESID) ? reinterpret_cast<const char*>(P.resid)
                                                   : reinterpret_cast<const char*>(P.out);
        const int line_elems = 128 / esz;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = quad * 32 + r8 + 8 * i;
          long long ro; bool ok;
          if (KIND == 0) { const int w = nt.w0 + r; ok = w < P.Wo; ro = nt.n * P.o_sn + nt.h * P.o_sh + w * P.o_sw + P.o_base; }
          else { const int co = nt.m_tile * kBM + r; ok = co < P.M_valid; ro = nt.split * P.o_split + co * P.o_sw + nt.tap * P.o_tap + P.o_base; }
          for (int c = cs * line_elems; c < kColsPerWarp; c += 4 * line_elems) {
            const int cg = nt.n_tile * BN + half * kColsPerWarp + c;
            if (ok && cg < P.N_valid)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(basep + (ro + cg) * esz));
          }
        }
      }
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * kColsPerWarp;
      if (P.flags & EPI_NOSTORE) {                         // measurement aid: main-loop-only timing
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        as ^= 1; if (as == 0) aphase ^= 1;
        continue;
      }
      uint32_t cur[16], nxt[16];
      tmem_ld16(taddr, cur);
      tmem_ld_wait();
#pragma unroll 1
      for (int c0 = 0; c0 < kColsPerWarp; c0 += 16) {
        const bool more = c0 + 16 < kColsPerWarp;
        if (more) tmem_ld16(taddr + c0 + 16, nxt);
        const int col = n0 + half * kColsPerWarp + c0;       // first global column of this chunk
        const int c4 = col + cs * 4;                          // first of this lane's 4 global columns (phase 2)
        const bool cok = c4 < P.N_valid;
        long long coff = c4;                                  // column part of the output offset
        long long qkv_which = 0; int qkv_hh = 0;
        if (P.flags & EPI_QKV) {
          const int D = P.qkv_H * P.qkv_hd;
          const int which = c4 / D, rem = c4 - which * D;
          qkv_hh = rem / P.qkv_hd;
          coff = rem - qkv_hh * P.qkv_hd;
          qkv_which = which;
        }
        long long offs[4];
        float4 pre[4];                                        // residual / accumulate operands, loaded early
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          offs[i] = roff[i] + coff;
          if (P.flags & EPI_QKV)
            offs[i] = (((qkv_which * P.NB + qb[i]) * P.qkv_H + qkv_hh) * P.qkv_T + qt[i]) * P.qkv_hd + coff;
          pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool ok = rok[i] && cok;
          if (ok && (P.flags & EPI_RESID)) pre[i] = *reinterpret_cast<const float4*>(P.resid + offs[i]);
          if (ok && (P.flags & EPI_ACCUM)) {
            if (P.flags & EPI_BF16) {
              const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(P.out) + offs[i]);
              const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y);
              pre[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
            } else {
              pre[i] = *reinterpret_cast<const float4*>(static_cast<const float*>(P.out) + offs[i]);
            }
          }
        }
        // ---- phase 1: thread = row; scale, stage to smem -------------------------------------------
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 f = make_float4(__uint_as_float(cur[i]) * P.alpha, __uint_as_float(cur[i + 1]) * P.alpha,
                                 __uint_as_float(cur[i + 2]) * P.alpha, __uint_as_float(cur[i + 3]) * P.alpha);
          *reinterpret_cast<float4*>(stg + lane * kStgStride + i) = f;
        }
        __syncwarp();
        // ---- phase 2: lane = (row group r8, column segment cs) --------------------------------------
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((P.flags & EPI_BIAS) && cok) bias4 = *reinterpret_cast<const float4*>(P.bias + c4);
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(stg + (r8 + 8 * i) * kStgStride + cs * 4);
          float v[4] = {x.x + bias4.x, x.y + bias4.y, x.z + bias4.z, x.w + bias4.w};
          const bool ok = rok[i] && cok;
          const long long off = offs[i];
          if (P.flags & EPI_RESID) { v[0] += pre[i].x; v[1] += pre[i].y; v[2] += pre[i].z; v[3] += pre[i].w; }
          if (P.flags & EPI_GELU) {
            if (ok) {
              uint2 u;
              u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
              *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(P.out2) + off) = u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = gelu_erf(v[k]);
          }
          if (P.flags & EPI_RELU) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = fmaxf(v[k], 0.f);
          }
          if (P.flags & EPI_ACCUM) { v[0] += pre[i].x; v[1] += pre[i].y; v[2] += pre[i].z; v[3] += pre[i].w; }
          if (P.flags & EPI_BF16) {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(P.out) + off;
            uint2 u;
            u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
            if (ok) *reinterpret_cast<uint2*>(o) = u;
            if (P.flags & EPI_STATS) {
              // statistics of exactly what is stored (bf16-rounded), zero for rows outside the image
              const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y);
              if (rok[i]) {
                s4[0] += f0.x; s4[1] += f0.y; s4[2] += f1.x; s4[3] += f1.y;
                q4[0] += f0.x * f0.x; q4[1] += f0.y * f0.y; q4[2] += f1.x * f1.x; q4[3] += f1.y * f1.y;
              }
            }
          } else {
            float* o = static_cast<float*>(P.out) + off;
            if (ok) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
        if (P.flags & EPI_STATS) {
          // lanes with equal cs hold the same 4 columns for different rows: 3 shuffle levels
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              s4[k] += __shfl_xor_sync(0xffffffffu, s4[k], o);
              q4[k] += __shfl_xor_sync(0xffffffffu, q4[k], o);
            }
          }
          if (lane < 4) {
            const int cc = half * kColsPerWarp + c0 + lane * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              stat_smem[(quad * 2 + 0) * BN + cc + k] = s4[k];
              stat_smem[(quad * 2 + 1) * BN + cc + k] = q4[k];
            }
          }
        }
        if (more) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (P.flags & EPI_STATS) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = threadIdx.x - 64;                       // 0..255
        for (int c = t; c < BN; c += 256) {
          if (n0 + c < P.N_valid) {
            float ssum = 0.f, qsum = 0.f;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) { ssum += stat_smem[(qd * 2) * BN + c]; qsum += stat_smem[(qd * 2 + 1) * BN + c]; }
            stat_acc[n0 + c] += ssum;                      // column n0+c is owned by exactly one thread per tile
            stat_acc[kStatCols + n0 + c] += qsum;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      as ^= 1; if (as == 0) aphase ^= 1;
    }
    if (P.flags & EPI_STATS) {                           // one partial row per CTA: stats[cta][2][N]
      float* dst = P.stats + static_cast<long long>(blockIdx.x) * 2 * P.N_valid;
      for (int c = threadIdx.x - 64; c < P.N_valid; c += 256) {
        dst[c] = stat_acc[c];
        dst[P.N_valid + c] = stat_acc[kStatCols + c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace htrvt

// =================================================================================================
// Host side: tensor maps + C-ABI entry points
// =================================================================================================
#include <cudaTypedefs.h>
#include <mutex>

using namespace htrvt;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 5-D bf16 tensor map, innermost dimension contiguous, SWIZZLE_128B, zero OOB fill.
// dims / box: innermost first.  strides: element strides of dims 1..4.
int make_map5(CUtensorMap* m, const void* ptr, const long long dims[5], const long long strides_elems[4],
              const int box[5]) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return HTRVT_ERR_DRIVER;
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return HTRVT_ERR_ALIGN;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < 5; ++i) {
    gd[i] = static_cast<cuuint64_t>(dims[i] < 1 ? 1 : dims[i]);
    bx[i] = static_cast<cuuint32_t>(box[i]);
    es[i] = 1;
  }
  for (int i = 0; i < 4; ++i) {
    gs[i] = static_cast<cuuint64_t>(strides_elems[i]) * 2;
    if (gs[i] & 15) return HTRVT_ERR_ALIGN;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HTRVT_OK : HTRVT_ERR_DRIVER;
}

// 2-D matrix [rows, cols] (cols contiguous, row stride ld) as a 5-D map (cols, rows, 1, 1, 1)
int make_map_matrix(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                    int box_rows) {
  const long long dims[5] = {cols, rows, 1, 1, 1};
  const long long st[4] = {ld, ld * rows, ld * rows, ld * rows};
  const int box[5] = {box_cols, box_rows, 1, 1, 1};
  return make_map5(m, ptr, dims, st, box);
}

// NHWC activation [N, H, W, C] viewed as (C, sw, W/sw, H, N): horizontal stride folded into a parity dim
int make_map_act(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int sw, int box_c, int box_w) {
  const long long dims[5] = {C, sw, W / sw, H, N};
  const long long st[4] = {C, static_cast<long long>(C) * sw, static_cast<long long>(C) * W,
                           static_cast<long long>(C) * W * H};
  const int box[5] = {box_c, 1, box_w, 1, 1};
  return make_map5(m, ptr, dims, st, box);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int KIND, bool B_MN>
int launch_one(const CUtensorMap& a, const CUtensorMap& b, const GemmP& P, int total_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  auto kern = tapgemm_kernel<BN, KIND, B_MN>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return HTRVT_ERR_LAUNCH;
    configured = true;
  }
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(a, b, P);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

template <int KIND, bool B_MN>
int launch_bn(int bn, const CUtensorMap& a, const CUtensorMap& b, const GemmP& P, int total, cudaStream_t s) {
  switch (bn) {
    case 128: return launch_one<128, KIND, B_MN>(a, b, P, total, s);
    case 192: return launch_one<192, KIND, B_MN>(a, b, P, total, s);
    case 256: return launch_one<256, KIND, B_MN>(a, b, P, total, s);
  }
  return HTRVT_ERR_SHAPE;
}

int pick_bn(int N) {
  if (N <= 128) return 128;
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0 && N < 512) return 128;
  return 256;
}

void fill_taps_conv(TapTab& t, int ks, int pad, int sw, int* n_taps) {
  int n = 0;
  for (int kh = 0; kh < ks; ++kh)
    for (int kw = 0; kw < ks; ++kw) {
      const int e = kw - pad;
      const int pw = ((e % sw) + sw) % sw;
      t.dh[n] = static_cast<int8_t>(kh - pad);
      t.pw[n] = static_cast<int8_t>(pw);
      t.dw[n] = static_cast<int8_t>((e - pw) / sw);
      t.widx[n] = static_cast<int8_t>(n);
      ++n;
    }
  *n_taps = n;
}

}  // namespace

// Y[M,N] = epilogue(alpha * X[M,K] W[N,K]^T): nn.Linear forward (both operands K-major).
extern "C" int htrvt_gemm_tn(const void* X, long long ldx, const void* W, long long ldw, int M, int N, int K,
                             int flags, const float* bias, const float* resid, void* out, long long ldo,
                             void* out2, float alpha, int qkv_B, int qkv_T, int qkv_H, int qkv_hd,
                             cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N & 7) || (ldo & 3)) return HTRVT_ERR_SHAPE;
  const int bn = pick_bn(N);
  CUtensorMap ta, tb;
  {
    const long long dims[5] = {K, 1, M, 1, 1};
    const long long st[4] = {ldx, ldx, ldx * M, ldx * M};
    const int box[5] = {kBK, 1, kBM, 1, 1};
    int r = make_map5(&ta, X, dims, st, box);
    if (r) return r;
    r = make_map_matrix(&tb, W, N, K, ldw, kBK, bn);
    if (r) return r;
  }
  GemmP P = {};
  P.kind = 0; P.Wo = M; P.Ho = 1; P.NB = (flags & EPI_QKV) ? qkv_B : 1;
  P.tiles_per_row = (M + kBM - 1) / kBM; P.tiles_m = P.tiles_per_row; P.tiles_n = (N + bn - 1) / bn;
  P.n_taps = 1; P.splits = 1; P.k_chunks = (K + kBK - 1) / kBK; P.a_sh = 1; P.b_tap_stride = 0;
  P.M_valid = M; P.N_valid = N; P.flags = flags;
  P.o_sn = 0; P.o_sh = 0; P.o_sw = ldo; P.o_base = 0;
  P.out = out; P.out2 = out2; P.bias = bias; P.resid = resid; P.alpha = alpha;
  P.qkv_T = qkv_T; P.qkv_H = qkv_H; P.qkv_hd = qkv_hd;
  return launch_bn<0, false>(bn, ta, tb, P, P.tiles_m * P.tiles_n, stream);
}

// dX[M,N] = dY[M,K] W[K,N]: nn.Linear input gradient (B operand MN-major, no transpose copy).
extern "C" int htrvt_gemm_nn(const void* dY, long long lddy, const void* W, long long ldw, int M, int N, int K,
                             int flags, void* out, long long ldo, float alpha, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N & 7) || (ldo & 3)) return HTRVT_ERR_SHAPE;
  const int bn = pick_bn(N);
  CUtensorMap ta, tb;
  {
    const long long dims[5] = {K, 1, M, 1, 1};
    const long long st[4] = {lddy, lddy, lddy * M, lddy * M};
    const int box[5] = {kBK, 1, kBM, 1, 1};
    int r = make_map5(&ta, dY, dims, st, box);
    if (r) return r;
    r = make_map_matrix(&tb, W, K, N, ldw, 64, 64);
    if (r) return r;
  }
  GemmP P = {};
  P.kind = 0; P.Wo = M; P.Ho = 1; P.NB = 1;
  P.tiles_per_row = (M + kBM - 1) / kBM; P.tiles_m = P.tiles_per_row; P.tiles_n = (N + bn - 1) / bn;
  P.n_taps = 1; P.splits = 1; P.k_chunks = (K + kBK - 1) / kBK; P.a_sh = 1; P.b_tap_stride = 0;
  P.M_valid = M; P.N_valid = N; P.flags = flags;
  P.o_sw = ldo; P.out = out; P.alpha = alpha;
  return launch_bn<0, true>(bn, ta, tb, P, P.tiles_m * P.tiles_n, stream);
}

namespace {
// Split-K factor for the weight-gradient GEMMs: minimise (waves x per-split tile time) + partial-sum traffic.
// base_tiles output tiles, q_total K chunks per tile, out_elems fp32 outputs.
int choose_splits(int base_tiles, long long q_total, long long out_elems = 0) {
  const int sms = num_sms();
  double best = 1e300;
  int best_s = 1;
  const double tile_us = static_cast<double>(q_total) * 0.23;          // ~0.23 us per 128xBNx64 k-chunk at full rate
  for (int s = 1; s <= 32 && s <= q_total; ++s) {
    const long long tiles = static_cast<long long>(base_tiles) * s;
    const double waves = static_cast<double>((tiles + sms - 1) / sms);
    double t = waves * tile_us / s;
    if (s > 1) t += 2.0 * s * out_elems * 4.0 / 5.0e6;                 // write + read of the partials at ~5 TB/s (us)
    if (t < best * 0.98) { best = t; best_s = s; }
  }
  return best_s;
}
}  // namespace

extern "C" size_t htrvt_wgrad_workspace_bytes(int Cout, int Cin, int n_taps, int M_pixels) {
  const int bn = pick_bn(Cin);
  const int base = n_taps * ((Cout + kBM - 1) / kBM) * ((Cin + bn - 1) / bn);
  (void)base; (void)M_pixels;
  return static_cast<size_t>(32) * Cout * n_taps * Cin * sizeof(float);   // upper bound (splits <= 32)
}

// Split-K partial sums [splits][Cout][taps][Cin] -> grad (+= or =); to_oihw permutes to [Cout][Cin][taps].
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int Cout, int taps, int Cin,
                                    float* __restrict__ grad, int accumulate, int to_oihw) {
  const long long total = static_cast<long long>(Cout) * taps * Cin;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[k * total + i];
    long long o = i;
    if (to_oihw) {
      const int ci = static_cast<int>(i % Cin);
      const long long r = i / Cin;
      const int tap = static_cast<int>(r % taps);
      const int co = static_cast<int>(r / taps);
      o = (static_cast<long long>(co) * Cin + ci) * taps + tap;
    }
    grad[o] = accumulate ? grad[o] + s : s;
  }
}

// dW[Nout, Kin] (+)= dY[M, Nout]^T X[M, Kin]  (both operands MN-major; split-K over M, fp32 partials)
extern "C" int htrvt_linear_wgrad(const void* dY, long long lddy, const void* X, long long ldx, int M, int Nout,
                                  int Kin, float* grad, int accumulate, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream) {
  if (M <= 0 || Nout <= 0 || Kin <= 0 || (Kin & 7) || (Nout & 7)) return HTRVT_ERR_SHAPE;
  const int bn = pick_bn(Kin);
  CUtensorMap ta, tb;
  int r;
  {
    const long long dimsA[5] = {Nout, 1, M, 1, 1};
    const long long stA[4] = {lddy, lddy, lddy * M, lddy * M};
    const int box[5] = {64, 1, 64, 1, 1};
    r = make_map5(&ta, dY, dimsA, stA, box);
    if (r) return r;
    const long long dimsB[5] = {Kin, 1, M, 1, 1};
    const long long stB[4] = {ldx, ldx, ldx * M, ldx * M};
    r = make_map5(&tb, X, dimsB, stB, box);
    if (r) return r;
  }
  GemmP P = {};
  P.kind = 1; P.Wo = M; P.Ho = 1; P.NB = 1; P.tiles_per_row = 1;
  P.tiles_m = (Nout + kBM - 1) / kBM; P.tiles_n = (Kin + bn - 1) / bn; P.n_taps = 1;
  P.k_chunks = (M + 63) / 64; P.a_sh = 1;
  P.splits = choose_splits(P.tiles_m * P.tiles_n, P.k_chunks, static_cast<long long>(Nout) * Kin);
  P.M_valid = Nout; P.N_valid = Kin; P.flags = 0; P.alpha = 1.f;
  const size_t need = static_cast<size_t>(P.splits) * Nout * Kin * sizeof(float);
  if (!workspace || workspace_bytes < need) return HTRVT_ERR_WORKSPACE;
  P.o_split = static_cast<long long>(Nout) * Kin; P.o_sw = Kin; P.o_tap = 0; P.o_base = 0;
  P.out = workspace;
  r = launch_bn<1, true>(bn, ta, tb, P, P.tiles_m * P.tiles_n * P.splits, stream);
  if (r) return r;
  const long long total = static_cast<long long>(Nout) * Kin;
  const int blocks = static_cast<int>((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), P.splits, Nout, 1, Kin, grad,
                                                  accumulate, 0);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// ---------------------------------------------------------------------------------------------
// Stem convolutions (NHWC bf16 activations, weights [Cout][ks*ks][Cin] bf16, bias-free), Cin % 64 == 0
// ---------------------------------------------------------------------------------------------
extern "C" int htrvt_conv_fwd(const void* x, int NB, int H, int W, int Cin, const void* w, int Cout, int ks,
                              int sh, int sw, void* y, float* stats_partial, int flags, cudaStream_t stream) {
  const int pad = ks / 2;
  if ((Cin % 64) || (Cout & 7) || (W % sw) || (ks != 1 && ks != 3)) return HTRVT_ERR_SHAPE;
  if (stats_partial && Cout > kStatCols) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cout);
  CUtensorMap ta, tb;
  int r = make_map_act(&ta, x, NB, H, W, Cin, sw, kBK, kBM);
  if (r) return r;
  r = make_map_matrix(&tb, w, Cout, static_cast<long long>(ks) * ks * Cin, static_cast<long long>(ks) * ks * Cin, kBK, bn);
  if (r) return r;
  GemmP P = {};
  P.kind = 0; P.Wo = Wo; P.Ho = Ho; P.NB = NB;
  P.tiles_per_row = (Wo + kBM - 1) / kBM; P.tiles_m = P.tiles_per_row * Ho * NB; P.tiles_n = (Cout + bn - 1) / bn;
  fill_taps_conv(P.tap, ks, pad, sw, &P.n_taps);
  P.splits = 1; P.k_chunks = Cin / kBK; P.a_sh = sh; P.b_tap_stride = Cin;
  P.M_valid = 0; P.N_valid = Cout; P.flags = EPI_BF16 | flags | (stats_partial ? EPI_STATS : 0);
  P.o_sn = static_cast<long long>(Ho) * Wo * Cout; P.o_sh = static_cast<long long>(Wo) * Cout; P.o_sw = Cout;
  P.out = y; P.stats = stats_partial; P.alpha = 1.f;
  return launch_bn<0, false>(bn, ta, tb, P, P.tiles_m * P.tiles_n, stream);
}

extern "C" int htrvt_conv_fwd_stats_rows(int NB, int H, int W, int ks, int sh, int sw) {
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  (void)NB; (void)Ho; (void)Wo;
  return num_sms();   // rows of the [ctas][2][Cout] partial-statistics buffer (upper bound)
}

// dx[NB,H,W,Cin] (= or +=) conv_transpose(dy[NB,Ho,Wo,Cout], w): one GEMM per output parity class.
extern "C" int htrvt_conv_dgrad(const void* dy, int NB, int H, int W, int Cin, const void* w, int Cout, int ks,
                                int sh, int sw, void* dx, int accumulate, cudaStream_t stream) {
  const int pad = ks / 2;
  if ((Cout % 8) || (Cin % 64) || (W % sw) || (H % sh) || (ks != 1 && ks != 3)) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cin);
  CUtensorMap ta, tb;
  int r = make_map_act(&ta, dy, NB, Ho, Wo, Cout, 1, kBK, kBM);
  if (r) return r;
  r = make_map_matrix(&tb, w, Cout, static_cast<long long>(ks) * ks * Cin, static_cast<long long>(ks) * ks * Cin, 64, 64);
  if (r) return r;
  for (int ph = 0; ph < sh; ++ph)
    for (int pw = 0; pw < sw; ++pw) {
      GemmP P = {};
      int n = 0;
      for (int kh = 0; kh < ks; ++kh)
        for (int kw = 0; kw < ks; ++kw) {
          const int eh = ph + pad - kh, ew = pw + pad - kw;
          if (((eh % sh) + sh) % sh != 0 || ((ew % sw) + sw) % sw != 0) continue;
          P.tap.dh[n] = static_cast<int8_t>(eh / sh);       // exact division (eh multiple of sh)
          P.tap.dw[n] = static_cast<int8_t>(ew / sw);
          P.tap.pw[n] = 0;
          P.tap.widx[n] = static_cast<int8_t>(kh * ks + kw);
          ++n;
        }
      const int Hq = (H - ph + sh - 1) / sh, Wq = (W - pw + sw - 1) / sw;
      if (n == 0) continue;                               // class receives no gradient (1x1 strided): caller zero-fills
      P.kind = 0; P.Wo = Wq; P.Ho = Hq; P.NB = NB;
      P.tiles_per_row = (Wq + kBM - 1) / kBM; P.tiles_m = P.tiles_per_row * Hq * NB; P.tiles_n = (Cin + bn - 1) / bn;
      P.n_taps = n; P.splits = 1; P.k_chunks = (Cout + kBK - 1) / kBK; P.a_sh = 1; P.b_tap_stride = Cin;
      P.N_valid = Cin; P.flags = EPI_BF16 | (accumulate ? EPI_ACCUM : 0);
      P.o_sn = static_cast<long long>(H) * W * Cin; P.o_sh = static_cast<long long>(sh) * W * Cin;
      P.o_sw = static_cast<long long>(sw) * Cin; P.o_base = (static_cast<long long>(ph) * W + pw) * Cin;
      P.out = dx; P.alpha = 1.f;
      r = launch_bn<0, true>(bn, ta, tb, P, P.tiles_m * P.tiles_n, stream);
      if (r) return r;
    }
  return HTRVT_OK;
}

// dw (+)= sum_pixels dy^T x_shifted; grad is fp32 OIHW (the nn.Conv2d parameter layout).
extern "C" int htrvt_conv_wgrad(const void* dy, const void* x, int NB, int H, int W, int Cin, int Cout, int ks,
                                int sh, int sw, float* grad_oihw, int accumulate, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream) {
  const int pad = ks / 2;
  if ((Cout % 8) || (Cin % 8) || (W % sw) || (ks != 1 && ks != 3)) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cin);
  CUtensorMap ta, tb;
  int r = make_map_act(&ta, dy, NB, Ho, Wo, Cout, 1, 64, 64);
  if (r) return r;
  r = make_map_act(&tb, x, NB, H, W, Cin, sw, 64, 64);
  if (r) return r;
  GemmP P = {};
  P.kind = 1; P.Wo = Wo; P.Ho = Ho; P.NB = NB; P.tiles_per_row = 1;
  P.tiles_m = (Cout + kBM - 1) / kBM; P.tiles_n = (Cin + bn - 1) / bn;
  fill_taps_conv(P.tap, ks, pad, sw, &P.n_taps);
  P.k_chunks = (Wo + 63) / 64; P.a_sh = sh;
  const long long Q = static_cast<long long>(NB) * Ho * P.k_chunks;
  P.splits = choose_splits(P.tiles_m * P.tiles_n * P.n_taps, Q, static_cast<long long>(Cout) * P.n_taps * Cin);
  P.M_valid = Cout; P.N_valid = Cin; P.flags = 0; P.alpha = 1.f;
  const long long per = static_cast<long long>(Cout) * P.n_taps * Cin;
  if (!workspace || workspace_bytes < static_cast<size_t>(P.splits) * per * sizeof(float)) return HTRVT_ERR_WORKSPACE;
  P.o_split = per; P.o_sw = static_cast<long long>(P.n_taps) * Cin; P.o_tap = Cin; P.o_base = 0;
  P.out = workspace;
  r = launch_bn<1, true>(bn, ta, tb, P, P.tiles_m * P.tiles_n * P.n_taps * P.splits, stream);
  if (r) return r;
  const int blocks = static_cast<int>((per + 255) / 256 < 2048 ? (per + 255) / 256 : 2048);
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), P.splits, Cout, P.n_taps, Cin,
                                                  grad_oihw, accumulate, 1);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}
