// Stem head forward on the tensor pipe: conv1 (1 -> C channels, 3x3, stride (2,1), pad 1) -> BatchNorm affine -> ReLU ->
// MaxPool(3, stride (2,1), pad 1), reference model_v1/model/resnet18.py:74-77, without materialising the conv output.
//
// stemhead.cu recomputes the K = 9 convolution on the FP32 pipe (22.5 FMAs per pooled output and channel: 0.88 ms per
// training step, 1.9 ms per 512 inference lines - instruction bound).  Here the convolution is a tcgen05 GEMM with the
// CHANNELS on the M axis (TMEM lane = channel) and the pixels of one conv row on the N axis (TMEM column = pixel), so
// that the 3x3 pooling window of a channel lies in ONE thread's registers: three accumulators (conv rows 2ho-1, 2ho,
// 2ho+1) of 80 columns serve 64 pooled outputs (8 columns of halo each side keep the TMEM loads 16-column aligned).
//   operands   fp16 with a hi/lo split, K = 32: patch row = [x_hi(9 taps) | x_lo(9) | x_hi(9) | 0(5)], weight row =
//              [w_hi(9) | w_hi(9) | w_lo(9) | 0(5)]  =>  sum_t (x_hi + x_lo) w_hi + x_hi w_lo: the fp32 convolution to
//              ~2^-21 relative (fp16 products are exact in the fp32 accumulator) - the train-mode parity of the fp32
//              kernel is kept.  Both tiles are K-major SWIZZLE_128B rows written by ordinary shared-memory stores
//              (16-byte chunk c of row r at chunk c ^ (r & 7)); the image is not a TMA operand (1 channel, im2col'd
//              by two producer warps from 7 staged image rows);
//   roles      warps 0-7: epilogue (warp w reads TMEM lane quadrant w & 3 = 32 channels, pooled columns 32 (w >> 2) ..
//              +31 of the tile), two im2col producer warps, one warp for TMEM allocation + MMA issue (one thread); for
//              C <= 192 the fourth lane quadrant is empty and those three roles sit on ITS scheduler (warps 3, 7, 11);
//   pipeline   work item = (unit, channel group): unit = 64 pooled outputs of one pooled row, group = C / 2 channels on
//              an M = 128 tile (C = 192: 96 real rows, the 4th lane quadrant idles).  TMEM holds two items (2 x 240
//              columns), shared memory two units' patch tiles: the MMAs of item k+1 and the im2col of unit u+1 run
//              under the epilogue of item k;
//   epilogue   y = acc * scale + shift per conv element, vertical max (+ first-maximum row), horizontal max with torch's
//              first-maximum-in-(kh, kw)-order rule, ReLU, 4-bit arg-max code (15 = no gradient), lane pairs exchange
//              one value per two outputs so every lane stores two adjacent channels (4 bytes) per store.
#include "common.cuh"

namespace htrvt {
namespace {

constexpr int kTcThreads = 384;                 // 12 warps
constexpr int kTcN = 80;                        // conv columns per tile
constexpr int kTcOut = 64;                      // pooled outputs per tile row
constexpr int kTcRowBytes = 128;                // SWIZZLE_128B K-major row pitch (64 halfs; K = 32 of them are used)
constexpr int kTcBTile = kTcN * kTcRowBytes;    // one conv row's patch tile: 10240 B (a multiple of 1024)
constexpr int kTcBUnit = 3 * kTcBTile;
constexpr int kTcATile = 128 * kTcRowBytes;     // one channel group's weights
constexpr int kTcAcc = 3 * kTcN;                // TMEM columns per item
constexpr int kTcStageCols = 84;                // staged image columns per row (82 used)
constexpr int kTcSmem = 2 * kTcATile + 2 * kTcBUnit + 7 * kTcStageCols * 4 + 128 + 1024;

// fp32 -> (hi, lo) fp16 pair in one word: low half = hi, high half = lo
__device__ __forceinline__ uint32_t split_hl(float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  return static_cast<uint32_t>(__half_as_ushort(h)) | (static_cast<uint32_t>(__half_as_ushort(l)) << 16);
}
// the 32 halfs of an operand row from nine (hi, lo) words: [a(9) | b(9) | c(9) | 0(5)] with a / b / c = hi or lo parts
template <bool WEIGHT>
__device__ __forceinline__ void build_row(const uint32_t (&v)[9], uint32_t (&o)[16]) {
  uint16_t h[32];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const uint16_t hi = static_cast<uint16_t>(v[t] & 0xffffu), lo = static_cast<uint16_t>(v[t] >> 16);
    h[t] = hi;                                  // patch: x_hi   weight: w_hi
    h[9 + t] = WEIGHT ? hi : lo;                // patch: x_lo   weight: w_hi
    h[18 + t] = WEIGHT ? lo : hi;               // patch: x_hi   weight: w_lo
  }
#pragma unroll
  for (int t = 27; t < 32; ++t) h[t] = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = static_cast<uint32_t>(h[2 * i]) | (static_cast<uint32_t>(h[2 * i + 1]) << 16);
}
__device__ __forceinline__ void store_row(uint32_t row_addr, int r, const uint32_t (&o)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t a = row_addr + (static_cast<uint32_t>(c ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o[4 * c]), "r"(o[4 * c + 1]),
                 "r"(o[4 * c + 2]), "r"(o[4 * c + 3])
                 : "memory");
  }
}
// fp16 x fp16 -> fp32, K-major operands
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// waits off the critical path poll with a back-off so that they do not take issue slots from the epilogue warps
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}

// u / d for u < 2^21, d <= 2048 with m = ceil(2^32 / d)
__device__ __forceinline__ int fast_div(int u, uint32_t m, int d) { return d == 1 ? u : static_cast<int>(__umulhi(static_cast<uint32_t>(u), m)); }

struct TcHeadP {
  const float* x; const float* w; const float* scale; const float* shift;
  void* out; void* out_bf; uint8_t* code;
  int B, H, W, C, Ho, Hc, ngroups, cg, tiles_w;
  uint32_t mul_tw, mul_ho;       // ceil(2^32 / tiles_w), ceil(2^32 / Ho): exact quotients for the unit counts admitted below
  int units;
};

template <bool CODE, int FMT, bool BFC>      // FMT: 0 = bf16 out, 1 = fp16 out; BFC: + a bf16 copy of an fp16 out
__global__ void __launch_bounds__(kTcThreads, 1) stem_head_tc_kernel(const TcHeadP P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                  // [2][128 rows][128 B]
  uint8_t* sB = smem + 2 * kTcATile;                   // [2 units][3 conv rows][80 rows][128 B]
  uint32_t* img = reinterpret_cast<uint32_t*>(sB + 2 * kTcBUnit);      // [7][84] (hi, lo) words
  uint64_t* bars = reinterpret_cast<uint64_t*>(img + 7 * kTcStageCols);
  uint64_t* bfull = bars;                              // [2] patch tiles of a unit written
  uint64_t* bempty = bars + 2;                         // [2] ... consumed by the MMAs
  uint64_t* tfull = bars + 4;                          // [2] accumulators of an item complete
  uint64_t* tempty = bars + 6;                         // [2] ... read by the 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = P.W, C = P.C, cg = P.cg;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bfull[i], 2); mbar_init(&bempty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * ((cg + 31) / 32)); }
    fence_barrier_init();
  }
  // weights: group g, row r = channel g * cg + r (rows >= cg are zero)
  for (int i = threadIdx.x; i < 2 * 128; i += kTcThreads) {
    const int g = i >> 7, r = i & 127;
    uint32_t v[9], o[16];
    const bool real = g < P.ngroups && r < cg;
    // channels with a negative BatchNorm scale get NEGATED weights: y = scale * conv + shift = |scale| * conv' + shift is
    // then non-decreasing in conv' for every channel, so the vertical maximum can be taken on the raw accumulators
    const float sgn = (real && __ldg(P.scale + g * cg + r) < 0.f) ? -1.f : 1.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = real ? split_hl(sgn * __ldg(P.w + (g * cg + r) * 9 + t)) : 0u;
    build_row<true>(v, o);
    store_row(smem_u32(sA) + g * kTcATile + r * kTcRowBytes, r, o);
  }
  fence_proxy_async();
  // C <= 192: the fourth lane quadrant of the M = 128 tile holds no channels, so its scheduler (warps 3, 7, 11) takes the
  // producers and the MMA issue instead of sharing the epilogue warps' issue slots
  const bool q3free = cg <= 96;
  const int prod_a = q3free ? 3 : 8, prod_b = q3free ? 7 : 9, mma_w = q3free ? 11 : 10;
  if (warp == mma_w) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == prod_a || warp == prod_b) {
    // =========================== im2col producers (64 threads) ===========================
    const int t = (warp == prod_a ? 0 : 32) + lane;
    // image rows 4 ho - 3 .. 4 ho + 3, columns w0 - 9 .. w0 + 72 of a unit (zero outside the image: the conv's padding):
    // thread t takes column t (and t + 64 for t < 18) of all seven rows.  The loads of unit u + 1 are issued before
    // the patch rows of unit u are built, so their latency hides under the build.
    float v0[7], v1[7];
    auto load_unit = [&](int u) {
      const int rw = fast_div(u, P.mul_tw, P.tiles_w), wt = u - rw * P.tiles_w;
      const int n = fast_div(rw, P.mul_ho, P.Ho), ho = rw - n * P.Ho;
      const int ww0 = wt * kTcOut - 9 + t, ww1 = ww0 + 64;
      const bool c0ok = ww0 >= 0 && ww0 < W, c1ok = t < 18 && ww1 < W;
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const int hh = 4 * ho - 3 + r;
        const bool rok = hh >= 0 && hh < P.H;
        const float* rowp = P.x + (static_cast<long long>(n) * P.H + (rok ? hh : 0)) * W;
        v0[r] = (rok && c0ok) ? __ldg(rowp + ww0) : 0.f;
        v1[r] = (rok && c1ok) ? __ldg(rowp + ww1) : 0.f;
      }
    };
    if (static_cast<int>(blockIdx.x) < P.units) load_unit(blockIdx.x);
    int it = 0;
    for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++it) {
      const int buf = it & 1;
      asm volatile("bar.sync 1, 64;" ::: "memory");     // the previous unit's tile build is done with img
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        img[r * kTcStageCols + t] = split_hl(v0[r]);
        if (t < 18) img[r * kTcStageCols + t + 64] = split_hl(v1[r]);
      }
      if (u + static_cast<int>(gridDim.x) < P.units) load_unit(u + gridDim.x);
      asm volatile("bar.sync 1, 64;" ::: "memory");
      if (it >= 2) mbar_wait_relaxed(&bempty[buf], ((it >> 1) - 1) & 1);
      // patch rows: conv row j (image rows 2j .. 2j+2 of the staged seven), pixel p (conv column w0 - 8 + p)
      const uint32_t bbase = smem_u32(sB) + buf * kTcBUnit;
      for (int pr = t; pr < 3 * kTcN; pr += 64) {
        const int j = pr / kTcN, p = pr - j * kTcN;
        uint32_t v[9], o[16];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) v[kh * 3 + kw] = img[(2 * j + kh) * kTcStageCols + p + kw];
        build_row<false>(v, o);
        store_row(bbase + j * kTcBTile + p * kTcRowBytes, p, o);
      }
      fence_proxy_async();                               // generic-proxy stores -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bfull[buf]);
    }
  } else if (warp == mma_w) {
    // =========================== MMA issue (one thread) ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_f16(128, kTcN);
      int it = 0, item = 0;
      for (int u = blockIdx.x; u < P.units; u += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait_relaxed(&bfull[buf], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t bbase = smem_u32(sB) + buf * kTcBUnit;
        for (int g = 0; g < P.ngroups; ++g, ++item) {
          const int tb = item & 1;
          if (item >= 2) { mbar_wait(&tempty[tb], ((item >> 1) - 1) & 1); tc_fence_after(); }
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t da = umma_desc_sw128(smem_u32(sA) + g * kTcATile + kk * 32, 16, 1024);
              const uint64_t db = umma_desc_sw128(bbase + j * kTcBTile + kk * 32, 16, 1024);
              umma_bf16(tmem_base + tb * kTcAcc + j * kTcN, da, db, idesc, kk);
            }
          umma_commit(&tfull[tb]);
        }
        umma_commit(&bempty[buf]);                       // every MMA that reads this unit's patch tiles has completed
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    const int q = warp & 3, hc = warp >> 2;
    int it = 0, item = 0;
    // a lane quadrant without real channels (C = 192: rows 96-127 of the M = 128 tile) has nothing to do and is not
    // counted in tempty
    for (int u = (warp < 8 && 32 * q < cg) ? static_cast<int>(blockIdx.x) : P.units; u < P.units; u += gridDim.x, ++it) {
      const int rw = fast_div(u, P.mul_tw, P.tiles_w), wt = u - rw * P.tiles_w;
      const int n = fast_div(rw, P.mul_ho, P.Ho), ho = rw - n * P.Ho;
      const int w0 = wt * kTcOut;
      for (int g = 0; g < P.ngroups; ++g, ++item) {
        const int tb = item & 1;
        mbar_wait(&tfull[tb], (item >> 1) & 1);
        tc_fence_after();
        {
          const int c = g * cg + 32 * q + lane;          // this thread's channel
          const float asc = fabsf(__ldg(P.scale + c)), sh = __ldg(P.shift + c);      // (sign folded into the weights)
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + tb * kTcAcc + 32 * hc;
          constexpr float kNeg = -3.0e38f;               // finite: the arg-max tag below lives in the low mantissa bits
          // a conv row outside the image (pool padding) is replaced by the middle row (always inside): the maximum and
          // its first-maximum position are unchanged
          uint32_t roff[3], rtag[3];                    // (the stand-in keeps the middle row's tag: a tie decodes to kh = 1)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int cr = 2 * ho - 1 + j;
            const int src = (cr >= 0 && cr < P.Hc) ? j : 1;
            roff[j] = static_cast<uint32_t>(src) * kTcN;
            rtag[j] = static_cast<uint32_t>(3 - src) << 2;
          }
          // column maxima m[i] of the conv columns w0 + 32 hc - 1 + i, i = 0 .. 33: vertical maximum of the raw
          // accumulators, then ONE affine per column.  With CODE the low four mantissa bits of every candidate carry
          // (3 - kh) << 2 | (3 - kw): a plain maximum then returns torch's first maximum in (kh, kw) order, and the
          // winner's tag is the arg-max code (values move by < 2^-19 relative: below fp16 / bf16 resolution)
          float m[34];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            uint32_t r0[16], r1[16], r2[16];
            tmem_ld16(taddr + roff[0] + 16 * ch, r0);
            tmem_ld16(taddr + roff[1] + 16 * ch, r1);
            tmem_ld16(taddr + roff[2] + 16 * ch, r2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int i = 16 * ch + e - 7;             // index into m[]
              if (i < 0 || i >= 34) continue;
              if (CODE) {
                const float x0 = __uint_as_float((r0[e] & ~15u) | rtag[0]);
                const float x1 = __uint_as_float((r1[e] & ~15u) | rtag[1]);
                const float x2 = __uint_as_float((r2[e] & ~15u) | rtag[2]);
                const float xm = fmaxf(x0, fmaxf(x1, x2));
                const float y = fmaf(xm, asc, sh);
                m[i] = __uint_as_float((__float_as_uint(y) & ~15u) | (__float_as_uint(xm) & 12u));
              } else {
                m[i] = fmaf(fmaxf(__uint_as_float(r0[e]), fmaxf(__uint_as_float(r1[e]), __uint_as_float(r2[e]))), asc, sh);
              }
            }
          }
          if (w0 + 32 * hc - 1 < 0) m[0] = kNeg;         // conv column -1 / W: pool padding
          if (w0 + 32 * hc + 32 >= W) m[33] = kNeg;
          // pooled outputs o = 0 .. 31 (conv columns o-1, o, o+1 = m[o], m[o+1], m[o+2]); pairs of outputs per store
          const long long gp0 = (static_cast<long long>(n) * P.Ho + ho) * W + w0 + 32 * hc;
          const int odd = lane & 1;
          char* po = static_cast<char*>(P.out) + ((gp0 + odd) * C + (c - odd)) * 2;
          char* pb = BFC ? static_cast<char*>(P.out_bf) + ((gp0 + odd) * C + (c - odd)) * 2 : nullptr;
          uint8_t* pc = CODE ? P.code + (gp0 + odd) * (C >> 1) + ((c - odd) >> 1) : nullptr;
          const long long ostep = 4LL * C;               // two pixels further, in bytes of a 16-bit tensor
          const uint32_t psel = odd ? 0x3276u : 0x5410u; // byte_perm: even lane (mine.lo, partner.lo), odd (partner.hi, mine.hi)
#pragma unroll
          for (int o = 0; o < 32; o += 2) {
            float val[2];
            uint32_t cd[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const int oo = o + s;
              if (CODE) {
                const float a0 = __uint_as_float(__float_as_uint(m[oo]) | 3u);
                const float a1 = __uint_as_float(__float_as_uint(m[oo + 1]) | 2u);
                const float a2 = __uint_as_float(__float_as_uint(m[oo + 2]) | 1u);
                const float best = fmaxf(a0, fmaxf(a1, a2));
                // tag = (3 - kh) << 2 | (3 - kw)  =>  code = 4 kh + kw = tag ^ 15; 15 = no gradient (ReLU inactive)
                cd[s] = best > 0.f ? ((__float_as_uint(best) & 15u) ^ 15u) : 15u;
                val[s] = fmaxf(best, 0.f);
              } else {
                val[s] = fmaxf(fmaxf(m[oo], fmaxf(m[oo + 1], m[oo + 2])), 0.f);
                cd[s] = 0u;
              }
            }
            // even lane keeps output o (channels c, c+1), odd lane output o+1 (channels c-1, c)
            if (CODE) {                                    // one float exchange serves both tensors and the codes
              const float send = odd ? val[0] : val[1];
              const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
              const float lo = odd ? recv : val[0], hi = odd ? val[1] : recv;
              *reinterpret_cast<uint32_t*>(po) = FMT == 1 ? pack_f16(lo, hi) : pack_bf16(lo, hi);
              if (BFC) *reinterpret_cast<uint32_t*>(pb) = pack_bf16(lo, hi);
              const uint32_t csend = odd ? cd[0] : cd[1];
              const uint32_t crecv = __shfl_xor_sync(0xffffffffu, csend, 1);
              const uint32_t byte = odd ? (crecv | (cd[1] << 4)) : (cd[0] | (crecv << 4));
              *pc = static_cast<uint8_t>(byte);
              pc += C;
            } else {                                       // pack first, exchange the packed word, pick halves by PRMT
              const uint32_t pk = FMT == 1 ? pack_f16(val[0], val[1]) : pack_bf16(val[0], val[1]);
              *reinterpret_cast<uint32_t*>(po) = __byte_perm(pk, __shfl_xor_sync(0xffffffffu, pk, 1), psel);
              if (BFC) {
                const uint32_t pk2 = pack_bf16(val[0], val[1]);
                *reinterpret_cast<uint32_t*>(pb) = __byte_perm(pk2, __shfl_xor_sync(0xffffffffu, pk2, 1), psel);
              }
            }
            po += ostep;
            if (BFC) pb += ostep;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[tb]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == mma_w) tmem_dealloc(tmem_base, 512);
}

int g_head_tc_mode = -1;            // -1: read HTRVT_STEMHEAD_TC (default on), 0 off, 1 on

template <bool CODE, int FMT, bool BFC>       // one function per kernel: HTRVT_ENSURE_SMEM remembers per call site
int tc_launch_one(const TcHeadP& P, int grid, cudaStream_t stream) {
  auto kern = stem_head_tc_kernel<CODE, FMT, BFC>;
  if (!HTRVT_ENSURE_SMEM(kern, kTcSmem)) return HTRVT_ERR_LAUNCH;
  kern<<<grid, kTcThreads, kTcSmem, stream>>>(P);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

}  // namespace

// 1 when the tensor-pipe stem head serves the shape (otherwise stemhead.cu's FP32-pipe kernel runs)
int stem_head_tc_supported(int B, int H, int W, int C, int out_fmt) {
  if (g_head_tc_mode == -1) {
    const char* e = getenv("HTRVT_STEMHEAD_TC");
    g_head_tc_mode = e ? (atoi(e) != 0) : 1;
  }
  if (!g_head_tc_mode) return 0;
  if (out_fmt != 0 && out_fmt != 1) return 0;
  if (B <= 0 || H < 4 || (H & 1) || W < kTcOut || (W % kTcOut) || C < 32 || C > 256) return 0;
  const int ng = C > 128 ? 2 : 1;
  if (C % ng) return 0;
  const int cg = C / ng;
  const int Ho = (H / 2 - 1) / 2 + 1;
  if (static_cast<long long>(B) * Ho * (W / kTcOut) >= (1LL << 21) || Ho > 2048 || W / kTcOut > 2048) return 0;   // fast_div range
  return (cg % 32) == 0;
}

int stem_head_tc_launch(const float* x, const float* w, const float* scale, const float* shift, void* out, void* out_bf,
                        void* code, int B, int H, int W, int C, int out_fmt, cudaStream_t stream) {
  TcHeadP P;
  P.x = x; P.w = w; P.scale = scale; P.shift = shift; P.out = out; P.out_bf = out_bf; P.code = static_cast<uint8_t*>(code);
  P.B = B; P.H = H; P.W = W; P.C = C; P.Hc = H / 2; P.Ho = (P.Hc - 1) / 2 + 1;
  P.ngroups = C > 128 ? 2 : 1; P.cg = C / P.ngroups; P.tiles_w = W / kTcOut;
  const long long units = static_cast<long long>(B) * P.Ho * P.tiles_w;
  if (units >= (1LL << 21) || P.tiles_w > 2048 || P.Ho > 2048) return HTRVT_ERR_SHAPE;
  P.mul_tw = static_cast<uint32_t>(((1ULL << 32) + P.tiles_w - 1) / P.tiles_w);
  P.mul_ho = static_cast<uint32_t>(((1ULL << 32) + P.Ho - 1) / P.Ho);
  P.units = static_cast<int>(units);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = P.units < sms ? static_cast<int>(P.units) : sms;
  if (out_fmt == 0) {                                  // bf16 output: it is its own bf16 copy
    P.out_bf = nullptr;
    return code ? tc_launch_one<true, 0, false>(P, grid, stream) : tc_launch_one<false, 0, false>(P, grid, stream);
  }
  if (code) return out_bf ? tc_launch_one<true, 1, true>(P, grid, stream) : tc_launch_one<true, 1, false>(P, grid, stream);
  return out_bf ? tc_launch_one<false, 1, true>(P, grid, stream) : tc_launch_one<false, 1, false>(P, grid, stream);
}

}  // namespace htrvt

// 0: FP32-pipe stem head only, 1: tensor-pipe stem head where the shape allows (default); returns the previous mode
extern "C" int htrvt_stem_head_set_mode(int mode) {
  const int prev = htrvt::g_head_tc_mode == -1 ? 1 : htrvt::g_head_tc_mode;
  htrvt::g_head_tc_mode = mode ? 1 : 0;
  return prev;
}
