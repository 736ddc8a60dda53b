// The sm_100a tap GEMM: persistent, warp-specialised tcgen05 kernel.
//   warp 0   : TMA producer (one elected lane) - 5-D tiled tensor maps, SWIZZLE_128B, OOB zero fill
//              implements conv padding, ragged tiles and K tails
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16, bf16 -> fp32 in TMEM)
// KIND: 0 = rows are output pixels (linear / conv fwd, dgrad); 1 = weight gradient, both operands pixel-major
// (MN-major); 2 = weight gradient with a K-major dY^T operand (tcgen05 runs ~1.4x slower with an MN-major A);
// 3 = weight gradient transposed: rows = (tap, 64-channel atom of x) - every 64-row half of a tile carries its own
// tap shift, so no rows are wasted when Cout is not a multiple of 128 - columns = Cout; when a CTA of a pair stages
// 96 columns the B operand uses SWIZZLE_64B (32-column atoms);
// 4 = transposed weight gradient of 3x3 stride-(.,1) convs with TWO live accumulators (256 rows = four (kernel row,
// channel atom, kw) atoms) that share the dY tile and read x through two staged 72-pixel windows: the three
// horizontal taps are the same window at row offsets 0 / 1 / 2 (the two 64-row halves of an MN-major A operand are
// just two start addresses, LBO apart - they may overlap).
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

// CL = 1: one CTA per 128 x BN tile.  CL = 2: a CTA PAIR (cta_group::2) per 256 x BN tile - each CTA stages its own
// 128 A rows and only HALF of B, so the bytes entering each SM per flop drop by ~1/3 (these GEMMs are bound by the
// ~68 B/clk L2 -> SM ingress: tensor-pipe utilisation tracks 68 / ((128 + BN) * 128 / (2 BN)) for every shape).
// RE (KIND 0, 3x3 stride-1 convs): one stage = ONE 136-pixel window of the activation (the three horizontal taps of a
// kernel row read it at row offsets 0 / 1 / 2 through the descriptor's base offset) + the three taps' weight tiles,
// i.e. K = 192 per stage and 1/3 of the A bytes: these GEMMs run at the ~40 B/clk/SM the L2 -> SM path delivers, so
// bytes per flop, not tensor-pipe issue, set their speed.
constexpr int kWinRows = 136;
constexpr int kWin4Rows = 72;                  // KIND 4: 64 pixels + 2 of halo, padded to whole 1024-byte swizzle atoms
constexpr int kWin4Bytes = kWin4Rows * 128;
// DUAL (fc1 in train mode): the epilogue stores TWO boxes per step (gelu(u) and the derivative gelu'(u)), so each epilogue
// warp gets four staging buffers instead of two and the ring gives up one stage.
// RESQ (fc2's input gradient times the saved gelu'; the BatchNorm-backward epilogue at BN <= 192): the second input tensor
// of the epilogue is staged in shared memory - every epilogue warp copies ITS [32 rows][BN / 2 columns] of the tile with
// coalesced 16-byte cp.async transfers before it waits for the accumulator (16-byte chunks XOR-swizzled by the row, or a
// row pitch of BN + 16 bytes when the row is not a power of two: the per-row reads of 8 consecutive rows fall on 8
// distinct 16-byte bank groups; the operand ring gives up one or two stages).
// Per-thread row loads (32 lanes x 64 B at a row pitch of several KB = 32 sectors per request) had left these GEMMs at
// 38 % / 58 % tensor-pipe activity with the L1 data pipe as their busiest unit and most of the epilogue's stall samples
// on those loads.
template <int BN, int CL, bool RE = false, bool W4 = false, bool DUAL = false, bool STG2 = false, bool RESQ = false>
struct GemmCfg {
  static constexpr int kABytes = W4 ? 2 * kWin4Bytes : RE ? kWinRows * kBK * 2 : kBM * kBK * 2;
  static constexpr int kBTile = (BN / CL) * kBK * 2;                  // one tap's B bytes staged by THIS CTA
  static constexpr int kBBytes = RE ? 3 * kBTile : kBTile;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // bytes per staged row: BN / 2 16-bit columns; a power-of-two row (BN = 256) is XOR-swizzled at 16-byte granularity,
  // any other gets one 16-byte pad
  static constexpr bool kResXor = ((BN / 2) & (BN / 2 - 1)) == 0;
  static constexpr int kResPitch = kResXor ? BN : BN + 16;
  static constexpr int kResBytes = RESQ ? 8 * 32 * kResPitch : 0;
  static constexpr int kStages = (W4 ? (BN >= 256 ? 3 : 4) : RE ? (BN >= 256 ? 2 : 3) : (CL == 2) ? 6 : ((BN <= 128) ? 6 : 4)) -
                                 (DUAL ? 1 : 0) - (RESQ ? (RE ? 1 : 2) : 0);
  static_assert(kStages >= 2, "operand ring");
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  // per epilogue warp: [32 rows][64 B] output boxes, two buffers; DUAL / STG2 (BatchNorm-backward epilogue: a second box
  // per step that only feeds the column sums) double them
  static constexpr int kWarpStaging = (DUAL || STG2) ? 4 * 2048 : 2 * 2048;
  static constexpr int kStagingBytes = 8 * kWarpStaging;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 256 /*barriers*/ + kResBytes;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert(kStageBytes % 1024 == 0, "SWIZZLE_128B stage alignment");
};

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

// EPI: compile-time epilogue variant, so that the common epilogue carries none of the others' code or registers
// (a runtime-flag version of the GELU / fp16 paths cost every KIND 0 kernel 15-40 % - measured):
//   0 plain (bias / ReLU / residual / statistics / accumulate; bf16 or fp32 out)
//   1 plain with IEEE fp16 output + residual (forward stem convolutions)
//   2 gelu(acc + bias), bf16 out            3 the same + gelu'(acc + bias) to a second tensor (DUAL)
//   4 acc * res, bf16 out (fc2 input gradient + activation backward: res = the saved gelu')
//   5 acc * relu_mask, bf16 out, + the BatchNorm-backward column sums of the layer in front (EPI_BN_BWD)
constexpr int kEpiPlain = 0, kEpiF16 = 1, kEpiGelu = 2, kEpiGeluDual = 3, kEpiGeluBwd = 4, kEpiBnBwd = 5;
template <int BN, int KIND, bool B_MN, int CL, bool RE = false, int EPI = 0>
__global__ void __launch_bounds__(kGemmThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
               const __grid_constant__ GemmP P) {
  constexpr bool W4 = (KIND == 4);
  constexpr bool DUAL = (EPI == kEpiGeluDual), F16 = (EPI == kEpiF16);
  constexpr bool GELU = (EPI == kEpiGelu || EPI == kEpiGeluDual), GELUB = (EPI == kEpiGeluBwd);
  constexpr bool BNB = (EPI == kEpiBnBwd);
  constexpr bool BOX2 = DUAL || BNB;                   // a second staged box per step
  constexpr bool RESQ = GELUB || (BNB && BN <= 192);   // second input tensor staged in shared memory by cp.async
  using Cfg = GemmCfg<BN, CL, RE, W4, DUAL, BNB, RESQ>;
  static_assert(EPI == 0 || KIND == 0, "epilogue variants exist for kind 0 only");
  static_assert(!(GELU || GELUB) || !RE, "GELU epilogues: linear layers only");
  static_assert(!RE || (KIND == 0 && !B_MN && CL == 2), "window reuse: kind 0, K-major weights, CTA pairs");
  static_assert(!W4 || (B_MN && CL == 1 && !RE), "two-accumulator weight gradient: single CTA, MN-major operands");
  constexpr bool A_MN = (KIND == 1 || KIND == 3 || KIND == 4);
  constexpr bool B_SW64 = B_MN && ((BN / CL) % 64) != 0;     // a 96-column half: 32-column atoms, SWIZZLE_64B
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem[];      // SWIZZLE_128B tiles need 1024-byte alignment
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* stage_base = smem;
  uint8_t* stage_out = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + Cfg::kStagingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = P.tiles_m * P.tiles_n * (P.kind == 0 ? 1 : P.n_taps * P.splits);
  // CL == 2: the CTA pair works on tiles (2p, 2p+1) = neighbouring m-tiles of the same n-tile (the host guarantees
  // tiles_m is even).  Each CTA runs its own producer and epilogue; only the leader (rank 0) issues the MMAs.
  const int crank = (CL == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_tile = (CL == 2) ? (static_cast<int>(blockIdx.x >> 1) * 2 + crank) : static_cast<int>(blockIdx.x);
  const int tile_step = (CL == 2) ? static_cast<int>(gridDim.x & ~1u) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (DUAL) tma_prefetch_desc(&tmC2);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8 * CL); }
    fence_barrier_init();
  }
  if (warp == 1) { if (CL == 2) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols); else tmem_alloc(tmem_slot, Cfg::kTmemCols); }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();           // the peer's barriers must be initialised before anything lands on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, the pair's
  // cluster sync) touches no global memory and may run under the tail of the kernel in front; from here on the kernel
  // reads and writes tensors, so it waits for that kernel's completion + flush.  The trigger right behind it lets the
  // NEXT launch do the same under this kernel's tail (its CTAs become resident as ours exit).  Both are no-ops for a
  // launch without the programmatic-serialisation attribute.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int id = first_tile; id < total_tiles; id += tile_step) {
        const TileCoord tc = decode_tile(P, id);
        const int n0 = tc.n_tile * BN, m0 = tc.m_tile * kBM;
        const int kiters = RE ? (P.n_taps / 3) * P.k_chunks
                              : (KIND == 0) ? P.n_taps * P.k_chunks : (tc.q_end - tc.q_begin);
        // running coordinates of the K loop (this one thread feeds the tensor pipe: no divisions per iteration)
        int tap = 0, cc = 0;                                  // KIND 0: tap and 64-channel chunk
        int wq = 0, ho = 0, n = 0;                            // KIND != 0: 64-pixel chunk of image row (n, ho)
        if (KIND != 0) {
          const int row = tc.q_begin / P.k_chunks;
          wq = tc.q_begin - row * P.k_chunks;
          n = row / P.Ho;
          ho = row - n * P.Ho;
        }
        // KIND 3: the tap window of each 64-channel half of the A tile is fixed per tile
        int a_c0[kBM / 64], a_pw[kBM / 64], a_dw[kBM / 64], a_dh[kBM / 64];
        bool a_ok[kBM / 64];
        if (KIND == 3) {
#pragma unroll
          for (int i = 0; i < kBM / 64; ++i) {
            const int atom = tc.m_tile * (kBM / 64) + i;
            const int t = atom / P.a_atoms_per_tap;
            a_ok[i] = t < P.a_taps;
            a_c0[i] = (atom - t * P.a_atoms_per_tap) * 64;
            a_pw[i] = a_ok[i] ? P.tap.pw[t] : 0;
            a_dw[i] = a_ok[i] ? P.tap.dw[t] : 0;
            a_dh[i] = a_ok[i] ? P.tap.dh[t] : 0;
          }
        }
        // KIND 4: the unit's four atoms live in two windows = the (kernel row, channel atom) of its first and last atom
        int w4_c[2] = {0, 0}, w4_dh[2] = {0, 0};
        if (W4) {
          const int ca3 = 3 * P.a_atoms_per_tap, n_atoms = 3 * ca3;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int a = min(4 * tc.m_tile + 3 * i, n_atoms - 1);
            const int kh = a / ca3;
            w4_c[i] = ((a - kh * ca3) / 3) * 64;
            w4_dh[i] = kh - 1;
          }
        }
        for (int k = 0; k < kiters; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          // CL == 2: every load of the pair reports to the LEADER's full barrier, which expects both CTAs' bytes
          if (CL == 1) mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
          else if (crank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
          auto load = [&](uint8_t* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
            if (CL == 2) tma_load_5d_pair(dst, map, &full[stage], c0, c1, c2, c3, c4);
            else tma_load_5d(dst, map, &full[stage], c0, c1, c2, c3, c4);
          };
          constexpr int BNL = BN / CL;                    // B columns staged by this CTA
          const int nb0 = n0 + crank * BNL;
          if (RE) {
            // `tap` counts kernel ROWS here: taps 3*tap .. 3*tap+2 share dh; the window starts one pixel to the left
            load(sa, &tmA, cc * kBK, 0, tc.w0 - 1, tc.h * P.a_sh + P.tap.dh[3 * tap], tc.n);
#pragma unroll
            for (int j = 0; j < 3; ++j)
              load(sb + j * Cfg::kBTile, &tmB, P.tap.widx[3 * tap + j] * P.b_tap_stride + cc * kBK, nb0, 0, 0, 0);
            if (++cc == P.k_chunks) { cc = 0; ++tap; }
          } else if (KIND == 0) {
            load(sa, &tmA, cc * kBK, P.tap.pw[tap], tc.w0 + P.tap.dw[tap], tc.h * P.a_sh + P.tap.dh[tap], tc.n);
            if (!B_MN) {
              load(sb, &tmB, P.tap.widx[tap] * P.b_tap_stride + cc * kBK, nb0, 0, 0, 0);
            } else {
#pragma unroll
              for (int i = 0; i < BNL / 64; ++i)
                load(sb + i * 8192, &tmB, P.tap.widx[tap] * P.b_tap_stride + nb0 + 64 * i, cc * kBK, 0, 0, 0);
            }
            if (++cc == P.k_chunks) { cc = 0; ++tap; }
          } else {
            const int w0c = wq * kBK;
            if (W4) {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                load(sa + i * kWin4Bytes, &tmA, w4_c[i], 0, w0c - 1, ho * P.a_sh + w4_dh[i], n);
#pragma unroll
              for (int i = 0; i < BNL / 64; ++i) load(sb + i * 8192, &tmB, nb0 + 64 * i, 0, w0c, ho, n);
            } else if (KIND == 3) {
              // A: two 64-channel atoms of the shifted activation, each with its own tap; B: dY, unshifted
#pragma unroll
              for (int i = 0; i < kBM / 64; ++i)
                load(sa + i * 8192, &tmA, a_c0[i], a_pw[i], w0c + a_dw[i], ho * P.a_sh + a_dh[i],
                     a_ok[i] ? n : P.NB);                     // past the last tap: an all-OOB box = zeros
              if (B_SW64) {
#pragma unroll
                for (int i = 0; i < BNL / 32; ++i) load(sb + i * 4096, &tmB, nb0 + 32 * i, 0, w0c, ho, n);
              } else {
#pragma unroll
                for (int i = 0; i < BNL / 64; ++i) load(sb + i * 8192, &tmB, nb0 + 64 * i, 0, w0c, ho, n);
              }
            } else {
              if (KIND == 1) {
#pragma unroll
                for (int i = 0; i < kBM / 64; ++i) load(sa + i * 8192, &tmA, m0 + 64 * i, 0, w0c, ho, n);
              } else {            // KIND 2: dY^T [n][ho][C][Wo] - pixels contiguous, a plain K-major tile
                load(sa, &tmA, w0c, m0, ho, n, 0);
              }
#pragma unroll
              for (int i = 0; i < BNL / 64; ++i)
                load(sb + i * 8192, &tmB, nb0 + 64 * i, P.tap.pw[tc.tap], w0c + P.tap.dw[tc.tap],
                     ho * P.a_sh + P.tap.dh[tc.tap], n);
            }
            if (++wq == P.k_chunks) { wq = 0; if (++ho == P.Ho) { ho = 0; ++n; } }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0 && crank == 0) {
      // operand formats live in bits [7,10) (A) and [10,13) (B) of the instruction descriptor: 1 = bf16, 0 = fp16
      const uint32_t idesc = umma_idesc_bf16(kBM * CL, BN, A_MN ? 1 : 0, B_MN ? 1 : 0) &
                             ~((P.a_f16 ? (1u << 7) : 0u) | (P.b_f16 ? (1u << 10) : 0u));
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int id = first_tile; id < total_tiles; id += tile_step) {
        const TileCoord tc = decode_tile(P, id);
        const int kiters = RE ? (P.n_taps / 3) * P.k_chunks
                              : (KIND == 0) ? P.n_taps * P.k_chunks : (tc.q_end - tc.q_begin);
        mbar_wait(&tempty[as], aphase ^ 1);
        if (W4) mbar_wait(&tempty[1], aphase ^ 1);             // a unit owns BOTH accumulators
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        // KIND 4: byte offset of each of the unit's four atoms inside the stage's A region (window, row = kw)
        uint32_t w4_off[4] = {0u, 0u, 0u, 0u};
        if (W4) {
          const int ca3 = 3 * P.a_atoms_per_tap, n_atoms = 3 * ca3;
          const int a0 = min(4 * tc.m_tile, n_atoms - 1), g0 = a0 / 3;       // g = (kernel row, channel atom) index
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int a = min(4 * tc.m_tile + j, n_atoms - 1), g = a / 3;
            w4_off[j] = static_cast<uint32_t>((g != g0 ? kWin4Bytes : 0) + (a - 3 * g) * 128);
          }
        }
        int re_row = 0, re_cc = 0;                             // RE: kernel row / channel chunk of this stage
        for (int k = 0; k < kiters; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          if (RE) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const uint32_t arow = sa + static_cast<uint32_t>(P.tap.dw[3 * re_row + j] + 1) * 128u;   // window row 0..2
#pragma unroll
              for (int kk = 0; kk < kBK / 16; ++kk) {
                // start address = window row: the swizzle phase follows the ABSOLUTE smem address bits (TMA wrote the
                // tile at a 1024-byte boundary), so no descriptor base offset is needed - verified tap by tap
                const uint64_t da = umma_desc_sw128(arow + kk * 32, 16, 1024);
                const uint64_t db = umma_desc_sw128(sb + j * Cfg::kBTile + kk * 32, 16, 1024);
                umma_bf16_pair(d_tmem, da, db, idesc, (k | j | kk) != 0 ? 1u : 0u);
              }
            }
            if (++re_cc == P.k_chunks) { re_cc = 0; ++re_row; }
          } else if (W4) {
#pragma unroll
            for (int acc = 0; acc < 2; ++acc) {
              const uint32_t a_lo = sa + w4_off[2 * acc];
              const uint32_t lbo = w4_off[2 * acc + 1] - w4_off[2 * acc];      // second 64-row half of the A tile
#pragma unroll
              for (int kk = 0; kk < kBK / 16; ++kk) {
                const uint64_t da = umma_desc_sw128(a_lo + kk * 2048, lbo, 1024);
                const uint64_t db = umma_desc_sw128(sb + kk * 2048, 8192, 1024);
                umma_bf16(tmem_base + acc * BN, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
              }
            }
          } else
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            const uint64_t da = A_MN ? umma_desc_sw128(sa + kk * 2048, 8192, 1024)
                                     : umma_desc_sw128(sa + kk * 32, 16, 1024);
            const uint64_t db = B_SW64 ? umma_desc_sw64(sb + kk * 1024, 4096, 512)
                                : B_MN ? umma_desc_sw128(sb + kk * 2048, 8192, 1024)
                                       : umma_desc_sw128(sb + kk * 32, 16, 1024);
            if (CL == 2) umma_bf16_pair(d_tmem, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, da, db, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          if (CL == 2) umma_commit_pair(&empty[stage], 3); else umma_commit(&empty[stage]);   // frees the slot in both CTAs
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (CL == 2) umma_commit_pair(&tfull[as], 3); else umma_commit(&tfull[as]);           // wakes both epilogues
        if (W4) { umma_commit(&tfull[1]); aphase ^= 1; continue; }                            // as stays 0
        as ^= 1; if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // =========================== epilogue (8 warps) ===========================
    // Deliberately light: only 8 warps run it, so anything beyond scale / bias / ReLU lives in separate
    // full-occupancy kernels (GELU, residual add) or in the memory system (accumulate = TMA reduce-add).
    // Every warp is autonomous: per 64-byte-wide box (32 bf16 or 16 fp32 columns) of ITS 32 rows it does
    // tcgen05.ld -> registers -> swizzled smem staging (two buffers per warp) -> lane 0 issues the TMA store.
    // TMA stores queue behind the producer's bulk loads inside the TMA unit (~1 us under load), so a warp never
    // waits for the store it just issued, only for the one issued two boxes ago (wait_group.read 1).
    // BatchNorm column sums are read back from the staged box and accumulated in registers across tiles.
    const int ew = warp - 2;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int half = ew >> 2;                // column half handled by this warp
    const int r = quad * 32 + lane;          // tile row owned in the register phase
    constexpr int kColsPerWarp = BN / 2;
    const bool out_bf16 = (EPI != kEpiPlain) || (P.flags & EPI_BF16) != 0;   // 16-bit output (bf16, or fp16 with EPI 1)
    const int box_cols = out_bf16 ? 32 : 16;
    uint8_t* stg = stage_out + ew * Cfg::kWarpStaging;      // 2 (DUAL: 2 x 2) x [32 rows][64 B], SWIZZLE_64B
    const int sw_r = (lane >> 1) & 3;
    int sbuf = 0;
    float ssum[4] = {0.f, 0.f, 0.f, 0.f}, qsum[4] = {0.f, 0.f, 0.f, 0.f};
    int stats_ntile = -1;
    // GELUB + EPI_COLSUM: column sums of the stored tile go straight into P.stats[N] (+=, atomics): fc1's bias gradient
    const bool colsum = (GELUB && (P.flags & EPI_COLSUM) != 0) || BNB;
    auto flush_stats = [&](int n_tile) {
      if (n_tile < 0) return;
      float* dst = P.stats + (colsum ? 0ll : static_cast<long long>(blockIdx.x * 4 + quad) * 2 * P.N_valid);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int col = n_tile * BN + half * kColsPerWarp + b * 32 + lane;
        if (b * 32 < kColsPerWarp && col < P.N_valid) {
          if (colsum) {
            atomicAdd(dst + col, ssum[b]);
            if (BNB) atomicAdd(dst + P.N_valid + col, qsum[b]);
          } else {
            dst[col] += ssum[b];
            dst[P.N_valid + col] += qsum[b];
          }
        }
        ssum[b] = 0.f; qsum[b] = 0.f;
      }
    };
    uint4 rq_nxt[4] = {};
    auto res_load = [&](const TileCoord& tc, int col, uint4 (&dst)[4]) {
      const int w = tc.w0 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(0u, 0u, 0u, 0u);
      if (w < P.Wo && col < P.N_valid) {
        const uint4* rp = reinterpret_cast<const uint4*>(
            static_cast<const __nv_bfloat16*>(P.res) +
            ((static_cast<long long>(tc.n) * P.Ho + tc.h) * P.Wo + w) * P.N_valid + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = __ldg(rp + i);
      }
    };
    uint32_t mw_nxt = 0u;
    auto mask_load = [&](const TileCoord& tc, int col) -> uint32_t {      // BNB: ReLU mask bits of 32 channels of a pixel
      const int w = tc.w0 + r;
      if (w < P.Wo && col < P.N_valid)
        return __ldg(reinterpret_cast<const uint32_t*>(
            P.mask + ((((static_cast<long long>(tc.n) * P.Ho + tc.h) * P.Wo + w) * P.N_valid + col) >> 3)));
      return 0u;
    };
    constexpr bool kResEpi = (GELUB || BNB) && !RESQ;   // epilogues that read a second tensor row by row from global
    // RESQ: this warp's [32 rows][kColsPerWarp] of the second tensor, row pitch Cfg::kResPitch bytes
    const uint32_t resq = RESQ ? smem_u32(stage_out + Cfg::kStagingBytes + 256) + ew * (32 * Cfg::kResPitch) : 0u;
    int as = 0; uint32_t aphase = 0;
    for (int id = first_tile; id < total_tiles; id += tile_step) {
      const TileCoord tc = decode_tile(P, id);
      const int n0 = tc.n_tile * BN;
      if (RESQ) {
        // requested before the wait for the accumulator: the round trip runs under the tile's MMAs.  Chunk 32 j + lane of
        // the slice (row-major, kChunksPerRow 16-byte chunks per row): a warp-wide copy reads whole contiguous row pieces
        constexpr int kChunksPerRow = kColsPerWarp / 8;
        __syncwarp();                                          // every lane is done with the previous tile's rows
#pragma unroll
        for (int j = 0; j < kChunksPerRow; ++j) {
          const int idx = 32 * j + lane;
          const int q = idx / kChunksPerRow, ch = idx - q * kChunksPerRow;     // row of the warp's 32, chunk of the row
          const int col = n0 + half * kColsPerWarp + ch * 8;
          const int w = tc.w0 + quad * 32 + q;
          const bool ok = w < P.Wo && col < P.N_valid;
          const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(P.res) +
              (ok ? ((static_cast<long long>(tc.n) * P.Ho + tc.h) * P.Wo + w) * P.N_valid + col : 0ll);
          const uint32_t dst = resq + q * Cfg::kResPitch + ((Cfg::kResXor ? (ch ^ (q & 7)) : ch) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (BNB) mw_nxt = mask_load(tc, n0 + half * kColsPerWarp);
      }
      if (kResEpi && KIND == 0) {
        // the first box's rows of the second tensor are requested BEFORE the wait for the accumulator: their L2 / DRAM
        // round trip runs under the tile's MMAs
        res_load(tc, n0 + half * kColsPerWarp, rq_nxt);
        if (BNB) mw_nxt = mask_load(tc, n0 + half * kColsPerWarp);
      }
      if (((P.flags & EPI_STATS) || colsum) && tc.n_tile != stats_ntile) { flush_stats(stats_ntile); stats_ntile = tc.n_tile; }
#pragma unroll 1
      for (int sub = 0; sub < (W4 ? 2 : 1); ++sub) {          // KIND 4: a unit = two accumulators = two 128-row tiles
      const int m_tile = W4 ? 2 * tc.m_tile + sub : tc.m_tile;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if (RESQ) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                                          // the rows were copied by other lanes
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + half * kColsPerWarp;
      const int nboxes = kColsPerWarp / box_cols;
#pragma unroll 1
      for (int b = 0; b < nboxes; ++b) {
        const int c_local = half * kColsPerWarp + b * box_cols;
        const int col = n0 + c_local;                        // first global column of this box
        uint32_t raw[32];
        tmem_ld16(taddr + b * box_cols, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
        if (out_bf16) tmem_ld16(taddr + b * box_cols + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
        tmem_ld_wait();
        if (b == nboxes - 1) {                               // accumulator fully read: hand TMEM back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CL == 2) mbar_arrive_leader(&tempty[as]); else mbar_arrive(&tempty[as]); }
        }
        if (P.flags & EPI_NOSTORE) continue;                 // measurement aid: main loop only
        if (((P.flags & EPI_STATS) || BNB) && KIND == 0 && tc.w0 + r >= P.Wo) {
          // rows past the end of the image row can pick up shifted-window data: keep them out of the statistics
#pragma unroll
          for (int i = 0; i < 32; ++i) raw[i] = 0u;
        }
        uint32_t packed[16];
        uint32_t packed2[BOX2 ? 16 : 1];
        if (out_bf16) {
          float bv[32];
          if (P.flags & EPI_BIAS) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {                // N_valid % 4 == 0: whole vectors are in or out
              const float4 t = (col + i < P.N_valid) ? __ldg(reinterpret_cast<const float4*>(P.bias + col + i))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
              bv[i] = t.x; bv[i + 1] = t.y; bv[i + 2] = t.z; bv[i + 3] = t.w;
            }
          }
          uint4 rq[4] = {};
          if (RESQ) {                                        // this thread's row of the staged slice, chunks 4b .. 4b + 3
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t a = resq + lane * Cfg::kResPitch + ((Cfg::kResXor ? ((4 * b + i) ^ (lane & 7)) : (4 * b + i)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(rq[i].x), "=r"(rq[i].y), "=r"(rq[i].z), "=r"(rq[i].w) : "r"(a));
            }
          } else
          if (KIND == 0 && (GELUB || BNB || (!GELU && (P.flags & EPI_RES)))) {   // this thread's output pixel, 32 consecutive channels
            // the 64 bytes of box b were requested one box earlier (rq_nxt): an L2 / DRAM round trip per box would
            // otherwise sit between the TMEM load and the store of every box of this warp
            if (b == 0 && !kResEpi) res_load(tc, col, rq);
            else {
#pragma unroll
              for (int i = 0; i < 4; ++i) rq[i] = rq_nxt[i];
            }
            if (b + 1 < nboxes) res_load(tc, col + box_cols, rq_nxt);
          }
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(rq);
          uint32_t mword = 0u;                               // BNB: ReLU mask bits of this pixel's 32 channels
          if (BNB) {
            mword = mw_nxt;
            if (b + 1 < nboxes) mw_nxt = mask_load(tc, col + box_cols);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a = __uint_as_float(raw[i]) * P.alpha, c = __uint_as_float(raw[i + 1]) * P.alpha;
            if (P.flags & EPI_BIAS) { a += bv[i]; c += bv[i + 1]; }
            if (GELU) {                                      // timm Mlp: fc1 -> nn.GELU (erf form), fused
              if (DUAL) {                                    // train: gelu'(u) goes to the second tensor - the backward's
                float ga, gc;                                // epilogue is then ONE multiply per element (evaluating gelu'
                gelu_val_grad(a, a, ga);                     // from a saved u there left fc2's input-gradient GEMM at 29 %
                gelu_val_grad(c, c, gc);                     // tensor-pipe activity: eight epilogue warps, instruction bound)
                packed2[i >> 1] = pack_bf16(ga, gc);
              } else {
                a = gelu_val(a); c = gelu_val(c);
              }
            }
            if (GELUB) {                                     // du = da * gelu'(u), gelu'(u) saved by the forward epilogue
              const float2 gd = unpack_bf16(rw[i >> 1]);
              a *= gd.x; c *= gd.y;
            }
            if (BNB) {                                       // g' = g * [y > 0]; q = g' * xhat for the column sums
              a = ((mword >> i) & 1u) ? a : 0.f;
              c = ((mword >> (i + 1)) & 1u) ? c : 0.f;
              const uint32_t pk = pack_bf16(a, c);
              const float2 gr = unpack_bf16(pk), xr = unpack_f16(rw[i >> 1]);
              const float2 mv = (col + i < P.N_valid) ? __ldg(reinterpret_cast<const float2*>(P.bn_mean + col + i)) : make_float2(0.f, 0.f);
              const float2 rv = (col + i < P.N_valid) ? __ldg(reinterpret_cast<const float2*>(P.bn_rstd + col + i)) : make_float2(0.f, 0.f);
              packed2[i >> 1] = pack_bf16(gr.x * ((xr.x - mv.x) * rv.x), gr.y * ((xr.y - mv.y) * rv.y));
            }
            if (KIND == 0 && !GELU && !GELUB && !BNB && (P.flags & EPI_RES)) {
              const float2 rr = F16 ? unpack_f16(rw[i >> 1]) : unpack_bf16(rw[i >> 1]);
              a += rr.x; c += rr.y;
            }
            if (!GELU && !GELUB && !BNB && (P.flags & EPI_RELU)) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
            packed[i >> 1] = F16 ? pack_f16(a, c) : pack_bf16(a, c);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((P.flags & EPI_BIAS) && col + i < P.N_valid) t = __ldg(reinterpret_cast<const float4*>(P.bias + col + i));
            float v[4] = {fmaf(__uint_as_float(raw[i]), P.alpha, t.x), fmaf(__uint_as_float(raw[i + 1]), P.alpha, t.y),
                          fmaf(__uint_as_float(raw[i + 2]), P.alpha, t.z), fmaf(__uint_as_float(raw[i + 3]), P.alpha, t.w)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (P.flags & EPI_RELU) v[k] = fmaxf(v[k], 0.f);
              packed[i + k] = __float_as_uint(v[k]);
            }
          }
        }
        // the store issued from this buffer two boxes ago must have finished READING it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        const uint32_t sbase = smem_u32(stg) + sbuf * (BOX2 ? 4096 : 2048);
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          const uint32_t addr = sbase + lane * 64 + ((c16 ^ sw_r) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * c16]),
                       "r"(packed[4 * c16 + 1]), "r"(packed[4 * c16 + 2]), "r"(packed[4 * c16 + 3])
                       : "memory");
          if (BOX2)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2048), "r"(packed2[4 * c16]),
                         "r"(packed2[4 * c16 + 1]), "r"(packed2[4 * c16 + 2]), "r"(packed2[4 * c16 + 3])
                         : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (col < P.N_valid) {
            int c1, c2, c3, c4;
            if (KIND == 0) { c1 = tc.w0 + quad * 32; c2 = tc.h; c3 = tc.n; c4 = 0; }
            else { c1 = m_tile * kBM + quad * 32; c2 = tc.tap; c3 = (P.flags & EPI_ACCUM) ? 0 : tc.split; c4 = 0; }
            if (P.flags & EPI_ACCUM)
              asm volatile(
                  "cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                  ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(sbase), "r"(col), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                  : "memory");
            else
              asm volatile(
                  "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                  ::"l"(reinterpret_cast<uint64_t>(&tmC)), "r"(sbase), "r"(col), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                  : "memory");
            if (DUAL)                                        // the pre-activation box, same coordinates, second tensor
              asm volatile(
                  "cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                  ::"l"(reinterpret_cast<uint64_t>(&tmC2)), "r"(sbase + 2048), "r"(col), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                  : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // (possibly empty) keeps the buffer <-> group pairing
        }
        if ((P.flags & EPI_STATS) || colsum) {
          // column `lane` of this warp's 32 staged rows (bf16 box = 32 columns); rows outside the image are
          // exact zeros (their A rows were TMA zero-filled and convolutions carry no bias)
          float sacc = 0.f, qacc = 0.f;
          const uint32_t base = sbase + (lane & 7) * 2;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            uint16_t hv;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(base + rr * 64 + (((lane >> 3) ^ ((rr >> 1) & 3)) << 4)));
            const float f = F16 ? f16_bits_to_float(hv) : __uint_as_float(static_cast<uint32_t>(hv) << 16);
            sacc += f;
            if (BNB) {                                       // second box: g' * xhat
              uint16_t hq;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hq) : "r"(base + 2048 + rr * 64 + (((lane >> 3) ^ ((rr >> 1) & 3)) << 4)));
              qacc += __uint_as_float(static_cast<uint32_t>(hq) << 16);
            } else {
              qacc = fmaf(f, f, qacc);
            }
          }
          // (static indices: `ssum[b]` with the runtime box index put both arrays in local memory - an LDL / STL round
          // trip per box behind the L1 queue of the epilogue's global loads)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (b == k) { ssum[k] += sacc; qsum[k] += qacc; }
        }
        sbuf ^= 1;
      }
      as ^= 1; if (as == 0) aphase ^= 1;
      }
    }
    if ((P.flags & EPI_STATS) || colsum) flush_stats(stats_ntile);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();           // no CTA may exit while its peer can still reach its smem / TMEM / barriers
  if (warp == 1) { if (CL == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

}  // namespace htrvt

// =================================================================================================
// Host side: tensor maps + C-ABI entry points
// =================================================================================================
#include <cudaTypedefs.h>
#include <cstdlib>
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

// 5-D tensor map, innermost dimension contiguous, zero OOB fill.  dims / box: innermost first; strides: element
// strides of dims 1..4.  Operand maps: bf16 + SWIZZLE_128B; output maps: bf16 or fp32 + SWIZZLE_64B.
int make_map5(CUtensorMap* m, const void* ptr, const long long dims[5], const long long strides_elems[4],
              const int box[5], int elem_bytes = 2, bool output = false) {
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
    gs[i] = static_cast<cuuint64_t>(strides_elems[i]) * elem_bytes;
    if (gs[i] & 15) return HTRVT_ERR_ALIGN;
  }
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                   const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   output ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
int make_map_act(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int sw, int box_c, int box_w,
                 bool swizzle64 = false) {
  const long long dims[5] = {C, sw, W / sw, H, N};
  const long long st[4] = {C, static_cast<long long>(C) * sw, static_cast<long long>(C) * W,
                           static_cast<long long>(C) * W * H};
  const int box[5] = {box_c, 1, box_w, 1, 1};
  return make_map5(m, ptr, dims, st, box, 2, swizzle64);      // (the "output" flavour of make_map5 = SWIZZLE_64B)
}

// Output map of a kind-0 GEMM: (cols, w, h, n, 1) with element strides (s_w, s_h, s_n); box = 64 bytes x 32 rows
int make_map_out(CUtensorMap* m, void* ptr, int elem_bytes, long long cols, long long W, long long H, long long N,
                 long long s_w, long long s_h, long long s_n) {
  const long long dims[5] = {cols, W, H, N, 1};
  const long long big = s_n * (N > 0 ? N : 1);
  const long long st[4] = {s_w, H > 1 ? s_h : s_w * W, N > 1 ? s_n : (H > 1 ? s_h * H : s_w * W), big > 0 ? big : s_w * W};
  const int box[5] = {64 / elem_bytes, 32, 1, 1, 1};          // one epilogue warp's rows
  return make_map5(m, ptr, dims, st, box, elem_bytes, true);
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

// Programmatic dependent launch of the tap GEMMs (HTRVT_PDL=0 or htrvt_set_pdl(0): plain stream order)
int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("HTRVT_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

template <int BN, int KIND, bool B_MN, int CL, bool RE = false, int EPI = 0>
int launch_one(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmP& P, int total_tiles,
               cudaStream_t stream, const CUtensorMap* c2p = nullptr) {
  using Cfg = GemmCfg<BN, CL, RE, KIND == 4, EPI == kEpiGeluDual, EPI == kEpiBnBwd,
                      EPI == kEpiGeluBwd || (EPI == kEpiBnBwd && BN <= 192)>;
  auto kern = tapgemm_kernel<BN, KIND, B_MN, CL, RE, EPI>;
  const CUtensorMap& c2 = c2p ? *c2p : c;
  if (!HTRVT_ENSURE_SMEM(kern, Cfg::kSmemBytes)) return HTRVT_ERR_LAUNCH;
  int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  if (CL == 2) grid &= ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {                       // the kernel's prologue may overlap the tail of the launch in front
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (cudaLaunchKernelEx(&cfg, kern, a, b, c, c2, P) != cudaSuccess) return HTRVT_ERR_LAUNCH;
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

template <int KIND, bool B_MN, int EPI = 0>
int launch_bn(int bn, int cl, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmP& P,
              int total, cudaStream_t s) {
  if (cl == 2) {
    switch (bn) {
      case 128: return launch_one<128, KIND, B_MN, 2, false, EPI>(a, b, c, P, total, s);
      case 192: return launch_one<192, KIND, B_MN, 2, false, EPI>(a, b, c, P, total, s);
      case 256: return launch_one<256, KIND, B_MN, 2, false, EPI>(a, b, c, P, total, s);
    }
    return HTRVT_ERR_SHAPE;
  }
  switch (bn) {
    case 128: return launch_one<128, KIND, B_MN, 1, false, EPI>(a, b, c, P, total, s);
    case 192: return launch_one<192, KIND, B_MN, 1, false, EPI>(a, b, c, P, total, s);
    case 256: return launch_one<256, KIND, B_MN, 1, false, EPI>(a, b, c, P, total, s);
  }
  return HTRVT_ERR_SHAPE;
}

int dbg_env(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

// 3x3 stride-1 convolution (forward, or input gradient with K-major transposed weights) as CTA pairs with the
// horizontal taps sharing one staged activation window (GemmCfg RE).  HTRVT_NOREUSE=1 disables it.
bool use_window_reuse(int ks, int sw, int cl, int bn, int n_taps) {
  static const int off = dbg_env("HTRVT_NOREUSE");
  return !off && ks == 3 && sw == 1 && cl == 2 && n_taps >= 3 && (n_taps % 3) == 0 && (bn == 192 || bn == 256);
}
template <int EPI = 0>
int launch_reuse(int bn, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmP& P, int total,
                 cudaStream_t s) {
  if (bn == 192) return launch_one<192, 0, false, 2, true, EPI>(a, b, c, P, total, s);
  return launch_one<256, 0, false, 2, true, EPI>(a, b, c, P, total, s);
}

// CTA pairs (cta_group::2) need tile pairs (2p, 2p+1) on the same n-tile: tiles_m even; an MN-major B operand is
// split by 64-column swizzle atoms, so each CTA's half (BN / 2) must be a multiple of 64.
int pick_cluster(int tiles_m, int total_tiles, int bn, bool b_mn) {
  static const int off = dbg_env("HTRVT_NOPAIR");
  if (off || (tiles_m % 2) != 0 || total_tiles < 4) return 1;
  if (b_mn && ((bn / 2) % 64) != 0) return 1;
  return 2;
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

// developer knobs for tools/profile_gemm.py (unset in production): HTRVT_DBG_SPLITS=n forces the split-K factor,
// HTRVT_DBG_WGRAD_NOSTORE=1 skips the weight-gradient epilogue stores (main loop timing)
// Split-K factor for the weight-gradient GEMMs: minimise (waves x per-split tile time) + partial-sum traffic.
int choose_splits(int base_tiles, long long q_total, long long out_elems) {
  const int sms = num_sms();
  if (const int forced = dbg_env("HTRVT_DBG_SPLITS")) return forced < q_total ? forced : static_cast<int>(q_total);
  double best = 1e300;
  int best_s = 1;
  const double tile_us = static_cast<double>(q_total) * 0.23;          // ~0.23 us per 128 x BN x 64 k-chunk
  for (int s = 1; s <= 32 && s <= q_total; ++s) {
    const long long tiles = static_cast<long long>(base_tiles) * s;
    const double waves = static_cast<double>((tiles + sms - 1) / sms);
    double t = waves * tile_us / s;
    if (s > 1) t += 2.0 * s * out_elems * 4.0 / 5.0e6;                 // write + read of the partials (us @ 5 TB/s)
    if (t < best * 0.98) { best = t; best_s = s; }
  }
  return best_s;
}

}  // namespace

// Programmatic dependent launch of every tap-GEMM launch (default on; HTRVT_PDL=0 in the environment turns it off):
// 1 = the kernel's prologue may run under the tail of the launch in front of it, 0 = plain stream order.  Returns the
// previous setting.  Results do not depend on it.
extern "C" int htrvt_set_pdl(int on) {
  const int old = pdl_enabled() ? 1 : 0;
  g_pdl = on ? 1 : 0;
  return old;
}

// Y[M,N] = epilogue(alpha * X[M,K] W[N,K]^T): nn.Linear forward (both operands K-major).
// flags: EPI_BF16 (else fp32 out), EPI_BIAS, EPI_RELU, EPI_ACCUM (out += via TMA reduce-add),
// EPI_GELU (4096; bf16 out): out = gelu(alpha * X W^T + bias) - timm Mlp's fc1 + nn.GELU in one kernel; with `pre`
// (nullable, bf16 [M, N], row stride ldp) the derivative gelu'(alpha * X W^T + bias) is stored too (train mode: the
// backward multiplies by it - htrvt_gemm_nn's gelu_u, or htrvt_mul_bf16).
extern "C" int htrvt_gemm_tn(const void* X, long long ldx, const void* W, long long ldw, int M, int N, int K,
                             int flags, const float* bias, void* out, long long ldo, float alpha, void* pre,
                             long long ldp, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (N & 3)) return HTRVT_ERR_SHAPE;
  if ((flags & EPI_GELU) && !(flags & EPI_BF16)) return HTRVT_ERR_SHAPE;
  if ((flags & EPI_GELU) && ((N % 256) != 0 || (flags & (EPI_ACCUM | EPI_RELU)))) return HTRVT_ERR_SHAPE;
  if (pre && !(flags & EPI_GELU)) return HTRVT_ERR_SHAPE;
  const int bn = pick_bn(N);
  const int esz = (flags & EPI_BF16) ? 2 : 4;
  const int tiles_m = (M + kBM - 1) / kBM, tiles_n = (N + bn - 1) / bn;
  const int cl = pick_cluster(tiles_m, tiles_m * tiles_n, bn, false);
  CUtensorMap ta, tb, tc;
  {
    const long long dims[5] = {K, 1, M, 1, 1};
    const long long st[4] = {ldx, ldx, ldx * M, ldx * M};
    const int box[5] = {kBK, 1, kBM, 1, 1};
    int r = make_map5(&ta, X, dims, st, box);
    if (r) return r;
    r = make_map_matrix(&tb, W, N, K, ldw, kBK, bn / cl);
    if (r) return r;
    r = make_map_out(&tc, out, esz, N, M, 1, 1, ldo, ldo * M, ldo * M);
    if (r) return r;
  }
  GemmP P = {};
  P.kind = 0; P.Wo = M; P.Ho = 1; P.NB = 1;
  P.tiles_per_row = tiles_m; P.tiles_m = tiles_m; P.tiles_n = tiles_n;
  P.n_taps = 1; P.splits = 1; P.k_chunks = (K + kBK - 1) / kBK; P.a_sh = 1; P.b_tap_stride = 0;
  P.M_valid = M; P.N_valid = N;
  P.flags = (flags & (EPI_BF16 | EPI_BIAS | EPI_RELU | EPI_ACCUM | EPI_NOSTORE | EPI_GELU)) | (pre ? EPI_DUAL : 0);
  P.bias = bias; P.alpha = alpha;
  const int total = P.tiles_m * P.tiles_n;
  if (pre) {                                               // N % 256 == 0 => bn == 256
    CUtensorMap tc2;
    int r = make_map_out(&tc2, pre, 2, N, M, 1, 1, ldp, ldp * M, ldp * M);
    if (r) return r;
    if (cl == 2) return launch_one<256, 0, false, 2, false, kEpiGeluDual>(ta, tb, tc, P, total, stream, &tc2);
    return launch_one<256, 0, false, 1, false, kEpiGeluDual>(ta, tb, tc, P, total, stream, &tc2);
  }
  if (flags & EPI_GELU) {
    if (bn != 256) return HTRVT_ERR_SHAPE;                 // the fused activation is built for 256-column tiles
    if (cl == 2) return launch_one<256, 0, false, 2, false, kEpiGelu>(ta, tb, tc, P, total, stream);
    return launch_one<256, 0, false, 1, false, kEpiGelu>(ta, tb, tc, P, total, stream);
  }
  return launch_bn<0, false>(bn, cl, ta, tb, tc, P, total, stream);
}

// dX[M,N] = dY[M,K] W[K,N]: nn.Linear input gradient (B operand MN-major, no transpose copy).
// gelu_u (nullable, bf16 [M, N] contiguous): dX = (dY W) * gelu_u - the backward of timm Mlp's activation fused
// into fc2's input-gradient GEMM (gelu_u = gelu'(pre-activation), saved by fc1's forward epilogue through `pre`).
// colsum (nullable, with gelu_u): fp32 [N] += column sums of the stored dX - fc1's bias gradient from the same epilogue.
extern "C" int htrvt_gemm_nn(const void* dY, long long lddy, const void* W, long long ldw, int M, int N, int K,
                             int flags, void* out, long long ldo, float alpha, const void* gelu_u, float* colsum,
                             cudaStream_t stream) {
  if (colsum && !gelu_u) return HTRVT_ERR_SHAPE;
  if (M <= 0 || N <= 0 || K <= 0 || (N & 7)) return HTRVT_ERR_SHAPE;
  if (gelu_u && (!(flags & EPI_BF16) || (flags & EPI_ACCUM) || (N % 256) || (reinterpret_cast<uintptr_t>(gelu_u) & 15)))
    return HTRVT_ERR_SHAPE;
  const int bn = pick_bn(N);
  const int esz = (flags & EPI_BF16) ? 2 : 4;
  const int tiles_m = (M + kBM - 1) / kBM, tiles_n = (N + bn - 1) / bn;
  const int cl = pick_cluster(tiles_m, tiles_m * tiles_n, bn, true);
  CUtensorMap ta, tb, tc;
  {
    const long long dims[5] = {K, 1, M, 1, 1};
    const long long st[4] = {lddy, lddy, lddy * M, lddy * M};
    const int box[5] = {kBK, 1, kBM, 1, 1};
    int r = make_map5(&ta, dY, dims, st, box);
    if (r) return r;
    r = make_map_matrix(&tb, W, K, N, ldw, 64, 64);
    if (r) return r;
    r = make_map_out(&tc, out, esz, N, M, 1, 1, ldo, ldo * M, ldo * M);
    if (r) return r;
  }
  GemmP P = {};
  P.kind = 0; P.Wo = M; P.Ho = 1; P.NB = 1;
  P.tiles_per_row = tiles_m; P.tiles_m = tiles_m; P.tiles_n = tiles_n;
  P.n_taps = 1; P.splits = 1; P.k_chunks = (K + kBK - 1) / kBK; P.a_sh = 1; P.b_tap_stride = 0;
  P.M_valid = M; P.N_valid = N;
  P.flags = (flags & (EPI_BF16 | EPI_ACCUM | EPI_NOSTORE)) | (gelu_u ? EPI_GELU_BWD : 0) | (colsum ? EPI_COLSUM : 0);
  P.alpha = alpha; P.res = gelu_u; P.stats = colsum;
  if (gelu_u) {                                            // N % 256 == 0 => bn == 256
    if (cl == 2) return launch_one<256, 0, true, 2, false, kEpiGeluBwd>(ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
    return launch_one<256, 0, true, 1, false, kEpiGeluBwd>(ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
  }
  return launch_bn<0, true>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
}

extern "C" size_t htrvt_wgrad_workspace_bytes(int Cout, int Cin, int n_taps, int M_pixels) {
  (void)M_pixels;
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

namespace {
// output map of a kind-1 (wgrad) GEMM: fp32 partials [splits][Cout][taps][Cin] as (Cin, Cout, taps, splits, 1)
int make_map_wgrad_out(CUtensorMap* m, void* ws, int Cout, int taps, int Cin, int splits) {
  const long long dims[5] = {Cin, Cout, taps, splits, 1};
  const long long per = static_cast<long long>(Cout) * taps * Cin;
  const long long st[4] = {static_cast<long long>(taps) * Cin, Cin, per, per * splits};
  const int box[5] = {16, 32, 1, 1, 1};
  return make_map5(m, ws, dims, st, box, 4, true);
}
}  // namespace

// dW[Nout, Kin] (+)= dY[M, Nout]^T X[M, Kin]  (both operands MN-major; split-K over M, fp32 partials)
extern "C" int htrvt_linear_wgrad(const void* dY, long long lddy, const void* X, long long ldx, int M, int Nout,
                                  int Kin, float* grad, int accumulate, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream) {
  if (M <= 0 || Nout <= 0 || Kin <= 0 || (Kin & 7) || (Nout & 7)) return HTRVT_ERR_SHAPE;
  const int bn = pick_bn(Kin);
  const int cl = pick_cluster((Nout + kBM - 1) / kBM, 4, bn, true);
  CUtensorMap ta, tb, tc;
  int r;
  {
    const long long dimsA[5] = {Nout, 1, M, 1, 1};
    const long long stA[4] = {lddy, lddy, lddy * M, lddy * M};
    const int box[5] = {64, 1, 64, 1, 1};
    r = make_map5(&ta, dY, dimsA, stA, box);
    if (r) return r;
    const long long dimsB[5] = {Kin, 1, M, 1, 1};
    const long long stB[4] = {ldx, ldx, ldx * M, ldx * M};
    const int boxB[5] = {64, 1, 64, 1, 1};
    r = make_map5(&tb, X, dimsB, stB, boxB);
    if (r) return r;
  }
  GemmP P = {};
  P.kind = 1; P.Wo = M; P.Ho = 1; P.NB = 1; P.tiles_per_row = 1;
  P.tiles_m = (Nout + kBM - 1) / kBM; P.tiles_n = (Kin + bn - 1) / bn; P.n_taps = 1;
  P.k_chunks = (M + 63) / 64; P.a_sh = 1;
  P.splits = choose_splits(P.tiles_m * P.tiles_n, P.k_chunks, static_cast<long long>(Nout) * Kin);
  // every split-K slice reduce-adds (TMA cp.reduce.async.bulk, fp32) straight into the gradient: no partial-sum
  // buffers, no reduce kernel; the summation order across slices is not fixed (like cuBLAS / cuDNN split-K)
  P.M_valid = Nout; P.N_valid = Kin; P.flags = EPI_ACCUM; P.alpha = 1.f;
  (void)workspace; (void)workspace_bytes;
  if (!accumulate && cudaMemsetAsync(grad, 0, static_cast<size_t>(Nout) * Kin * sizeof(float), stream) != cudaSuccess)
    return HTRVT_ERR_LAUNCH;
  r = make_map_wgrad_out(&tc, grad, Nout, 1, Kin, 1);
  if (r) return r;
  return launch_bn<1, true>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n * P.splits, stream);
}

// ---------------------------------------------------------------------------------------------
// Stem convolutions (NHWC bf16 activations, weights [Cout][ks*ks][Cin] bf16, bias-free), Cin % 64 == 0
// ---------------------------------------------------------------------------------------------
extern "C" int htrvt_conv_fwd_stats_rows(int NB, int H, int W, int ks, int sh, int sw) {
  (void)NB; (void)H; (void)W; (void)ks; (void)sh; (void)sw;
  return 4 * num_sms();     // rows of the [ctas * 4 quadrants][2][Cout] partial-statistics buffer (zero-initialised):
                            // per-CTA rows + a fixed-order finalise keep the train-mode FORWARD bit-reproducible
}

// bias (nullable, fp32 [Cout]): added in the epilogue before the optional ReLU (eval mode: the folded BatchNorm shift);
// res (nullable, 16-bit [NB,Ho,Wo,Cout]): residual added there too (eval mode: the block's skip connection).
// flags: EPI_RELU (128), EPI_NOSTORE (256), EPI_F16 (1024): x, w, y and res are IEEE fp16 instead of bf16 - the
// storage format of the forward stem (the engine's choice; bf16 is kept for callers / tests that want it);
// 2048: y is fp32 (no res / stats), with EPI_ACCUM (16) the launch adds into y (split-operand fp32-parity mode)
extern "C" int htrvt_conv_fwd(const void* x, int NB, int H, int W, int Cin, const void* w, int Cout, int ks,
                              int sh, int sw, void* y, float* stats_partial, int flags, const float* bias,
                              const void* res, cudaStream_t stream) {
  const int pad = ks / 2;
  if ((Cin % 64) || (Cout & 7) || (W % sw) || (ks != 1 && ks != 3)) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cout);
  const int tiles_m_all = ((Wo + kBM - 1) / kBM) * Ho * NB;
  const int cl = pick_cluster(tiles_m_all, tiles_m_all * ((Cout + bn - 1) / bn), bn, false);
  const bool reuse = use_window_reuse(ks, sw, cl, bn, ks * ks);
  CUtensorMap ta, tb, tc;
  int r = make_map_act(&ta, x, NB, H, W, Cin, sw, kBK, reuse ? kWinRows : kBM);
  if (r) return r;
  r = make_map_matrix(&tb, w, Cout, static_cast<long long>(ks) * ks * Cin, static_cast<long long>(ks) * ks * Cin, kBK,
                      bn / cl);
  if (r) return r;
  const bool y_f32 = (flags & 2048) != 0;                 // fp32-parity mode: raw fp32 output, optionally accumulated
  r = make_map_out(&tc, y, y_f32 ? 4 : 2, Cout, Wo, Ho, NB, Cout, static_cast<long long>(Wo) * Cout,
                   static_cast<long long>(Ho) * Wo * Cout);
  if (r) return r;
  GemmP P = {};
  P.kind = 0; P.Wo = Wo; P.Ho = Ho; P.NB = NB;
  P.tiles_per_row = (Wo + kBM - 1) / kBM; P.tiles_m = P.tiles_per_row * Ho * NB; P.tiles_n = (Cout + bn - 1) / bn;
  fill_taps_conv(P.tap, ks, pad, sw, &P.n_taps);
  P.splits = 1; P.k_chunks = Cin / kBK; P.a_sh = sh; P.b_tap_stride = Cin;
  P.M_valid = 0; P.N_valid = Cout;
  if (bias && ((Cout & 3) || stats_partial)) return HTRVT_ERR_SHAPE;
  if (res && ((Cout & 31) || stats_partial || (reinterpret_cast<uintptr_t>(res) & 15))) return HTRVT_ERR_SHAPE;
  if (y_f32 && (res || stats_partial)) return HTRVT_ERR_SHAPE;
  P.flags = (y_f32 ? (flags & EPI_ACCUM) : EPI_BF16) | (flags & (EPI_RELU | EPI_NOSTORE | EPI_F16)) |
            (stats_partial ? EPI_STATS : 0) | (bias ? EPI_BIAS : 0) | (res ? EPI_RES : 0);
  P.a_f16 = P.b_f16 = (flags & EPI_F16) ? 1 : 0;        // forward stem tensors: x, w, y (and res) are fp16
  if (y_f32) P.flags &= ~EPI_F16;                       // (the operand format stays in a_f16 / b_f16)
  P.stats = stats_partial; P.bias = bias; P.res = res; P.alpha = 1.f;
  if ((P.flags & EPI_F16) != 0) {                          // fp16 output / residual: its own epilogue instantiation
    if (reuse) return launch_reuse<kEpiF16>(bn, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
    return launch_bn<0, false, kEpiF16>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
  }
  if (reuse) return launch_reuse(bn, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
  return launch_bn<0, false>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
}

// dx[NB,H,W,Cin] (= or +=) conv_transpose(dy[NB,Ho,Wo,Cout], w): one GEMM per output parity class.
// w_t (optional): the weights as bf16 [Cin][ks*ks][Cout] (htrvt_pack_weights transposed mode): B becomes a K-major
// operand, which lets BN = 192 tiles run as cta_group::2 pairs (an MN-major B cannot be split at 96 columns).
// dy, w, w_t and dx are bf16: gradients keep bf16's range, and tcgen05 kind::f16 wants both operands of an MMA in the
// SAME 16-bit format (a bf16 x fp16 instruction descriptor raised "illegal instruction" on B200) - so the backward
// GEMMs never see the forward pass's fp16 tensors (the engine keeps bf16 copies of what they need).
extern "C" int htrvt_conv_dgrad(const void* dy, int NB, int H, int W, int Cin, const void* w, const void* w_t, int Cout,
                                int ks, int sh, int sw, void* dx, int accumulate, cudaStream_t stream) {
  const int pad = ks / 2;
  if ((Cout % 8) || (Cin % 64) || (W % sw) || (H % sh) || (ks != 1 && ks != 3)) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cin);
  CUtensorMap ta, tb;
  int r = make_map_act(&ta, dy, NB, Ho, Wo, Cout, 1, kBK, kBM);
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
      const bool kmajor = w_t != nullptr && (Cout % 64) == 0;
      P.n_taps = n; P.splits = 1; P.k_chunks = (Cout + kBK - 1) / kBK; P.a_sh = 1; P.b_tap_stride = kmajor ? Cout : Cin;
      P.N_valid = Cin; P.flags = EPI_BF16 | (accumulate ? EPI_ACCUM : 0);
      P.alpha = 1.f;
      const int cl = pick_cluster(P.tiles_m, P.tiles_m * P.tiles_n, bn, !kmajor);
      // (a vertical stride only thins the kernel rows of a parity class: 3 or 6 taps, still whole rows of 3)
      const bool reuse = kmajor && use_window_reuse(ks, sw, cl, bn, n);
      if (reuse) {
        r = make_map_act(&ta, dy, NB, Ho, Wo, Cout, 1, kBK, kWinRows);
        if (r) return r;
      }
      if (kmajor)
        r = make_map_matrix(&tb, w_t, Cin, static_cast<long long>(ks) * ks * Cout, static_cast<long long>(ks) * ks * Cout,
                            kBK, bn / cl);
      else
        r = make_map_matrix(&tb, w, Cout, static_cast<long long>(ks) * ks * Cin, static_cast<long long>(ks) * ks * Cin,
                            64, 64);
      if (r) return r;
      CUtensorMap tc;                                     // parity class (ph, pw) of dx as a strided tensor
      r = make_map_out(&tc, static_cast<__nv_bfloat16*>(dx) + (static_cast<long long>(ph) * W + pw) * Cin, 2, Cin, Wq,
                       Hq, NB, static_cast<long long>(sw) * Cin, static_cast<long long>(sh) * W * Cin,
                       static_cast<long long>(H) * W * Cin);
      if (r) return r;
      r = reuse ? launch_reuse(bn, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream)
          : kmajor ? launch_bn<0, false>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream)
                   : launch_bn<0, true>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
      if (r) return r;
    }
  return HTRVT_OK;
}

// Input gradient of a 3x3 stride-1 convolution FUSED with the first half of the BatchNorm backward of the layer in front
// (conv -> BN -> ReLU -> [this conv]): the epilogue masks the gradient with that layer's ReLU bits (dx = g' = g * [y > 0])
// and accumulates sums[0][Cin] += sum g', sums[1][Cin] += sum g' * xhat (xhat from raw = the layer's raw fp16 conv output,
// mean / rstd its batch statistics) - exactly what htrvt_bn_bwd's reduction pass would compute from dx, so
// htrvt_bn_bwd_apply can follow directly.  Only the window-sharing CTA-pair kernel carries this epilogue: returns
// HTRVT_ERR_SHAPE for shapes it does not serve (the caller then runs htrvt_conv_dgrad + htrvt_bn_bwd).
extern "C" int htrvt_conv_dgrad_bn(const void* dy, int NB, int H, int W, int Cin, const void* w_t, int Cout, void* dx,
                                   const void* raw_f16, const void* relu_mask_bits, const float* mean, const float* rstd,
                                   float* sums, cudaStream_t stream) {
  const int ks = 3;
  if ((Cout % 64) || (Cin % 64) || !w_t || !raw_f16 || !relu_mask_bits || !mean || !rstd || !sums) return HTRVT_ERR_SHAPE;
  const int bn = pick_bn(Cin);
  GemmP P = {};
  int n = 0;
  for (int kh = 0; kh < ks; ++kh)
    for (int kw = 0; kw < ks; ++kw) {
      P.tap.dh[n] = static_cast<int8_t>(1 - kh);
      P.tap.dw[n] = static_cast<int8_t>(1 - kw);
      P.tap.pw[n] = 0;
      P.tap.widx[n] = static_cast<int8_t>(kh * ks + kw);
      ++n;
    }
  P.kind = 0; P.Wo = W; P.Ho = H; P.NB = NB;
  P.tiles_per_row = (W + kBM - 1) / kBM; P.tiles_m = P.tiles_per_row * H * NB; P.tiles_n = (Cin + bn - 1) / bn;
  P.n_taps = n; P.splits = 1; P.k_chunks = (Cout + kBK - 1) / kBK; P.a_sh = 1; P.b_tap_stride = Cout;
  P.N_valid = Cin; P.flags = EPI_BF16 | EPI_BN_BWD;
  P.alpha = 1.f;
  P.res = raw_f16; P.mask = static_cast<const uint8_t*>(relu_mask_bits); P.bn_mean = mean; P.bn_rstd = rstd; P.stats = sums;
  const int cl = pick_cluster(P.tiles_m, P.tiles_m * P.tiles_n, bn, false);
  if (!use_window_reuse(ks, 1, cl, bn, n) || (Cin % 32)) return HTRVT_ERR_SHAPE;
  CUtensorMap ta, tb, tc;
  int r = make_map_act(&ta, dy, NB, H, W, Cout, 1, kBK, kWinRows);
  if (r) return r;
  r = make_map_matrix(&tb, w_t, Cin, static_cast<long long>(ks) * ks * Cout, static_cast<long long>(ks) * ks * Cout, kBK,
                      bn / cl);
  if (r) return r;
  r = make_map_out(&tc, dx, 2, Cin, W, H, NB, static_cast<long long>(Cin), static_cast<long long>(W) * Cin,
                   static_cast<long long>(H) * W * Cin);
  if (r) return r;
  return launch_reuse<kEpiBnBwd>(bn, ta, tb, tc, P, P.tiles_m * P.tiles_n, stream);
}

// dw (+)= sum_pixels dy^T x_shifted.  Two output modes:
//   grad_oihw    : fp32 OIHW (the nn.Conv2d parameter layout) through split-K partials in `workspace` and a
//                  deterministic reduce / permute kernel;
//   grad_tapmajor: fp32 [Cout][taps][Cin], every split-K slice reduce-adds (TMA) into it - no partials, no reduce
//                  kernel; htrvt_unpack_conv_grads permutes all tap-major gradients of a step into OIHW at once.
// dy_t (optional): the same gradient stored [NB][Ho][Cout][Wo] (htrvt_transpose_px) - selects the K-major-A kernel.
static int conv_wgrad_impl(const void* dy, const void* dy_t, const void* x, int NB, int H, int W, int Cin, int Cout,
                           int ks, int sh, int sw, float* grad_oihw, int accumulate, float* grad_tapmajor,
                           void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int pad = ks / 2;
  if ((Cout % 8) || (Cin % 8) || (W % sw) || (ks != 1 && ks != 3)) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cin);
  CUtensorMap ta, tb, tc;
  const bool a_kmajor = dy_t != nullptr && (Wo % 8) == 0;
  int r;
  if (a_kmajor) {
    const long long dims[5] = {Wo, Cout, Ho, NB, 1};
    const long long st[4] = {Wo, static_cast<long long>(Cout) * Wo, static_cast<long long>(Ho) * Cout * Wo,
                             static_cast<long long>(NB) * Ho * Cout * Wo};
    const int box[5] = {64, kBM, 1, 1, 1};
    r = make_map5(&ta, dy_t, dims, st, box);
  } else {
    if (!dy) return HTRVT_ERR_SHAPE;
    r = make_map_act(&ta, dy, NB, Ho, Wo, Cout, 1, 64, 64);
  }
  if (r) return r;
  const int cl = pick_cluster((Cout + kBM - 1) / kBM, 4, bn, true);
  r = make_map_act(&tb, x, NB, H, W, Cin, sw, 64, 64);
  if (r) return r;
  GemmP P = {};
  P.kind = 1; P.Wo = Wo; P.Ho = Ho; P.NB = NB; P.tiles_per_row = 1;
  P.tiles_m = (Cout + kBM - 1) / kBM; P.tiles_n = (Cin + bn - 1) / bn;
  fill_taps_conv(P.tap, ks, pad, sw, &P.n_taps);
  P.k_chunks = (Wo + 63) / 64; P.a_sh = sh;
  const long long Q = static_cast<long long>(NB) * Ho * P.k_chunks;
  const long long per = static_cast<long long>(Cout) * P.n_taps * Cin;
  P.splits = choose_splits(P.tiles_m * P.tiles_n * P.n_taps, Q, grad_tapmajor ? per / 2 : per);
  P.M_valid = Cout; P.N_valid = Cin; P.alpha = 1.f;
  P.flags = (grad_tapmajor ? EPI_ACCUM : 0) | (dbg_env("HTRVT_DBG_WGRAD_NOSTORE") ? EPI_NOSTORE : 0);
  if (grad_tapmajor) {
    r = make_map_wgrad_out(&tc, grad_tapmajor, Cout, P.n_taps, Cin, 1);
  } else {
    if (!workspace || workspace_bytes < static_cast<size_t>(P.splits) * per * sizeof(float)) return HTRVT_ERR_WORKSPACE;
    r = make_map_wgrad_out(&tc, workspace, Cout, P.n_taps, Cin, P.splits);
  }
  if (r) return r;
  r = a_kmajor ? launch_bn<2, true>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n * P.n_taps * P.splits, stream)
               : launch_bn<1, true>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n * P.n_taps * P.splits, stream);
  if (r || grad_tapmajor) return r;
  const int blocks = static_cast<int>((per + 255) / 256 < 2048 ? (per + 255) / 256 : 2048);
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), P.splits, Cout, P.n_taps, Cin,
                                                  grad_oihw, accumulate, 1);
  HTRVT_LAUNCH_CHECK();
  return HTRVT_OK;
}

// (all weight-gradient entry points: dy and x are bf16 - see htrvt_conv_dgrad)
extern "C" int htrvt_conv_wgrad(const void* dy, const void* dy_t, const void* x, int NB, int H, int W, int Cin,
                                int Cout, int ks, int sh, int sw, float* grad_oihw, int accumulate, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream) {
  if (!grad_oihw) return HTRVT_ERR_SHAPE;
  return conv_wgrad_impl(dy, dy_t, x, NB, H, W, Cin, Cout, ks, sh, sw, grad_oihw, accumulate, nullptr, workspace,
                         workspace_bytes, stream);
}

extern "C" int htrvt_conv_wgrad_acc(const void* dy, const void* dy_t, const void* x, int NB, int H, int W, int Cin,
                                    int Cout, int ks, int sh, int sw, float* grad_tapmajor, cudaStream_t stream) {
  if (!grad_tapmajor) return HTRVT_ERR_SHAPE;
  return conv_wgrad_impl(dy, dy_t, x, NB, H, W, Cin, Cout, ks, sh, sw, nullptr, 1, grad_tapmajor, nullptr, 0, stream);
}

// Transposed weight gradient: grad_tco[taps][Cin][Cout] (fp32) += x_shifted^T dy.  The GEMM's M axis is the flattened
// (tap, input channel) index in 64-channel atoms - each half tile loads its own tap window - and N = Cout, so Cout =
// 192 / 384 waste no MMA rows and every shape runs as cta_group::2 pairs (a 96-column half of dY is staged as three
// 32-column SWIZZLE_64B atoms).  Split-K slices reduce-add in place.
extern "C" int htrvt_conv_wgrad_acc_t(const void* dy, const void* x, int NB, int H, int W, int Cin, int Cout, int ks,
                                      int sh, int sw, float* grad_tco, cudaStream_t stream) {
  const int pad = ks / 2;
  if (!dy || !x || !grad_tco || (Cout % 8) || (Cin % 64) || (W % sw) || (ks != 1 && ks != 3)) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cout);
  GemmP P = {};
  fill_taps_conv(P.tap, ks, pad, sw, &P.a_taps);
  P.a_atoms_per_tap = Cin / 64;
  const int atoms = P.a_taps * P.a_atoms_per_tap;
  int tiles_m = (atoms + 1) / 2;
  // pairs whose halves are 96 columns (three SWIZZLE_64B atoms) are correct but measured ~5 % SLOWER than the single-CTA
  // SWIZZLE_128B kernel on the 192/384-channel layers: opt-in (HTRVT_SW64=1), halves of 64 / 128 columns pair up
  static const int nopair = dbg_env("HTRVT_NOPAIR"), sw64_ok = dbg_env("HTRVT_SW64");
  const int cl = (!nopair && atoms >= 3 && (((bn / 2) % 64) == 0 || (sw64_ok && ((bn / 2) % 32) == 0))) ? 2 : 1;
  if (cl == 2) tiles_m = (tiles_m + 1) & ~1;
  const bool sw64 = cl == 2 && ((bn / 2) % 64) != 0;
  CUtensorMap ta, tb, tc;
  int r = make_map_act(&ta, x, NB, H, W, Cin, sw, 64, 64);
  if (r) return r;
  r = make_map_act(&tb, dy, NB, Ho, Wo, Cout, 1, sw64 ? 32 : 64, 64, sw64);
  if (r) return r;
  {
    const long long rows = static_cast<long long>(P.a_taps) * Cin;
    const long long dims[5] = {Cout, rows, 1, 1, 1};
    const long long big = rows * Cout;
    const long long st[4] = {Cout, big, big, big};
    const int box[5] = {16, 32, 1, 1, 1};
    r = make_map5(&tc, grad_tco, dims, st, box, 4, true);
    if (r) return r;
  }
  P.kind = 3; P.Wo = Wo; P.Ho = Ho; P.NB = NB; P.tiles_per_row = 1;
  P.tiles_m = tiles_m; P.tiles_n = (Cout + bn - 1) / bn; P.n_taps = 1;
  P.k_chunks = (Wo + 63) / 64; P.a_sh = sh;
  const long long Q = static_cast<long long>(NB) * Ho * P.k_chunks;
  const long long per = static_cast<long long>(Cout) * P.a_taps * Cin;
  P.splits = choose_splits(P.tiles_m * P.tiles_n, Q, per / 2);
  P.M_valid = P.a_taps * Cin; P.N_valid = Cout; P.alpha = 1.f;
  P.flags = EPI_ACCUM | (dbg_env("HTRVT_DBG_WGRAD_NOSTORE") ? EPI_NOSTORE : 0);
  return launch_bn<3, true>(bn, cl, ta, tb, tc, P, P.tiles_m * P.tiles_n * P.splits, stream);
}

// Weight gradient of a 3x3 convolution with horizontal stride 1, two live accumulators and shared x windows (KIND 4).
// grad_atoms: fp32 [3 kh][Cin / 64][3 kw][64][Cout] (+=) - the GEMM's row order; htrvt_unpack_conv_grads takes it with
// taps[i] = -1009.  Bytes staged per flop are ~half of htrvt_conv_wgrad_acc_t's (dY tile shared by 256 rows, the three
// horizontal taps read from one window).
extern "C" int htrvt_conv_wgrad_acc_w(const void* dy, const void* x, int NB, int H, int W, int Cin, int Cout, int sh,
                                      float* grad_atoms, cudaStream_t stream) {
  const int ks = 3, pad = 1, sw = 1;
  if (!dy || !x || !grad_atoms || (Cout % 8) || (Cin % 64)) return HTRVT_ERR_SHAPE;
  const int Ho = (H + 2 * pad - ks) / sh + 1, Wo = (W + 2 * pad - ks) / sw + 1;
  const int bn = pick_bn(Cout);
  GemmP P = {};
  P.a_taps = 9;
  P.a_atoms_per_tap = Cin / 64;
  const int n_atoms = 9 * P.a_atoms_per_tap;
  CUtensorMap ta, tb, tc;
  int r = make_map_act(&ta, x, NB, H, W, Cin, 1, 64, kWin4Rows);
  if (r) return r;
  r = make_map_act(&tb, dy, NB, Ho, Wo, Cout, 1, 64, 64);
  if (r) return r;
  {
    const long long rows = static_cast<long long>(n_atoms) * 64;
    const long long dims[5] = {Cout, rows, 1, 1, 1};
    const long long big = rows * Cout;
    const long long st[4] = {Cout, big, big, big};
    const int box[5] = {16, 32, 1, 1, 1};
    r = make_map5(&tc, grad_atoms, dims, st, box, 4, true);
    if (r) return r;
  }
  P.kind = 4; P.Wo = Wo; P.Ho = Ho; P.NB = NB; P.tiles_per_row = 1;
  P.tiles_m = (n_atoms + 3) / 4; P.tiles_n = (Cout + bn - 1) / bn; P.n_taps = 1;
  P.k_chunks = (Wo + 63) / 64; P.a_sh = sh;
  const long long Q = static_cast<long long>(NB) * Ho * P.k_chunks;
  const long long per = static_cast<long long>(Cout) * 9 * Cin;
  P.splits = choose_splits(P.tiles_m * P.tiles_n, 2 * Q, per / 2);          // a unit's k-chunk is two tiles' worth of MMA
  P.M_valid = n_atoms * 64; P.N_valid = Cout; P.alpha = 1.f;
  P.flags = EPI_ACCUM | (dbg_env("HTRVT_DBG_WGRAD_NOSTORE") ? EPI_NOSTORE : 0);
  const int total = P.tiles_m * P.tiles_n * P.splits;
  switch (bn) {
    case 128: return launch_one<128, 4, true, 1>(ta, tb, tc, P, total, stream);
    case 192: return launch_one<192, 4, true, 1>(ta, tb, tc, P, total, stream);
    case 256: return launch_one<256, 4, true, 1>(ta, tb, tc, P, total, stream);
  }
  return HTRVT_ERR_SHAPE;
}

// dst_oihw[i] += permute(src[i]) for n <= 64 conv weights per launch: taps[i] > 0: src [Cout][taps][Cin];
// taps[i] < 0: src [|taps|][Cin][Cout] (the transposed weight-gradient GEMM)
namespace htrvt {
constexpr int kMaxUnpack = 64;
struct UnpackTable {
  const float* src[kMaxUnpack];
  float* dst[kMaxUnpack];
  long long numel[kMaxUnpack];
  int cin[kMaxUnpack];
  int taps[kMaxUnpack];
};
__global__ void __launch_bounds__(256) unpack_conv_grads_kernel(const __grid_constant__ UnpackTable T) {
  extern __shared__ float up_stage[];                   // one output channel's [taps][Cin] block
  const int t = blockIdx.y;
  const float* __restrict__ src = T.src[t];
  float* __restrict__ dst = T.dst[t];
  const long long n = T.numel[t];
  const int Cin = T.cin[t], taps = T.taps[t];
  if (taps < 0) {
    // src [tp][Cin][Cout] (htrvt_conv_wgrad_acc_t) -> dst [Cout][Cin][tp]: tiles of 8 co x 64 ci x tp taps through
    // smem - 32-byte read segments along co, contiguous RMW of 64 * tp floats per output channel
    const bool atoms = taps <= -1000;                 // [kh][Cin/64][kw][64][Cout] (htrvt_conv_wgrad_acc_w)
    const int tp = atoms ? -taps - 1000 : -taps, per = Cin * tp;
    const int Cout = static_cast<int>(n / per);
    const int tiles_ci = Cin / 64, tiles = (Cout / 8) * tiles_ci, cnt = tp * 64 * 8;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int co0 = (tile / tiles_ci) * 8, ci0 = (tile % tiles_ci) * 64;
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int co = i & 7, ci = (i >> 3) & 63, tap = i >> 9;
        const long long row = atoms ? ((static_cast<long long>(tap / 3) * (Cin / 64) + ci0 / 64) * 3 + tap % 3) * 64 + ci
                                    : static_cast<long long>(tap) * Cin + ci0 + ci;
        up_stage[(tap * 64 + ci) * 9 + co] = src[row * Cout + co0 + co];
      }
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int tap = i % tp, r = i / tp, ci = r & 63, co = r >> 6;
        dst[static_cast<long long>(co0 + co) * per + (ci0 + ci) * tp + tap] += up_stage[(tap * 64 + ci) * 9 + co];
      }
    }
    return;
  }
  const int per = Cin * taps;
  const int Cout = static_cast<int>(n / per);
  // per output channel a [taps][Cin] -> [Cin][taps] transpose through smem: contiguous read, contiguous RMW
  for (int co = blockIdx.x; co < Cout; co += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < per; i += blockDim.x) up_stage[i] = src[static_cast<long long>(co) * per + i];
    __syncthreads();
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
      const int ci = i / taps, tap = i - ci * taps;
      dst[static_cast<long long>(co) * per + i] += up_stage[tap * Cin + ci];
    }
  }
}
}  // namespace htrvt

extern "C" int htrvt_unpack_conv_grads(int n, const void* const* src, void* const* dst, const long long* numel,
                                       const int* cin, const int* taps, cudaStream_t stream) {
  for (int base = 0; base < n; base += kMaxUnpack) {
    UnpackTable T = {};
    const int cnt = n - base < kMaxUnpack ? n - base : kMaxUnpack;
    for (int i = 0; i < cnt; ++i) {
      T.src[i] = static_cast<const float*>(src[base + i]);
      T.dst[i] = static_cast<float*>(dst[base + i]);
      T.numel[i] = numel[base + i];
      T.cin[i] = cin[base + i];
      T.taps[i] = taps[base + i];
      if (T.cin[i] <= 0 || T.taps[i] == 0 || (T.taps[i] < 0 && (T.cin[i] % 64))) return HTRVT_ERR_SHAPE;
    }
    int smem = 0;
    for (int i = 0; i < cnt; ++i) {
      const int need = T.taps[i] > 0 ? T.cin[i] * T.taps[i] * 4 : (T.taps[i] <= -1000 ? 9 : -T.taps[i]) * 64 * 9 * 4;
      if (need > smem) smem = need;
    }
    if (smem > 48 * 1024) return HTRVT_ERR_SHAPE;
    dim3 grid(4 * 148, cnt);                          // several 256-thread blocks per SM: the pass is pure HBM streaming
    unpack_conv_grads_kernel<<<grid, 256, smem, stream>>>(T);
    HTRVT_LAUNCH_CHECK();
  }
  return HTRVT_OK;
}
