// "Row-GEMM + shift-add" convolution on the sm_100a tensor cores, for convs with FEW output channels: the second-layer
// 7x7 head convs of KGnet (KGnet.py:161-209, Cout = 5 / 10 / 40 per head, three heads per scale fused in one launch).
//
// Why: with the plain implicit GEMM (tc_conv.cu) such a layer issues one M=128 x N=16 x K=16 tcgen05.mma per filter tap
// and k-step.  Each of those reads its 4 KiB A slab from shared memory, which costs ~64-96 cycles however small N is,
// so a 7x7 conv re-reads every activation 49 times and runs at < 10 % of the tensor pipe (measured: 13 ms for the three
// c0 heads).  Here the S horizontal taps are folded into N instead:
//
//     Z[p, (s, co)] = sum_{r, c} X[y + r - pad, p, c] * Wt[r][s][c][co]         (one GEMM, N = S * Cout, K = R * Cin)
//     out[y, x, co] = sum_s Z[x + s - pad, (s, co)]                              (shift-add, done in the epilogue)
//
// so every activation tile is read by the tensor core once per filter ROW (7x instead of 49x) and every MMA has
// N = 64..160.  The accumulator Z of a 128-pixel tile lives in TMEM (<= 512 fp32 columns: all three heads at once).
// The epilogue warps pull Z tap by tap out of TMEM and add it, shifted by (pad - s) pixels, into an fp32 output row
// buffer in shared memory; tiles of one image row are walked left to right by the same CTA, so the 2*pad pixels that
// straddle a tile edge are simply carried in that buffer (ring-indexed) to the next tile.  Finished pixels get
// 1/scale, bias, (sigmoid) and go to the fp32 NCHW head outputs.
//
//   roles   warp 0: TMA producer (A tile 128 px x 64 ch + the unit's weight slab per slot), warp 1: MMA issuer,
//           warps 2-9: epilogue (two per TMEM lane quadrant, alternate 16-channel batches).  One smem ring of NS uniform
//           slots, full/empty mbarriers; one or two TMEM accumulator sets (tmem_full / tmem_empty hand-off).
//   64-channel 3x3 convs (NHWC out): 1 / 2 / 3 split-fp16 passes, weights resident in smem when they fit, shift-add in
//           registers by warp shuffle for whole-row tiles, outputs staged in smem and written by TMA stores.
#include "tc_shift.cuh"
#include "tc_conv.cuh"
#include "tc_ptx.cuh"

#include <cmath>
#include <cstring>
#include <type_traits>
#include <vector>

namespace kg {

constexpr int SH_THREADS = 320;              // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue (two per TMEM lane quadrant)
// The 7x7 head kernel runs FOUR epilogue warps per TMEM lane quadrant (576 threads): its 464 accumulator columns fit TMEM once, so
// the tensor core idles while the epilogue holds the accumulators (ncu: 40 % tensor pipe, epilogue warps 23 % of their samples waiting
// for the MMA and vice versa); the shift-add phase is latency-bound, more warps shorten it.
__host__ __device__ constexpr int sh_epi_warps(int outmode, int taps) { return (outmode == 0 && taps == 7) ? 16 : 8; }
constexpr int SH_MAX_UNITS = 6, SH_MAX_TAPS = 8, SH_MAX_CHUNKS = 16;
constexpr int SH_BK = 64;
constexpr int SH_A_TILE = 128 * SH_BK * 2;   // 16 KiB
constexpr int SH_MAX_N = 192;                // rows of one weight slab (N of one MMA)
constexpr int SH_MAX_SMEM = 227 * 1024;

struct ShGroup { const float* bias; float* out32; int in_coff; float inv_scale; int sigmoid; };
struct ShParams {
  CUtensorMap a_map[2];                       // activation planes (hi, lo)
  CUtensorMap w_map[SH_MAX_UNITS][2];         // weight slab of each unit, planes (hi, lo)
  CUtensorMap o_map[2];                       // NHWC output planes, boxes of 64 channels x 128 pixels (TMA-store epilogue)
  unsigned stage_bytes;                       // 0, or 32 KiB of output staging (hi 16 KiB + lo 16 KiB) in front of everything else
  ShGroup grp[SH_MAX_GROUPS];
  __half* out_hi; __half* out_lo;             // NHWC output mode (conv 0)
  const uint8_t* mask;
  int relu;
  int N, H, W, pad, kchunks;
  int BW, BH, tiles_x, rows_y, num_work;
  int row_mode, RB;
  int NS;
  unsigned slot_bytes, w_slab, tmem_cols;     // w_slab: bytes reserved per weight plane inside a slot
  unsigned wres_bytes;                        // resident-weights mode: bytes of the weight region in front of the ring
};

// Compile-time geometry of one fused launch: up to three convs with O0 / O1 / O2 output channels, TAPS x TAPS filters.
// A conv's S taps sit side by side along N with a stride of align8(n_out) columns; a "unit" is a run of taps that fits one
// MMA (N <= SH_MAX_N); TMEM columns == rows of the packed weight tensor, allocated unit after unit.
template <int O0, int O1, int O2, int TAPS>
struct ShCfg {
  static constexpr int NG = (O0 > 0) + (O1 > 0) + (O2 > 0);
  static constexpr int n_out(int g) { return g == 0 ? O0 : (g == 1 ? O1 : O2); }
  static constexpr int stride(int g) { return (n_out(g) + 7) / 8 * 8; }
  static constexpr int per(int g) { return SH_MAX_N / stride(g) < TAPS ? SH_MAX_N / stride(g) : TAPS; }   // taps per unit
  static constexpr int units_of(int g) { return (TAPS + per(g) - 1) / per(g); }
  static constexpr int unit_taps(int g, int k) { return TAPS - k * per(g) < per(g) ? TAPS - k * per(g) : per(g); }
  static constexpr int unit_n(int g, int k) { return (unit_taps(g, k) * stride(g) + 15) / 16 * 16; }
  static constexpr int gcols(int g) { int c = 0; for (int k = 0; k < units_of(g); ++k) c += unit_n(g, k); return c; }
  static constexpr int col0(int g) { int c = 0; for (int i = 0; i < g; ++i) c += gcols(i); return c; }
  static constexpr int bufcol(int g) { int c = 0; for (int i = 0; i < g; ++i) c += n_out(i); return c; }
  static constexpr int nchunks(int g) { return (n_out(g) + 7) / 8; }                     // 8-column TMEM loads per tap
  static constexpr int chunk0(int g) { int c = 0; for (int i = 0; i < g; ++i) c += nchunks(i); return c; }
  // epilogue work list: batches of up to 16 channels of one conv; batch b is handled by epilogue half (b & 1)
  static constexpr int nbatch(int g) { return (n_out(g) + 15) / 16; }
  static constexpr int NBT = (O0 > 0 ? nbatch(0) : 0) + (O1 > 0 ? nbatch(1) : 0) + (O2 > 0 ? nbatch(2) : 0);
  static constexpr int b_group(int b) { int g = 0; while (b >= nbatch(g)) { b -= nbatch(g); ++g; } return g; }
  static constexpr int b_lo(int b) { int g = 0; while (b >= nbatch(g)) { b -= nbatch(g); ++g; } return b * 16; }
  static constexpr int b_n(int b) { return n_out(b_group(b)) - b_lo(b) < 16 ? n_out(b_group(b)) - b_lo(b) : 16; }
  static constexpr int last_batch_of_half(int h) { int l = -1; for (int b = 0; b < NBT; ++b) if ((b & 1) == h) l = b; return l; }
  static constexpr int last_batch_of_part(int q, int nparts) { int l = -1; for (int b = 0; b < NBT; ++b) if (b % nparts == q) l = b; return l; }
  static constexpr int NU = (O0 > 0 ? units_of(0) : 0) + (O1 > 0 ? units_of(1) : 0) + (O2 > 0 ? units_of(2) : 0);
  static constexpr int COLS = col0(NG);
  static constexpr int NOUT = bufcol(NG);
  static constexpr int RS = NOUT | 1;         // odd row stride of the output buffer: conflict-free across pixels
  // unit u -> (group, index inside the group)
  static constexpr int u_group(int u) { int g = 0; while (u >= units_of(g)) { u -= units_of(g); ++g; } return g; }
  static constexpr int u_index(int u) { int g = 0; while (u >= units_of(g)) { u -= units_of(g); ++g; } return u; }
  static constexpr int u_n(int u) { return unit_n(u_group(u), u_index(u)); }
  static constexpr int u_col(int u) { int c = col0(u_group(u)); for (int k = 0; k < u_index(u); ++k) c += unit_n(u_group(u), k); return c; }
  static constexpr bool u_share(int u) { return u_index(u) == 1; }                      // 2nd unit of a conv reuses the 1st one's A tile
  static constexpr bool u_keep(int u) { return u_index(u) == 0 && units_of(u_group(u)) > 1; }
  // the accumulator column of (group, tap) must be linear in the tap for the epilogue: every non-final unit is unpadded
  static constexpr bool linear() {
    for (int g = 0; g < NG; ++g)
      for (int k = 0; k + 1 < units_of(g); ++k)
        if (unit_n(g, k) != unit_taps(g, k) * stride(g)) return false;
    return true;
  }
  static_assert(COLS <= 512 && NU <= SH_MAX_UNITS && TAPS <= SH_MAX_TAPS, "configuration does not fit TMEM / the unit table");
};

// compile-time loop: f(std::integral_constant<int, I>) for I in [I0, N) -- the geometry functions of ShCfg are then evaluated
// by the compiler (constexpr locals), never at run time
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}

template <int NTHREADS = 256>
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// OUTMODE 0: fp32 NCHW per conv (+sigmoid); OUTMODE 1: conv 0 -> split-fp16 NHWC (+ReLU, +mask)
// PASSES 1: fp16 x fp16; 2: (hi + lo activations) x hi weights; 3: split-fp16 on both sides (hi*hi + lo*hi + hi*lo).
// WRES: the whole weight tensor stays resident in shared memory (loaded once per CTA); the ring then carries activations only.
template <int O0, int O1, int O2, int TAPS, int PASSES, int OUTMODE, bool WRES>
__global__ void __launch_bounds__(64 + 32 * sh_epi_warps(OUTMODE, TAPS), 1) tc_shift_kernel(const __grid_constant__ ShParams p) {
  using Cfg = ShCfg<O0, O1, O2, TAPS>;
  constexpr int EW = sh_epi_warps(OUTMODE, TAPS);            // epilogue warps
  constexpr int NPART = EW / 4;                              // epilogue warps per TMEM lane quadrant: batch b belongs to part b % NPART
  static_assert(Cfg::linear(), "tap columns must be linear");
  static_assert(OUTMODE == 0 || (Cfg::NG == 1 && O0 % 8 == 0), "NHWC output: one conv, Cout a multiple of 8");
  constexpr int NU = Cfg::NU, NG = Cfg::NG, RS = Cfg::RS, COLS = Cfg::COLS;
  constexpr int NPA = PASSES >= 2 ? 2 : 1;                   // activation planes
  constexpr int NPW = PASSES == 3 ? 2 : 1;                   // weight planes
  static_assert(!WRES || NU == 1, "resident weights: single-unit configurations only");
  constexpr int ACC = 2 * COLS <= 512 ? 2 : 1;               // TMEM accumulator sets (2: epilogue of tile i overlaps the MMAs of tile i+1)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t smem0 = (smem_base + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage = smem0;                              // output staging of the register-shuffle epilogue (p.stage_bytes, may be 0)
  const uint32_t wres = smem0 + p.stage_bytes;               // resident weights: [r][chunk][plane] slabs of w_slab bytes
  const uint32_t ring = wres + (WRES ? p.wres_bytes : 0u);
  const uint32_t bars = ring + (uint32_t)p.NS * p.slot_bytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (p.NS + s); };
  const uint32_t tbar = bars + 16u * p.NS;
  auto tmem_full = [&](int a) { return tbar + 8u * a; };
  auto tmem_empty = [&](int a) { return tbar + 16u + 8u * a; };
  const uint32_t wres_bar = tbar + 32u;
  const uint32_t tmem_slot = tbar + 40u;
  float* buf = reinterpret_cast<float*>(smem_raw + (smem0 - smem_base) + p.stage_bytes + (WRES ? p.wres_bytes : 0u) + (size_t)p.NS * p.slot_bytes +
                                        16u * p.NS + 64u);
  const uint32_t w_off = (uint32_t)NPA * SH_A_TILE;          // weight planes follow the activation planes inside a slot
  float* s_bias = buf + p.RB * RS;                           // biases of all convs (a global load per output pixel stalls the emission)
  float* s_edge = s_bias + ((Cfg::NOUT + 3) & ~3);           // register-shuffle epilogue: warp / tile edge exchange (832 floats)

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int pl = 0; pl < NPA; ++pl) prefetch_tmap(&p.a_map[pl]);
#pragma unroll
    for (int pl = 0; pl < NPW; ++pl)
#pragma unroll
      for (int u = 0; u < NU; ++u) prefetch_tmap(&p.w_map[u][pl]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.NS; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), EW); }
    mbar_init(wres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int e = threadIdx.x - 64; e < p.RB * RS; e += 32 * EW) buf[e] = 0.f;
    static_for<0, NG>([&](auto Gx) __attribute__((always_inline)) {
      constexpr int g = decltype(Gx)::value;
      for (int e = threadIdx.x - 64; e < Cfg::n_out(g); e += 32 * EW) s_bias[Cfg::bufcol(g) + e] = p.grp[g].bias[e];
    });
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===== TMA producer =====
    const bool leader = elect_one();
    int slot = 0; uint32_t ph = 0;
    if (WRES && leader) {                                    // the whole weight tensor, once
      mbar_expect_tx(wres_bar, (uint32_t)(TAPS * p.kchunks * NPW) * (uint32_t)Cfg::u_n(0) * 128u);
      for (int r = 0; r < TAPS; ++r)
        for (int ch = 0; ch < p.kchunks; ++ch)
#pragma unroll
          for (int pl = 0; pl < NPW; ++pl)
            tma_load_3d(wres + (uint32_t)((r * p.kchunks + ch) * NPW + pl) * p.w_slab, &p.w_map[0][pl], ch * SH_BK, 0, r, wres_bar);
    }
    __syncwarp();
    for (int work = blockIdx.x; work < p.num_work; work += gridDim.x) {
      const int n = work / p.rows_y, y0 = (work - n * p.rows_y) * p.BH;
      for (int tx = 0; tx < p.tiles_x; ++tx)
        for (int r = 0; r < TAPS; ++r)
          for (int ch = 0; ch < p.kchunks; ++ch) {
            static_for<0, NU>([&](auto U) __attribute__((always_inline)) {
              constexpr int u = decltype(U)::value;
              constexpr int g = Cfg::u_group(u), wrow = Cfg::u_col(u);
              constexpr bool share = Cfg::u_share(u);
              constexpr uint32_t tx_bytes = (uint32_t)NPA * (share ? 0u : (uint32_t)SH_A_TILE) + (WRES ? 0u : (uint32_t)NPW * (uint32_t)Cfg::u_n(u) * 128u);
              mbar_wait(empty(slot), ph ^ 1u);
              if (leader) {
                const uint32_t base = ring + (uint32_t)slot * p.slot_bytes;
                mbar_expect_tx(full(slot), tx_bytes);
                if (!share) {
#pragma unroll
                  for (int pl = 0; pl < NPA; ++pl)
                    tma_load_4d(base + (uint32_t)pl * SH_A_TILE, &p.a_map[pl], p.grp[g].in_coff + ch * SH_BK, tx * p.BW, y0 + r - p.pad, n, full(slot));
                }
                if (!WRES) {
#pragma unroll
                  for (int pl = 0; pl < NPW; ++pl) tma_load_3d(base + w_off + (uint32_t)pl * p.w_slab, &p.w_map[u][pl], ch * SH_BK, wrow, r, full(slot));
                }
              }
              __syncwarp();
              if (++slot == p.NS) { slot = 0; ph ^= 1u; }
            });
          }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool leader = elect_one();
    int slot = 0; uint32_t ph = 0;
    uint32_t tile_it = 0;
    if (WRES) { mbar_wait(wres_bar, 0u); tc_fence_after(); }
    for (int work = blockIdx.x; work < p.num_work; work += gridDim.x)
      for (int tx = 0; tx < p.tiles_x; ++tx, ++tile_it) {
        const uint32_t acc = ACC == 2 ? (tile_it & 1u) : 0u;
        const uint32_t acc_phase = ACC == 2 ? ((tile_it >> 1) & 1u) : (tile_it & 1u);
        mbar_wait(tmem_empty(acc), acc_phase ^ 1u);          // the epilogue has read this accumulator set
        tc_fence_after();
        const uint32_t dbase = tmem_base + acc * (uint32_t)COLS;
        for (int r = 0; r < TAPS; ++r)
          for (int ch = 0; ch < p.kchunks; ++ch) {
            const uint32_t accum0 = (r | ch) != 0 ? 1u : 0u;
            static_for<0, NU>([&](auto U) __attribute__((always_inline)) {
              constexpr int u = decltype(U)::value;
              constexpr bool share = Cfg::u_share(u), keep = Cfg::u_keep(u);
              constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(Cfg::u_n(u) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, K-major, M = 128
              constexpr uint32_t dcol = (uint32_t)Cfg::u_col(u);
              mbar_wait(full(slot), ph);
              tc_fence_after();
              const uint32_t sbase = ring + (uint32_t)slot * p.slot_bytes;
              const int pslot = slot == 0 ? p.NS - 1 : slot - 1;
              const uint32_t abase = share ? ring + (uint32_t)pslot * p.slot_bytes : sbase;
              uint64_t adesc[NPA], bdesc[NPW];
#pragma unroll
              for (int pl = 0; pl < NPA; ++pl) adesc[pl] = umma_desc(abase + (uint32_t)pl * SH_A_TILE);
#pragma unroll
              for (int pl = 0; pl < NPW; ++pl)
                bdesc[pl] = umma_desc(WRES ? wres + (uint32_t)((r * p.kchunks + ch) * NPW + pl) * p.w_slab : sbase + w_off + (uint32_t)pl * p.w_slab);
              if (leader) {
#pragma unroll
                for (int k = 0; k < SH_BK / 16; ++k)
#pragma unroll
                  for (int ps = 0; ps < PASSES; ++ps) {
                    constexpr int kAPl[3] = {0, 1, 0}, kWPl[3] = {0, 0, 1};     // hi*hi, lo*hi, hi*lo
                    umma_f16(dbase + dcol, adesc[kAPl[ps] % NPA] + (uint64_t)(2 * k), bdesc[kWPl[ps] % NPW] + (uint64_t)(2 * k), idesc,
                             (k | ps) == 0 ? accum0 : 1u);
                  }
                if (share) umma_commit(empty(pslot));        // the shared A tile's slot is released with this unit
                if (!keep) umma_commit(empty(slot));
              }
              __syncwarp();
              if (++slot == p.NS) { slot = 0; ph ^= 1u; }
            });
          }
        if (leader) umma_commit(tmem_full(acc));
        __syncwarp();
      }
  } else {
    // ===== epilogue: warps 2..9; warp w owns TMEM lanes 32 * (w % 4) .. + 31 (thread t <-> tile position t); the two warps of
    // a quadrant ("halves") take alternate 16-channel batches of the work list =====
    // (launch parameters used in the loops below are copied into registers once: the epilogue is instruction-bound, and reloading them
    // from the parameter bank after every asm barrier cost a third of its instructions in the implicit-GEMM kernel)
    struct { int H, W, pad, BW, BH, tiles_x, rows_y, num_work, row_mode, RB, relu; unsigned stage_bytes; const uint8_t* mask;
             __half* out_hi; __half* out_lo; } const P_ = {p.H, p.W, p.pad, p.BW, p.BH, p.tiles_x, p.rows_y, p.num_work, p.row_mode, p.RB, p.relu,
                                                       p.stage_bytes, p.mask, p.out_hi, p.out_lo};
    const ShGroup grp_[SH_MAX_GROUPS] = {p.grp[0], p.grp[1], p.grp[2]};
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                     // which of the NPART warps of this quadrant ("half" in the two-part kernels)
    const int part = half;
    const int t = quad * 32 + lane;
    const int tj = P_.row_mode ? 0 : t / P_.BW, txx = P_.row_mode ? t : t - tj * P_.BW;
    const long long cs = (long long)P_.H * P_.W;
    uint32_t tile_it = 0;
    int off = 0;                                           // ring offset of logical buffer row 0
    // register-shuffle epilogue (64-channel 3x3, whole-row tiles): this warp's 32 biases live in registers for the whole kernel
    // (one shared-memory load per output value otherwise)
    constexpr bool kRegEpi = OUTMODE == 1 && TAPS == 3 && O0 == 64 && NG == 1;
    float r_bias[kRegEpi ? 32 : 1];
    if constexpr (kRegEpi) {
#pragma unroll
      for (int c = 0; c < 32; ++c) r_bias[c] = s_bias[half * 32 + c];
    }
    for (int work = blockIdx.x; work < P_.num_work; work += gridDim.x) {
      const int n = work / P_.rows_y, y0 = (work - n * P_.rows_y) * P_.BH;
      for (int tx = 0; tx < P_.tiles_x; ++tx, ++tile_it) {
        const uint32_t acc = ACC == 2 ? (tile_it & 1u) : 0u;
        const uint32_t acc_phase = ACC == 2 ? ((tile_it >> 1) & 1u) : (tile_it & 1u);
        mbar_wait(tmem_full(acc), acc_phase);
        tc_fence_after();
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (uint32_t)COLS;
        if constexpr (OUTMODE == 1 && TAPS == 3 && O0 == 64 && NG == 1) {
          if (P_.row_mode) {
            // ---- 3x3, 64 channels, whole-row tiles: shift-add in REGISTERS.  out[x] = Z[x-1][tap 0] + Z[x][tap 1] + Z[x+1][tap 2]:
            // the neighbours' values come by warp shuffle, only warp / tile edges go through shared memory.  (The smem row
            // buffer of the generic path below competes with the tensor core for the shared-memory pipe: measured 48 % LSU +
            // 24 % UMMA wavefronts on c0_conv.2.)  The last pixel of a tile waits for the next tile's first pixel ("pend").
            const bool first = tx == 0, last = tx == P_.tiles_x - 1;
            const uint32_t par = tile_it & 1u;
            float* e0 = s_edge + par * 256u;                 // tap-0 values of lane 31 of every quadrant, [quad][64], double-buffered
            const float* e0_prev = s_edge + (par ^ 1u) * 256u;
            float* e2 = s_edge + 512;                        // tap-2 values of lane 0 of every quadrant
            float* pend = s_edge + 768;                      // partial sum of the previous tile's last pixel
            const int cb = half * 32;                        // this warp's 32 channels
            const ShGroup& G = grp_[0];
            // Finished pixels leave through a staging tile in shared memory and ONE TMA store per plane and tile: a lane-per-pixel
            // direct store writes 16 B per lane at a 128-byte stride, i.e. 32 partial-sector L2 transactions per instruction, and
            // that alone cost half of the kernel time (c0_conv.2: 2.25 ms with, 1.14 ms without the stores).  Staging row j holds
            // pixel x0 + j for j < 127; the tile's last pixel waits in "pend" and is stored directly by the next tile.
            const bool staged = P_.stage_bytes != 0;
            auto emit32 = [&](const float* v, long long pix, int srow) __attribute__((always_inline)) {
              const bool has_mask = P_.mask != nullptr;                    // uniform: the forward_seg atlases only
              const bool keep = !has_mask || P_.mask[pix] != 0;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint4 h4, l4;
                uint32_t* hh = reinterpret_cast<uint32_t*>(&h4);
                uint32_t* ll = reinterpret_cast<uint32_t*>(&l4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float a = fmaf(v[8 * k + 2 * e], G.inv_scale, r_bias[8 * k + 2 * e]);
                  float bq = fmaf(v[8 * k + 2 * e + 1], G.inv_scale, r_bias[8 * k + 2 * e + 1]);
                  if (P_.relu) { a = fmaxf(a, 0.f); bq = fmaxf(bq, 0.f); }
                  if (has_mask) { a = keep ? a : 0.f; bq = keep ? bq : 0.f; }
                  if (P_.out_lo != nullptr) split_f16x2(a, bq, hh[e], ll[e]);
                  else hh[e] = f16x2_sat(a, bq);
                }
                if (srow >= 0) {                               // 128-byte swizzle of the store's tensor map: 16-byte piece j of row r at (j ^ (r & 7))
                  const uint32_t a0 = stage + (uint32_t)srow * 128u + ((((uint32_t)(half * 4 + k)) ^ ((uint32_t)srow & 7u)) << 4);
                  st_shared_v4(a0, h4);
                  if (P_.out_lo != nullptr) st_shared_v4(a0 + 16384u, l4);
                } else {
                  *reinterpret_cast<uint4*>(P_.out_hi + pix * 64 + cb + 8 * k) = h4;
                  if (P_.out_lo != nullptr) *reinterpret_cast<uint4*>(P_.out_lo + pix * 64 + cb + 8 * k) = l4;
                }
              }
            };
            if (staged && warp == 2 && lane == 0) bulk_wait_read0();   // the previous tile's TMA stores have read the staging tile
            __syncwarp();                                              // (tcgen05.ld below is warp-collective)
            const long long pix0 = ((long long)n * P_.H + y0) * P_.W + (long long)tx * 128;
            float z0[32], z2[32];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tmem_ld8_nowait(lane_addr + (uint32_t)(cb + 8 * c), reinterpret_cast<uint32_t*>(&z0[8 * c]));
              tmem_ld8_nowait(lane_addr + (uint32_t)(128 + cb + 8 * c), reinterpret_cast<uint32_t*>(&z2[8 * c]));
            }
            tmem_ld_wait();
            // edge exchange with the neighbouring quadrants: 128-bit shared stores / loads (one lane is active, the instruction count
            // is what matters)
            if (lane == 31) {
              float4* d4 = reinterpret_cast<float4*>(e0 + quad * 64 + cb);
#pragma unroll
              for (int c = 0; c < 8; ++c) d4[c] = make_float4(z0[4 * c], z0[4 * c + 1], z0[4 * c + 2], z0[4 * c + 3]);
            }
            if (lane == 0) {
              float4* d4 = reinterpret_cast<float4*>(e2 + quad * 64 + cb);
#pragma unroll
              for (int c = 0; c < 8; ++c) d4[c] = make_float4(z2[4 * c], z2[4 * c + 1], z2[4 * c + 2], z2[4 * c + 3]);
            }
            epi_bar();                                       // (also orders the staging writes below after bulk_wait_read0 above)
            if (lane == 0 && !first && quad == 0) {          // the previous tile's last pixel is complete now
              float v[32];
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = pend[cb + c] + z2[c];
              emit32(v, pix0 - 1, -1);                         // one pixel per tile: direct store
            }
            __syncwarp();                                    // reconverge before the warp-collective shuffles / tcgen05.ld below
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              z0[c] = __shfl_up_sync(0xffffffffu, z0[c], 1);       // left neighbour's tap 0
              z2[c] = __shfl_down_sync(0xffffffffu, z2[c], 1);     // right neighbour's tap 2
            }
            if (lane == 0) {
              const float4* s4 = reinterpret_cast<const float4*>(quad == 0 ? e0_prev + 3 * 64 + cb : e0 + (quad - 1) * 64 + cb);
              const bool zero = quad == 0 && first;
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 f = zero ? make_float4(0.f, 0.f, 0.f, 0.f) : s4[c];
                z0[4 * c] = f.x; z0[4 * c + 1] = f.y; z0[4 * c + 2] = f.z; z0[4 * c + 3] = f.w;
              }
            }
            if (lane == 31) {
              const float4* s4 = reinterpret_cast<const float4*>(e2 + (quad == 3 ? 0 : quad + 1) * 64 + cb);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 f = quad == 3 ? make_float4(0.f, 0.f, 0.f, 0.f) : s4[c];
                z2[4 * c] = f.x; z2[4 * c + 1] = f.y; z2[4 * c + 2] = f.z; z2[4 * c + 3] = f.w;
              }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t z1[8];
              tmem_ld8_nowait(lane_addr + (uint32_t)(64 + cb + 8 * c), z1);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) z0[8 * c + e] = __uint_as_float(z1[e]) + z0[8 * c + e] + z2[8 * c + e];
            }
            tc_fence_before();                               // Z fully read: hand the accumulators back to the MMA issuer
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty(acc)) : "memory");
            if (quad == 3 && lane == 31 && !last) {
#pragma unroll
              for (int c = 0; c < 32; ++c) pend[cb + c] = z0[c];
            } else {
              emit32(z0, pix0 + t, (staged && t < 127) ? t : -1);       // the row's very last pixel goes out directly
            }
            if (staged) fence_proxy_async_smem();
            epi_bar();
            if (staged && warp == 2 && lane == 0) {
              tma_store_4d(&p.o_map[0], stage, 0, tx * 128, y0, n);                  // pixels x0 .. x0 + 126
              if (P_.out_lo != nullptr) tma_store_4d(&p.o_map[1], stage + 16384u, 0, tx * 128, y0, n);
              bulk_commit();
            }
            __syncwarp();
            continue;
          }
        }
        // ---- shift-add: tap s of position q goes to output pixel q - s + pad (everything below is compile-time unrolled) ----
        static_for<0, TAPS>([&](auto Sx) __attribute__((always_inline)) {
          constexpr int s = decltype(Sx)::value;
          int L; bool valid = true;
          if (P_.row_mode) { L = t - s + 2 * P_.pad; }
          else { const int x2 = txx - s + P_.pad; valid = x2 >= 0 && x2 < P_.BW; L = tj * P_.BW + x2; }
          int phys = L + off; if (phys >= P_.RB) phys -= P_.RB;
          float* row = buf + (valid ? phys : 0) * RS;
          // batch by batch (<= 16 channels): TMEM loads -> wait -> smem loads -> adds -> stores
          static_for<0, Cfg::NBT>([&](auto Bx) __attribute__((always_inline)) {
            constexpr int b = decltype(Bx)::value;
            constexpr int g = Cfg::b_group(b), c_lo = Cfg::b_lo(b), c_n = Cfg::b_n(b), b0 = Cfg::bufcol(g), nch = (c_n + 7) / 8;
            if ((b % NPART) == part) {
              uint32_t raw[nch][8];
              static_for<0, nch>([&](auto Cx) __attribute__((always_inline)) {
                constexpr int c = decltype(Cx)::value;
                constexpr uint32_t col = (uint32_t)(Cfg::col0(g) + Cfg::stride(g) * s + c_lo + 8 * c);
                tmem_ld8_nowait(lane_addr + col, raw[c]);
              });
              tmem_ld_wait();
              if (s == TAPS - 1 && b == Cfg::last_batch_of_part(b % NPART, NPART)) {   // this warp has read all of its columns of Z
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty(acc)) : "memory");
              }
              if (valid) {
                float cur[c_n];
#pragma unroll
                for (int e = 0; e < c_n; ++e) cur[e] = row[b0 + c_lo + e];        // all loads first: the smem latency is paid once
#pragma unroll
                for (int e = 0; e < c_n; ++e) row[b0 + c_lo + e] = cur[e] + __uint_as_float(raw[e / 8][e % 8]);
              }
            }
          });
          if (s == TAPS - 1) {                                  // a part without work still releases the accumulators
            static_for<0, NPART>([&](auto Qx) __attribute__((always_inline)) {
              constexpr int q = decltype(Qx)::value;
              if (Cfg::last_batch_of_part(q, NPART) < 0 && part == q) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty(acc)) : "memory");
              }
            });
          }
          epi_bar<32 * EW>();
        });
        // ---- emit the finished pixels (and clear their buffer rows) ----
        const bool last = tx == P_.tiles_x - 1;
        const int nL = P_.row_mode ? (last ? P_.RB : 128) : 128;
        for (int L = t; L < nL; L += 128) {
          int phys = L + off; if (phys >= P_.RB) phys -= P_.RB;
          float* row = buf + phys * RS;
          int y, x; bool inside;
          if (P_.row_mode) { y = y0; x = tx * 128 + L - P_.pad; inside = x >= 0 && x < P_.W; }
          else { const int jj = L / P_.BW; y = y0 + jj; x = L - jj * P_.BW; inside = y < P_.H; }
          const long long pix = ((long long)n * P_.H + y) * P_.W + x;
          const bool keep = OUTMODE == 1 && inside && (P_.mask == nullptr || P_.mask[pix] != 0);
          static_for<0, Cfg::NBT>([&](auto Bx) __attribute__((always_inline)) {
            constexpr int b = decltype(Bx)::value;
            constexpr int g = Cfg::b_group(b), c_lo = Cfg::b_lo(b), c_n = Cfg::b_n(b), b0 = Cfg::bufcol(g), no = Cfg::n_out(g);
            if ((b % NPART) == part) {
              float v[c_n];
#pragma unroll
              for (int e = 0; e < c_n; ++e) v[e] = row[b0 + c_lo + e];
#pragma unroll
              for (int e = 0; e < c_n; ++e) row[b0 + c_lo + e] = 0.f;
              const ShGroup& G = grp_[g];
              if (inside) {
                if (OUTMODE == 0) {
                  float* dst = G.out32 + (((long long)n * no + c_lo) * P_.H + y) * P_.W + x;
#pragma unroll
                  for (int e = 0; e < c_n; ++e) {
                    float o = fmaf(v[e], G.inv_scale, s_bias[b0 + c_lo + e]);
                    if (G.sigmoid) o = 1.f / (1.f + expf(-o));
                    *dst = o;
                    dst += cs;
                  }
                } else {
                  // split-fp16 NHWC: c_n is a multiple of 8 here
                  static_for<0, c_n / 8>([&](auto Cx) __attribute__((always_inline)) {
                    constexpr int c0 = decltype(Cx)::value * 8;
                    uint4 h4, l4;
                    uint32_t* hh = reinterpret_cast<uint32_t*>(&h4);
                    uint32_t* ll = reinterpret_cast<uint32_t*>(&l4);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      float a = fmaf(v[c0 + 2 * e], G.inv_scale, s_bias[b0 + c_lo + c0 + 2 * e]);
                      float bq = fmaf(v[c0 + 2 * e + 1], G.inv_scale, s_bias[b0 + c_lo + c0 + 2 * e + 1]);
                      if (P_.relu) { a = fmaxf(a, 0.f); bq = fmaxf(bq, 0.f); }
                      if (!keep) { a = 0.f; bq = 0.f; }
                      split_f16x2(a, bq, hh[e], ll[e]);
                    }
                    *reinterpret_cast<uint4*>(P_.out_hi + pix * no + c_lo + c0) = h4;
                    if (P_.out_lo != nullptr) *reinterpret_cast<uint4*>(P_.out_lo + pix * no + c_lo + c0) = l4;
                  });
                }
              }
            }
          });
        }
        if (P_.row_mode) { off += 128; if (off >= P_.RB) off -= P_.RB; }
        epi_bar<32 * EW>();
      }
    }
  }
  if (p.stage_bytes != 0 && warp == 2 && lane == 0) bulk_wait_all();   // every TMA store has landed before the CTA retires
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// the instantiated configurations
using HeadsCfg = ShCfg<5, 10, 40, 7>;   // KGnet's second-layer heads: kp 5 (sigmoid), short offsets 10, mid offsets 40; 7x7
using C64Cfg = ShCfg<64, 0, 0, 3>;      // one 3x3 conv with 64 output channels (c0_conv.2, c1/c2_up_conv, layer1 conv2, mask branch)
using C1Cfg = ShCfg<1, 0, 0, 3>;        // seg_head.2: 3x3, one output channel
enum ShKernel { SHK_NONE = -1, SHK_HEADS = 0, SHK_C64, SHK_C1 };

static ShKernel pick_kernel(int R, int S, int n_groups, const int* n_out, bool nhwc_out) {
  if (R != S) return SHK_NONE;
  if (S == 7 && n_groups == 3 && n_out[0] == 5 && n_out[1] == 10 && n_out[2] == 40 && !nhwc_out) return SHK_HEADS;
  if (S == 3 && n_groups == 1 && n_out[0] == 64 && nhwc_out) return SHK_C64;
  if (S == 3 && n_groups == 1 && n_out[0] == 1 && !nhwc_out) return SHK_C1;
  return SHK_NONE;
}

bool tc_shift_supported(int H, int W, int R, int S, int pad, int Cin, int n_groups, const int* n_out, bool nhwc_out) {
  if (!tc_available()) return false;
  const char* off = getenv("KG_TC_SHIFT");
  if (off && off[0] == '0') return false;
  if (H < 1 || 2 * pad != S - 1 || Cin % SH_BK != 0 || n_groups < 1 || n_groups > SH_MAX_GROUPS) return false;
  if (!((W >= 128 && W % 128 == 0) || (W >= 8 && W < 128 && is_pow2(W)))) return false;
  const ShKernel k = pick_kernel(R, S, n_groups, n_out, nhwc_out);
  if (k == SHK_C64 || k == SHK_C1) { const char* c = getenv("KG_TC_SHIFT_CONV"); if (c && c[0] == '0') return false; }
  return k != SHK_NONE;
}

template <class Cfg>
static int shift_pack_t(const TcShiftOp* op, TcShiftPacked* out) {
  const int rows = Cfg::COLS, S = op->S, R = op->R;
  const size_t ne = (size_t)R * rows * op->Cin;
  std::vector<__half> hi(ne, __float2half_rn(0.f)), lo;
  if (op->passes == 3) lo.assign(ne, __float2half_rn(0.f));
  // row of (conv g, tap s, channel co) = col0(g) + s * stride(g) + co; each conv is scaled by a power of two so that small
  // weights stay in fp16's normal range (undone by inv_scale in the epilogue)
  for (int g = 0; g < Cfg::NG; ++g) {
    const TcShiftGroup& G = op->g[g];
    KG_REQUIRE(G.h_w && G.n_out == Cfg::n_out(g), "tc_shift_pack: conv %d: null weights or unexpected Cout", g);
    float mx = 0.f;
    const size_t nw = (size_t)R * S * op->Cin * G.n_out;
    for (size_t i = 0; i < nw; ++i) mx = fmaxf(mx, fabsf(G.h_w[i]));
    int e = 0;
    if (mx > 0.f && std::isfinite(mx)) { int ex; frexpf(mx, &ex); e = 10 - ex; }
    e = std::max(-40, std::min(40, e));
    const float scale = ldexpf(1.f, e);
    out->inv_scale[g] = ldexpf(1.f, -e);
    for (int r = 0; r < R; ++r)
      for (int t = 0; t < S; ++t)
        for (int ci = 0; ci < op->Cin; ++ci)
          for (int co = 0; co < G.n_out; ++co) {
            const float v = G.h_w[(((size_t)r * S + t) * op->Cin + ci) * G.n_out + co] * scale;
            const size_t o = ((size_t)r * rows + Cfg::col0(g) + t * Cfg::stride(g) + co) * op->Cin + ci;
            const __half h = __float2half_rn(v);
            hi[o] = h;
            if (op->passes == 3) lo[o] = __float2half_rn(v - __half2float(h));
          }
  }
  __half* d = nullptr;
  KG_CUDA_CHECK(cudaMalloc(&d, ne * sizeof(__half)));
  out->d_hi = std::shared_ptr<void>(d, [](void* q) { cudaFree(q); });
  KG_CUDA_CHECK(cudaMemcpy(d, hi.data(), ne * sizeof(__half), cudaMemcpyHostToDevice));
  if (op->passes == 3) {
    KG_CUDA_CHECK(cudaMalloc(&d, ne * sizeof(__half)));
    out->d_lo = std::shared_ptr<void>(d, [](void* q) { cudaFree(q); });
    KG_CUDA_CHECK(cudaMemcpy(d, lo.data(), ne * sizeof(__half), cudaMemcpyHostToDevice));
  }
  out->passes = op->passes;
  return KG_OK;
}

template <class Cfg>
static int shift_prepare_t(TcShiftOp* op, EncodeTiledFn encode) {
  std::shared_ptr<ShParams> sp(new ShParams());
  ShParams& p = *sp;
  memset(&p, 0, sizeof(p));
  const int NPA = op->passes >= 2 ? 2 : 1, NPW = op->passes == 3 ? 2 : 1;
  p.N = op->N; p.H = op->H; p.W = op->W; p.pad = op->pad; p.kchunks = op->Cin / SH_BK;
  p.row_mode = op->W >= 128 ? 1 : 0;
  p.BW = p.row_mode ? 128 : op->W; p.BH = 128 / p.BW;
  p.tiles_x = p.row_mode ? op->W / 128 : 1;
  p.rows_y = ceil_div(op->H, p.BH);
  p.num_work = op->N * p.rows_y;
  p.RB = p.row_mode ? 128 + 2 * op->pad : 128;
  if (std::is_same<Cfg, C64Cfg>::value && p.row_mode && op->out_hi != nullptr) p.RB = 0;   // register-shuffle epilogue: no row buffer
  p.out_hi = op->out_hi; p.out_lo = op->out_lo; p.mask = op->mask; p.relu = op->relu ? 1 : 0;
  const int rows = Cfg::COLS, R = op->R;
  const int acc_sets = 2 * Cfg::COLS <= 512 ? 2 : 1;
  unsigned tc = 32;
  while (tc < (unsigned)(acc_sets * Cfg::COLS)) tc *= 2;
  p.tmem_cols = tc;
  if (!op->packed.valid() || (op->passes == 3 && op->packed.d_lo == nullptr)) KG_TRY(shift_pack_t<Cfg>(op, &op->packed));
  for (int g = 0; g < Cfg::NG; ++g) {
    const TcShiftGroup& G = op->g[g];
    KG_REQUIRE(G.d_bias && G.n_out == Cfg::n_out(g), "tc_shift_prepare: conv %d: null bias or unexpected Cout", g);
    ShGroup& D = p.grp[g];
    D.bias = G.d_bias; D.in_coff = G.in_coff; D.sigmoid = G.sigmoid ? 1 : 0; D.inv_scale = op->packed.inv_scale[g];
  }
  for (int pl = 0; pl < NPA; ++pl) {
    const __half* base = pl == 0 ? op->in_hi : op->in_lo;
    KG_REQUIRE(base != nullptr, "tc_shift_prepare: input plane %d is null", pl);
    cuuint64_t dims[4] = {(cuuint64_t)op->in_C, (cuuint64_t)op->W, (cuuint64_t)op->H, (cuuint64_t)op->N};
    cuuint64_t strides[3] = {(cuuint64_t)op->in_C * 2, (cuuint64_t)op->W * op->in_C * 2, (cuuint64_t)op->H * op->W * op->in_C * 2};
    cuuint32_t box[4] = {(cuuint32_t)SH_BK, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&p.a_map[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_shift_prepare: cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return KG_ERR_CUDA; }
  }
  int maxN = 0;
  for (int u = 0; u < Cfg::NU; ++u) {
    maxN = std::max(maxN, Cfg::u_n(u));
    for (int pl = 0; pl < NPW; ++pl) {
      void* base = pl == 0 ? op->packed.d_hi.get() : op->packed.d_lo.get();
      cuuint64_t dims[3] = {(cuuint64_t)op->Cin, (cuuint64_t)rows, (cuuint64_t)R};
      cuuint64_t strides[2] = {(cuuint64_t)op->Cin * 2, (cuuint64_t)rows * op->Cin * 2};
      cuuint32_t box[3] = {(cuuint32_t)SH_BK, (cuuint32_t)Cfg::u_n(u), 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = encode(&p.w_map[u][pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("tc_shift_prepare: cuTensorMapEncodeTiled(weights, unit %d) failed: %d", u, (int)r); return KG_ERR_CUDA; }
    }
  }
  p.w_slab = (unsigned)align_up((size_t)maxN * 128, 1024);
  const size_t buf_bytes = ((size_t)p.RB * Cfg::RS + Cfg::NOUT + 4 + 832) * sizeof(float);
  // TMA-store epilogue of the register-shuffle path (64-channel NHWC output, whole-row tiles)
  { const char* e = getenv("KG_SH_STAGE"); p.stage_bytes = (std::is_same<Cfg, C64Cfg>::value && p.row_mode && op->out_hi != nullptr && !(e && e[0] == '0')) ? 32768u : 0u; }
  if (p.stage_bytes) {
    for (int pl = 0; pl < 2; ++pl) {
      __half* base = pl == 0 ? op->out_hi : op->out_lo;
      if (base == nullptr) continue;
      cuuint64_t dims[4] = {64, (cuuint64_t)op->W, (cuuint64_t)op->H, (cuuint64_t)op->N};
      cuuint64_t strides[3] = {128, (cuuint64_t)op->W * 128, (cuuint64_t)op->H * op->W * 128};
      cuuint32_t box[4] = {64, 127, 1, 1};                   // pixels x0 .. x0 + 126 of a tile; the 128th is finished by the next tile
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = encode(&p.o_map[pl], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("tc_shift_prepare: cuTensorMapEncodeTiled(output) failed: %d", (int)r); return KG_ERR_CUDA; }
    }
  }
  const size_t fixed = 1024 + 16 * 16 + 64 + buf_bytes + 64 + p.stage_bytes;
  // resident weights: the whole [R][chunks][planes] weight tensor stays in smem when it leaves room for >= 3 activation slots
  const size_t wres_total = (size_t)R * p.kchunks * NPW * p.w_slab;
  const char* wr = getenv("KG_TC_SHIFT_WRES");
  op->wres = Cfg::NU == 1 && !(wr && wr[0] == '0') && fixed + wres_total + 3 * (size_t)NPA * SH_A_TILE <= (size_t)SH_MAX_SMEM;
  p.wres_bytes = op->wres ? (unsigned)wres_total : 0u;
  p.slot_bytes = (unsigned)(NPA * SH_A_TILE + (op->wres ? 0 : NPW * p.w_slab));
  int ns = (int)((SH_MAX_SMEM - fixed - p.wres_bytes) / p.slot_bytes);
  if (ns > 8) ns = 8;
  KG_REQUIRE(ns >= 2, "tc_shift_prepare: tile does not fit in shared memory");
  p.NS = ns;
  op->smem_bytes = (unsigned)(1024 + p.stage_bytes + p.wres_bytes + (size_t)ns * p.slot_bytes + 16 * ns + 64 + buf_bytes + 64);
  op->grid = (unsigned)std::min(p.num_work, tc_num_sms());
  op->params = sp;
  if (getenv("KG_TC_DEBUG"))
    fprintf(stderr, "[tc_shift] N%d %dx%d Cin%d k%dx%d passes%d wres%d | units%d cols%d acc%d BW%d BH%d tiles_x%d work%d NS%d slot%u RB%d RS%d smem%u tmem%u\n",
            op->N, op->H, op->W, op->Cin, R, op->S, op->passes, (int)op->wres, Cfg::NU, Cfg::COLS, acc_sets, p.BW, p.BH, p.tiles_x, p.num_work, p.NS, p.slot_bytes, p.RB,
            Cfg::RS, op->smem_bytes, p.tmem_cols);
  return KG_OK;
}

static ShKernel kernel_of(const TcShiftOp* op) {
  int n_out[SH_MAX_GROUPS] = {0, 0, 0};
  for (int g = 0; g < op->n_groups && g < SH_MAX_GROUPS; ++g) n_out[g] = op->g[g].n_out;
  if (!tc_shift_supported(op->H, op->W, op->R, op->S, op->pad, op->Cin, op->n_groups, n_out, op->out_hi != nullptr)) return SHK_NONE;
  return pick_kernel(op->R, op->S, op->n_groups, n_out, op->out_hi != nullptr);
}

int tc_shift_pack(const TcShiftOp* op, TcShiftPacked* out) {
  KG_REQUIRE(op && out, "tc_shift_pack: null argument");
  switch (kernel_of(op)) {
    case SHK_HEADS: return shift_pack_t<HeadsCfg>(op, out);
    case SHK_C64: return shift_pack_t<C64Cfg>(op, out);
    case SHK_C1: return shift_pack_t<C1Cfg>(op, out);
    default: set_error("tc_shift_pack: unsupported shape"); return KG_ERR_INVALID;
  }
}

int tc_shift_prepare(TcShiftOp* op) {
  KG_REQUIRE(op != nullptr, "tc_shift_prepare: null op");
  KG_REQUIRE(op->passes >= 1 && op->passes <= 3, "tc_shift_prepare: passes=%d", op->passes);
  const ShKernel k = kernel_of(op);
  KG_REQUIRE(k != SHK_NONE, "tc_shift_prepare: unsupported shape");
  KG_REQUIRE(k != SHK_HEADS || op->passes == 1, "tc_shift_prepare: the fused head kernel is single-pass only");
  EncodeTiledFn encode = (EncodeTiledFn)tc_encode_tiled_fn();
  KG_REQUIRE(encode != nullptr, "tc_shift_prepare: cuTensorMapEncodeTiled unavailable");
  static bool attr_set = false;
  if (!attr_set) {
#define KG_SH_ATTR(...) KG_CUDA_CHECK(cudaFuncSetAttribute(tc_shift_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_MAX_SMEM))
    KG_SH_ATTR(5, 10, 40, 7, 1, 0, false);
    KG_SH_ATTR(64, 0, 0, 3, 1, 1, false); KG_SH_ATTR(64, 0, 0, 3, 2, 1, false); KG_SH_ATTR(64, 0, 0, 3, 3, 1, false);
    KG_SH_ATTR(64, 0, 0, 3, 1, 1, true); KG_SH_ATTR(64, 0, 0, 3, 2, 1, true); KG_SH_ATTR(64, 0, 0, 3, 3, 1, true);
    KG_SH_ATTR(1, 0, 0, 3, 1, 0, false); KG_SH_ATTR(1, 0, 0, 3, 3, 0, false);
    KG_SH_ATTR(1, 0, 0, 3, 1, 0, true); KG_SH_ATTR(1, 0, 0, 3, 3, 0, true);
#undef KG_SH_ATTR
    attr_set = true;
  }
  switch (k) {
    case SHK_HEADS: return shift_prepare_t<HeadsCfg>(op, encode);
    case SHK_C64: return shift_prepare_t<C64Cfg>(op, encode);
    default: return shift_prepare_t<C1Cfg>(op, encode);
  }
}

int tc_shift_launch(const TcShiftOp* op, float* const* out32, cudaStream_t stream) {
  KG_REQUIRE(op && op->params, "tc_shift_launch: op not prepared");
  ShParams p = *reinterpret_cast<const ShParams*>(op->params.get());
  const bool nhwc = op->out_hi != nullptr;
  if (!nhwc) {
    KG_REQUIRE(out32 != nullptr, "tc_shift_launch: null outputs");
    for (int g = 0; g < op->n_groups; ++g) {
      KG_REQUIRE(out32[g] != nullptr, "tc_shift_launch: output %d is null", g);
      p.grp[g].out32 = out32[g];
    }
  }
#define KG_SH_LAUNCH(...) tc_shift_kernel<__VA_ARGS__><<<op->grid, SH_THREADS, op->smem_bytes, stream>>>(p)
  switch (kernel_of(op)) {
    case SHK_HEADS: tc_shift_kernel<5, 10, 40, 7, 1, 0, false><<<op->grid, 64 + 32 * sh_epi_warps(0, 7), op->smem_bytes, stream>>>(p); break;
    case SHK_C64:
      if (op->wres) { if (op->passes == 3) KG_SH_LAUNCH(64, 0, 0, 3, 3, 1, true); else if (op->passes == 2) KG_SH_LAUNCH(64, 0, 0, 3, 2, 1, true); else KG_SH_LAUNCH(64, 0, 0, 3, 1, 1, true); }
      else { if (op->passes == 3) KG_SH_LAUNCH(64, 0, 0, 3, 3, 1, false); else if (op->passes == 2) KG_SH_LAUNCH(64, 0, 0, 3, 2, 1, false); else KG_SH_LAUNCH(64, 0, 0, 3, 1, 1, false); }
      break;
    case SHK_C1:
      KG_REQUIRE(op->passes != 2, "tc_shift_launch: the one-channel kernel has no 2-pass variant");
      if (op->wres) { if (op->passes == 3) KG_SH_LAUNCH(1, 0, 0, 3, 3, 0, true); else KG_SH_LAUNCH(1, 0, 0, 3, 1, 0, true); }
      else { if (op->passes == 3) KG_SH_LAUNCH(1, 0, 0, 3, 3, 0, false); else KG_SH_LAUNCH(1, 0, 0, 3, 1, 0, false); }
      break;
    default: set_error("tc_shift_launch: unsupported shape"); return KG_ERR_INVALID;
  }
#undef KG_SH_LAUNCH
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

}  // namespace kg
