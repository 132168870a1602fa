// Implicit-GEMM convolution on the 5th-gen tensor cores of sm_100a (KGnet.py: every stride-1 conv with
// Cin % 64 == 0 — backbone 1x1/3x3, decoder 3x3 / concat-1x1, the 7x7 heads).
//
//   GEMM view      M = output pixels (a BH x BW patch of one image = 128 rows), N = output channels (<= 256 per CTA),
//                  K = taps x input channels, walked as (tap, 64-channel chunk) steps.
//   A operand      no im2col buffer: for filter tap (r, s) the 128 x 64 activation tile is ONE 4-D TMA box load of the
//                  NHWC tensor at (c0, x0 + s - pad, y0 + r - pad, n); out-of-image coordinates are zero-filled by
//                  TMA, which is exactly the conv's zero padding.  The box lands in shared memory as 128 rows of
//                  128 B with the 128-byte swizzle, i.e. the canonical K-major UMMA layout.
//   B operand      weights pre-packed [tap][cout][cin] fp16, one 3-D TMA box (64 x BN x 1) per step.
//   MMA            tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN, K = 16 x 4 per step, fp32 accumulators in
//                  TMEM (MT accumulators of BN columns; MT = 2 shares each weight tile between two pixel tiles).
//   precision      passes = 1: fp16 x fp16.  passes = 3: split-fp16 (hi*hi + lo*hi + hi*lo), ~fp32 accuracy.
//   concat         torch.cat((a, b), 1) feeding a conv = two TMA sources walked back to back along K.
//   epilogue       tcgen05.ld -> * 1/scale + bias (+ residual) -> ReLU / sigmoid -> split-fp16 NHWC and/or fp32 NCHW.
//   roles          warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue (two warps per TMEM lane quadrant);
//                  smem full/empty mbarrier rings between producer and MMA, tmem_full / tmem_empty barriers per
//                  accumulator buffer between MMA issuer and epilogue.  Persistent CTAs (one per SM).
//   CTA pair       tc_conv2_kernel below: the same conv as ONE M = 256 MMA over two SMs (cta_group::2) for the
//                  single-pass first-layer head convs, where the shared-memory pipe bounds the one-CTA kernel.
#include "tc_conv.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace kg {

constexpr int TC_BM = 128;          // rows (pixels) per accumulator tile
constexpr int TC_BK = 64;           // channels per K step (one 128-byte swizzle row)
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;   // 16 KiB
constexpr int TC_THREADS = 320;        // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue (two warps per TMEM lane quadrant)
constexpr int TC_MAX_SMEM = 227 * 1024;

struct TcParams {
  CUtensorMap a_map[2][2];          // [source][plane hi/lo]
  CUtensorMap w_map[2];             // [plane hi/lo]
  const float* bias;
  __half* out_hi; __half* out_lo;
  float* out32;
  const __half* res_hi; const __half* res_lo;
  const uint8_t* mask;
  float inv_scale;
  int N, H, W, Cout;
  int R, S, pad, stride;
  int chunks0, chunks1, coff0;
  int BN, BW, BH, tiles_x, tiles_y, m_tiles;
  int MT, NPL, passes;
  int NA, NW;                       // slots of the activation ring / weight ring
  int n_tiles, num_work;            // N tiles; work items = ceil(m_tiles / MT) * n_tiles, walked persistently
  int acc_stages;                   // TMEM accumulator buffers (2 = epilogue of tile i overlaps the MMAs of tile i+1)
  int KS;                           // independent accumulator chains per pixel tile (partials are summed in the epilogue):
                                    // back-to-back MMAs into ONE accumulator serialise on its latency when N is small
  int strip;                        // 1: A slot = (BW + S - 1)-pixel strip shared by the S taps of a filter row
  int wg;                           // weight tap group: taps per W slot (1, or S: the S taps of a filter row arrive in ONE TMA box
                                    // behind ONE barrier round trip -- narrow-N layers are bound by those round trips)
  unsigned a_tile_bytes, a_tx_bytes;   // smem bytes reserved per A tile (1024-aligned) / bytes TMA delivers per A tile
  int base_offset_mode;
  int two_cta;                      // CTA-pair kernel: BN is the FULL N of the pair's MMA, each CTA holds BN / 2 weight rows
  int relu, sigmoid;
  unsigned tmem_cols;
};

// ---- the kernel ---------------------------------------------------------------------------------
// Two independent smem rings: the A ring holds activation tiles, the W ring holds weight tiles.  In "strip" mode
// (tile = one image row of 128 pixels, filter wider than 1) one A slot holds the 128+S-1 pixel strip of filter row r and
// is reused by the S horizontal taps: tap s reads it through a descriptor whose start address is advanced by s rows
// (s * 128 B), so the activations are fetched from L2 once per filter ROW instead of once per tap.
template <int MT, int PASSES>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(const __grid_constant__ TcParams p) {
  constexpr int NPA = PASSES >= 2 ? 2 : 1;      // activation planes (hi, lo)
  constexpr int NPW = PASSES == 3 ? 2 : 1;      // weight planes: 2-pass = split activations x single-plane weights
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;     // 128B swizzle atoms need 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_tile = p.a_tile_bytes;                             // one plane of one pixel tile (or strip)
  const uint32_t a_slot = (uint32_t)(MT * NPA) * a_tile;
  const uint32_t w_tile = (uint32_t)p.BN * 128u;
  const uint32_t w_plane = (uint32_t)p.wg * w_tile;                   // one plane of one W slot: wg taps back to back
  const uint32_t w_slot = (uint32_t)NPW * w_plane;
  const uint32_t a_ring = smem0, w_ring = a_ring + (uint32_t)p.NA * a_slot;
  const uint32_t bars = w_ring + (uint32_t)p.NW * w_slot;
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (p.NA + s); };
  auto w_full = [&](int s) { return bars + 16u * p.NA + 8u * s; };
  auto w_empty = [&](int s) { return bars + 16u * p.NA + 8u * (p.NW + s); };
  const uint32_t bar_tmem = bars + 16u * (p.NA + p.NW);
  auto tmem_full = [&](int a) { return bar_tmem + 8u * a; };
  auto tmem_empty = [&](int a) { return bar_tmem + 16u + 8u * a; };
  const uint32_t tmem_slot = bar_tmem + 32u;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.a_map[0][0]); prefetch_tmap(&p.w_map[0]);
    if (NPA == 2) prefetch_tmap(&p.a_map[0][1]);
    if (NPW == 2) prefetch_tmap(&p.w_map[1]);
    if (p.chunks1 > 0) { prefetch_tmap(&p.a_map[1][0]); if (NPA == 2) prefetch_tmap(&p.a_map[1][1]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.NA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.NW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 8); }   // 8 epilogue warps release a buffer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int chunks = p.chunks0 + p.chunks1;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t acc_cols = (uint32_t)(MT * p.KS * p.BN);

  if (warp == 0) {
    {
      // ===== TMA producer: converged warp, one elected lane issues =====
      const bool leader = elect_one();
      int ai = 0, wi = 0; uint32_t aph = 0, wph = 0;
      for (int work = blockIdx.x; work < p.num_work; work += gridDim.x) {
      const int m_first = (work / p.n_tiles) * MT;
      const int n0 = (work % p.n_tiles) * p.BN;
      int tn[2], ty0[2], tx0[2];
      for (int mt = 0; mt < MT; ++mt) {
        const int m = m_first + mt;
        if (m < p.m_tiles) {
          const int n = m / tiles_per_img, rem = m - n * tiles_per_img;
          const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
          tn[mt] = n; ty0[mt] = ty * p.BH; tx0[mt] = tx * p.BW;
        } else { tn[mt] = p.N; ty0[mt] = 0; tx0[mt] = 0; }       // out-of-range tile: TMA zero-fills
      }
      for (int r = 0; r < p.R; ++r)
        for (int ch = 0; ch < chunks; ++ch) {
          const int src = ch < p.chunks0 ? 0 : 1;
          const int c = src == 0 ? p.coff0 + ch * TC_BK : (ch - p.chunks0) * TC_BK;
          for (int s = 0; s < p.S; ++s) {
            if (!p.strip || s == 0) {
              mbar_wait(a_empty(ai), aph ^ 1u);
              if (leader) {
                mbar_expect_tx(a_full(ai), (uint32_t)(MT * NPA) * p.a_tx_bytes);
                const uint32_t abase = a_ring + (uint32_t)ai * a_slot;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                  for (int pl = 0; pl < NPA; ++pl)
                    tma_load_4d(abase + (uint32_t)(mt * NPA + pl) * a_tile, &p.a_map[src][pl], c,
                                tx0[mt] * p.stride + (p.strip ? 0 : s) - p.pad, ty0[mt] * p.stride + r - p.pad, tn[mt], a_full(ai));
              }
              __syncwarp();
              if (++ai == p.NA) { ai = 0; aph ^= 1u; }
            }
            if (p.wg == 1 || s == 0) {
              mbar_wait(w_empty(wi), wph ^ 1u);
              if (leader) {
                mbar_expect_tx(w_full(wi), w_slot);
                const uint32_t wbase = w_ring + (uint32_t)wi * w_slot;
#pragma unroll
                for (int pl = 0; pl < NPW; ++pl) tma_load_3d(wbase + (uint32_t)pl * w_plane, &p.w_map[pl], ch * TC_BK, n0, r * p.S + s, w_full(wi));
              }
              __syncwarp();
              if (++wi == p.NW) { wi = 0; wph ^= 1u; }
            }
          }
        }
      }   // work loop
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer: the whole warp walks the loop (converged), one elected lane issues =====
      const bool leader = elect_one();
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);   // f16 x f16 -> f32, K-major A and B
      int ai = 0, wi = 0; uint32_t aph = 0, wph = 0;
      int it = 0;
      for (int work = blockIdx.x; work < p.num_work; work += gridDim.x, ++it) {
      const int acc = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
      mbar_wait(tmem_empty(acc), acc_phase ^ 1u);      // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t acc_base = tmem_base + (uint32_t)acc * acc_cols;
      // The single issuing thread is the critical resource: keep the per-MMA instruction count minimal (descriptor
      // constants hoisted, everything unrolled at compile time).
      uint32_t cnt = 0;                                  // MMAs issued per pixel tile of this work item
      const uint32_t ks_mask = (uint32_t)p.KS - 1u;      // KS is a power of two
      const uint32_t bn = (uint32_t)p.BN;
      for (int r = 0; r < p.R; ++r)
        for (int ch = 0; ch < chunks; ++ch)
          for (int s = 0; s < p.S; ++s) {
            if (!p.strip || s == 0) mbar_wait(a_full(ai), aph);
            if (p.wg == 1 || s == 0) mbar_wait(w_full(wi), wph);
            tc_fence_after();
            const uint32_t abase = a_ring + (uint32_t)ai * a_slot + (p.strip ? (uint32_t)s * 128u : 0u);
            const uint32_t wbase = w_ring + (uint32_t)wi * w_slot + (p.wg > 1 ? (uint32_t)s * w_tile : 0u);
            const bool w_last = p.wg == 1 || s == p.S - 1;
            uint64_t adesc[MT][NPA], bdesc[NPW];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int pl = 0; pl < NPA; ++pl) adesc[mt][pl] = umma_desc(abase + (uint32_t)(mt * NPA + pl) * a_tile);
#pragma unroll
            for (int pl = 0; pl < NPW; ++pl) bdesc[pl] = umma_desc(wbase + (uint32_t)pl * w_plane);
            if (leader) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
#pragma unroll
              for (int ps = 0; ps < PASSES; ++ps) {
                constexpr int kAPl[3] = {0, 1, 0}, kWPl[3] = {0, 0, 1};     // hi*hi, lo*hi, hi*lo
                const uint32_t chain = cnt & ks_mask;
                const uint32_t accum = cnt > ks_mask ? 1u : 0u;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)                              // innermost: consecutive MMAs hit different accumulators
                  umma_f16(acc_base + ((uint32_t)mt * (uint32_t)p.KS + chain) * bn, adesc[mt][kAPl[ps] % NPA] + (uint64_t)(2 * k),
                           bdesc[kWPl[ps] % NPW] + (uint64_t)(2 * k), idesc, accum);
                ++cnt;
              }
            }
            if (w_last) umma_commit(w_empty(wi));     // frees the weight slot once these MMAs have read it
            if (!p.strip || s == p.S - 1) umma_commit(a_empty(ai));
            }
            __syncwarp();
            if (w_last) { if (++wi == p.NW) { wi = 0; wph ^= 1u; } }
            if (!p.strip || s == p.S - 1) {
              if (++ai == p.NA) { ai = 0; aph ^= 1u; }
            }
          }
      if (leader) umma_commit(tmem_full(acc));
      __syncwarp();
      }   // work loop
    }
  } else {
    // ===== epilogue: warps 2..9, TMEM lane quadrant = warp % 4; the two warps of a quadrant take alternate 16-column chunks.
    // The epilogue is instruction-bound (ncu source view, c0_cat_refine: 9 100 warp-instructions per 128-pixel tile, a third of
    // them parameter reloads, index arithmetic and branches), so every launch parameter it needs is hoisted into registers here,
    // per-tile addresses are formed once, and the chunk loop steps by a constant. =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int H = p.H, W = p.W, HW = H * W, Cout = p.Cout, BN = p.BN, KS = p.KS, BW = p.BW, BH = p.BH, tiles_x = p.tiles_x;
    const int m_tiles = p.m_tiles, n_tiles = p.n_tiles, num_work = p.num_work;
    const bool two_acc = p.acc_stages == 2, relu = p.relu != 0, sigm = p.sigmoid != 0;
    const float inv_scale = p.inv_scale;
    const float* __restrict__ bias = p.bias;
    __half* const out_hi = p.out_hi; __half* const out_lo = p.out_lo;
    float* const out32 = p.out32;
    const __half* const res_hi = p.res_hi; const __half* const res_lo = p.res_lo;
    const uint8_t* const mask = p.mask;
    const int ry = row / BW, rx = row - ry * BW;                       // this lane's pixel inside the tile
    const uint32_t lane_base = ((uint32_t)(quad * 32) << 16);
    int it = 0;
    for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++it) {
    const int m_first = (work / n_tiles) * MT;
    const int n0 = (work % n_tiles) * BN;
    const int acc = two_acc ? (it & 1) : 0;
    const uint32_t acc_phase = two_acc ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u);
    mbar_wait(tmem_full(acc), acc_phase);
    tc_fence_after();
    const uint32_t acc_base = tmem_base + (uint32_t)acc * acc_cols + lane_base;
    for (int mt = 0; mt < MT; ++mt) {
      const int m = m_first + mt;
      if (m >= m_tiles) break;
      const int n = m / tiles_per_img, rem = m - n * tiles_per_img;
      const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
      const int oy = ty * BH + ry, ox = tx * BW + rx;
      const bool valid = oy < H && ox < W;
      const long long pix = (long long)n * HW + (long long)oy * W + ox;
      const bool keep = valid && (mask == nullptr || mask[pix] != 0);
      const bool has_res = res_hi != nullptr && valid;
      // per-tile bases: channel n0 of this lane's pixel
      const long long obase = pix * Cout + n0;
      float* o32 = out32 != nullptr ? out32 + ((long long)n * Cout + n0) * HW + (long long)oy * W + ox : nullptr;
      const uint32_t tbase = acc_base + (uint32_t)(mt * KS * BN);
      const int c_first = half * 16;
      // the residual and the accumulator columns of chunk i + 1 are requested before chunk i is processed (a lane-per-pixel
      // 32-byte residual load per plane waits ~1 us on HBM; the TMEM round trip is shorter but was equally exposed)
      uint4 rh_next[2], rl_next[2];
      uint32_t raw_next[16];
      if (has_res && c_first < BN && n0 + c_first + 16 <= Cout) {
        ld_global_nc_v8(res_hi + obase + c_first, rh_next[0], rh_next[1]);
        ld_global_nc_v8(res_lo + obase + c_first, rl_next[0], rl_next[1]);
      }
      if (KS == 1 && c_first < BN) tmem_ld16_nowait(tbase + (uint32_t)c_first, raw_next);
      for (int c0 = c_first; c0 < BN; c0 += 32) {
        float v[16];
        const uint4 rhv[2] = {rh_next[0], rh_next[1]}, rlv[2] = {rl_next[0], rl_next[1]};
        const int cn = c0 + 32;
        if (has_res && cn < BN && n0 + cn + 16 <= Cout) {
          ld_global_nc_v8(res_hi + obase + cn, rh_next[0], rh_next[1]);
          ld_global_nc_v8(res_lo + obase + cn, rl_next[0], rl_next[1]);
        }
        if (KS == 1) {
          tmem_ld_wait16(raw_next);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw_next[j]);
          if (cn < BN) tmem_ld16_nowait(tbase + (uint32_t)cn, raw_next);
        } else {
          // the KS partial accumulators of a narrow-N tile: all loads are issued before the single wait
          uint32_t part[4][16];
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < KS) tmem_ld16_nowait(tbase + (uint32_t)(ks * BN + c0), part[ks]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(part[0][j]);
#pragma unroll
          for (int ks = 1; ks < 4; ++ks)
            if (ks < KS) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(part[ks][j]);
            }
        }
        const int co0 = n0 + c0;
        if (!valid || co0 >= Cout) continue;
        if (co0 + 16 <= Cout) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + co0 + j));
            v[j] = fmaf(v[j], inv_scale, b.x); v[j + 1] = fmaf(v[j + 1], inv_scale, b.y);
            v[j + 2] = fmaf(v[j + 2], inv_scale, b.z); v[j + 3] = fmaf(v[j + 3], inv_scale, b.w);
          }
          if (has_res) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const __half2* ah = reinterpret_cast<const __half2*>(&rhv[q]);
              const __half2* bh = reinterpret_cast<const __half2*>(&rlv[q]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 fa = __half22float2(ah[e]), fb = __half22float2(bh[e]);
                v[q * 8 + 2 * e] += fa.x + fb.x; v[q * 8 + 2 * e + 1] += fa.y + fb.y;
              }
            }
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (sigm) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
          }
          if (mask != nullptr && !keep) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
          if (out_hi != nullptr) {
            uint4 hi4[2], lo4[2];
            uint32_t* hh = reinterpret_cast<uint32_t*>(hi4);
            uint32_t* ll = reinterpret_cast<uint32_t*>(lo4);
            if (out_lo != nullptr) {
#pragma unroll
              for (int e = 0; e < 8; ++e) split_f16x2(v[2 * e], v[2 * e + 1], hh[e], ll[e]);
              st_global_v8(out_hi + obase + c0, hi4[0], hi4[1]);
              st_global_v8(out_lo + obase + c0, lo4[0], lo4[1]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) hh[e] = f16x2_sat(v[2 * e], v[2 * e + 1]);
              st_global_v8(out_hi + obase + c0, hi4[0], hi4[1]);
            }
          }
          if (o32 != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o32[(long long)(c0 + j) * HW] = v[j];
          }
        } else {
          // ragged tail of output channels (head convs: Cout = 5 / 10 / 40): fp32 NCHW output only
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = co0 + j;
            if (co < Cout) {
              float x = fmaf(v[j], inv_scale, __ldg(bias + co));
              if (relu) x = fmaxf(x, 0.f);
              if (sigm) x = 1.f / (1.f + expf(-x));
              if (!keep) x = 0.f;
              if (o32 != nullptr) o32[(long long)(c0 + j) * HW] = x;
            }
          }
        }
      }
    }
    // this warp has read its quadrant of the accumulator buffer: hand it back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty(acc)) : "memory");
    }   // work loop
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- CTA-pair variant (cta_group::2) ------------------------------------------------------------------------------
// First-layer head convs (single pass, ReLU, hi-plane NHWC output, no residual, 7x7 with N = 192 / 256).  In the one-CTA kernel
// every MMA re-reads A (4 KiB) AND the full weight tile (N x 32 B) from shared memory while TMA refills the weight ring at
// ~0.55 wavefronts/clk: ~1.4 shared-memory wavefronts per clock are requested where the pipe delivers one (ncu: 54-62 %
// tensor-pipe active at c0/c1).  Two CTAs of a cluster run ONE M = 256 MMA: each CTA contributes its own 128-pixel A tile and
// HALF of the weight rows, so per SM the weight reads, the weight TMA fill and the L2 weight traffic all halve.
//   - both CTAs run a TMA producer (own A tile + own half of W); every load signals the LEADER's (rank 0) full barriers
//   - only the leader issues tcgen05.mma.cta_group::2; its commits are multicast to the empty / tmem_full barriers of both CTAs
//   - both CTAs run the epilogue on their own TMEM (rows 0-127 / 128-255 of D); all 16 epilogue warps release the leader's
//     tmem_empty barrier
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) tc_conv2_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const uint32_t a_slot = p.a_tile_bytes;
  const uint32_t w_tile = (uint32_t)(p.BN / 2) * 128u;               // this CTA's half of the weight tile
  const uint32_t w_slot = w_tile;
  const uint32_t a_ring = smem0, w_ring = a_ring + (uint32_t)p.NA * a_slot;
  const uint32_t bars = w_ring + (uint32_t)p.NW * w_slot;
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (p.NA + s); };
  auto w_full = [&](int s) { return bars + 16u * p.NA + 8u * s; };
  auto w_empty = [&](int s) { return bars + 16u * p.NA + 8u * (p.NW + s); };
  const uint32_t bar_tmem = bars + 16u * (p.NA + p.NW);
  auto tmem_full = [&](int a) { return bar_tmem + 8u * a; };
  auto tmem_empty = [&](int a) { return bar_tmem + 16u + 8u * a; };
  const uint32_t tmem_slot = bar_tmem + 32u;

  if (warp == 0 && lane == 0) { prefetch_tmap(&p.a_map[0][0]); prefetch_tmap(&p.w_map[0]); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.NA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.NW; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 16); }   // 8 epilogue warps of each CTA
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                                 // barriers of both CTAs are initialised before any remote signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int chunks = p.chunks0;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t acc_cols = (uint32_t)p.BN;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own pixel tile, own half of the weight rows; completion -> leader's full barriers =====
    const bool leader = elect_one();
    int ai = 0, wi = 0; uint32_t aph = 0, wph = 0;
    for (int work = cluster_id; work < p.num_work; work += n_clusters) {
      const int m = (work / p.n_tiles) * 2 + (int)rank;
      const int n0 = (work % p.n_tiles) * p.BN + (int)rank * (p.BN / 2);
      int tn, ty0, tx0;
      if (m < p.m_tiles) {
        const int n = m / tiles_per_img, rem = m - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        tn = n; ty0 = ty * p.BH; tx0 = tx * p.BW;
      } else { tn = p.N; ty0 = 0; tx0 = 0; }                          // out-of-range tile: TMA zero-fills
      for (int r = 0; r < p.R; ++r)
        for (int ch = 0; ch < chunks; ++ch) {
          const int c = p.coff0 + ch * TC_BK;
          for (int s = 0; s < p.S; ++s) {
            if (!p.strip || s == 0) {
              mbar_wait(a_empty(ai), aph ^ 1u);
              if (leader) {
                if (is_leader) mbar_expect_tx(a_full(ai), 2u * p.a_tx_bytes);       // both CTAs' tiles land on the leader's barrier
                tma_load_4d_2sm(a_ring + (uint32_t)ai * a_slot, &p.a_map[0][0], c, tx0 + (p.strip ? 0 : s) - p.pad, ty0 + r - p.pad, tn,
                                mapa_cluster(a_full(ai), 0));
              }
              __syncwarp();
              if (++ai == p.NA) { ai = 0; aph ^= 1u; }
            }
            mbar_wait(w_empty(wi), wph ^ 1u);
            if (leader) {
              if (is_leader) mbar_expect_tx(w_full(wi), 2u * w_slot);
              tma_load_3d_2sm(w_ring + (uint32_t)wi * w_slot, &p.w_map[0], ch * TC_BK, n0, r * p.S + s, mapa_cluster(w_full(wi), 0));
            }
            __syncwarp();
            if (++wi == p.NW) { wi = 0; wph ^= 1u; }
          }
        }
    }
  } else if (warp == 1) {
    if (is_leader) {
      // ===== MMA issuer (leader CTA only): M = 256 over the pair =====
      const bool leader = elect_one();
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // f16 x f16 -> f32, K-major A and B
      int ai = 0, wi = 0; uint32_t aph = 0, wph = 0;
      int it = 0;
      for (int work = cluster_id; work < p.num_work; work += n_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tmem_empty(acc), acc_phase ^ 1u);                  // both CTAs' epilogues have drained this accumulator buffer
        tc_fence_after();
        const uint32_t acc_base = tmem_base + (uint32_t)acc * acc_cols;
        uint32_t cnt = 0;
        for (int r = 0; r < p.R; ++r)
          for (int ch = 0; ch < chunks; ++ch)
            for (int s = 0; s < p.S; ++s) {
              if (!p.strip || s == 0) mbar_wait(a_full(ai), aph);
              mbar_wait(w_full(wi), wph);
              tc_fence_after();
              const uint64_t adesc = umma_desc(a_ring + (uint32_t)ai * a_slot + (p.strip ? (uint32_t)s * 128u : 0u));
              const uint64_t bdesc = umma_desc(w_ring + (uint32_t)wi * w_slot);
              if (leader) {
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                  umma_f16_2sm(acc_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, cnt > 0 ? 1u : 0u);
                  ++cnt;
                }
                umma_commit_2sm(w_empty(wi));                         // frees the weight slot in both CTAs
                if (!p.strip || s == p.S - 1) umma_commit_2sm(a_empty(ai));
              } else {
                cnt += TC_BK / 16;
              }
              __syncwarp();
              if (++wi == p.NW) { wi = 0; wph ^= 1u; }
              if (!p.strip || s == p.S - 1) { if (++ai == p.NA) { ai = 0; aph ^= 1u; } }
            }
        if (leader) umma_commit_2sm(tmem_full(acc));
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue (both CTAs): warps 2..9, TMEM lane quadrant = warp % 4, alternate 16-column chunks per half =====
    struct { int H, W, BH, BW, tiles_x, m_tiles, n_tiles, BN, Cout, relu, num_work; float inv_scale; const float* bias; __half* out_hi; } const P_ =
        {p.H, p.W, p.BH, p.BW, p.tiles_x, p.m_tiles, p.n_tiles, p.BN, p.Cout, p.relu, p.num_work, p.inv_scale, p.bias, p.out_hi};   // registers, not parameter-bank reloads
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int HW = P_.H * P_.W;
    int it = 0;
    for (int work = cluster_id; work < P_.num_work; work += n_clusters, ++it) {
      const int m = (work / P_.n_tiles) * 2 + (int)rank;
      const int n0 = (work % P_.n_tiles) * P_.BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tmem_full(acc), acc_phase);
      tc_fence_after();
      const uint32_t acc_base = tmem_base + (uint32_t)acc * acc_cols;
      if (m < P_.m_tiles) {
        const int n = m / tiles_per_img, rem = m - n * tiles_per_img;
        const int ty = rem / P_.tiles_x, tx = rem - ty * P_.tiles_x;
        const int oy = ty * P_.BH + row / P_.BW, ox = tx * P_.BW + row % P_.BW;
        const bool valid = oy < P_.H && ox < P_.W;
        const long long pix = (long long)n * HW + (long long)oy * P_.W + ox;
        for (int c0 = half * 16; c0 < P_.BN; c0 += 32) {
          uint32_t raw[16];
          tmem_ld16(acc_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, raw);
          const int co0 = n0 + c0;
          if (!valid || co0 >= P_.Cout) continue;
          uint4 hi4[2];
          uint32_t* hh = reinterpret_cast<uint32_t*>(hi4);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(P_.bias + co0 + j));
            const float v0 = fmaf(__uint_as_float(raw[j]), P_.inv_scale, b.x), v1 = fmaf(__uint_as_float(raw[j + 1]), P_.inv_scale, b.y);
            const float v2 = fmaf(__uint_as_float(raw[j + 2]), P_.inv_scale, b.z), v3 = fmaf(__uint_as_float(raw[j + 3]), P_.inv_scale, b.w);
            // ReLU + clamp to the finite fp16 range + convert: one F2FP per pair
            hh[j / 2] = P_.relu ? f16x2_relu_sat(v0, v1) : f16x2_sat(v0, v1);
            hh[j / 2 + 1] = P_.relu ? f16x2_relu_sat(v2, v3) : f16x2_sat(v2, v3);
          }
          st_global_v8(P_.out_hi + pix * P_.Cout + co0, hi4[0], hi4[1]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(tmem_empty(acc), 0));   // release on the LEADER's barrier
    }
  }
  tc_fence_before();
  cluster_sync_all();                                                 // both CTAs are done with TMEM and with each other's barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_tc_state = 0;   // 0 unknown, 1 ok, -1 unavailable
static int g_num_sms = 148;
static char g_tc_msg[256] = "not initialised";

static void tc_init() {
  if (g_tc_state != 0) return;
  const char* off = getenv("KG_DISABLE_TC");
  if (off && off[0] == '1') { g_tc_state = -1; snprintf(g_tc_msg, sizeof(g_tc_msg), "disabled by KG_DISABLE_TC"); return; }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    g_tc_state = -1; snprintf(g_tc_msg, sizeof(g_tc_msg), "no CUDA device"); return;
  }
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10) { g_tc_state = -1; snprintf(g_tc_msg, sizeof(g_tc_msg), "compute capability %d.x is not sm_100", major); return; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
      qres != cudaDriverEntryPointSuccess) {
    g_tc_state = -1; snprintf(g_tc_msg, sizeof(g_tc_msg), "cuTensorMapEncodeTiled not found in the driver"); return;
  }
  g_encode = (EncodeTiledFn)fn;
  if (cudaFuncSetAttribute(tc_conv_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(tc_conv_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(tc_conv_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(tc_conv_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(tc_conv_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(tc_conv_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(tc_conv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess) {
    g_tc_state = -1; snprintf(g_tc_msg, sizeof(g_tc_msg), "cannot raise the dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
    return;
  }
  g_tc_state = 1; snprintf(g_tc_msg, sizeof(g_tc_msg), "ok");
}

bool tc_available() { tc_init(); return g_tc_state == 1; }
bool tc_stride2_enabled() { const char* v = getenv("KG_TC_STRIDE2"); return !(v && v[0] == '0'); }
const char* tc_status() { tc_init(); return g_tc_msg; }
void* tc_encode_tiled_fn() { tc_init(); return g_tc_state == 1 ? (void*)g_encode : nullptr; }
int tc_num_sms() { tc_init(); return g_num_sms; }
bool tc_layer_supported(int cin, int cout, int R, int S) { return cin % TC_BK == 0 && cout >= 1 && R >= 1 && S >= 1 && R * S <= 64; }

void tc_free_weights(TcWeights& w) {
  if (w.d_hi) cudaFree(w.d_hi);
  if (w.d_lo) cudaFree(w.d_lo);
  w = TcWeights();
}

int tc_pack_weights(const float* w, int cin, int cout, int R, int S, TcWeights* out) {
  tc_free_weights(*out);
  const int taps = R * S;
  const int cout_pad = cout <= 256 ? (int)align_up(cout, 16) : (int)align_up(cout, 256);
  float mx = 0.f;
  const size_t n = (size_t)taps * cin * cout;
  for (size_t i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(w[i]));
  int e = 0;
  if (mx > 0.f && std::isfinite(mx)) { int ex; frexpf(mx, &ex); e = 10 - ex; }   // max |w| * 2^e in [512, 1024)
  if (e > 40) e = 40;
  if (e < -40) e = -40;
  const float scale = ldexpf(1.f, e);
  std::vector<__half> hi((size_t)taps * cout_pad * cin, __float2half_rn(0.f)), lo(hi.size(), __float2half_rn(0.f));
  for (int t = 0; t < taps; ++t)
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co) {
        const float v = w[((size_t)t * cin + ci) * cout + co] * scale;
        const __half h = __float2half_rn(v);
        const size_t o = ((size_t)t * cout_pad + co) * cin + ci;
        hi[o] = h;
        lo[o] = __float2half_rn(v - __half2float(h));
      }
  KG_CUDA_CHECK(cudaMalloc(&out->d_hi, hi.size() * sizeof(__half)));
  KG_CUDA_CHECK(cudaMalloc(&out->d_lo, lo.size() * sizeof(__half)));
  KG_CUDA_CHECK(cudaMemcpy(out->d_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice));
  KG_CUDA_CHECK(cudaMemcpy(out->d_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice));
  out->cin = cin; out->cout = cout; out->cout_pad = cout_pad; out->taps = taps; out->inv_scale = ldexpf(1.f, -e);
  out->valid = true;
  return KG_OK;
}

static int encode_act_map(CUtensorMap* m, const __half* base, int C, int W, int H, int N, int BW, int BH, int stride = 1) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // strided convs: the box spans stride*B input pixels and TMA picks every stride-th one (elementStrides)
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(BW * stride), (cuuint32_t)(BH * stride), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations C=%d W=%d H=%d N=%d box %dx%d) failed: %d", C, W, H, N, BW, BH, (int)r); return KG_ERR_CUDA; }
  return KG_OK;
}

static int encode_w_map(CUtensorMap* m, const __half* base, int cin, int cout_pad, int taps, int BN, int wg) {
  cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout_pad, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cout_pad * cin * 2};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)BN, (cuuint32_t)wg};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights cin=%d cout=%d taps=%d BN=%d) failed: %d", cin, cout_pad, taps, BN, (int)r); return KG_ERR_CUDA; }
  return KG_OK;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

int tc_conv_prepare(TcConvOp* op) {
  if (!tc_available()) { set_error("tensor-core path unavailable: %s", tc_status()); return KG_ERR_STATE; }
  KG_REQUIRE(op && op->w && op->w->valid, "tc_conv_prepare: weights not packed");
  KG_REQUIRE(op->C0 % TC_BK == 0 && op->C1 % TC_BK == 0 && op->C0 + op->C1 == op->w->cin, "tc_conv_prepare: channel split %d+%d vs cin %d",
             op->C0, op->C1, op->w->cin);
  KG_REQUIRE(op->passes >= 1 && op->passes <= 3, "tc_conv_prepare: passes=%d", op->passes);
  KG_REQUIRE(op->passes == 1 || (op->in0_lo != nullptr && (op->C1 == 0 || op->in1_lo != nullptr)), "tc_conv_prepare: 2- / 3-pass need lo planes");
  KG_REQUIRE(op->out_hi == nullptr || op->Cout % 16 == 0, "tc_conv_prepare: NHWC output needs Cout %% 16 == 0 (Cout=%d)", op->Cout);
  std::shared_ptr<TcParams> sp(new TcParams());
  TcParams& p = *sp;
  memset(&p, 0, sizeof(p));
  p.N = op->N; p.H = op->H; p.W = op->W; p.Cout = op->Cout; p.R = op->R; p.S = op->S; p.pad = op->pad;
  p.stride = op->stride;
  const int Hin = op->stride == 1 ? op->H : op->Hin, Win = op->stride == 1 ? op->W : op->Win;
  KG_REQUIRE(op->stride == 1 || (op->stride == 2 && op->C1 == 0 && Hin > 0 && Win > 0), "tc_conv_prepare: unsupported stride %d", op->stride);
  p.chunks0 = op->C0 / TC_BK; p.chunks1 = op->C1 / TC_BK; p.coff0 = op->in0_coff;
  p.BN = op->w->cout_pad <= 256 ? op->w->cout_pad : 256;
  int bw = 1;
  while (bw * 2 <= op->W && bw * 2 <= TC_BM) bw *= 2;
  p.BW = bw; p.BH = TC_BM / bw;
  p.tiles_x = ceil_div(op->W, p.BW); p.tiles_y = ceil_div(op->H, p.BH);
  p.m_tiles = p.tiles_x * p.tiles_y * op->N;
  p.passes = op->passes; p.NPL = op->passes >= 2 ? 2 : 1;      // NPL: activation planes; weight planes: npw below
  const int npw = op->passes == 3 ? 2 : 1;
  // TMEM budget: acc_stages x MT x BN fp32 columns <= 512.  Two accumulator buffers let the epilogue of one tile overlap
  // the MMAs of the next; MT = 2 (two pixel tiles share each weight tile) only where that still fits.
  p.acc_stages = env_int("KG_TC_ACC", 2) >= 2 ? 2 : 1;
  int mt = (op->passes == 1 && p.acc_stages * 2 * p.BN <= 512) ? 2 : 1;
  // deep-K single-pass layers with a full 256-wide N tile (first-layer heads at c2 / c3): the weight stream from L2 is the
  // limiter, so two pixel tiles share every weight tile (MT = 2) at the price of a single accumulator buffer
  // (measured: c2 7.0 -> 6.7 ms, c3 7.5 -> 6.7 ms)
  if (op->passes == 1 && p.BN == 256 && op->R * op->S * (op->C0 + op->C1) >= 49 * 256 && getenv("KG_TC_ACC") == nullptr &&
      getenv("KG_TC_MT") == nullptr) { p.acc_stages = 1; mt = 2; }
  mt = env_int("KG_TC_MT", mt);
  if (mt < 1) mt = 1;
  if (mt > 2) mt = 2;
  if (p.acc_stages * mt * p.BN > 512) mt = 1;
  p.MT = mt;
  p.KS = 1;
  // strip mode: tile = one 128-pixel image row, the S taps of a filter row share one (128 + S - 1)-pixel strip
  const bool strip = op->stride == 1 && op->S > 1 && p.BW == TC_BM && env_int("KG_TC_STRIP", 1) != 0;
  p.strip = strip ? 1 : 0;
  p.base_offset_mode = env_int("KG_TC_BASEOFF", 0);
  const int strip_px = strip ? TC_BM + op->S - 1 : TC_BM;
  p.a_tx_bytes = (unsigned)(strip ? strip_px * 128 : TC_A_TILE);
  p.a_tile_bytes = (unsigned)align_up(p.a_tx_bytes, 1024);
  // (A TMA-store epilogue through a 64 KiB staging area existed in round 1 (KG_TC_OTMA); measured slower -- decoder 12.6 -> 17.7 ms,
  // the ring loses 64 KiB -- and removed.)
  const size_t budget = TC_MAX_SMEM - 2048;
  // weight tap group: all S taps of a filter row in one W slot when that slot stays small (narrow-N layers)
  p.wg = 1;
  if (op->S > 1 && (size_t)op->S * npw * p.BN * 128 <= (size_t)env_int("KG_TC_WG_MAXKB", 48) * 1024 && env_int("KG_TC_WG", 1) != 0) p.wg = op->S;
  auto fit = [&](int mtv, int* na, int* nw) {
    const size_t a_slot = (size_t)mtv * p.NPL * p.a_tile_bytes, w_slot = (size_t)p.wg * npw * p.BN * 128;
    if (strip || p.wg > 1) {
      // A slots and W slots are consumed at different rates (one A strip per filter row / one W group per filter row)
      *na = strip ? 2 : 4;
      if ((size_t)*na * a_slot + 2 * w_slot > budget) { *na = 2; if (2 * a_slot + 2 * w_slot > budget) return false; }
      *nw = (int)((budget - (size_t)*na * a_slot) / w_slot);
      if (*nw > 8) *nw = 8;
      const int na_more = strip ? 3 : 8;
      if (*nw >= (p.wg > 1 ? 3 : 6)) {
        // spend spare room on more A slots while keeping >= 3 (grouped) / 4 W slots
        const int keep_w = p.wg > 1 ? 3 : 4;
        int na2 = (int)((budget - (size_t)keep_w * w_slot) / a_slot);
        if (na2 > na_more) na2 = na_more;
        if (na2 > *na) { *na = na2; *nw = (int)((budget - (size_t)*na * a_slot) / w_slot); if (*nw > 8) *nw = 8; }
      }
      return *nw >= 2;
    }
    int st = (int)(budget / (a_slot + w_slot));
    if (st > 8) st = 8;
    *na = *nw = st;
    return st >= 2;
  };
  int na = 0, nw = 0;
  if (!fit(p.MT, &na, &nw) && p.MT == 2) { p.MT = 1; }
  if (!fit(p.MT, &na, &nw) && p.wg > 1) { p.wg = 1; }
  if (!fit(p.MT, &na, &nw)) {
    const size_t a_slot = (size_t)p.MT * p.NPL * p.a_tile_bytes, w_slot = (size_t)p.wg * npw * p.BN * 128;
    KG_REQUIRE(a_slot + w_slot <= budget, "tc_conv_prepare: tile does not fit in shared memory");
    na = nw = 1;
  }
  const int cap = env_int("KG_TC_STAGES", 0);
  if (cap > 0) { if (na > cap) na = cap; if (nw > cap) nw = cap; }
  p.NA = na; p.NW = nw;
  {
    int ks = 512 / (p.acc_stages * p.MT * p.BN);
    int want = p.BN <= 64 ? 4 / p.MT : (p.BN <= 128 ? 2 : 1);           // target ~4 independent chains when N is small
    if (p.BN <= 64) want = env_int("KG_TC_KSWANT", want);
    if (ks > want) ks = want;
    const int ks_env = env_int("KG_TC_KS", 0);
    if (ks_env > 0 && ks_env < ks) ks = ks_env;
    if (ks < 1) ks = 1;
    const int total_mmas = op->R * op->S * (p.chunks0 + p.chunks1) * 4 * p.passes;
    while (ks > 1 && total_mmas < ks) --ks;                               // every chain must receive at least one MMA
    while (ks & (ks - 1)) --ks;                                           // power of two (chain = counter & (KS - 1))
    if (ks > 4) ks = 4;                                                   // the epilogue sums at most four partial accumulators
    p.KS = ks;
  }
  // CTA-pair kernel (tc_conv2_kernel): single-pass, hi-plane NHWC output, no residual / mask / second source, N = 192 or 256
  p.two_cta = (op->passes == 1 && (p.BN == 192 || p.BN == 256) && op->w->cout_pad % p.BN == 0 && op->C1 == 0 && op->stride == 1 &&
               op->res_hi == nullptr && op->mask == nullptr && op->out_hi != nullptr && op->out_lo == nullptr && !op->sigmoid &&
               op->Cout % 16 == 0 && p.m_tiles >= 2 && env_int("KG_TC_2CTA", 1) != 0) ? 1 : 0;
  if (p.two_cta) {
    p.MT = 1; p.KS = 1; p.acc_stages = 2; p.wg = 1;
    const size_t budget2 = TC_MAX_SMEM - 2048;
    const size_t a_slot = p.a_tile_bytes, w_slot = (size_t)(p.BN / 2) * 128;
    if (strip) { na = 3; nw = (int)((budget2 - 3 * a_slot) / w_slot); if (nw > 8) nw = 8; }
    else { int st = (int)(budget2 / (a_slot + w_slot)); if (st > 8) st = 8; na = nw = st; }
    KG_REQUIRE(na >= 2 && nw >= 2, "tc_conv_prepare: CTA-pair tile does not fit in shared memory");
    p.NA = na; p.NW = nw;
  }
  unsigned cols = 32;
  while (cols < (unsigned)(p.acc_stages * p.MT * p.KS * p.BN)) cols *= 2;
  p.tmem_cols = cols;
  p.n_tiles = op->w->cout_pad / p.BN;
  p.num_work = ceil_div(p.m_tiles, p.two_cta ? 2 : p.MT) * p.n_tiles;
  p.bias = op->bias; p.inv_scale = op->w->inv_scale;
  p.out_hi = op->out_hi; p.out_lo = op->out_lo; p.res_hi = op->res_hi; p.res_lo = op->res_lo;
  p.relu = op->relu; p.sigmoid = op->sigmoid; p.mask = op->mask;
  const int box_w = strip ? strip_px : p.BW;
  KG_TRY(encode_act_map(&p.a_map[0][0], op->in0_hi, op->in0_C, Win, Hin, op->N, box_w, p.BH, op->stride));
  if (p.NPL == 2) KG_TRY(encode_act_map(&p.a_map[0][1], op->in0_lo, op->in0_C, Win, Hin, op->N, box_w, p.BH, op->stride));
  if (op->C1 > 0) {
    KG_TRY(encode_act_map(&p.a_map[1][0], op->in1_hi, op->in1_C, op->W, op->H, op->N, box_w, p.BH));
    if (p.NPL == 2) KG_TRY(encode_act_map(&p.a_map[1][1], op->in1_lo, op->in1_C, op->W, op->H, op->N, box_w, p.BH));
  }
  KG_TRY(encode_w_map(&p.w_map[0], op->w->d_hi, op->w->cin, op->w->cout_pad, op->w->taps, p.two_cta ? p.BN / 2 : p.BN, p.wg));
  if (npw == 2) KG_TRY(encode_w_map(&p.w_map[1], op->w->d_lo, op->w->cin, op->w->cout_pad, op->w->taps, p.BN, p.wg));
  const int persist = env_int("KG_TC_CTAS", g_num_sms);
  op->grid_x = (unsigned)std::min(p.num_work, std::max(1, persist));
  if (p.two_cta) op->grid_x = (unsigned)std::min(2 * p.num_work, std::max(2, persist & ~1));   // whole CTA pairs
  op->grid_y = 1;
  op->smem_bytes = (unsigned)((size_t)p.NA * p.MT * p.NPL * p.a_tile_bytes + (size_t)p.NW * p.wg * npw * p.BN * 128 + 16 * (p.NA + p.NW) + 128 + 1024);
  if (p.two_cta) op->smem_bytes = (unsigned)((size_t)p.NA * p.a_tile_bytes + (size_t)p.NW * (p.BN / 2) * 128 + 16 * (p.NA + p.NW) + 128 + 1024);
  KG_REQUIRE(op->smem_bytes <= (unsigned)TC_MAX_SMEM, "tc_conv_prepare: smem %u > %d", op->smem_bytes, TC_MAX_SMEM);
  op->params = sp;
  if (env_int("KG_TC_DEBUG", 0))
    fprintf(stderr, "[tc] N%d %dx%d C%d+%d->%d k%dx%d s%d passes%d | BN%d BW%d BH%d MT%d KS%d acc%d strip%d wg%d 2cta%d NA%d NW%d work%d grid%u smem%u tmem%u\n",
            op->N, op->H, op->W, op->C0, op->C1, op->Cout, op->R, op->S, op->stride, op->passes, p.BN, p.BW, p.BH, p.MT, p.KS, p.acc_stages,
            p.strip, p.wg, p.two_cta, p.NA, p.NW, p.num_work, op->grid_x, op->smem_bytes, p.tmem_cols);
  return KG_OK;
}

int tc_conv_launch(const TcConvOp* op, float* out32, cudaStream_t stream) {
  KG_REQUIRE(op && op->params, "tc_conv_launch: op not prepared");
  TcParams p = *reinterpret_cast<const TcParams*>(op->params.get());
  p.out32 = out32;
  dim3 grid(op->grid_x, op->grid_y, 1);
  if (p.two_cta && out32 == nullptr) {
    tc_conv2_kernel<<<grid, TC_THREADS, op->smem_bytes, stream>>>(p);
    KG_CUDA_CHECK(cudaGetLastError());
    return KG_OK;
  }
  KG_REQUIRE(!p.two_cta, "tc_conv_launch: the CTA-pair kernel has no fp32 NCHW output");
  if (p.MT == 1 && p.passes == 1) tc_conv_kernel<1, 1><<<grid, TC_THREADS, op->smem_bytes, stream>>>(p);
  else if (p.MT == 2 && p.passes == 1) tc_conv_kernel<2, 1><<<grid, TC_THREADS, op->smem_bytes, stream>>>(p);
  else if (p.MT == 1 && p.passes == 2) tc_conv_kernel<1, 2><<<grid, TC_THREADS, op->smem_bytes, stream>>>(p);
  else if (p.MT == 2 && p.passes == 2) tc_conv_kernel<2, 2><<<grid, TC_THREADS, op->smem_bytes, stream>>>(p);
  else if (p.MT == 1 && p.passes == 3) tc_conv_kernel<1, 3><<<grid, TC_THREADS, op->smem_bytes, stream>>>(p);
  else tc_conv_kernel<2, 3><<<grid, TC_THREADS, op->smem_bytes, stream>>>(p);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

}  // namespace kg
