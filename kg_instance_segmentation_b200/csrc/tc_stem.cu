// The two convs that read the raw image, on the tensor cores: c0_conv.0 (3 -> 64, 3x3 / s1, KGnet.py:139-141) and
// conv1 + bn1 (3 -> 64, 7x7 / s2, KGnet.py:131-133, 278-280), ReLU fused, fp32 NCHW in, split-fp16 NHWC out.
//
// Cin = 3 rules out the TMA implicit GEMM (a 6-byte pixel is below TMA's 16-byte granule), and on CUDA cores these two
// layers cost 27 x 64 / 147 x 64 FFMAs per pixel.  Here every thread builds ONE row of the im2col tile directly in shared
// memory -- K = taps x 3 padded to 32 / 160, split into fp16 hi / lo planes, in the 128-byte-swizzled K-major layout
// tcgen05.mma expects -- and one thread issues the (2 or 10 k-steps) x 3 split passes of M = 128 x N = 64 MMAs.  The
// weights sit in shared memory for the whole CTA as a pre-swizzled image built on the host.  No warp specialisation:
// build -> MMA -> epilogue run back to back, several CTAs per SM overlap each other's phases.
#include "tc_stem.cuh"
#include "tc_conv.cuh"
#include "tc_ptx.cuh"

#include <cmath>
#include <vector>

namespace kg {

constexpr int ST_THREADS = 128;

struct StemParams {
  const float* x; const __half* w_img; const float* bias;
  __half* out_hi; __half* out_lo;
  int N, H, W, Ho, Wo;
  long long total;         // output pixels
  int tiles;
};

// K x K conv, stride STRIDE, 3 input channels.  k index = (r * K + s) * 3 + c, padded to KP (a multiple of 16).
template <int K, int STRIDE>
__global__ void __launch_bounds__(ST_THREADS) tc_stem_kernel(const StemParams p) {
  constexpr int KR = K * K * 3;
  constexpr int KP = (KR + 15) / 16 * 16;
  constexpr int NCH = (KP + 63) / 64;              // 64-element K chunks (one 128-byte swizzle row each)
  constexpr int PAD = K / 2;
  constexpr uint32_t A_PLANE = NCH * 16384u, W_PLANE = NCH * 8192u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t smem0 = (smem_base + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem0 - smem_base);
  const uint32_t a_smem = smem0;                                   // [plane][chunk][128 rows x 128 B]
  const uint32_t w_smem = smem0 + 2u * A_PLANE;                    // [plane][chunk][64 rows x 128 B]
  const uint32_t bar = w_smem + 2u * W_PLANE;
  const uint32_t tmem_slot = bar + 8u;
  float* s_bias = reinterpret_cast<float*>(sm + 2u * A_PLANE + 2u * W_PLANE + 16u);
  const int t = threadIdx.x, warp = t >> 5;

  // weights: the host-built swizzled image, copied verbatim
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img);
    uint4* dst = reinterpret_cast<uint4*>(sm + 2u * A_PLANE);
    for (int e = t; e < (int)(2u * W_PLANE / 16u); e += ST_THREADS) dst[e] = __ldg(src + e);
    if (t < 64) s_bias[t] = p.bias[t];
  }
  if (t == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32, K-major, M = 128, N = 64
  const uint32_t sw = (uint32_t)(t & 7);
  const uint32_t a_row = a_smem + (uint32_t)t * 128u;

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
    // ---- build this thread's im2col row (pixel g) ----
    const long long g = (long long)tile * 128 + t;
    const bool live = g < p.total;
    int ox = 0, oy = 0, n = 0;
    if (live) { ox = (int)(g % p.Wo); oy = (int)((g / p.Wo) % p.Ho); n = (int)(g / ((long long)p.Wo * p.Ho)); }
    const float* xn = p.x + (long long)n * 3 * p.H * p.W;
    const int iy0 = oy * STRIDE - PAD, ix0 = ox * STRIDE - PAD;
#pragma unroll
    for (int j = 0; j < KP / 8; ++j) {                             // 8 consecutive k = one 16-byte piece of the row
      uint4 h4, l4;
      __half2* hh = reinterpret_cast<__half2*>(&h4);
      __half2* ll = reinterpret_cast<__half2*>(&l4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int k = j * 8 + 2 * e + q;                         // compile-time after unrolling
          v[q] = 0.f;
          if (k < KR) {
            const int tap = k / 3, c = k - tap * 3, r = tap / K, s = tap - r * K;
            const int iy = iy0 + r, ix = ix0 + s;
            if (live && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v[q] = __ldg(xn + ((long long)c * p.H + iy) * p.W + ix);
          }
        }
        const __half2 h = __floats2half2_rn(v[0], v[1]);
        const float2 hf = __half22float2(h);
        hh[e] = h;
        ll[e] = __floats2half2_rn(v[0] - hf.x, v[1] - hf.y);
      }
      const uint32_t off = (uint32_t)(j >> 3) * 16384u + ((((uint32_t)(j & 7)) ^ sw) << 4);
      st_shared_v4(a_row + off, h4);
      st_shared_v4(a_row + A_PLANE + off, l4);
    }
    fence_proxy_async_smem();                                      // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    // ---- MMAs: 3 split passes (hi*hi, lo*hi, hi*lo) per 16-wide k-step ----
    if (t == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < KP / 16; ++ks) {
        const uint32_t ch = (uint32_t)(ks >> 2), kk = (uint32_t)(ks & 3);
        const uint64_t a_hi = umma_desc(a_smem + ch * 16384u) + (uint64_t)(2 * kk), a_lo = umma_desc(a_smem + A_PLANE + ch * 16384u) + (uint64_t)(2 * kk);
        const uint64_t w_hi = umma_desc(w_smem + ch * 8192u) + (uint64_t)(2 * kk), w_lo = umma_desc(w_smem + W_PLANE + ch * 8192u) + (uint64_t)(2 * kk);
        umma_f16(tmem_base, a_hi, w_hi, idesc, ks == 0 ? 0u : 1u);
        umma_f16(tmem_base, a_lo, w_hi, idesc, 1u);
        umma_f16(tmem_base, a_hi, w_lo, idesc, 1u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, it & 1u);
    tc_fence_after();
    // ---- epilogue: thread = output pixel (TMEM lane), bias + ReLU, split fp16, 128 B per plane ----
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t raw[16];
      tmem_ld16(lane_addr + (uint32_t)c0, raw);
      if (live) {
        uint4 hi4[2], lo4[2];
        __half2* hh = reinterpret_cast<__half2*>(hi4);
        __half2* ll = reinterpret_cast<__half2*>(lo4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a = fminf(fmaxf(__uint_as_float(raw[2 * e]) + s_bias[c0 + 2 * e], 0.f), 65504.f);
          const float b = fminf(fmaxf(__uint_as_float(raw[2 * e + 1]) + s_bias[c0 + 2 * e + 1], 0.f), 65504.f);
          const __half2 h = __floats2half2_rn(a, b);
          const float2 hf = __half22float2(h);
          hh[e] = h;
          ll[e] = __floats2half2_rn(a - hf.x, b - hf.y);
        }
        st_global_v8(p.out_hi + g * 64 + c0, hi4[0], hi4[1]);
        st_global_v8(p.out_lo + g * 64 + c0, lo4[0], lo4[1]);
      }
    }
    tc_fence_before();
    __syncthreads();                                               // accumulator and A tile are free for the next tile
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

// ---- uint8 input ------------------------------------------------------------------------------------------------------------------
// The same two convs fed by the camera image itself: uint8 NHWC (cv2 BGR), i.e. BEFORE `x / 255 - 0.5` (test.py:92).  With
// a = k - 128 (an integer in [-128, 127]: exact in fp16) the normalised pixel is (a + 0.5) / 255, hence
//     sum_taps w * x = sum_taps (w / 255) * a + 0.5 * sum_{taps inside the image} (w / 255)
// and a tap outside the image (zero padding of the NORMALISED image) is the exact fp16 value a = -0.5 when the constant
// 0.5 * sum_{all taps} (w / 255) is folded into the bias.  The im2col tile therefore needs ONE fp16 plane (no hi / lo split, no
// rounding at all on the activation side), two MMA passes (a x w_hi, a x w_lo) instead of three, half the shared memory (two CTAs
// per SM for the 7x7 kernel), and a filter row of the tile is 3 K contiguous bytes of the image instead of 3 K strided floats;
// the separate normalisation kernel and its fp32 NCHW tensor disappear.  w / 255 is scaled by a power of two into fp16's normal
// range (inv_scale undoes it in the epilogue).
struct StemU8Params {
  const uint8_t* img; const __half* w_img; const float* bias;
  __half* out_hi; __half* out_lo;
  int N, H, W, Ho, Wo;
  long long total;
  int tiles;
  float inv_scale;
};

template <int K, int STRIDE>
__global__ void __launch_bounds__(ST_THREADS) tc_stem_u8_kernel(const StemU8Params p) {
  constexpr int KR = K * K * 3;
  constexpr int KP = (KR + 15) / 16 * 16;
  constexpr int NCH = (KP + 63) / 64;
  constexpr int PAD = K / 2;
  constexpr uint32_t A_PLANE = NCH * 16384u, W_PLANE = NCH * 8192u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t smem0 = (smem_base + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (smem0 - smem_base);
  const uint32_t a_smem = smem0;                                   // [chunk][128 rows x 128 B]
  const uint32_t w_smem = smem0 + A_PLANE;                         // [plane][chunk][64 rows x 128 B]
  const uint32_t bar = w_smem + 2u * W_PLANE;
  const uint32_t tmem_slot = bar + 8u;
  float* s_bias = reinterpret_cast<float*>(sm + A_PLANE + 2u * W_PLANE + 16u);
  const int t = threadIdx.x, warp = t >> 5;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img);
    uint4* dst = reinterpret_cast<uint4*>(sm + A_PLANE);
    for (int e = t; e < (int)(2u * W_PLANE / 16u); e += ST_THREADS) dst[e] = __ldg(src + e);
    if (t < 64) s_bias[t] = p.bias[t];
  }
  if (t == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t sw = (uint32_t)(t & 7);
  const uint32_t a_row = a_smem + (uint32_t)t * 128u;
  const float inv_scale = p.inv_scale;

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
    const long long g = (long long)tile * 128 + t;
    const bool live = g < p.total;
    int ox = 0, oy = 0, n = 0;
    if (live) { ox = (int)(g % p.Wo); oy = (int)((g / p.Wo) % p.Ho); n = (int)(g / ((long long)p.Wo * p.Ho)); }
    const uint8_t* xn = p.img + (long long)n * p.H * p.W * 3;
    const int iy0 = oy * STRIDE - PAD, ix0 = ox * STRIDE - PAD;
    // a filter row of this pixel is 3 K contiguous bytes of the image (HWC): one base pointer and one validity bit per row / column,
    // every tap then is a byte load at a compile-time offset
    const uint8_t* rowp[K];
    unsigned row_ok = 0u, col_ok = 0u;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int iy = iy0 + r;
      const bool ok = live && iy >= 0 && iy < p.H;
      row_ok |= (ok ? 1u : 0u) << r;
      rowp[r] = xn + ((long long)(ok ? iy : 0) * p.W + ix0) * 3;
    }
#pragma unroll
    for (int s2 = 0; s2 < K; ++s2) col_ok |= ((ix0 + s2 >= 0 && ix0 + s2 < p.W) ? 1u : 0u) << s2;
#pragma unroll
    for (int j = 0; j < KP / 8; ++j) {
      uint4 h4;
      __half2* hh = reinterpret_cast<__half2*>(&h4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int k = j * 8 + 2 * e + q;                         // compile-time after unrolling
          v[q] = 0.f;                                              // K padding: the weight image is zero there
          if (k < KR) {
            const int tap = k / 3, c = k - tap * 3, r = tap / K, s2 = tap - r * K;
            const bool ok = ((row_ok >> r) & (col_ok >> s2) & 1u) != 0u;
            v[q] = ok ? (float)((int)__ldg(rowp[r] + 3 * s2 + c) - 128) : -0.5f;
          }
        }
        hh[e] = __floats2half2_rn(v[0], v[1]);                     // exact: integers of magnitude <= 128, or -0.5
      }
      const uint32_t off = (uint32_t)(j >> 3) * 16384u + ((((uint32_t)(j & 7)) ^ sw) << 4);
      st_shared_v4(a_row + off, h4);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < KP / 16; ++ks) {
        const uint32_t ch = (uint32_t)(ks >> 2), kk = (uint32_t)(ks & 3);
        const uint64_t a = umma_desc(a_smem + ch * 16384u) + (uint64_t)(2 * kk);
        const uint64_t w_hi = umma_desc(w_smem + ch * 8192u) + (uint64_t)(2 * kk), w_lo = umma_desc(w_smem + W_PLANE + ch * 8192u) + (uint64_t)(2 * kk);
        umma_f16(tmem_base, a, w_hi, idesc, ks == 0 ? 0u : 1u);
        umma_f16(tmem_base, a, w_lo, idesc, 1u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, it & 1u);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t raw[16];
      tmem_ld16(lane_addr + (uint32_t)c0, raw);
      if (live) {
        uint4 hi4[2], lo4[2];
        __half2* hh = reinterpret_cast<__half2*>(hi4);
        __half2* ll = reinterpret_cast<__half2*>(lo4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float a = fminf(fmaxf(fmaf(__uint_as_float(raw[2 * e]), inv_scale, s_bias[c0 + 2 * e]), 0.f), 65504.f);
          const float b = fminf(fmaxf(fmaf(__uint_as_float(raw[2 * e + 1]), inv_scale, s_bias[c0 + 2 * e + 1]), 0.f), 65504.f);
          const __half2 h = __floats2half2_rn(a, b);
          const float2 hf = __half22float2(h);
          hh[e] = h;
          ll[e] = __floats2half2_rn(a - hf.x, b - hf.y);
        }
        st_global_v8(p.out_hi + g * 64 + c0, hi4[0], hi4[1]);
        st_global_v8(p.out_lo + g * 64 + c0, lo4[0], lo4[1]);
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

template <int K>
static size_t stem_u8_smem_bytes() {
  constexpr int KP = (K * K * 3 + 15) / 16 * 16, NCH = (KP + 63) / 64;
  return 1024 + (size_t)NCH * 16384 + 2 * (size_t)NCH * 8192 + 16 + 64 * sizeof(float) + 64;
}

template <int K>
static size_t stem_smem_bytes() {
  constexpr int KP = (K * K * 3 + 15) / 16 * 16, NCH = (KP + 63) / 64;
  return 1024 + 2 * (size_t)NCH * 16384 + 2 * (size_t)NCH * 8192 + 16 + 64 * sizeof(float) + 64;
}

// host: the weight image as it sits in shared memory: [plane hi/lo][chunk][64 rows (cout) x 128 B], 16-byte piece j of
// row n stored at piece (j ^ (n & 7)) (128-byte swizzle), element k of the row = tap * 3 + c
int tc_stem_pack(const float* w_tap_cin_cout, int K, TcStemWeights* out) {
  KG_REQUIRE(w_tap_cin_cout && out && (K == 3 || K == 7), "tc_stem_pack: bad arguments");
  const int KR = K * K * 3, KP = (KR + 15) / 16 * 16, NCH = (KP + 63) / 64;
  const size_t plane = (size_t)NCH * 64 * 64;                      // halfs
  std::vector<__half> img(2 * plane, __float2half_rn(0.f));
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < KR; ++k) {
      const float v = w_tap_cin_cout[(size_t)k * 64 + n];          // [tap][cin=3][cout=64] == [k][n]
      const __half h = __float2half_rn(v);
      const int ch = k / 64, kk = k % 64, piece = kk / 8, e = kk % 8;
      const size_t o = (size_t)ch * 64 * 64 + (size_t)n * 64 + (size_t)((piece ^ (n & 7)) * 8 + e);
      img[o] = h;
      img[plane + o] = __float2half_rn(v - __half2float(h));
    }
  __half* d = nullptr;
  KG_CUDA_CHECK(cudaMalloc(&d, img.size() * sizeof(__half)));
  out->d_img = std::shared_ptr<void>(d, [](void* q) { cudaFree(q); });
  KG_CUDA_CHECK(cudaMemcpy(d, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  out->K = K;
  return KG_OK;
}

// uint8-input variant: the weight image holds w / 255 * 2^e (hi / lo), the bias the folded constant (see tc_stem_u8_kernel)
int tc_stem_pack_u8(const float* w_tap_cin_cout, const float* bias64, int K, TcStemWeights* out) {
  KG_REQUIRE(w_tap_cin_cout && bias64 && out && (K == 3 || K == 7), "tc_stem_pack_u8: bad arguments");
  const int KR = K * K * 3;
  double wmax = 0.;
  for (int i = 0; i < KR * 64; ++i) wmax = std::max(wmax, std::fabs((double)w_tap_cin_cout[i]) / 255.);
  int e = 0;
  if (wmax > 0.) { e = (int)std::floor(std::log2(16384. / wmax)); e = std::max(-24, std::min(30, e)); }
  const double scale = std::ldexp(1., e);
  std::vector<float> scaled((size_t)KR * 64);
  std::vector<float> bias(64);
  for (int n = 0; n < 64; ++n) {
    double sum = 0.;
    for (int k = 0; k < KR; ++k) {
      const double w = (double)w_tap_cin_cout[(size_t)k * 64 + n] / 255.;
      scaled[(size_t)k * 64 + n] = (float)(w * scale);            // |.| <= 16384: hi and lo = w - hi are both normal fp16 numbers
      sum += w;
    }
    bias[n] = (float)((double)bias64[n] + 0.5 * sum);
  }
  KG_TRY(tc_stem_pack(scaled.data(), K, out));
  float* db = nullptr;
  KG_CUDA_CHECK(cudaMalloc(&db, 64 * sizeof(float)));
  out->d_bias_u8 = std::shared_ptr<void>(db, [](void* q) { cudaFree(q); });
  KG_CUDA_CHECK(cudaMemcpy(db, bias.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
  out->inv_scale = (float)std::ldexp(1., -e);
  return KG_OK;
}

bool tc_stem_supported(int K, int stride) {
  if (!tc_available()) return false;
  const char* off = getenv("KG_TC_STEM");
  if (off && off[0] == '0') return false;
  return (K == 3 && stride == 1) || (K == 7 && stride == 2);
}

int tc_stem_launch(const float* x, const TcStemWeights* w, const float* bias, __half* out_hi, __half* out_lo, int N, int H, int W, int K,
                   int stride, cudaStream_t s) {
  KG_REQUIRE(x && w && w->d_img && w->K == K && bias && out_hi && out_lo, "tc_stem_launch: bad arguments");
  KG_REQUIRE(tc_stem_supported(K, stride), "tc_stem_launch: unsupported conv %dx%d / s%d", K, K, stride);
  static bool attr_set = false;
  if (!attr_set) {
    KG_CUDA_CHECK(cudaFuncSetAttribute(tc_stem_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem_smem_bytes<3>()));
    KG_CUDA_CHECK(cudaFuncSetAttribute(tc_stem_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem_smem_bytes<7>()));
    attr_set = true;
  }
  const int pad = K / 2;
  StemParams p{};
  p.x = x; p.w_img = reinterpret_cast<const __half*>(w->d_img.get()); p.bias = bias; p.out_hi = out_hi; p.out_lo = out_lo;
  p.N = N; p.H = H; p.W = W; p.Ho = (H + 2 * pad - K) / stride + 1; p.Wo = (W + 2 * pad - K) / stride + 1;
  p.total = (long long)N * p.Ho * p.Wo;
  p.tiles = (int)((p.total + 127) / 128);
  const int sms = tc_num_sms();
  if (K == 3) {
    const int grid = std::min(p.tiles, sms * 4);                  // 48 KiB smem, 64 TMEM columns: four CTAs per SM
    tc_stem_kernel<3, 1><<<grid, ST_THREADS, stem_smem_bytes<3>(), s>>>(p);
  } else {
    const int grid = std::min(p.tiles, sms);
    tc_stem_kernel<7, 2><<<grid, ST_THREADS, stem_smem_bytes<7>(), s>>>(p);
  }
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

int tc_stem_launch_u8(const uint8_t* img, const TcStemWeights* w, __half* out_hi, __half* out_lo, int N, int H, int W, int K, int stride,
                      cudaStream_t s) {
  KG_REQUIRE(img && w && w->d_img && w->d_bias_u8 && w->K == K && out_hi && out_lo, "tc_stem_launch_u8: bad arguments");
  KG_REQUIRE(tc_stem_supported(K, stride), "tc_stem_launch_u8: unsupported conv %dx%d / s%d", K, K, stride);
  static bool attr_set = false;
  if (!attr_set) {
    KG_CUDA_CHECK(cudaFuncSetAttribute(tc_stem_u8_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem_u8_smem_bytes<3>()));
    KG_CUDA_CHECK(cudaFuncSetAttribute(tc_stem_u8_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stem_u8_smem_bytes<7>()));
    attr_set = true;
  }
  const int pad = K / 2;
  StemU8Params p{};
  p.img = img; p.w_img = reinterpret_cast<const __half*>(w->d_img.get()); p.bias = reinterpret_cast<const float*>(w->d_bias_u8.get());
  p.out_hi = out_hi; p.out_lo = out_lo; p.inv_scale = w->inv_scale;
  p.N = N; p.H = H; p.W = W; p.Ho = (H + 2 * pad - K) / stride + 1; p.Wo = (W + 2 * pad - K) / stride + 1;
  p.total = (long long)N * p.Ho * p.Wo;
  p.tiles = (int)((p.total + 127) / 128);
  const int sms = tc_num_sms();
  if (K == 3) {
    tc_stem_u8_kernel<3, 1><<<std::min(p.tiles, sms * 6), ST_THREADS, stem_u8_smem_bytes<3>(), s>>>(p);     // 33 KiB smem: six CTAs per SM
  } else {
    tc_stem_u8_kernel<7, 2><<<std::min(p.tiles, sms * 2), ST_THREADS, stem_u8_smem_bytes<7>(), s>>>(p);     // 97 KiB smem: two CTAs per SM
  }
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

}  // namespace kg
