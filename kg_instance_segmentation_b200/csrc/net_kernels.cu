// CUDA-core kernels of the KGnet forward: generic (ragged, windowed, two-source) FFMA convolution, bilinear
// resize, 3x3/s2 max pool and layout import/export.  These cover the layers the tcgen05 implicit-GEMM kernel
// (tc_conv.cu) does not take: Cin=3 stems, stride-2 convs, and the variable-size per-box mask branch.
// They are also the on-device fp32 reference the tensor-core kernel is validated against.
#include "net.cuh"

namespace kg {

__device__ __forceinline__ float ld_split(const __half* hi, const __half* lo, long long i) {
  float v = __half2float(hi[i]);
  if (lo != nullptr) v += __half2float(lo[i]);
  return v;
}

__device__ __forceinline__ void st_split(__half* hi, __half* lo, long long i, float v) {
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  const __half h = __float2half_rn(v);
  if (hi != nullptr) hi[i] = h;
  if (lo != nullptr) lo[i] = __float2half_rn(v - __half2float(h));
}

// ------------------------------------------------------------------------------------------------
// Generic convolution (KGnet.py: every nn.Conv2d of the model).  Tile: 64 output pixels x 64 output channels
// per CTA, 4x4 per thread, K chunks of 16 input channels per filter tap.  fp32 accumulation in tap-major,
// channel-minor order.
constexpr int FT_P = 64, FT_C = 64, FT_K = 16;

template <bool X32>
__global__ void __launch_bounds__(256) conv_ffma_kernel(const ConvArgs a) {
  const ConvProb pb = a.probs[blockIdx.z];
  const int npix = pb.Hout * pb.Wout;
  const int p0 = blockIdx.x * FT_P;
  if (p0 >= npix) return;
  const int co0 = blockIdx.y * FT_C;
  __shared__ float sA[FT_K][FT_P + 4];
  __shared__ __align__(16) float sB[FT_K][FT_C];
  const int tid = threadIdx.x;
  const int lp = tid >> 2, lc = (tid & 3) * 4;
  const int p = p0 + lp;
  const bool pv = p < npix;
  const int oy = pv ? p / pb.Wout : 0, ox = pv ? p - oy * pb.Wout : 0;
  const int ty = tid >> 4, tx = tid & 15;
  const int Cin = a.C0 + a.C1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ntaps = a.R * a.S;
  for (int tap = 0; tap < ntaps; ++tap) {
    const int r = tap / a.S, s = tap - r * a.S;
    const int iy = oy * a.stride - a.pad + r, ix = ox * a.stride - a.pad + s;
    const bool inb = pv && iy >= 0 && iy < pb.Hin && ix >= 0 && ix < pb.Win;
    for (int c0 = 0; c0 < Cin; c0 += FT_K) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (inb) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + lc + j;
          if (c < Cin) {
            if (X32) {
              v[j] = __ldg(a.x32 + pb.in0_off + ((long long)c * pb.Hin + iy) * pb.Win + ix);
            } else if (c < a.C0) {
              v[j] = ld_split(a.in0_hi, a.in0_lo, pb.in0_off + (long long)iy * pb.in0_pitch + (long long)ix * a.in0_ps + c);
            } else {
              v[j] = ld_split(a.in1_hi, a.in1_lo, pb.in1_off + (long long)iy * pb.in1_pitch + (long long)ix * a.in1_ps + (c - a.C0));
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) sA[lc + j][lp] = v[j];
      {
        const int k = tid >> 4, n = (tid & 15) * 4;
        const int c = c0 + k;
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cin) {
          const float* wp = a.w + ((long long)tap * Cin + c) * a.Cout + co0 + n;
          if (co0 + n + 3 < a.Cout && (a.Cout & 3) == 0) {
            wv = __ldg(reinterpret_cast<const float4*>(wp));
          } else {
            if (co0 + n < a.Cout) wv.x = __ldg(wp);
            if (co0 + n + 1 < a.Cout) wv.y = __ldg(wp + 1);
            if (co0 + n + 2 < a.Cout) wv.z = __ldg(wp + 2);
            if (co0 + n + 3 < a.Cout) wv.w = __ldg(wp + 3);
          }
        }
        *reinterpret_cast<float4*>(&sB[k][n]) = wv;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < FT_K; ++k) {
        const float4 av = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&sB[k][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = p0 + ty * 4 + i;
    if (q >= npix) continue;
    const int qy = q / pb.Wout, qx = q - qy * pb.Wout;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co >= a.Cout) continue;
      float v = acc[i][j] + (a.bias != nullptr ? __ldg(a.bias + co) : 0.f);
      if (a.res_hi != nullptr)
        v += ld_split(a.res_hi, a.res_lo, pb.res_off + (long long)qy * pb.res_pitch + (long long)qx * a.res_ps + co);
      if (a.relu) v = fmaxf(v, 0.f);
      if (a.sigmoid) v = 1.f / (1.f + expf(-v));
      if (a.out_hi != nullptr || a.out_lo != nullptr)
        st_split(a.out_hi, a.out_lo, pb.out_off + (long long)qy * pb.out_pitch + (long long)qx * a.out_ps + co, v);
      if (a.out32 != nullptr) a.out32[pb.out32_off + ((long long)co * pb.Hout + qy) * pb.Wout + qx] = v;
    }
  }
}

int launch_conv_ffma(const ConvArgs& a, int nprob, int max_pix, cudaStream_t s) {
  if (nprob <= 0 || max_pix <= 0) return KG_OK;
  dim3 grid(ceil_div(max_pix, FT_P), ceil_div(a.Cout, FT_C), nprob);
  KG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv_ffma: grid too large (%u, %u)", grid.y, grid.z);
  if (a.x32 != nullptr) conv_ffma_kernel<true><<<grid, 256, 0, s>>>(a);
  else conv_ffma_kernel<false><<<grid, 256, 0, s>>>(a);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
// Stem convolutions on the raw image (KGnet.py:139-141 c0_conv.0: 3->64 3x3/s1, KGnet.py:131-133 conv1+bn1: 3->64
// 7x7/s2), fp32 NCHW in, split-fp16 NHWC out, ReLU fused.  Cin = 3 makes these HBM-bound on the output write, so one
// thread owns one output pixel and all 64 output channels (128 B contiguous per plane); weights are broadcast from smem.
// PX horizontally adjacent output pixels per thread: every weight float4 fetched from smem feeds 4 * PX FFMAs (the
// one-pixel version issued one LDS.128 per 4 FFMAs and was bound by that), and the PX pixels share their input columns.
template <int K, int STRIDE, int PX>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, __half* __restrict__ out_hi,
                                                        __half* __restrict__ out_lo, int N, int H, int W, int Ho, int Wo) {
  __shared__ __align__(16) float sw[K * K * 3 * 64];
  for (int e = threadIdx.x; e < K * K * 3 * 64; e += 128) sw[e] = w[e];
  __syncthreads();
  const int wq = (Wo + PX - 1) / PX;                    // pixel groups per output row
  const long long total = (long long)N * Ho * wq;
  const long long g = (long long)blockIdx.x * 128 + threadIdx.x;
  if (g >= total) return;
  const int ox0 = (int)(g % wq) * PX;
  const int oy = (int)((g / wq) % Ho);
  const int n = (int)(g / ((long long)wq * Ho));
  constexpr int PAD = K / 2;
  constexpr int NIN = K + STRIDE * (PX - 1);            // input columns the PX pixels touch in one filter row
  float acc[PX][64];
#pragma unroll
  for (int q = 0; q < PX; ++q)
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[q][j] = __ldg(bias + j);
  const float* xn = x + (long long)n * 3 * H * W;
  const int ix0 = ox0 * STRIDE - PAD;
  for (int r = 0; r < K; ++r) {
    const int iy = oy * STRIDE - PAD + r;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float in[NIN];
      const float* row = xn + ((long long)c * H + iy) * W;
#pragma unroll
      for (int i = 0; i < NIN; ++i) { const int ix = ix0 + i; in[i] = (ix >= 0 && ix < W) ? __ldg(row + ix) : 0.f; }
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const float4* wr = reinterpret_cast<const float4*>(&sw[((r * K + s) * 3 + c) * 64]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 w4 = wr[j];
#pragma unroll
          for (int q = 0; q < PX; ++q) {
            const float v = in[s + q * STRIDE];
            acc[q][4 * j] = fmaf(v, w4.x, acc[q][4 * j]); acc[q][4 * j + 1] = fmaf(v, w4.y, acc[q][4 * j + 1]);
            acc[q][4 * j + 2] = fmaf(v, w4.z, acc[q][4 * j + 2]); acc[q][4 * j + 3] = fmaf(v, w4.w, acc[q][4 * j + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PX; ++q) {
    if (ox0 + q >= Wo) break;
    const long long p = ((long long)n * Ho + oy) * Wo + ox0 + q;
    uint4* oh = reinterpret_cast<uint4*>(out_hi + p * 64);
    uint4* ol = reinterpret_cast<uint4*>(out_lo + p * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint4 h4, l4;
      __half2* hh = reinterpret_cast<__half2*>(&h4);
      __half2* ll = reinterpret_cast<__half2*>(&l4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = fminf(fmaxf(acc[q][k * 8 + 2 * e], 0.f), 65504.f), b = fminf(fmaxf(acc[q][k * 8 + 2 * e + 1], 0.f), 65504.f);
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        hh[e] = h;
        ll[e] = __floats2half2_rn(a - hf.x, b - hf.y);
      }
      oh[k] = h4; ol[k] = l4;
    }
  }
}

int launch_stem_conv(const float* x, const float* w, const float* bias, __half* out_hi, __half* out_lo, int N, int H, int W, int K,
                     int stride, cudaStream_t s) {
  const int pad = K / 2;
  const int Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1;
  constexpr int PX = 2;
  const long long total = (long long)N * Ho * ((Wo + PX - 1) / PX);
  const unsigned grid = (unsigned)((total + 127) / 128);
  if (K == 3 && stride == 1) stem_conv_kernel<3, 1, PX><<<grid, 128, 0, s>>>(x, w, bias, out_hi, out_lo, N, H, W, Ho, Wo);
  else if (K == 7 && stride == 2) stem_conv_kernel<7, 2, PX><<<grid, 128, 0, s>>>(x, w, bias, out_hi, out_lo, N, H, W, Ho, Wo);
  else { set_error("stem conv: unsupported k=%d stride=%d", K, stride); return KG_ERR_INVALID; }
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
// F.interpolate(mode='bilinear', align_corners=False) to an explicit size (KGnet.py:110,288-297), fp32 math.
__device__ __forceinline__ void ld8_split(const __half* hi, const __half* lo, long long i, float* v) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + i));
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(ah[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  if (lo != nullptr) {
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + i));
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(bh[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
  }
}

// hi = fp16(v) with saturation to +-65504 (one packed convert, F2FP.SATFINITE: replaces an explicit clamp), lo = fp16(v - hi)
__device__ __forceinline__ void split2_sat(float a, float b, __half2& hi, __half2& lo) {
  uint32_t h, l;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(b - hf.y), "f"(a - hf.x));
  hi = *reinterpret_cast<const __half2*>(&h);
  lo = *reinterpret_cast<const __half2*>(&l);
}

__global__ void __launch_bounds__(256) bilinear_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                                                       int in_ps, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                       int out_ps, int C, const ResizeProb* __restrict__ probs) {
  const ResizeProb pb = probs[blockIdx.z];
  const int cg = C >> 3;   // channel octets: one 16-byte load per plane per corner
  const int fr = pb.frame;                                   // 0 / 1: output rectangle grown by a frame of zeros
  const int Wf = pb.Wout + 2 * fr, Hf = pb.Hout + 2 * fr;
  const unsigned total = (unsigned)Hf * (unsigned)Wf * (unsigned)cg;   // < 2^31 (checked by the launcher): 32-bit index math
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int q = (int)(e / (unsigned)cg);
  const int c = (int)(e - (unsigned)q * (unsigned)cg) * 8;
  const int oy = q / Wf - fr, ox = q - (q / Wf) * Wf - fr;
  if (oy < 0 || ox < 0 || oy >= pb.Hout || ox >= pb.Wout) {   // frame pixel
    const long long o = pb.out_off + (long long)oy * pb.out_pitch + (long long)ox * out_ps + c;
    *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(0u, 0u, 0u, 0u);
    if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float rh = (float)pb.Hin / (float)pb.Hout, rw = (float)pb.Win / (float)pb.Wout;
  float sy = rh * ((float)oy + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
  float sx = rw * ((float)ox + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
  const int y0 = (int)sy, x0 = (int)sx;
  const int yp = y0 < pb.Hin - 1 ? 1 : 0, xp = x0 < pb.Win - 1 ? 1 : 0;
  const float ly1 = sy - (float)y0, ly0 = 1.f - ly1, lx1 = sx - (float)x0, lx0 = 1.f - lx1;
  const long long b00 = pb.in_off + (long long)y0 * pb.in_pitch + (long long)x0 * in_ps + c;
  const long long b01 = b00 + (long long)xp * in_ps, b10 = b00 + (long long)yp * pb.in_pitch, b11 = b10 + (long long)xp * in_ps;
  const long long o = pb.out_off + (long long)oy * pb.out_pitch + (long long)ox * out_ps + c;
  float v00[8], v01[8], v10[8], v11[8];
  ld8_split(in_hi, in_lo, b00, v00); ld8_split(in_hi, in_lo, b01, v01);
  ld8_split(in_hi, in_lo, b10, v10); ld8_split(in_hi, in_lo, b11, v11);
  uint4 h4, l4;
  __half2* hh = reinterpret_cast<__half2*>(&h4);
  __half2* ll = reinterpret_cast<__half2*>(&l4);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float r[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int k = 2 * j + t;
      r[t] = ly0 * (lx0 * v00[k] + lx1 * v01[k]) + ly1 * (lx0 * v10[k] + lx1 * v11[k]);
    }
    split2_sat(r[0], r[1], hh[j], ll[j]);
  }
  *reinterpret_cast<uint4*>(out_hi + o) = h4;
  if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + o) = l4;
}

// Row-tiled variant for the many small resize problems of forward_seg (one per box and level, KGnet.py:110): the generic kernel above
// gives every thread ONE output vector behind two dependent global round trips (problem record, then the four corners) in CTAs that
// live for a few microseconds -- measured at 1.5 TB/s.  Here a CTA owns two (framed) output rows of one problem and its threads loop
// over the row's vectors with the eight corner loads of both rows in flight; row coordinates are CTA-uniform.  Same arithmetic per
// output as bilinear_kernel.
constexpr int BL_ROWS = 2;
__global__ void __launch_bounds__(256) bilinear_rows_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                                                            int in_ps, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                                            int out_ps, int C, const ResizeProb* __restrict__ probs) {
  const ResizeProb pb = probs[blockIdx.z];
  const int fr = pb.frame;
  const int Wf = pb.Wout + 2 * fr, Hf = pb.Hout + 2 * fr;
  const int r0 = blockIdx.x * BL_ROWS;
  if (r0 >= Hf) return;
  const int cg = C >> 3;
  const int vecs = Wf * cg;
  const float rh = (float)pb.Hin / (float)pb.Hout, rw = (float)pb.Win / (float)pb.Wout;
  bool row_on[BL_ROWS], row_in[BL_ROWS];
  long long in_row[BL_ROWS], out_row[BL_ROWS];
  int ystep[BL_ROWS];
  float ly1[BL_ROWS];
#pragma unroll
  for (int r = 0; r < BL_ROWS; ++r) {
    const int oy = r0 + r - fr;
    row_on[r] = r0 + r < Hf;
    row_in[r] = row_on[r] && oy >= 0 && oy < pb.Hout;
    float sy = rh * ((float)oy + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
    const int y0 = row_in[r] ? (int)sy : 0;
    ystep[r] = y0 < pb.Hin - 1 ? pb.in_pitch : 0;
    ly1[r] = sy - (float)y0;
    in_row[r] = pb.in_off + (long long)y0 * pb.in_pitch;
    out_row[r] = pb.out_off + (long long)oy * pb.out_pitch;
  }
  for (int v = threadIdx.x; v < vecs; v += 256) {
    const int xq = v / cg;
    const int c = (v - xq * cg) * 8;
    const int ox = xq - fr;
    const bool x_in = ox >= 0 && ox < pb.Wout;
    float sx = rw * ((float)ox + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
    const int x0 = x_in ? (int)sx : 0;
    const int xstep = x0 < pb.Win - 1 ? in_ps : 0;
    const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
    const long long xoff = (long long)x0 * in_ps + c;
    float v00[BL_ROWS][8], v01[BL_ROWS][8], v10[BL_ROWS][8], v11[BL_ROWS][8];
#pragma unroll
    for (int r = 0; r < BL_ROWS; ++r)
      if (row_in[r] && x_in) {
        const long long b = in_row[r] + xoff;
        ld8_split(in_hi, in_lo, b, v00[r]); ld8_split(in_hi, in_lo, b + xstep, v01[r]);
        ld8_split(in_hi, in_lo, b + ystep[r], v10[r]); ld8_split(in_hi, in_lo, b + ystep[r] + xstep, v11[r]);
      }
#pragma unroll
    for (int r = 0; r < BL_ROWS; ++r) {
      if (!row_on[r]) continue;
      const long long o = out_row[r] + (long long)ox * out_ps + c;
      uint4 h4 = make_uint4(0u, 0u, 0u, 0u), l4 = make_uint4(0u, 0u, 0u, 0u);       // frame pixels stay zero
      if (row_in[r] && x_in) {
        __half2* hh = reinterpret_cast<__half2*>(&h4);
        __half2* ll = reinterpret_cast<__half2*>(&l4);
        const float ly0 = 1.f - ly1[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float q[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int k = 2 * j + t;
            q[t] = ly0 * (lx0 * v00[r][k] + lx1 * v01[r][k]) + ly1[r] * (lx0 * v10[r][k] + lx1 * v11[r][k]);
          }
          split2_sat(q[0], q[1], hh[j], ll[j]);
        }
      }
      *reinterpret_cast<uint4*>(out_hi + o) = h4;
      if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + o) = l4;
    }
  }
}

// Exact x2 upsampling of a dense NHWC tensor (the four F.interpolate calls of the decoder, KGnet.py:288-297).  One thread owns one
// INPUT pixel x 8 channels and produces its 2 x 2 output pixels from the 3 x 3 input neighbourhood: 9 loads per plane for 4 outputs
// where the generic kernel issues 16, and the weights are the constants 0.25 / 0.75 (align_corners=False at scale 2: source
// coordinate (o + 0.5) / 2 - 0.5, clamped at 0 like the generic kernel / ATen).  Same expression per output as bilinear_kernel.
__global__ void __launch_bounds__(256) bilinear2x_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                                                         __half* __restrict__ out_hi, __half* __restrict__ out_lo, int Hin, int Win, int C,
                                                         unsigned total) {
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int cg = C >> 3;
  const int c = (int)(e % (unsigned)cg) * 8;
  unsigned q = e / (unsigned)cg;
  const int j = (int)(q % (unsigned)Win); q /= (unsigned)Win;
  const int i = (int)(q % (unsigned)Hin);
  const int n = (int)(q / (unsigned)Hin);
  // neighbourhood rows / columns (i-1, i, i+1) clamped to the image: a clamped entry repeats the edge value, which is exactly what
  // the generic kernel reads there (index clamp at the far edge, source coordinate clamp at 0 with weights (1, 0) at the near edge)
  const int r[3] = {max(i - 1, 0), i, min(i + 1, Hin - 1)}, cc[3] = {max(j - 1, 0), j, min(j + 1, Win - 1)};
  const long long img = (long long)n * Hin * Win;
  float v[3][3][8];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) ld8_split(in_hi, in_lo, (img + (long long)r[a] * Win + cc[b]) * C + c, v[a][b]);
  // output parity 0 blends (i-1, i) with (0.25, 0.75) -- (1, 0) at i == 0 --, parity 1 blends (i, i+1) with (0.75, 0.25)
  const float wy[2][2] = {{i > 0 ? 0.25f : 1.f, i > 0 ? 0.75f : 0.f}, {0.75f, 0.25f}};
  const float wx[2][2] = {{j > 0 ? 0.25f : 1.f, j > 0 ? 0.75f : 0.f}, {0.75f, 0.25f}};
  const int Wout = 2 * Win;
  const long long oimg = (long long)n * (2 * Hin) * Wout;
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      uint4 h4, l4;
      __half2* hh = reinterpret_cast<__half2*>(&h4);
      __half2* ll = reinterpret_cast<__half2*>(&l4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float o[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int ch = 2 * k + t;
          o[t] = wy[py][0] * (wx[px][0] * v[py][px][ch] + wx[px][1] * v[py][px + 1][ch]) +
                 wy[py][1] * (wx[px][0] * v[py + 1][px][ch] + wx[px][1] * v[py + 1][px + 1][ch]);
        }
        split2_sat(o[0], o[1], hh[k], ll[k]);
      }
      const long long op = (oimg + (long long)(2 * i + py) * Wout + (2 * j + px)) * C + c;
      *reinterpret_cast<uint4*>(out_hi + op) = h4;
      if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + op) = l4;
    }
}

int launch_bilinear2x(const __half* in_hi, const __half* in_lo, __half* out_hi, __half* out_lo, int N, int Hin, int Win, int C, cudaStream_t s) {
  KG_REQUIRE((C & 7) == 0, "bilinear2x: channel count must be a multiple of 8 (C=%d)", C);
  const long long total = (long long)N * Hin * Win * (C >> 3);
  KG_REQUIRE(total < (1ll << 32), "bilinear2x: problem too large");
  bilinear2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in_hi, in_lo, out_hi, out_lo, Hin, Win, C, (unsigned)total);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

int launch_bilinear(const __half* in_hi, const __half* in_lo, int in_ps, __half* out_hi, __half* out_lo, int out_ps, int C,
                    const ResizeProb* probs, int nprob, int max_pix, cudaStream_t s) {
  if (nprob <= 0 || max_pix <= 0) return KG_OK;
  KG_REQUIRE((C & 7) == 0 && (in_ps & 7) == 0 && (out_ps & 7) == 0, "bilinear: channel counts must be multiples of 8 (C=%d)", C);
  KG_REQUIRE((long long)max_pix * (C >> 3) < (1ll << 31), "bilinear: problem too large (%d pixels x %d channels)", max_pix, C);
  // (max_pix must count the frame pixels of framed problems)
  dim3 grid((unsigned)(((long long)max_pix * (C >> 3) + 255) / 256), 1, nprob);
  bilinear_kernel<<<grid, 256, 0, s>>>(in_hi, in_lo, in_ps, out_hi, out_lo, out_ps, C, probs);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

int launch_bilinear_rows(const __half* in_hi, const __half* in_lo, int in_ps, __half* out_hi, __half* out_lo, int out_ps, int C,
                         const ResizeProb* probs, int nprob, int max_rows, cudaStream_t s) {
  if (nprob <= 0 || max_rows <= 0) return KG_OK;
  KG_REQUIRE((C & 7) == 0 && (in_ps & 7) == 0 && (out_ps & 7) == 0, "bilinear: channel counts must be multiples of 8 (C=%d)", C);
  KG_REQUIRE(nprob <= 65535, "bilinear: too many problems (%d)", nprob);
  dim3 grid((unsigned)ceil_div(max_rows, BL_ROWS), 1, (unsigned)nprob);
  bilinear_rows_kernel<<<grid, 256, 0, s>>>(in_hi, in_lo, in_ps, out_hi, out_lo, out_ps, C, probs);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
// Rectangle copies between NHWC tensors (forward_seg: the feature crops of every box and level into the atlases).  A rectangle is Hout
// rows of Wout * C contiguous halfs on both sides: one CTA copies 4 rows with 16-byte vectors.
constexpr int CR_ROWS = 4;
__global__ void __launch_bounds__(256) copy_rects_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int in_ps,
                                                         __half* __restrict__ out_hi, __half* __restrict__ out_lo, int out_ps, int C,
                                                         const ResizeProb* __restrict__ probs) {
  const ResizeProb pb = probs[blockIdx.z];
  const int r0 = blockIdx.x * CR_ROWS;
  if (r0 >= pb.Hout) return;
  const int cg = C >> 3;
  const int vecs = pb.Wout * cg;                              // 16-byte vectors per row and plane
  const int rows = min(CR_ROWS, pb.Hout - r0);
  for (int e = threadIdx.x; e < rows * vecs; e += 256) {
    const int r = e / vecs, v = e - r * vecs;
    const int x = v / cg, c = (v - x * cg) * 8;
    const long long i = pb.in_off + (long long)(r0 + r) * pb.in_pitch + (long long)x * in_ps + c;
    const long long o = pb.out_off + (long long)(r0 + r) * pb.out_pitch + (long long)x * out_ps + c;
    *reinterpret_cast<uint4*>(out_hi + o) = __ldg(reinterpret_cast<const uint4*>(in_hi + i));
    if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + o) = __ldg(reinterpret_cast<const uint4*>(in_lo + i));
  }
}

int launch_copy_rects(const __half* in_hi, const __half* in_lo, int in_ps, __half* out_hi, __half* out_lo, int out_ps, int C,
                      const ResizeProb* probs, int nprob, int max_h, cudaStream_t s) {
  if (nprob <= 0 || max_h <= 0) return KG_OK;
  KG_REQUIRE((C & 7) == 0 && (in_ps & 7) == 0 && (out_ps & 7) == 0, "copy_rects: channel counts must be multiples of 8 (C=%d)", C);
  KG_REQUIRE(nprob <= 65535, "copy_rects: too many rectangles (%d)", nprob);
  dim3 grid((unsigned)ceil_div(max_h, CR_ROWS), 1, (unsigned)nprob);
  copy_rects_kernel<<<grid, 256, 0, s>>>(in_hi, in_lo, in_ps, out_hi, out_lo, out_ps, C, probs);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
// validity mask of the forward_seg atlas: 1 inside every packed box rectangle
__global__ void __launch_bounds__(128) fill_rects_kernel(uint8_t* __restrict__ mask, const RectProb* __restrict__ rects) {
  const RectProb r = rects[blockIdx.x];
  for (int e = threadIdx.x; e < r.h * r.w; e += 128) {
    const int y = e / r.w, x = e - y * r.w;
    mask[r.off + (long long)y * r.pitch + x] = 1;
  }
}

int launch_fill_rects(uint8_t* mask, const RectProb* rects, int nrect, cudaStream_t s) {
  if (nrect <= 0) return KG_OK;
  fill_rects_kernel<<<nrect, 128, 0, s>>>(mask, rects);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
// nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (KGnet.py:134,282).
// One thread = one output pixel x 8 channels (16-byte loads / stores per plane; the scalar version moved 2 bytes per lane).
__global__ void __launch_bounds__(256) maxpool_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                                                      __half* __restrict__ out_hi, __half* __restrict__ out_lo, int N, int Hin,
                                                      int Win, int Hout, int Wout, int C) {
  const int cg = C >> 3;
  const long long total = (long long)N * Hout * Wout * cg;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int c = (int)(e % cg) * 8;
  long long q = e / cg;
  const int ox = (int)(q % Wout); q /= Wout;
  const int oy = (int)(q % Hout);
  const int n = (int)(q / Hout);
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = oy * 2 + dy;
    if (iy < 0 || iy >= Hin) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = ox * 2 + dx;
      if (ix < 0 || ix >= Win) continue;
      float v[8];
      ld8_split(in_hi, in_lo, (((long long)n * Hin + iy) * Win + ix) * C + c, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
    }
  }
  uint4 h4, l4;
  __half2* hh = reinterpret_cast<__half2*>(&h4);
  __half2* ll = reinterpret_cast<__half2*>(&l4);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = fminf(fmaxf(m[2 * j], -65504.f), 65504.f), b2 = fminf(fmaxf(m[2 * j + 1], -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b2);
    const float2 hf = __half22float2(h);
    hh[j] = h;
    ll[j] = __floats2half2_rn(a - hf.x, b2 - hf.y);
  }
  const long long o = (((long long)n * Hout + oy) * Wout + ox) * C + c;
  *reinterpret_cast<uint4*>(out_hi + o) = h4;
  if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + o) = l4;
}

int launch_maxpool3x3s2(const __half* in_hi, const __half* in_lo, __half* out_hi, __half* out_lo, int N, int Hin, int Win, int C,
                        cudaStream_t s) {
  KG_REQUIRE((C & 7) == 0, "maxpool: channel count must be a multiple of 8 (C=%d)", C);
  const int Hout = (Hin + 2 - 3) / 2 + 1, Wout = (Win + 2 - 3) / 2 + 1;
  const long long total = (long long)N * Hout * Wout * (C >> 3);
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in_hi, in_lo, out_hi, out_lo, N, Hin, Win, Hout, Wout, C);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
// layout conversion at the API edge: split-fp16 NHWC <-> fp32 NCHW, 32 pixels x 32 channels per CTA via smem.
__global__ void __launch_bounds__(256) export_nchw_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo,
                                                          float* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (p < HW && c < C) ? ld_split(in_hi, in_lo, ((long long)n * HW + p) * C + c) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    if (p < HW && c < C) out[((long long)n * C + c) * HW + p] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256) import_nchw_kernel(const float* __restrict__ in, __half* __restrict__ out_hi,
                                                          __half* __restrict__ out_lo, int HW, int C) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (p < HW && c < C) ? in[((long long)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) st_split(out_hi, out_lo, ((long long)n * HW + p) * C + c, tile[tx][i]);
  }
}

int launch_export_nchw(const __half* in_hi, const __half* in_lo, float* out, int N, int HW, int C, cudaStream_t s) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N);
  export_nchw_kernel<<<grid, 256, 0, s>>>(in_hi, in_lo, out, HW, C);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

int launch_import_nchw(const float* in, __half* out_hi, __half* out_lo, int N, int HW, int C, cudaStream_t s) {
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N);
  import_nchw_kernel<<<grid, 256, 0, s>>>(in, out_hi, out_lo, HW, C);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

}  // namespace kg
