// Device-side decode path of KGnet (sm_100a):
//   Hough vote -> Gaussian blur -> peaks -> conf-sorted greedy keypoint-graph grouping -> refine ->
//   boxes -> NMS, batched over N images x up to 4 scales.
// Replaces postprocessing.py:8-261 and nms.py:4-53 of the reference (file:line cited per kernel).
//
// Arithmetic contract: everything after the vote accumulation follows the reference's fp64 operation
// order exactly (this file is compiled with -fmad=false; sqrt/div are IEEE).  The vote accumulation
// itself is done in 2^-44 fixed point with integer atomics: integer addition is associative, so the
// result is run-to-run deterministic and independent of the scatter order, at the price of an
// absolute error <= 2^-45 per vote against the reference's sequential fp64 sum (coo_matrix.todense()).
#include "decode.cuh"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cstdlib>

namespace kg {

// scipy.ndimage._filters._gaussian_kernel1d(sigma=2, order=0, radius=8) (postprocessing.py:144);
// symmetric, taps 0..8 (tap 8 = centre).  tests/test_cabi_cpu.py (test_gaussian_taps_and_constants_match_scipy) checks these against SciPy.
__constant__ double c_gauss[9] = {0x1.18aad19e4159bp-14, 0x1.c98b8c5d0dda5p-12, 0x1.227362b5fc92dp-9,
                                  0x1.1f30504e20207p-7,  0x1.ba4d4125ffd2ap-6,  0x1.0941b71ceef37p-4,
                                  0x1.ef9093fc46e5ap-4,  0x1.68856f9ab1982p-3,  0x1.98862a07ae7b4p-3};
// np.pi * KP_RADIUS**2 (postprocessing.py:51)
constexpr double KG_PI_R2 = 0x1.3a28c59d5433bp+6;
constexpr double KG_FIX = 0x1p44;
constexpr double KG_UNFIX = 0x1p-44;
constexpr int GAUSS_R = 8;

// directed-edge index m of (seed s -> target t) in EDGES + reversed(EDGES) (postprocessing.py:89,108)
__constant__ int c_mid_index[5][5] = {{-1, 0, 1, 2, 3}, {10, -1, 4, 5, 6}, {11, 14, -1, 7, 8},
                                      {12, 15, 17, -1, 9}, {13, 16, 18, 19, -1}};

// ------------------------------------------------------------------------------------------------
// K1: Hough vote.  compute_heatmaps + accumulate_votes (postprocessing.py:16-53).
// One CTA per 64 x 32 tile of source pixels of one (image, keypoint channel) plane.  Votes land within a few pixels of
// their source (short offsets are trained inside the radius-5 discs), so the four bilinear splats of every pixel are
// accumulated in SHARED memory over an (64+16) x (32+16) window around the tile, and the window is flushed with one
// coalesced global reduction per non-zero cell (~1 per pixel instead of 4 scattered ones: the first version of this kernel
// ran at the L2 atomic rate).  A vote outside the window goes to global memory directly.
// The 64-bit fixed-point cell is kept as two 32-bit words: sm_100 serialises 64-bit shared atomics lane by lane (measured:
// 0.85 shared wavefronts per lane-atomic), 32-bit ones run a conflict-free warp per wavefront.  The low word is added with
// the returning form, an unsigned wrap is the carry into the high word -- exact for any number of votes per cell.
// Integer accumulation is associative: the result does not depend on any of this.
constexpr int VT_W = 64, VT_H = 64, VT_HALO = 5;
constexpr int VR_W = VT_W + 2 * VT_HALO, VR_H = VT_H + 2 * VT_HALO;     // 74 x 74 cells, 2 x 4 B each: 43 KiB (five CTAs per SM)

// one splat of the generic path: image-bounds and window tests per cell
__device__ __forceinline__ void vote_cell(long long q, int gy, int gx, int H, int W, int wx0, int wy0, unsigned* __restrict__ s_lo,
                                          unsigned* __restrict__ s_hi, unsigned long long* __restrict__ plane) {
  if (q == 0 || (unsigned)gy >= (unsigned)H || (unsigned)gx >= (unsigned)W) return;      // good_inds (:34-35)
  const int ly = gy - wy0, lx = gx - wx0;
  if ((unsigned)ly < (unsigned)VR_H && (unsigned)lx < (unsigned)VR_W) {
    const int c = ly * VR_W + lx;
    const unsigned lo = (unsigned)q;
    unsigned hi = (unsigned)((unsigned long long)q >> 32);
    const unsigned old = atomicAdd(s_lo + c, lo);
    hi += (old + lo < old) ? 1u : 0u;
    if (hi != 0u) atomicAdd(s_hi + c, hi);
  } else {
    atomicAdd(plane + (size_t)gy * W + gx, (unsigned long long)q);
  }
}

// rare path (image border, or a vote that leaves the tile's window): kept out of line so that the hot loop stays small
__device__ __noinline__ void vote_slow(long long q0, long long q1, long long q2, long long q3, int iy, int ix, int H, int W, int wx0, int wy0,
                                       unsigned* __restrict__ s_lo, unsigned* __restrict__ s_hi, unsigned long long* __restrict__ plane) {
  vote_cell(q0, iy, ix, H, W, wx0, wy0, s_lo, s_hi, plane);          // tl
  vote_cell(q1, iy, ix + 1, H, W, wx0, wy0, s_lo, s_hi, plane);      // tr
  vote_cell(q2, iy + 1, ix, H, W, wx0, wy0, s_lo, s_hi, plane);      // bl
  vote_cell(q3, iy + 1, ix + 1, H, W, wx0, wy0, s_lo, s_hi, plane);  // br
}

// splat into window cell c (no tests): low word with the returning add, carry + high word only when non-zero
__device__ __forceinline__ void vote_fast(long long q, int c, unsigned* __restrict__ s_lo, unsigned* __restrict__ s_hi) {
  const unsigned lo = (unsigned)q;
  unsigned hi = (unsigned)((unsigned long long)q >> 32);
  const unsigned old = atomicAdd(s_lo + c, lo);
  hi += (old + lo < old) ? 1u : 0u;
  if (hi != 0u) atomicAdd(s_hi + c, hi);
}

__device__ __forceinline__ void vote_one(float kpv, float sx, float sy, double xd, double yd, int H, int W, int wx0, int wy0,
                                         unsigned* __restrict__ s_lo, unsigned* __restrict__ s_hi,
                                         unsigned long long* __restrict__ plane) {
  const double xs = xd + (double)sx;              // int64 + f32 -> f64 (:49)
  const double ys = yd + (double)sy;
  const double fx = floor(xs), fy = floor(ys);
  // a vote can only land in the image if floor is in [-1, W): everything else (incl. NaN / inf offsets) is dropped like the
  // reference's good_inds mask does
  if (!(fx >= -1. && fx < (double)W && fy >= -1. && fy < (double)H)) return;
  const double dx = xs - fx, dy = ys - fy;
  const double omdx = 1. - dx, omdy = 1. - dy;
  // 2^44 folded into p first: a power-of-two scale commutes with every rounding below
  const double p44 = (double)kpv * KG_FIX;
  const double a = p44 * omdx, b = p44 * dx, c = p44 * dy;     // (p*(1-dx)), (p*dx), (p*dy): the reference's left-to-right products (:27-30)
  const int ix = (int)fx, iy = (int)fy;
  const long long q0 = __double2ll_rn(a * omdy), q1 = __double2ll_rn(b * omdy), q2 = __double2ll_rn(c * omdx), q3 = __double2ll_rn(c * dx);
  // ceil = floor + 1 whenever the fractional part is non-zero; when it is zero the splat's weight (dx or dy) is zero and
  // the vote adds nothing, so floor + 1 can be used unconditionally
  const int lx = ix - wx0, ly = iy - wy0;
  if (ix >= 0 && ix + 1 < W && iy >= 0 && iy + 1 < H && (unsigned)lx < (unsigned)(VR_W - 1) && (unsigned)ly < (unsigned)(VR_H - 1)) {
    const int cell = ly * VR_W + lx;                  // all four cells inside the image and inside the window
    vote_fast(q0, cell, s_lo, s_hi);
    vote_fast(q1, cell + 1, s_lo, s_hi);
    vote_fast(q2, cell + VR_W, s_lo, s_hi);
    vote_fast(q3, cell + VR_W + 1, s_lo, s_hi);
  } else {
    vote_slow(q0, q1, q2, q3, iy, ix, H, W, wx0, wy0, s_lo, s_hi, plane);
  }
}

__global__ void __launch_bounds__(256, 4) vote_kernel(const float* __restrict__ kp, const float* __restrict__ sh,
                                                   unsigned long long* __restrict__ acc, int H, int W) {
  __shared__ unsigned s_lo[VR_H * VR_W], s_hi[VR_H * VR_W];
  const int plane_id = blockIdx.z;                      // n * 5 + i
  const int n = plane_id / 5, i = plane_id - n * 5;
  const size_t hw = (size_t)H * W;
  const int x0 = blockIdx.x * VT_W, y0 = blockIdx.y * VT_H;
  const int wx0 = x0 - VT_HALO, wy0 = y0 - VT_HALO;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < VR_H * VR_W; e += 256) { s_lo[e] = 0u; s_hi[e] = 0u; }
  __syncthreads();
  const float* kpp = kp + ((size_t)n * 5 + i) * hw;
  const float* sxp = sh + ((size_t)n * 10 + 2 * i) * hw;
  const float* syp = sxp + hw;
  unsigned long long* plane = acc + (size_t)plane_id * hw;
  // lane <-> consecutive x: the splats of one instruction go to consecutive cells (distinct banks); warp w owns 8 rows.
  // All 24 loads of a half-row group are issued before the first vote (the kernel is otherwise bound by their latency).
  constexpr int RPW = VT_H / 8;
  const int yb = y0 + warp * RPW;
  const int rows = min(RPW, H - yb);                      // uniform per warp
#pragma unroll
  for (int half = 0; half < VT_W / 32; ++half) {
    const int x = x0 + half * 32 + lane;
    if (x >= W || rows <= 0) continue;
    const double xd = (double)x;
    const int o = yb * W + x;                               // 32-bit indexing: a plane has < 2^31 cells (checked by the launcher)
    const float* kq = kpp + o; const float* xq = sxp + o; const float* yq = syp + o;
    float kv[RPW], ax[RPW], ay[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int ro = r < rows ? r * W : 0;
      kv[r] = __ldg(kq + ro); ax[r] = __ldg(xq + ro); ay[r] = __ldg(yq + ro);
    }
    double yd = (double)yb;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (r < rows) vote_one(kv[r], ax[r], ay[r], xd, yd, H, W, wx0, wy0, s_lo, s_hi, plane);
      yd += 1.0;
    }
  }
  __syncthreads();
  // flush: one coalesced global reduction per non-zero cell (cells were only ever hit by in-image votes)
  for (int ly = warp; ly < VR_H; ly += 8) {
    unsigned long long* row = plane + (size_t)(wy0 + ly) * W + wx0;
#pragma unroll
    for (int lx = lane; lx < VR_W; lx += 32) {
      const unsigned lo = s_lo[ly * VR_W + lx], hi = s_hi[ly * VR_W + lx];
      if ((lo | hi) != 0u) atomicAdd(row + lx, ((unsigned long long)hi << 32) + (unsigned long long)lo);
    }
  }
}

__device__ __forceinline__ int reflect_index(int i, int n) {
  // scipy 'reflect' (d c b a | a b c d | d c b a), valid for any offset
  const int period = 2 * n;
  int m = i % period;
  if (m < 0) m += period;
  return m >= n ? period - 1 - m : m;
}

// ------------------------------------------------------------------------------------------------
// K2: heat = acc / (pi r^2); gaussian_filter(sigma=2) (postprocessing.py:143-144, scipy correlate1d symmetric branch:
// axis 0 then axis 1, paired taps, no FMA); get_keypoints (:56-64): cross-footprint local maximum and conf > peak_thresh.
//
// Streaming separable filter.  A CTA owns a strip of CW output columns and a segment of rows of one (image, channel)
// plane and walks DOWN the rows:
//   * vertical pass in registers: thread c owns column x0 - 9 + c and keeps the 17 input rows around the current row in a
//     register window, so every accumulator cell is read from global memory ONCE per segment (coalesced 8-byte loads,
//     three rows in flight) and never staged in shared memory.  The window is a 20-slot circular buffer whose rotation is
//     unrolled at compile time (slots are named registers, no moves);
//   * horizontal pass through one shared row: the vertical results of the row are written to smem, and the first
//     ceil((CW+2)/4) threads each produce 4 adjacent outputs from 20 values (10 x LDS.128 per 4 outputs instead of 17 loads
//     per output);
//   * peak test of the PREVIOUS row from registers (up / centre / down) and two shared loads (left / right neighbours).
// (The first version staged a 50 x 50 halo tile per 32 x 32 outputs and re-read every value 17 times per axis: ~350
// instructions per output, 6 % of HBM peak.)
//
// acc -> heat is an fp64 division by the constant pi*25.  Correctly rounded constant division in three operations
// (Markstein): q = x * y, r = fma(-q, d, x), q' = fma(r, y, q) with y = RN(1/d); the 2^-44 fixed-point scale is a power
// of two and folds into the constants exactly.  Exhaustively equal to x / d on 3e5 random accumulators (host check with
// exact rationals) and bit-identical peaks on every fixture.
constexpr double KG_RCP_PI_R2 = 1.0 / KG_PI_R2;                 // RN(1/d), evaluated by the host compiler (IEEE)
constexpr double KG_Y44 = KG_RCP_PI_R2 * KG_UNFIX;              // y * 2^-44 (exact scaling)
constexpr double KG_D44 = KG_PI_R2 * KG_FIX;                    // d * 2^44  (exact scaling)
constexpr int BW_SLOTS = 20;                                    // 17-row window + 3 rows in flight
constexpr int BP_HALO = GAUSS_R + 1;                            // 9: blur radius + 1 column / row for the peak test

__device__ __forceinline__ double acc_to_heat(double raw_bits) {
  const double Q = (double)__double_as_longlong(raw_bits);      // exact below 2^53
  const double q0 = Q * KG_Y44;
  const double r = fma(-q0, KG_D44, Q);                          // (x - q0 * d) * 2^44, exact
  return fma(r, KG_Y44, q0);
}

struct BlurParams {
  const unsigned long long* acc;
  int H, W, CW, RH, n_strips;
  double peak_thresh;
  int list_index_base, n_scales, max_peaks;
  double* peak_conf; int* peak_key; int* peak_count;
  double* out_vote; double* out_heat;
  int* status;
  int emit_peaks;                 // 0: only export the fp64 maps (out_vote / out_heat); the peak list comes from the prefilter path
  unsigned long long* cand; int* cand_count; int cand_cap; int scale;    // prefilter path: candidate list (all scales share it)
};

// vertical pass of one row for window phase PH: slot (PH + i) % 20 holds input row y - 8 + i.  Converts the row that
// enters the window, issues the load of row y + 11 into the slot that just left it, returns the 17-tap paired sum.
template <int PH>
__device__ __forceinline__ double blur_vstep(double (&win)[BW_SLOTS], const unsigned long long* __restrict__ col, int y, int H, int W,
                                             double& centre) {
  constexpr int s_new = (PH + 16) % BW_SLOTS, s_load = (PH + 19) % BW_SLOTS;
  win[s_new] = acc_to_heat(win[s_new]);
  win[s_load] = __longlong_as_double((long long)__ldg(col + (size_t)reflect_index(y + GAUSS_R + 3, H) * W));
  centre = win[(PH + 8) % BW_SLOTS];
  double t = centre * c_gauss[GAUSS_R];
#pragma unroll
  for (int j = 0; j < GAUSS_R; ++j) t += (win[(PH + j) % BW_SLOTS] + win[(PH + 16 - j) % BW_SLOTS]) * c_gauss[j];
  return t;
}

template <int NT>
__global__ void __launch_bounds__(NT, NT > 160 ? 2 : 3) blur_peak_kernel(const BlurParams p) {
  __shared__ __align__(16) double s_tmp[2][NT + 8];
  __shared__ __align__(16) double s_h[2][NT + 8];               // blurred column b of a row at index b + 2 (b = -1 .. CW + 4)
  const int plane_id = blockIdx.z;                               // n * 5 + i
  const int n = plane_id / 5, ch = plane_id - n * 5;
  const int H = p.H, W = p.W, CW = p.CW;
  const size_t hw = (size_t)H * W;
  const unsigned long long* plane = p.acc + (size_t)plane_id * hw;
  const int x0 = blockIdx.x * CW;
  const int ybeg = blockIdx.y * p.RH, yend = min(H, ybeg + p.RH);   // output rows [ybeg, yend)
  const int tid = threadIdx.x;
  // vertical role: tmp column c = tid <-> image column x0 - 9 + c (reflected for the loads)
  const int gxc = x0 - BP_HALO + tid;
  const unsigned long long* col = plane + reflect_index(gxc, W);
  const bool vote_out = p.out_vote != nullptr && gxc >= x0 && gxc < min(W, x0 + CW);
  // horizontal role: blurred columns b = 4 tid .. 4 tid + 3 <-> image columns x0 - 1 + b
  const int nq = (CW + 2 + 3) >> 2;
  const bool hrole = tid < nq;
  const int b0 = 4 * tid;
  for (int e = tid; e < 2 * (NT + 8); e += NT) { (&s_tmp[0][0])[e] = 0.; (&s_h[0][0])[e] = 0.; }

  const int yfirst = ybeg - 1, ylast = yend;                      // blurred rows ybeg - 1 .. yend feed the peak test of ybeg .. yend - 1
  double win[BW_SLOTS];
  // phase 0 <-> row yfirst: slots 0..15 = rows yfirst - 8 .. yfirst + 7 (converted), 16..18 = rows yfirst + 8 .. + 10 (raw, in flight)
#pragma unroll
  for (int k = 0; k < 19; ++k) {
    const double raw = __longlong_as_double((long long)__ldg(col + (size_t)reflect_index(yfirst - GAUSS_R + k, H) * W));
    win[k] = k < 16 ? acc_to_heat(raw) : raw;
  }
  win[19] = 0.;
  double hu[4] = {0., 0., 0., 0.}, hc[4] = {0., 0., 0., 0.};      // blurred rows y - 2 and y - 1 of this thread's 4 columns
  __syncthreads();

  int phase = 0;
  for (int y = yfirst; y <= ylast; ++y) {
    double t, centre;
    switch (phase) {
#define KG_VSTEP(P) case P: t = blur_vstep<P>(win, col, y, H, W, centre); break;
      KG_VSTEP(0) KG_VSTEP(1) KG_VSTEP(2) KG_VSTEP(3) KG_VSTEP(4) KG_VSTEP(5) KG_VSTEP(6) KG_VSTEP(7) KG_VSTEP(8) KG_VSTEP(9)
      KG_VSTEP(10) KG_VSTEP(11) KG_VSTEP(12) KG_VSTEP(13) KG_VSTEP(14) KG_VSTEP(15) KG_VSTEP(16) KG_VSTEP(17) KG_VSTEP(18)
      default: t = blur_vstep<19>(win, col, y, H, W, centre); break;
#undef KG_VSTEP
    }
    phase = phase == BW_SLOTS - 1 ? 0 : phase + 1;
    const int par = y & 1;
    s_tmp[par][tid] = t;
    if (vote_out && y >= ybeg && y < yend) p.out_vote[(size_t)plane_id * hw + (size_t)y * W + gxc] = centre;
    __syncthreads();
    if (!hrole) continue;
    // ---- horizontal pass: 4 adjacent outputs from 20 values of the shared row ----
    double v[20];
    {
      const double2* src = reinterpret_cast<const double2*>(&s_tmp[par][b0]);
#pragma unroll
      for (int k = 0; k < 10; ++k) { const double2 d = src[k]; v[2 * k] = d.x; v[2 * k + 1] = d.y; }
    }
    double hd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double a = v[k + GAUSS_R] * c_gauss[GAUSS_R];
#pragma unroll
      for (int j = 0; j < GAUSS_R; ++j) a += (v[k + j] + v[k + 16 - j]) * c_gauss[j];
      hd[k] = a;
    }
    {
      double2* dst = reinterpret_cast<double2*>(&s_h[par][b0 + 2]);      // index b + 2: the pairs stay 16-byte aligned
      dst[0] = make_double2(hd[0], hd[1]); dst[1] = make_double2(hd[2], hd[3]);
    }
    // ---- peak test of row yr = y - 1 (centre hc, up hu, down hd; left / right from the previous row in smem) ----
    const int yr = y - 1;
    if (yr >= ybeg && yr < yend) {
      const double* prev = s_h[par ^ 1];
      const double left_edge = prev[b0 + 1], right_edge = prev[b0 + 6];    // columns b0 - 1 and b0 + 4 of row yr
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = b0 + k, gx = x0 - 1 + b;
        if (b < 1 || b > CW || gx >= W) continue;
        const double h = hc[k];
        if (p.out_heat != nullptr) p.out_heat[(size_t)plane_id * hw + (size_t)yr * W + gx] = h;
        double m = h;
        if (yr > 0) m = fmax(m, hu[k]);
        if (yr < H - 1) m = fmax(m, hd[k]);
        if (gx > 0) m = fmax(m, k == 0 ? left_edge : hc[k - 1]);
        if (gx < W - 1) m = fmax(m, k == 3 ? right_edge : hc[k + 1]);
        if (p.emit_peaks && m == h && h > p.peak_thresh) {
          const int list = n * p.n_scales + p.list_index_base;
          const int slot = atomicAdd(p.peak_count + list, 1);
          if (slot < p.max_peaks) {
            p.peak_conf[(size_t)list * p.max_peaks + slot] = h;
            p.peak_key[(size_t)list * p.max_peaks + slot] = ch * (int)hw + yr * W + gx;
          } else {
            atomicOr(p.status, 1);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { hu[k] = hc[k]; hc[k] = hd[k]; }
  }
}

// ------------------------------------------------------------------------------------------------
// K2a / K2b: the production path of blur + peaks.  The fp64 filter above is exact but occupies the fp64 pipe for ~60
// operations per pixel; only ~0.1 % of the pixels are peaks.  So:
//   K2a  blur32_candidates_kernel: the same separable filter in fp32 (FMA allowed) and a CONSERVATIVE peak test.  All terms are
//        non-negative, so the fp32 result has a relative error E <= 25 * 2^-24 = 1.5e-6 (one rounding per operation on the path
//        accumulator -> blurred value, including the rounded constants).  A true peak (h >= every neighbour, h > T) therefore
//        satisfies h32 * F >= n32 and h32 * F >= T with F = 1 + 6e-6 > (1 + E) / (1 - E): the candidate set is a superset
//        of the peak set.  Warp-private strips: lane l owns columns 4l .. 4l+3 of a 128-column strip and their 17-row
//        register windows; the horizontal pass takes the neighbours' vertical sums by warp shuffle.  No shared memory, no
//        barrier -- every warp streams down its rows independently.
//   K2b  exact_peaks_kernel: one warp per candidate recomputes the blurred value at the candidate and its four neighbours
//        in fp64 in the reference's exact operation order (same arithmetic as blur_peak_kernel) and applies the exact test.
// The peak list is bit-identical to the all-fp64 kernel's (tests/test_decode_gpu.py runs both).
constexpr float KG_CAND_F = 1.000006f;
constexpr int B32_CW = 110;                 // output columns per warp strip: 128 - 2 * 9

__constant__ float c_gauss32[9] = {(float)0x1.18aad19e4159bp-14, (float)0x1.c98b8c5d0dda5p-12, (float)0x1.227362b5fc92dp-9,
                                   (float)0x1.1f30504e20207p-7,  (float)0x1.ba4d4125ffd2ap-6,  (float)0x1.0941b71ceef37p-4,
                                   (float)0x1.ef9093fc46e5ap-4,  (float)0x1.68856f9ab1982p-3,  (float)0x1.98862a07ae7b4p-3};

struct Blur32Row { float t[4]; };

// one row of the vertical pass for window phase PH (18-slot circular window: slot (PH + i) % 18 = input row y - 8 + i);
// the raw accumulator row y + 8 was loaded one iteration earlier (a row iteration of a warp takes longer than the load latency)
template <int PH>
__device__ __forceinline__ Blur32Row blur32_vstep(float (&win)[4][18], unsigned long long (&raw)[4],
                                                  const unsigned long long* __restrict__ next_row, const int (&off)[4], float cscale) {
  constexpr int s_new = (PH + 16) % 18;
  Blur32Row r;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    win[k][s_new] = (float)(long long)raw[k] * cscale;
    raw[k] = __ldg(next_row + off[k]);
    float t = win[k][(PH + 8) % 18] * c_gauss32[GAUSS_R];
#pragma unroll
    for (int j = 0; j < GAUSS_R; ++j) t = fmaf(win[k][(PH + j) % 18] + win[k][(PH + 16 - j) % 18], c_gauss32[j], t);
    r.t[k] = t;
  }
  return r;
}

// scipy 'reflect' for a row / column index that is at most n outside [0, n): one reflection, no division
__device__ __forceinline__ int reflect_near(int i, int n) {
  i = i < 0 ? -1 - i : i;
  return i >= n ? 2 * n - 1 - i : i;
}

__global__ void __launch_bounds__(128, 4) blur32_candidates_kernel(const BlurParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int plane_id = blockIdx.y;
  const int H = p.H, W = p.W;
  const size_t hw = (size_t)H * W;
  const unsigned long long* plane = p.acc + (size_t)plane_id * hw;
  const int item = blockIdx.x * 4 + warp;                      // (row segment, column strip) of this warp
  const int seg = item / p.n_strips, strip = item - seg * p.n_strips;
  const int ybeg = seg * p.RH;
  if (ybeg >= H) return;
  const int yend = min(H, ybeg + p.RH);
  const int x0 = strip * B32_CW;
  const bool small = H < 2 * BP_HALO + 4;                       // tiny maps: an index can be reflected more than once
  // this lane's columns: strip column c = 4 lane + k <-> image column x0 - 9 + c, reflected at the image border (so the four
  // cells are consecutive only in the interior: per-column offsets from the first one)
  const int g0 = reflect_index(x0 - BP_HALO + 4 * lane, W);
  int off[4];
  bool col_in[4], col_out[4];                                  // inside the image / a column this strip reports candidates for
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = 4 * lane + k, gx = x0 - BP_HALO + c;
    off[k] = reflect_index(gx, W) - g0;
    col_in[k] = gx >= 0 && gx < W;
    col_out[k] = c >= BP_HALO && c < BP_HALO + B32_CW && gx < W;
  }
  const unsigned long long* col0 = plane + g0;
  const float cscale = (float)(KG_UNFIX / KG_PI_R2);
  float win[4][18];
  unsigned long long raw[4];
  auto row_ptr = [&](int gy) { return col0 + (size_t)(small ? reflect_index(gy, H) : reflect_near(gy, H)) * W; };
  auto load_row = [&](int gy, unsigned long long (&dst)[4]) {
    const unsigned long long* rowp = row_ptr(gy);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = __ldg(rowp + off[k]);
  };
  const int yfirst = ybeg - 1, ylast = yend;
  {
    unsigned long long tmp[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      load_row(yfirst - GAUSS_R + i, tmp);
#pragma unroll
      for (int k = 0; k < 4; ++k) win[k][i] = (float)(long long)tmp[k] * cscale;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { win[k][16] = 0.f; win[k][17] = 0.f; }
    load_row(yfirst + 8, raw);
  }
  // heat >= 0 everywhere, so a neighbour outside the image is represented by 0: it can never beat the centre
  float hu[4] = {0.f, 0.f, 0.f, 0.f}, hc[4] = {0.f, 0.f, 0.f, 0.f};
  const float thr = (float)p.peak_thresh;
  int phase = 0;
  for (int y = yfirst; y <= ylast; ++y) {
    const unsigned long long* next_row = row_ptr(y + GAUSS_R + 1);
    Blur32Row r;
    switch (phase) {
#define KG_V32(P) case P: r = blur32_vstep<P>(win, raw, next_row, off, cscale); break;
      KG_V32(0) KG_V32(1) KG_V32(2) KG_V32(3) KG_V32(4) KG_V32(5) KG_V32(6) KG_V32(7) KG_V32(8)
      KG_V32(9) KG_V32(10) KG_V32(11) KG_V32(12) KG_V32(13) KG_V32(14) KG_V32(15) KG_V32(16)
      default: r = blur32_vstep<17>(win, raw, next_row, off, cscale); break;
#undef KG_V32
    }
    phase = phase == 17 ? 0 : phase + 1;
    // ---- horizontal pass: this lane's 4 outputs need the vertical sums of lanes l-2 .. l+2 ----
    float v[20];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = __shfl_sync(0xffffffffu, r.t[k], (lane + 30) & 31);
      v[4 + k] = __shfl_sync(0xffffffffu, r.t[k], (lane + 31) & 31);
      v[8 + k] = r.t[k];
      v[12 + k] = __shfl_sync(0xffffffffu, r.t[k], (lane + 1) & 31);
      v[16 + k] = __shfl_sync(0xffffffffu, r.t[k], (lane + 2) & 31);
    }
    const bool row_in = y >= 0 && y < H;
    float hd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = v[k + GAUSS_R] * c_gauss32[GAUSS_R];
#pragma unroll
      for (int j = 0; j < GAUSS_R; ++j) a = fmaf(v[k + j] + v[k + 16 - j], c_gauss32[j], a);
      hd[k] = (row_in && col_in[k]) ? a : 0.f;
    }
    // ---- conservative peak test of row yr = y - 1 ----
    const int yr = y - 1;
    const float left_edge = __shfl_sync(0xffffffffu, hc[3], (lane + 31) & 31);
    const float right_edge = __shfl_sync(0xffffffffu, hc[0], (lane + 1) & 31);
    if (yr >= ybeg && yr < yend) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float hF = hc[k] * KG_CAND_F;
        const float m = fmaxf(fmaxf(hu[k], hd[k]), fmaxf(k == 0 ? left_edge : hc[k - 1], k == 3 ? right_edge : hc[k + 1]));
        if (col_out[k] && hF >= m && hF >= thr) {
          const int slot = atomicAdd(p.cand_count, 1);
          if (slot < p.cand_cap)
            p.cand[slot] = ((unsigned long long)p.scale << 56) | ((unsigned long long)plane_id << 32) | ((unsigned long long)yr << 16) |
                           (unsigned long long)(x0 - BP_HALO + 4 * lane + k);
          else
            atomicOr(p.status, 1);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { hu[k] = hc[k]; hc[k] = hd[k]; }
  }
}

struct ExactParams {
  const unsigned long long* acc[KG_MAX_SCALES];
  int H[KG_MAX_SCALES], W[KG_MAX_SCALES];
  const unsigned long long* cand; const int* cand_count; int cand_cap;
  double peak_thresh;
  int n_scales, max_peaks;
  double* peak_conf; int* peak_key; int* peak_count;
  int* status;
};

// K2b: exact fp64 re-evaluation of the candidates (postprocessing.py:56-64,143-144; same operation order as blur_peak_kernel).
// Lane c (0..18) owns image column x - 9 + c: 19 input rows y - 9 .. y + 9 -> vertical sums of rows y - 1, y, y + 1; lanes
// 0..4 then each evaluate one of the five horizontal sums from the warp's shared scratch.
__global__ void __launch_bounds__(128) exact_peaks_kernel(const ExactParams p) {
  __shared__ double s_t[4][3][20];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int total = min(*p.cand_count, p.cand_cap);
  const int warps_total = gridDim.x * 4;
  for (int ci = blockIdx.x * 4 + warp; ci < total; ci += warps_total) {
    const unsigned long long e = p.cand[ci];
    const int s = (int)(e >> 56), plane_id = (int)((e >> 32) & 0xffffffu), y = (int)((e >> 16) & 0xffffu), x = (int)(e & 0xffffu);
    const int H = p.H[s], W = p.W[s];
    const size_t hw = (size_t)H * W;
    const unsigned long long* plane = p.acc[s] + (size_t)plane_id * hw;
    double tv[3] = {0., 0., 0.};
    if (lane < 19) {
      const unsigned long long* col = plane + reflect_index(x - 9 + lane, W);
      double a[19];
#pragma unroll
      for (int i = 0; i < 19; ++i) a[i] = __longlong_as_double((long long)__ldg(col + (size_t)reflect_index(y - 9 + i, H) * W));
#pragma unroll
      for (int i = 0; i < 19; ++i) a[i] = acc_to_heat(a[i]);
#pragma unroll
      for (int r = 0; r < 3; ++r) {                    // rows y - 1, y, y + 1: centre at a[8 + r]
        double t = a[8 + r] * c_gauss[GAUSS_R];
#pragma unroll
        for (int j = 0; j < GAUSS_R; ++j) t += (a[r + j] + a[r + 16 - j]) * c_gauss[j];
        tv[r] = t;
      }
      s_t[warp][0][lane] = tv[0]; s_t[warp][1][lane] = tv[1]; s_t[warp][2][lane] = tv[2];
    }
    __syncwarp();
    // lane 0: h(y, x); 1: h(y - 1, x); 2: h(y + 1, x); 3: h(y, x - 1); 4: h(y, x + 1)
    double h = 0.;
    if (lane < 5) {
      const int row = lane == 1 ? 0 : (lane == 2 ? 2 : 1);
      const int cc = lane == 3 ? 8 : (lane == 4 ? 10 : 9);          // scratch column of the output's centre (x <-> 9)
      const double* t = s_t[warp][row];
      h = t[cc] * c_gauss[GAUSS_R];
#pragma unroll
      for (int j = 0; j < GAUSS_R; ++j) h += (t[cc - 8 + j] + t[cc + 8 - j]) * c_gauss[j];
    }
    const double hc = __shfl_sync(0xffffffffu, h, 0), hup = __shfl_sync(0xffffffffu, h, 1), hdn = __shfl_sync(0xffffffffu, h, 2);
    const double hl = __shfl_sync(0xffffffffu, h, 3), hr = __shfl_sync(0xffffffffu, h, 4);
    if (lane == 0) {
      double m = hc;
      if (y > 0) m = fmax(m, hup);
      if (y < H - 1) m = fmax(m, hdn);
      if (x > 0) m = fmax(m, hl);
      if (x < W - 1) m = fmax(m, hr);
      if (m == hc && hc > p.peak_thresh) {
        const int n = plane_id / 5, ch = plane_id - n * 5;
        const int list = n * p.n_scales + s;
        const int slot = atomicAdd(p.peak_count + list, 1);
        if (slot < p.max_peaks) {
          p.peak_conf[(size_t)list * p.max_peaks + slot] = hc;
          p.peak_key[(size_t)list * p.max_peaks + slot] = ch * (int)hw + y * W + x;
        } else {
          atomicOr(p.status, 1);
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// K3: per (image, scale): sort peaks (conf desc, generation order asc == Python's stable
// list.sort(reverse=True), postprocessing.py:87), greedy grouping (:98-124), refine_skeleton
// (:150-159) and skeleton_to_box (:164-242).
struct GroupParams {
  const float* mid[KG_MAX_SCALES];
  int H[KG_MAX_SCALES], W[KG_MAX_SCALES], box_scale[KG_MAX_SCALES];
  int n_scales, max_peaks;
};

__device__ __forceinline__ bool peak_before(double ca, int ka, double cb, int kb) {
  return ca > cb || (ca == cb && ka < kb);
}

// refine_skeleton + skeleton_to_box for one skeleton.  sk = 15 doubles (x,y,conf)x5.  Returns
// bit0: refine keeps it, bit1: a box was produced (box[5] = y1,x1,y2,x2,conf).
__device__ int skeleton_box(const double* __restrict__ sk_in, double scale, bool apply_refine, double* box) {
  double x[5], y[5], c[5];
  bool m[5];
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    x[k] = sk_in[3 * k] * scale; y[k] = sk_in[3 * k + 1] * scale; c[k] = sk_in[3 * k + 2];   // skeleton[:, :2] *= scale (:169)
    m[k] = x[k] > 0.;                                                                      // (:175)
    cnt += m[k];
  }
  const bool keep = cnt >= 3 || (m[0] && m[3]) || (m[1] && m[2]);                            // (:153-157)
  if (apply_refine && !keep) return 0;
  int res = keep ? 1 : 0;
  double csum = 0.;
  bool first = true;
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (m[k]) { csum = first ? c[k] : csum + c[k]; first = false; }
  const double conf = csum / (double)cnt;                                                  // skeleton[mask,2].mean()
  const int nc = m[0] + m[1] + m[2] + m[3];
  double y1, x1, y2, x2;
  if (nc == 4) {
    y1 = fmin(y[0], y[1]); y2 = fmax(y[2], y[3]); x1 = fmin(x[0], x[2]); x2 = fmax(x[1], x[3]);
  } else if (nc == 3) {
    y1 = (m[0] && m[1]) ? fmin(y[0], y[1]) : fmax(y[0], y[1]);
    y2 = fmax(y[2], y[3]);
    x1 = (m[0] && m[2]) ? fmin(x[0], x[2]) : fmax(x[0], x[2]);
    x2 = fmax(x[1], x[3]);
  } else if (nc == 2) {
    if (m[0] && m[3]) { y1 = y[0]; y2 = y[3]; x1 = x[0]; x2 = x[3]; }
    else if (m[1] && m[2]) { y1 = y[1]; y2 = y[2]; x1 = x[2]; x2 = x[1]; }
    else if (m[0] && m[1] && m[4]) { y1 = fmin(y[0], y[1]); y2 = y1 + (y[4] - y1) * 2.; x1 = x[0]; x2 = x[1]; }
    else if (m[0] && m[2] && m[4]) { y1 = y[0]; y2 = y[2]; x1 = fmin(x[0], x[2]); x2 = x1 + (x[4] - x1) * 2.; }
    else if (m[1] && m[3] && m[4]) { y1 = y[1]; y2 = y[3]; x2 = fmax(x[1], x[3]); x1 = x2 - (x2 - x[4]) * 2.; }
    else if (m[2] && m[3] && m[4]) { y2 = fmax(y[2], y[3]); y1 = y2 - (y2 - y[4]) * 2.; x1 = x[2]; x2 = x[3]; }
    else return res;
  } else {
    return res;
  }
  box[0] = y1; box[1] = x1; box[2] = y2; box[3] = x2; box[4] = conf;
  return res | 2;
}

// ordered compaction of per-skeleton boxes; called by the whole CTA
__device__ void boxes_from_skeletons(const double* __restrict__ skel, int nskel, double scale, bool apply_refine,
                                     double* __restrict__ boxes, int* __restrict__ n_boxes, uint8_t* __restrict__ keep) {
  __shared__ int s_warp_cnt[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  int base = 0;
  for (int q0 = 0; q0 < nskel; q0 += blockDim.x) {
    const int q = q0 + tid;
    double box[5];
    int r = 0;
    if (q < nskel) {
      r = skeleton_box(skel + (size_t)q * 15, scale, apply_refine, box);
      if (keep != nullptr) keep[q] = (uint8_t)(r & 1);
    }
    const bool has = (r & 2) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, has);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = 0, total = 0;
    for (int w = 0; w < nwarps; ++w) {
      const int c = s_warp_cnt[w];
      if (w < warp) off += c;
      total += c;
    }
    if (has) {
      double* dst = boxes + (size_t)(base + off + __popc(bal & ((1u << lane) - 1u))) * 5;
#pragma unroll
      for (int k = 0; k < 5; ++k) dst[k] = box[k];
    }
    base += total;
    __syncthreads();
  }
  if (tid == 0) *n_boxes = base;
}

// Spatial hash of the peaks of one list: cells of 8 x 8 pixels per keypoint type, chained through s_next.  Both neighbourhood
// queries of the greedy loop are radius searches around a point -- partner within KP_RADIUS + 1 = 6 px of the proposal (:113-114),
// skeleton keypoint of the seed's type within 10 px of the seed (:100) -- so they only have to visit 3 x 3 resp. 5 x 5 cells
// instead of every peak of the list (the first version scanned all K peaks per seed and target: O(K^2) with three block-wide
// barriers per scan, 30 ms for a 2 500-peak list).  With the scans gone the loop is latency-bound and runs on ONE warp
// (__syncwarp instead of __syncthreads).
constexpr int GH_CELL_SHIFT = 3;

__device__ __forceinline__ unsigned group_hash(int id, int cy, int cx, unsigned mask) {
  const unsigned key = ((unsigned)id * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)cx * 83492791u);
  return (key ^ (key >> 15)) & mask;
}

__global__ void __launch_bounds__(256) group_kernel(GroupParams gp, double* __restrict__ peak_conf_g,
                                                    int* __restrict__ peak_key_g, int* __restrict__ peak_count_g,
                                                    double* __restrict__ skel_g, int* __restrict__ skel_xy_g,
                                                    int* __restrict__ skel_count_g, uint8_t* __restrict__ skel_keep_g,
                                                    double* __restrict__ sbox_g, int* __restrict__ sbox_count_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = gp.max_peaks;
  double* s_conf = reinterpret_cast<double*>(smem_raw);
  int* s_key = reinterpret_cast<int*>(s_conf + P);
  unsigned short* s_px = reinterpret_cast<unsigned short*>(s_key + P);
  unsigned short* s_py = s_px + P;
  unsigned short* s_next = s_py + P;                   // hash chain: next peak of the same bucket (0xffff = end)
  unsigned short* s_head = s_next + P;                 // [2 * P] bucket heads
  unsigned char* s_id = reinterpret_cast<unsigned char*>(s_head + 2 * P);
  unsigned char* s_state = s_id + P;                   // bit0: alive (not yet consumed), bit1: keypoint of an existing skeleton

  const int list = blockIdx.x;
  const int n = list / gp.n_scales, s = list - n * gp.n_scales;
  const int H = gp.H[s], W = gp.W[s], hw = H * W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = min(peak_count_g[list], P);
  double* gconf = peak_conf_g + (size_t)list * P;
  int* gkey = peak_key_g + (size_t)list * P;

  int Kp = 1;
  while (Kp < K) Kp <<= 1;
  for (int e = tid; e < Kp; e += blockDim.x) {
    s_conf[e] = e < K ? gconf[e] : -1.;
    s_key[e] = e < K ? gkey[e] : INT_MAX;
  }
  __syncthreads();
  for (int k = 2; k <= Kp; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = tid; e < Kp; e += blockDim.x) {
        const int p = e ^ j;
        if (p > e) {
          const double ca = s_conf[e], cb = s_conf[p];
          const int ka = s_key[e], kb = s_key[p];
          const bool up = (e & k) == 0;
          const bool in_order = peak_before(ca, ka, cb, kb);
          if (up != in_order) { s_conf[e] = cb; s_conf[p] = ca; s_key[e] = kb; s_key[p] = ka; }
        }
      }
      __syncthreads();
    }
  }
  unsigned hmask = 1;
  while ((int)hmask < 2 * max(K, 1)) hmask <<= 1;
  hmask = min(hmask, (unsigned)(2 * P)) - 1u;
  for (int e = tid; e <= (int)hmask; e += blockDim.x) s_head[e] = 0xffffu;
  for (int e = tid; e < K; e += blockDim.x) {
    const int key = s_key[e];
    const int id = key / hw, rem = key - id * hw;
    const int y = rem / W;
    s_id[e] = (unsigned char)id; s_py[e] = (unsigned short)y; s_px[e] = (unsigned short)(rem - y * W);
    s_state[e] = 1;
    gconf[e] = s_conf[e];     // export the sorted order (doubles as the caller-visible peak list)
    gkey[e] = key;
  }
  __syncthreads();
  // chains are built by one thread per BUCKET-independent insertion order: a serial pass keeps them deterministic (rank-ascending
  // from the head is not required: every query applies its own explicit tie-break)
  if (tid == 0) {
    for (int e = K - 1; e >= 0; --e) {
      const unsigned b = group_hash(s_id[e], s_py[e] >> GH_CELL_SHIFT, s_px[e] >> GH_CELL_SHIFT, hmask);
      s_next[e] = s_head[b];
      s_head[b] = (unsigned short)e;
    }
  }
  __syncthreads();

  // ---- greedy grouping (postprocessing.py:98-124), one warp ----
  double* skel = skel_g + (size_t)list * P * 15;
  int* skel_xy = skel_xy_g + (size_t)list * P * 5;      // kept for the debug export of integer skeletons
  const float* mid = gp.mid[s] + (size_t)n * 40 * hw;
  __shared__ int s_nskel;
  if (warp == 0) {
    int nskel = 0;
    int n_missing = 0;                                   // lane t: skeletons without a keypoint of type t -- their s[t] sits at (0, 0) (:100 quirk)
    // The seed's four mid offsets (8 floats, scattered over 8 channel planes) are on the critical path of the sequential loop:
    // they are fetched four ranks AHEAD, one value per lane (lane = 8 * (peak - base) + 2 * kk + component), and handed out by
    // shuffle when their peak becomes the seed.
    auto fetch_mid = [&](int base) -> float {
      const int pk = base + (lane >> 3);
      if (pk >= K) return 0.f;
      const int id = s_id[pk], kk = (lane >> 1) & 3, t = kk + (kk >= id ? 1 : 0);
      const int m = c_mid_index[id][t];
      return __ldg(mid + (size_t)(2 * m + (lane & 1)) * hw + (int)s_py[pk] * W + (int)s_px[pk]);          // (:110-112)
    };
    float mid_cur = fetch_mid(0), mid_next = fetch_mid(4);
    for (int i = 0; i < K; ++i) {
      if (i > 0 && (i & 3) == 0) { mid_cur = mid_next; mid_next = fetch_mid(i + 4); }
      if (!(s_state[i] & 1)) continue;                   // consumed earlier (keypoints.pop(matches[0][0]))
      const int sid = s_id[i], sx = s_px[i], sy = s_py[i];
      // -- any(norm(kp.xy - s[kp.id,:2]) <= 10) over the existing skeletons: exact in integers --
      int hit = (__shfl_sync(0xffffffffu, n_missing, sid) > 0 && sx * sx + sy * sy <= 100) ? 1 : 0;
      if (lane < 25) {
        const int cy = (sy >> GH_CELL_SHIFT) + lane / 5 - 2, cx = (sx >> GH_CELL_SHIFT) + lane % 5 - 2;
        if (cy >= 0 && cx >= 0 && (cy << GH_CELL_SHIFT) < H && (cx << GH_CELL_SHIFT) < W) {
          for (unsigned j = s_head[group_hash(sid, cy, cx, hmask)]; j != 0xffffu; j = s_next[j]) {
            if ((s_state[j] & 2) && s_id[j] == sid) {
              const int ddx = sx - (int)s_px[j], ddy = sy - (int)s_py[j];
              hit |= (ddx * ddx + ddy * ddy <= 100);
            }
          }
        }
      }
      if (__any_sync(0xffffffffu, hit)) { if (lane == 0) s_state[i] = 0; __syncwarp(); continue; }   // keypoints.pop(0) happened (:99)
      // -- the four partner searches (targets in ascending id, BFS order over K5): lane = 8 * kk + c handles target kk, cell c of the
      //    3 x 3 neighbourhood of the proposal (lane c == 0 also takes the ninth cell); an 8-lane butterfly picks the nearest --
      const int kk = lane >> 3, tt = kk + (kk >= sid ? 1 : 0);
      const float mox = __shfl_sync(0xffffffffu, mid_cur, 8 * (i & 3) + 2 * kk);
      const float moy = __shfl_sync(0xffffffffu, mid_cur, 8 * (i & 3) + 2 * kk + 1);
      const double prx = (double)sx + (double)mox, pry = (double)sy + (double)moy;
      double best = DBL_MAX;
      int bestj = INT_MAX;
      {
        // candidates lie within 6 px of the proposal: cells floor(pr / 8) - 1 .. + 1 cover them
        const double fcx = floor(prx * 0.125), fcy = floor(pry * 0.125);
        if (fcx >= -1. && fcy >= -1. && fcx <= 8192. && fcy <= 8192.) {
          const int ncell = (lane & 7) == 0 ? 2 : 1;
          for (int q = 0; q < ncell; ++q) {
            const int cell = q == 0 ? (lane & 7) : 8;
            const int cy = (int)fcy + cell / 3 - 1, cx = (int)fcx + cell % 3 - 1;
            if (cy >= 0 && cx >= 0 && (cy << GH_CELL_SHIFT) < H && (cx << GH_CELL_SHIFT) < W) {
              for (unsigned j = s_head[group_hash(tt, cy, cx, hmask)]; j != 0xffffu; j = s_next[j]) {
                if ((int)j > i && (s_state[j] & 1) && s_id[j] == tt && (s_py[j] >> GH_CELL_SHIFT) == cy && (s_px[j] >> GH_CELL_SHIFT) == cx) {
                  const double ddx = prx - (double)s_px[j], ddy = pry - (double)s_py[j];
                  const double d = sqrt(ddx * ddx + ddy * ddy);                                       // np.linalg.norm (:114,117)
                  if (d <= 6.0 && (d < best || (d == best && (int)j < bestj))) { best = d; bestj = (int)j; }
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {                      // min over the 8 lanes of the target: distance, then rank (stable sort -> first minimum)
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
        if (ob < best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
      }
      // Lane t (0..4) owns keypoint type t of the new skeleton: (x, y, conf), all zero when missing.
      double kx = 0., ky = 0., kc = 0.;
      if (lane == sid) { kx = (double)sx; ky = (double)sy; kc = s_conf[i]; }
      {
        const int my_kk = lane - (lane > sid ? 1 : 0);        // lane t != sid <-> target index kk(t)
        const int bj = __shfl_sync(0xffffffffu, bestj, 8 * (my_kk & 3));
        if (lane < 5 && lane != sid) {
          if (bj != INT_MAX) { kx = (double)s_px[bj]; ky = (double)s_py[bj]; kc = s_conf[bj]; s_state[bj] = 2; }   // popped (:120), now a skeleton keypoint
          else ++n_missing;
        }
      }
      if (lane == 0) s_state[i] = 2;
      if (lane < 5) {
        double* dst = skel + (size_t)nskel * 15 + 3 * lane;
        dst[0] = kx; dst[1] = ky; dst[2] = kc;
        skel_xy[nskel * 5 + lane] = (int)kx | ((int)ky << 16);
      }
      ++nskel;
      __syncwarp();
    }
    if (lane == 0) { skel_count_g[list] = nskel; s_nskel = nskel; }
  }
  __syncthreads();
  boxes_from_skeletons(skel, s_nskel, (double)gp.box_scale[s], true, sbox_g + (size_t)list * P * 5, sbox_count_g + list,
                       skel_keep_g ? skel_keep_g + (size_t)list * P : nullptr);
}

// ------------------------------------------------------------------------------------------------
constexpr int NMS_MASK_CAP = 2048;     // lists up to this length use the all-pairs bit-matrix path (512 KiB of workspace per image)

// K4: per image: gather_skeleton (postprocessing.py:255-261: scale 0..3 concatenated) and
// non_maximum_suppression_numpy (nms.py:4-53).  argsort ties: (conf, index) ascending.
__global__ void __launch_bounds__(256) nms_kernel(const double* __restrict__ sbox_g, const int* __restrict__ sbox_count_g,
                                                  int n_lists, int list_cap, int max_boxes, double nms_thresh,
                                                  double* __restrict__ boxes_g, int* __restrict__ box_count_g,
                                                  double* __restrict__ dets_g, int* __restrict__ det_count_g,
                                                  int* __restrict__ status, double* __restrict__ packed_g, int packed_k,
                                                  unsigned* __restrict__ mask_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_conf = reinterpret_cast<double*>(smem_raw);
  int* s_idx = reinterpret_cast<int*>(s_conf + max_boxes);
  unsigned char* s_supp = reinterpret_cast<unsigned char*>(s_idx + max_boxes);
  const int n = blockIdx.x, tid = threadIdx.x;
  double* boxes = boxes_g + (size_t)n * max_boxes * 5;
  double* dets = dets_g + (size_t)n * max_boxes * 5;

  int B = 0;
  for (int l = 0; l < n_lists; ++l) {
    const int cnt = sbox_count_g[n * n_lists + l];
    const double* src = sbox_g + (size_t)(n * n_lists + l) * list_cap * 5;
    const int room = max(0, min(cnt, max_boxes - B));
    for (int e = tid; e < room * 5; e += blockDim.x) boxes[(size_t)B * 5 + e] = src[e];
    if (cnt > room && tid == 0) atomicOr(status, 2);
    B += room;
  }
  if (tid == 0) box_count_g[n] = B;
  __syncthreads();
  int Bp = 1;
  while (Bp < B) Bp <<= 1;
  for (int e = tid; e < Bp; e += blockDim.x) {
    s_conf[e] = e < B ? boxes[(size_t)e * 5 + 4] : -DBL_MAX;
    s_idx[e] = e < B ? e : -1;
    s_supp[e] = 0;
  }
  __syncthreads();
  // descending (conf, index): the reference pops argsort(conf)[-1] first
  for (int k = 2; k <= Bp; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = tid; e < Bp; e += blockDim.x) {
        const int p = e ^ j;
        if (p > e) {
          const double ca = s_conf[e], cb = s_conf[p];
          const int ia = s_idx[e], ib = s_idx[p];
          const bool up = (e & k) == 0;
          const bool in_order = ca > cb || (ca == cb && ia > ib);
          if (up != in_order) { s_conf[e] = cb; s_conf[p] = ca; s_idx[e] = ib; s_idx[p] = ia; }
        }
      }
      __syncthreads();
    }
  }
  int nkeep = 0;
  if (mask_g != nullptr && B > 64 && B <= NMS_MASK_CAP) {
    // Dense lists: the suppression relation of ALL pairs first (bit b of row a: IoU(a, b) > thresh, b after a in the sorted order;
    // same fp64 expression as below, every pair independent), then one warp walks the sorted list with the removed-set in
    // registers.  The greedy loop below spends a block-wide scan + barrier per KEPT box: 0.7 ms for a 700-box image.
    const int words = (B + 31) >> 5;
    unsigned* mask = mask_g + (size_t)n * NMS_MASK_CAP * (NMS_MASK_CAP / 32);
    // the sorted boxes are staged in shared memory: the pair loop is latency-bound on the box loads otherwise
    double* s_box = reinterpret_cast<double*>(smem_raw + (((size_t)max_boxes * 13 + 15) & ~(size_t)15));
    for (int e = tid; e < B * 4; e += blockDim.x) s_box[e] = boxes[(size_t)s_idx[e >> 2] * 5 + (e & 3)];
    __syncthreads();
    for (int e = tid; e < B * words; e += blockDim.x) {
      const int a = e / words, w = e - a * words;
      unsigned bits = 0u;
      if (w * 32 + 31 > a) {
        const double* ca = s_box + 4 * a;
        const double cy1 = ca[0], cx1 = ca[1], cy2 = ca[2], cx2 = ca[3];
        const double carea = (cx2 - cx1) * (cy2 - cy1);
        for (int k = 0; k < 32; ++k) {
          const int b = w * 32 + k;
          if (b <= a || b >= B) continue;
          const double* o = s_box + 4 * b;
          const double yy1 = fmax(o[0], cy1), xx1 = fmax(o[1], cx1), yy2 = fmin(o[2], cy2), xx2 = fmin(o[3], cx2);
          const double iw = fmax(0., xx2 - xx1), ih = fmax(0., yy2 - yy1);
          const double inter = iw * ih;
          const double oarea = (o[3] - o[1]) * (o[2] - o[0]);
          const double iou = inter / ((oarea - inter) + carea);
          if (!(iou <= nms_thresh)) bits |= 1u << k;
        }
      }
      mask[(size_t)a * words + w] = bits;
    }
    __syncthreads();
    if (tid < 32) {
      unsigned removed[NMS_MASK_CAP / 1024];               // lane l owns words l, l + 32, ...
#pragma unroll
      for (int q = 0; q < NMS_MASK_CAP / 1024; ++q) removed[q] = 0u;
      for (int a = 0; a < B; ++a) {
        const int w = a >> 5;
        unsigned mine = 0u;
#pragma unroll
        for (int q = 0; q < NMS_MASK_CAP / 1024; ++q) if (q == (w >> 5)) mine = removed[q];
        const unsigned word = __shfl_sync(0xffffffffu, mine, w & 31);
        if ((word >> (a & 31)) & 1u) continue;
        if (tid < 5) dets[(size_t)nkeep * 5 + tid] = boxes[(size_t)s_idx[a] * 5 + tid];
        ++nkeep;
#pragma unroll
        for (int q = 0; q < NMS_MASK_CAP / 1024; ++q) {
          const int ww = q * 32 + tid;
          if (ww < words) removed[q] |= mask[(size_t)a * words + ww];
        }
      }
    }
    nkeep = __shfl_sync(0xffffffffu, nkeep, 0);            // (warp 0 only has the count)
    __shared__ int s_nkeep;
    if (tid == 0) s_nkeep = nkeep;
    __syncthreads();
    nkeep = s_nkeep;
  } else
  for (int a = 0; a < B; ++a) {
    if (s_supp[a]) continue;
    const int c = s_idx[a];
    const double cy1 = boxes[(size_t)c * 5], cx1 = boxes[(size_t)c * 5 + 1], cy2 = boxes[(size_t)c * 5 + 2],
                 cx2 = boxes[(size_t)c * 5 + 3];
    if (tid < 5) dets[(size_t)nkeep * 5 + tid] = boxes[(size_t)c * 5 + tid];
    ++nkeep;
    const double carea = (cx2 - cx1) * (cy2 - cy1);                                  // (:15)
    for (int b = a + 1 + tid; b < B; b += blockDim.x) {
      if (s_supp[b]) continue;
      const double* o = boxes + (size_t)s_idx[b] * 5;
      const double yy1 = fmax(o[0], cy1), xx1 = fmax(o[1], cx1), yy2 = fmin(o[2], cy2), xx2 = fmin(o[3], cx2);   // (:33-36)
      const double w = fmax(0., xx2 - xx1), h = fmax(0., yy2 - yy1);
      const double inter = w * h;
      const double oarea = (o[3] - o[1]) * (o[2] - o[0]);
      const double iou = inter / ((oarea - inter) + carea);                          // (:47-48)
      if (!(iou <= nms_thresh)) s_supp[b] = 1;                                       // keeps IoU<=thr; NaN is dropped (:49)
    }
    __syncthreads();
  }
  if (tid == 0) det_count_g[n] = nkeep;
  if (packed_g != nullptr) {
    // fixed-size record of this image for the data-parallel all-gather: row 0 = (count, kept rows, 0, 0, 0), rows 1.. = detections
    double* rec = packed_g + (size_t)n * (packed_k + 1) * 5;
    const int kept = min(nkeep, packed_k);
    __syncthreads();                                     // dets rows written by threads 0..4 above
    if (tid < 5) rec[tid] = tid == 0 ? (double)nkeep : (tid == 1 ? (double)kept : 0.);
    for (int e = tid; e < kept * 5; e += blockDim.x) rec[5 + e] = dets[e];
    if (nkeep > packed_k && tid == 0) atomicOr(status, 4);
  }
}

// nms.py:4-53 for lists that do not fit the shared-memory kernel (kg_nms_host with > 8192 boxes): one CTA, sort keys and
// suppression flags in global scratch ([cap] f64 conf, [cap] i32 index, [cap] u8 flag).  Same order and arithmetic as nms_kernel.
__global__ void __launch_bounds__(1024) nms_big_kernel(const double* __restrict__ boxes, int B, int Bp, double nms_thresh,
                                                       unsigned char* __restrict__ scratch, double* __restrict__ dets,
                                                       int* __restrict__ det_count) {
  double* s_conf = reinterpret_cast<double*>(scratch);
  int* s_idx = reinterpret_cast<int*>(s_conf + Bp);
  unsigned char* s_supp = reinterpret_cast<unsigned char*>(s_idx + Bp);
  const int tid = threadIdx.x;
  for (int e = tid; e < Bp; e += blockDim.x) {
    s_conf[e] = e < B ? boxes[(size_t)e * 5 + 4] : -DBL_MAX;
    s_idx[e] = e < B ? e : -1;
    s_supp[e] = 0;
  }
  __syncthreads();
  for (int k = 2; k <= Bp; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = tid; e < Bp; e += blockDim.x) {
        const int p = e ^ j;
        if (p > e) {
          const double ca = s_conf[e], cb = s_conf[p];
          const int ia = s_idx[e], ib = s_idx[p];
          const bool up = (e & k) == 0;
          const bool in_order = ca > cb || (ca == cb && ia > ib);
          if (up != in_order) { s_conf[e] = cb; s_conf[p] = ca; s_idx[e] = ib; s_idx[p] = ia; }
        }
      }
      __syncthreads();
    }
  int nkeep = 0;
  for (int a = 0; a < B; ++a) {
    if (s_supp[a]) continue;
    const int c = s_idx[a];
    const double cy1 = boxes[(size_t)c * 5], cx1 = boxes[(size_t)c * 5 + 1], cy2 = boxes[(size_t)c * 5 + 2], cx2 = boxes[(size_t)c * 5 + 3];
    if (tid < 5) dets[(size_t)nkeep * 5 + tid] = boxes[(size_t)c * 5 + tid];
    ++nkeep;
    const double carea = (cx2 - cx1) * (cy2 - cy1);
    for (int b = a + 1 + tid; b < B; b += blockDim.x) {
      if (s_supp[b]) continue;
      const double* o = boxes + (size_t)s_idx[b] * 5;
      const double yy1 = fmax(o[0], cy1), xx1 = fmax(o[1], cx1), yy2 = fmin(o[2], cy2), xx2 = fmin(o[3], cx2);
      const double w = fmax(0., xx2 - xx1), h = fmax(0., yy2 - yy1);
      const double inter = w * h;
      const double oarea = (o[3] - o[1]) * (o[2] - o[0]);
      const double iou = inter / ((oarea - inter) + carea);
      if (!(iou <= nms_thresh)) s_supp[b] = 1;
    }
    __syncthreads();
  }
  if (tid == 0) *det_count = nkeep;
}

// standalone refine/box kernel for host lists (kg_skeletons_to_boxes_host)
__global__ void __launch_bounds__(256) skeleton_box_kernel(const double* __restrict__ skel, int nskel, double scale,
                                                           int apply_refine, double* __restrict__ boxes,
                                                           int* __restrict__ n_boxes, uint8_t* __restrict__ keep) {
  boxes_from_skeletons(skel, nskel, scale, apply_refine != 0, boxes, n_boxes, keep);
}

// ------------------------------------------------------------------------------------------------
static int num_sms() {
  static int sms = 0;
  if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  return sms;
}

// rows per segment: every segment re-reads 18 halo rows, every wave of the grid costs one segment time -> minimise waves * (RH + 18)
static int pick_rows(int H, long long units_per_row_segment, long long slots) {
  int best_rh = H; double best_cost = 1e30;
  for (int rh = 16; rh <= H; rh += 8) {
    const long long units = units_per_row_segment * ceil_div(H, rh);
    const double waves = (double)((units + slots - 1) / slots);
    const double cost = waves * (rh + 2 * BP_HALO);
    if (cost < best_cost) { best_cost = cost; best_rh = rh; }
  }
  if (H < 16) best_rh = H;
  const char* env = getenv("KG_BLUR_RH");
  if (env && atoi(env) > 0) best_rh = std::min(H, atoi(env));
  return best_rh;
}

// All-fp64 streaming filter (blur_peak_kernel): exports the fp64 maps, and the peak list when bp.emit_peaks is set.
// Column strips of at most 270 outputs (288 threads, two CTAs per SM at 96 registers).
static int launch_blur_peak(BlurParams bp, int N, cudaStream_t stream) {
  const int W = bp.W, H = bp.H;
  bp.n_strips = ceil_div(W, 270);
  bp.CW = ceil_div(W, bp.n_strips);
  const int need = bp.CW + 2 * BP_HALO;
  const int nt = need <= 96 ? 96 : need <= 160 ? 160 : 288;
  const int per_sm = std::max(1, 65536 / (112 * nt));
  bp.RH = pick_rows(H, (long long)N * 5 * bp.n_strips, (long long)num_sms() * per_sm);
  dim3 grid(bp.n_strips, ceil_div(H, bp.RH), N * 5);
  KG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "blur_peak: grid too large");
  switch (nt) {
    case 96: blur_peak_kernel<96><<<grid, 96, 0, stream>>>(bp); break;
    case 160: blur_peak_kernel<160><<<grid, 160, 0, stream>>>(bp); break;
    default: blur_peak_kernel<288><<<grid, 288, 0, stream>>>(bp); break;
  }
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

// fp32 prefilter (blur32_candidates_kernel): warp-private strips of 110 output columns, 4 warps per CTA, 16 warps per SM at 128 registers
static int launch_blur32(BlurParams bp, int N, cudaStream_t stream) {
  bp.n_strips = ceil_div(bp.W, B32_CW);
  bp.RH = pick_rows(bp.H, (long long)N * 5 * bp.n_strips, (long long)num_sms() * 16);
  const int items = bp.n_strips * ceil_div(bp.H, bp.RH);
  dim3 grid(ceil_div(items, 4), N * 5, 1);
  KG_REQUIRE(grid.y <= 65535, "blur32: grid too large");
  blur32_candidates_kernel<<<grid, 128, 0, stream>>>(bp);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

struct DecodeWorkspace {
  unsigned long long* acc[KG_MAX_SCALES];
  double* peak_conf; int* peak_key; int* peak_count;
  unsigned long long* cand; int* cand_count; int cand_cap;
  double* skel; int* skel_xy; int* skel_count;
  double* sbox; int* sbox_count;
  double* boxes; int* box_count;
  double* dets; int* det_count;
  unsigned* nms_mask;                      // [N, NMS_MASK_CAP, NMS_MASK_CAP / 32] suppression bit matrix of the dense-list NMS path
  char* zero_begin; size_t zero_bytes;     // region cleared at the start of every call
  size_t total;
};

static DecodeWorkspace carve(const kg_decode_config& cfg, const kg_decode_scale* sc, void* base) {
  DecodeWorkspace w{};
  Arena a(base, (size_t)-1);
  const size_t lists = (size_t)cfg.N * cfg.n_scales, P = cfg.max_peaks;
  // zeroed region first: vote accumulators + counters
  char* z0 = a.take<char>(0);
  for (int s = 0; s < cfg.n_scales; ++s) w.acc[s] = a.take<unsigned long long>((size_t)cfg.N * 5 * sc[s].H * sc[s].W);
  w.peak_count = a.take<int>(lists);
  w.cand_count = a.take<int>(1);
  char* z1 = a.take<char>(0);
  w.zero_begin = z0; w.zero_bytes = (size_t)(z1 - z0);
  w.peak_conf = a.take<double>(lists * P);
  w.peak_key = a.take<int>(lists * P);
  w.cand_cap = (int)std::min<size_t>(lists * P, (size_t)1 << 26);
  w.cand = a.take<unsigned long long>((size_t)w.cand_cap);
  w.skel = a.take<double>(lists * P * 15);
  w.skel_xy = a.take<int>(lists * P * 5);
  w.skel_count = a.take<int>(lists);
  w.sbox = a.take<double>(lists * P * 5);
  w.sbox_count = a.take<int>(lists);
  w.boxes = a.take<double>((size_t)cfg.N * cfg.max_boxes * 5);
  w.box_count = a.take<int>(cfg.N);
  w.dets = a.take<double>((size_t)cfg.N * cfg.max_boxes * 5);
  w.det_count = a.take<int>(cfg.N);
  w.nms_mask = a.take<unsigned>((size_t)cfg.N * NMS_MASK_CAP * (NMS_MASK_CAP / 32));
  w.total = align_up(a.off, 256);
  return w;
}

static int check_config(const kg_decode_config* cfg, const kg_decode_scale* sc) {
  KG_REQUIRE(cfg != nullptr && sc != nullptr, "kg_decode: null config/scales");
  KG_REQUIRE(cfg->N >= 1 && cfg->N <= 65535, "kg_decode: N=%d out of range", cfg->N);
  KG_REQUIRE(cfg->n_scales >= 1 && cfg->n_scales <= KG_MAX_SCALES, "kg_decode: n_scales=%d", cfg->n_scales);
  KG_REQUIRE(is_pow2(cfg->max_peaks) && cfg->max_peaks >= 64 && cfg->max_peaks <= 8192,
             "kg_decode: max_peaks=%d must be a power of two in [64, 8192]", cfg->max_peaks);
  KG_REQUIRE(is_pow2(cfg->max_boxes) && cfg->max_boxes >= 64 && cfg->max_boxes <= 8192,
             "kg_decode: max_boxes=%d must be a power of two in [64, 8192]", cfg->max_boxes);
  for (int s = 0; s < cfg->n_scales; ++s) {
    KG_REQUIRE(sc[s].H >= 1 && sc[s].W >= 1 && sc[s].H <= 65535 && sc[s].W <= 65535 &&
                   (long long)sc[s].H * sc[s].W * 5 < (long long)INT_MAX,
               "kg_decode: scale %d has unsupported size %dx%d", s, sc[s].H, sc[s].W);
    KG_REQUIRE(sc[s].box_scale >= 1, "kg_decode: scale %d box_scale=%d", s, sc[s].box_scale);
  }
  return KG_OK;
}

size_t decode_workspace_bytes(const kg_decode_config* cfg, const kg_decode_scale* sc) {
  if (check_config(cfg, sc) != KG_OK) return 0;
  return carve(*cfg, sc, nullptr).total;
}

static size_t group_smem(int P) { return (size_t)P * 24 + 16; }
static size_t nms_smem(int B) { return (((size_t)B * 13 + 15) & ~(size_t)15) + (size_t)std::min(B, NMS_MASK_CAP) * 32; }   // + staged boxes of the bit-matrix path

int decode_launch(const kg_decode_config* cfg, const kg_decode_scale* sc, const kg_decode_outputs* out, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream, int* n_launches) {
  KG_TRY(check_config(cfg, sc));
  KG_REQUIRE(out != nullptr && out->d_status != nullptr, "kg_decode: outputs / d_status must be non-null");
  KG_REQUIRE(out->d_det_packed == nullptr || out->det_packed_k >= 1, "kg_decode: det_packed_k=%d", out->det_packed_k);
  KG_REQUIRE(workspace != nullptr, "kg_decode: null workspace");
  DecodeWorkspace w = carve(*cfg, sc, workspace);
  if (w.total > workspace_bytes) {
    set_error("kg_decode: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
    return KG_ERR_WORKSPACE;
  }
  for (int s = 0; s < cfg->n_scales; ++s)
    KG_REQUIRE(sc[s].d_kp && sc[s].d_short && sc[s].d_mid, "kg_decode: scale %d has a null head pointer", s);
  int launches = 0;
  const int N = cfg->N, S = cfg->n_scales, P = cfg->max_peaks;
  KG_CUDA_CHECK(cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, stream));
  KG_CUDA_CHECK(cudaMemsetAsync(out->d_status, 0, sizeof(int), stream));
  // KG_DECODE_FP64=1: the all-fp64 streaming filter produces the peak list (A/B and cross-check of the prefilter path)
  static const bool all_fp64 = getenv("KG_DECODE_FP64") != nullptr && getenv("KG_DECODE_FP64")[0] == '1';
  double* peak_conf = out->d_peak_conf ? out->d_peak_conf : w.peak_conf;
  int* peak_key = out->d_peak_key ? out->d_peak_key : w.peak_key;
  int* peak_count = w.peak_count;
  {
    // stage 0 = vote + prefilter of all scales, timed as one region.  (Running the scales' chains on concurrent side streams was
    // measured: 0.855 vs 0.853 ms -- the coarse scales' small grids are not what the time goes to.)
    StageScope t(0, stream);
    for (int s = 0; s < S; ++s) {
      const int H = sc[s].H, W = sc[s].W;
      dim3 g1(ceil_div(W, VT_W), ceil_div(H, VT_H), N * 5);
      vote_kernel<<<g1, 256, 0, stream>>>(sc[s].d_kp, sc[s].d_short, w.acc[s], H, W);
      BlurParams bp{};
      bp.acc = w.acc[s]; bp.H = H; bp.W = W; bp.peak_thresh = cfg->peak_thresh; bp.list_index_base = s; bp.n_scales = S; bp.max_peaks = P;
      bp.peak_conf = peak_conf; bp.peak_key = peak_key; bp.peak_count = peak_count;
      bp.out_vote = out->d_vote[s]; bp.out_heat = out->d_heat[s]; bp.status = out->d_status;
      bp.cand = w.cand; bp.cand_count = w.cand_count; bp.cand_cap = w.cand_cap; bp.scale = s;
      bp.emit_peaks = all_fp64 ? 1 : 0;
      if (all_fp64 || bp.out_vote != nullptr || bp.out_heat != nullptr) { KG_TRY(launch_blur_peak(bp, N, stream)); ++launches; }
      if (!all_fp64) { KG_TRY(launch_blur32(bp, N, stream)); ++launches; }
      launches += 1;
    }
  }
  if (!all_fp64) {
    StageScope t(1, stream);
    ExactParams ep{};
    for (int s = 0; s < S; ++s) { ep.acc[s] = w.acc[s]; ep.H[s] = sc[s].H; ep.W[s] = sc[s].W; }
    ep.cand = w.cand; ep.cand_count = w.cand_count; ep.cand_cap = w.cand_cap; ep.peak_thresh = cfg->peak_thresh;
    ep.n_scales = S; ep.max_peaks = P; ep.peak_conf = peak_conf; ep.peak_key = peak_key; ep.peak_count = peak_count; ep.status = out->d_status;
    exact_peaks_kernel<<<num_sms() * 4, 128, 0, stream>>>(ep);
    ++launches;
  }
  GroupParams gp{};
  for (int s = 0; s < S; ++s) { gp.mid[s] = sc[s].d_mid; gp.H[s] = sc[s].H; gp.W[s] = sc[s].W; gp.box_scale[s] = sc[s].box_scale; }
  gp.n_scales = S; gp.max_peaks = P;
  double* skel = out->d_skeletons ? out->d_skeletons : w.skel;
  int* skel_count = out->d_skel_count ? out->d_skel_count : w.skel_count;
  KG_CUDA_CHECK(cudaFuncSetAttribute(group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)group_smem(8192)));
  {
    StageScope t(2, stream);
    group_kernel<<<N * S, 256, group_smem(P), stream>>>(gp, peak_conf, peak_key, peak_count, skel, w.skel_xy, skel_count,
                                                        out->d_skel_keep, w.sbox, w.sbox_count);
  }
  KG_CUDA_CHECK(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem(8192)));
  {
    StageScope t(3, stream);
    nms_kernel<<<N, 256, nms_smem(cfg->max_boxes), stream>>>(w.sbox, w.sbox_count, S, P, cfg->max_boxes, cfg->nms_thresh,
                                                             out->d_boxes ? out->d_boxes : w.boxes,
                                                             out->d_box_count ? out->d_box_count : w.box_count,
                                                             out->d_dets ? out->d_dets : w.dets,
                                                             out->d_det_count ? out->d_det_count : w.det_count, out->d_status,
                                                             out->d_det_packed, out->det_packed_k, w.nms_mask);
  }
  launches += 2;
  if (out->d_peak_count != nullptr) {
    KG_CUDA_CHECK(cudaMemcpyAsync(out->d_peak_count, peak_count, sizeof(int) * N * S, cudaMemcpyDeviceToDevice, stream));
  }
  KG_CUDA_CHECK(cudaGetLastError());
  if (n_launches) *n_launches = launches;
  return KG_OK;
}

// ---- host-list helpers ------------------------------------------------------------------------
// (convenience entry points: allocate, copy and synchronise on a private stream; buffers are released on every path)
namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct OwnStream {
  cudaStream_t s = nullptr;
  ~OwnStream() { if (s) cudaStreamDestroy(s); }
};
}  // namespace

int skeletons_to_boxes_host(const double* h_skel, int n, int box_scale, int apply_refine, uint8_t* h_keep,
                            double* h_boxes, int* n_boxes) {
  KG_REQUIRE(n >= 0 && n_boxes != nullptr, "kg_skeletons_to_boxes_host: bad arguments");
  *n_boxes = 0;
  if (n == 0) return KG_OK;
  KG_REQUIRE(h_skel && h_boxes, "kg_skeletons_to_boxes_host: null buffer");
  DevBuf skel, boxes, cnt, keep;
  OwnStream st;
  KG_CUDA_CHECK(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
  KG_CUDA_CHECK(cudaMalloc(&skel.p, sizeof(double) * 15 * n));
  KG_CUDA_CHECK(cudaMalloc(&boxes.p, sizeof(double) * 5 * n));
  KG_CUDA_CHECK(cudaMalloc(&cnt.p, sizeof(int)));
  KG_CUDA_CHECK(cudaMalloc(&keep.p, n));
  KG_CUDA_CHECK(cudaMemcpyAsync(skel.p, h_skel, sizeof(double) * 15 * n, cudaMemcpyHostToDevice, st.s));
  skeleton_box_kernel<<<1, 256, 0, st.s>>>(skel.as<double>(), n, (double)box_scale, apply_refine, boxes.as<double>(), cnt.as<int>(),
                                           keep.as<uint8_t>());
  KG_CUDA_CHECK(cudaGetLastError());
  KG_CUDA_CHECK(cudaMemcpyAsync(n_boxes, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st.s));
  KG_CUDA_CHECK(cudaStreamSynchronize(st.s));
  if (*n_boxes > 0) KG_CUDA_CHECK(cudaMemcpyAsync(h_boxes, boxes.p, sizeof(double) * 5 * *n_boxes, cudaMemcpyDeviceToHost, st.s));
  if (h_keep) KG_CUDA_CHECK(cudaMemcpyAsync(h_keep, keep.p, n, cudaMemcpyDeviceToHost, st.s));
  KG_CUDA_CHECK(cudaStreamSynchronize(st.s));
  return KG_OK;
}

// nms.py:4-53 for a host list of any length: up to 8192 boxes in the shared-memory kernel, more through the
// global-memory variant (nms_big_kernel: same order, same arithmetic).
int nms_host(const double* h_boxes, int n, double nms_thresh, double* h_out, int* n_out) {
  KG_REQUIRE(n >= 0 && n_out != nullptr, "kg_nms_host: bad arguments");
  *n_out = 0;
  if (n == 0) return KG_OK;
  KG_REQUIRE(h_boxes && h_out, "kg_nms_host: null buffer");
  int cap = 64;
  while (cap < n) cap <<= 1;
  KG_REQUIRE(cap <= (1 << 20), "kg_nms_host: at most 2^20 boxes (got %d)", n);
  DevBuf in, boxes, dets, ints, scratch;
  OwnStream st;
  KG_CUDA_CHECK(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
  KG_CUDA_CHECK(cudaMalloc(&in.p, sizeof(double) * 5 * cap));
  KG_CUDA_CHECK(cudaMalloc(&boxes.p, sizeof(double) * 5 * cap));
  KG_CUDA_CHECK(cudaMalloc(&dets.p, sizeof(double) * 5 * cap));
  KG_CUDA_CHECK(cudaMalloc(&ints.p, sizeof(int) * 4));
  int h_ints[4] = {n, 0, 0, 0};   // [0] list count, [1] box_count, [2] det_count, [3] status
  KG_CUDA_CHECK(cudaMemcpyAsync(ints.p, h_ints, sizeof(h_ints), cudaMemcpyHostToDevice, st.s));
  KG_CUDA_CHECK(cudaMemcpyAsync(in.p, h_boxes, sizeof(double) * 5 * n, cudaMemcpyHostToDevice, st.s));
  int* d_ints = ints.as<int>();
  if (cap <= 8192) {
    DevBuf maskb;                                        // dense lists take the bit-matrix path, like kg_decode
    if (n <= NMS_MASK_CAP) KG_CUDA_CHECK(cudaMalloc(&maskb.p, (size_t)NMS_MASK_CAP * (NMS_MASK_CAP / 32) * sizeof(unsigned)));
    KG_CUDA_CHECK(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem(8192)));
    nms_kernel<<<1, 256, nms_smem(cap), st.s>>>(in.as<double>(), d_ints, 1, cap, cap, nms_thresh, boxes.as<double>(), d_ints + 1,
                                                dets.as<double>(), d_ints + 2, d_ints + 3, nullptr, 0, maskb.as<unsigned>());
    KG_CUDA_CHECK(cudaGetLastError());
    KG_CUDA_CHECK(cudaStreamSynchronize(st.s));          // maskb is released at the end of this scope
  } else {
    KG_CUDA_CHECK(cudaMalloc(&scratch.p, (size_t)cap * 13));
    nms_big_kernel<<<1, 1024, 0, st.s>>>(in.as<double>(), n, cap, nms_thresh, scratch.as<unsigned char>(), dets.as<double>(), d_ints + 2);
  }
  KG_CUDA_CHECK(cudaGetLastError());
  KG_CUDA_CHECK(cudaMemcpyAsync(h_ints, ints.p, sizeof(h_ints), cudaMemcpyDeviceToHost, st.s));
  KG_CUDA_CHECK(cudaStreamSynchronize(st.s));
  *n_out = h_ints[2];
  if (*n_out > 0) KG_CUDA_CHECK(cudaMemcpyAsync(h_out, dets.p, sizeof(double) * 5 * *n_out, cudaMemcpyDeviceToHost, st.s));
  KG_CUDA_CHECK(cudaStreamSynchronize(st.s));
  return KG_OK;
}

}  // namespace kg
