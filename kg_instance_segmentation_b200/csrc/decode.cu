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

#include <cfloat>
#include <climits>

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
// One thread per (image, keypoint channel, source pixel); four bilinear splats each.
__global__ void __launch_bounds__(256) vote_kernel(const float* __restrict__ kp, const float* __restrict__ sh,
                                                   unsigned long long* __restrict__ acc, int H, int W) {
  const int hw = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  const int i = blockIdx.y, n = blockIdx.z;
  const int y = p / W, x = p - y * W;
  const double ps = (double)__ldg(kp + ((size_t)n * 5 + i) * hw + p);
  const double xs = (double)x + (double)__ldg(sh + ((size_t)n * 10 + 2 * i) * hw + p);      // int64 + f32 -> f64 (:49)
  const double ys = (double)y + (double)__ldg(sh + ((size_t)n * 10 + 2 * i + 1) * hw + p);
  const double fx = floor(xs), fy = floor(ys), cx = ceil(xs), cy = ceil(ys);
  const double dx = xs - fx, dy = ys - fy;
  const double omdx = 1. - dx, omdy = 1. - dy;
  const double v[4] = {ps * omdx * omdy, ps * dx * omdy, ps * dy * omdx, ps * dy * dx};   // tl, tr, bl, br (:27-30)
  const double ty[4] = {fy, fy, cy, cy};
  const double tx[4] = {fx, cx, fx, cx};
  unsigned long long* plane = acc + ((size_t)n * 5 + i) * hw;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (ty[k] >= 0. && ty[k] < (double)H && tx[k] >= 0. && tx[k] < (double)W) {   // good_inds (:34-35)
      const long long q = __double2ll_rn(v[k] * KG_FIX);
      if (q != 0) atomicAdd(plane + (int)ty[k] * W + (int)tx[k], (unsigned long long)q);
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
// K2: heat = acc / (pi r^2); gaussian_filter(sigma=2) (postprocessing.py:143-144, scipy correlate1d
// symmetric branch: axis 0 then axis 1, paired taps, no FMA); get_keypoints (:56-64): cross-footprint
// local maximum and conf > peak_thresh.  One CTA per 32x32 output tile of one (image, channel) plane.
constexpr int BT = 32;                 // output tile edge
constexpr int BB = BT + 2;             // blurred tile (+1 halo for the peak test)
constexpr int BI = BB + 2 * GAUSS_R;   // input tile
__global__ void __launch_bounds__(256) blur_peak_kernel(const unsigned long long* __restrict__ acc, int H, int W,
                                                        double peak_thresh, int list_index_base, int n_scales,
                                                        int max_peaks, double* __restrict__ peak_conf,
                                                        int* __restrict__ peak_key, int* __restrict__ peak_count,
                                                        double* __restrict__ out_vote, double* __restrict__ out_heat,
                                                        int* __restrict__ status) {
  __shared__ double s_in[BI][BI];
  __shared__ double s_tmp[BB][BI];
  __shared__ double s_blur[BB][BB + 1];
  const int plane_id = blockIdx.z;                 // n*5 + i
  const int n = plane_id / 5, ch = plane_id - n * 5;
  const int hw = H * W;
  const unsigned long long* plane = acc + (size_t)plane_id * hw;
  const int x0 = blockIdx.x * BT, y0 = blockIdx.y * BT;
  const int tid = threadIdx.x;

  const bool interior = y0 >= GAUSS_R + 1 && x0 >= GAUSS_R + 1 && y0 - (GAUSS_R + 1) + BI <= H && x0 - (GAUSS_R + 1) + BI <= W;
  for (int e = tid; e < BI * BI; e += 256) {
    const int r = e / BI, c = e - r * BI;
    int gy = y0 - (GAUSS_R + 1) + r, gx = x0 - (GAUSS_R + 1) + c;
    if (!interior) { gy = reflect_index(gy, H); gx = reflect_index(gx, W); }
    const long long q = (long long)__ldg(plane + gy * W + gx);
    s_in[r][c] = ((double)q * KG_UNFIX) / KG_PI_R2;
  }
  __syncthreads();
  if (out_vote != nullptr) {
    for (int e = tid; e < BT * BT; e += 256) {
      const int r = e / BT, c = e - r * BT;
      const int gy = y0 + r, gx = x0 + c;
      if (gy < H && gx < W) out_vote[(size_t)plane_id * hw + gy * W + gx] = s_in[r + GAUSS_R + 1][c + GAUSS_R + 1];
    }
  }
  // axis-0 pass
  for (int e = tid; e < BB * BI; e += 256) {
    const int r = e / BI, c = e - r * BI;
    const int rc = r + GAUSS_R;
    double t = s_in[rc][c] * c_gauss[GAUSS_R];
#pragma unroll
    for (int j = -GAUSS_R; j < 0; ++j) t += (s_in[rc + j][c] + s_in[rc - j][c]) * c_gauss[j + GAUSS_R];
    s_tmp[r][c] = t;
  }
  __syncthreads();
  // axis-1 pass
  for (int e = tid; e < BB * BB; e += 256) {
    const int r = e / BB, c = e - r * BB;
    const int cc = c + GAUSS_R;
    double t = s_tmp[r][cc] * c_gauss[GAUSS_R];
#pragma unroll
    for (int j = -GAUSS_R; j < 0; ++j) t += (s_tmp[r][cc + j] + s_tmp[r][cc - j]) * c_gauss[j + GAUSS_R];
    s_blur[r][c] = t;
  }
  __syncthreads();
  for (int e = tid; e < BT * BT; e += 256) {
    const int r = e / BT, c = e - r * BT;
    const int gy = y0 + r, gx = x0 + c;
    if (gy >= H || gx >= W) continue;
    const double h = s_blur[r + 1][c + 1];
    if (out_heat != nullptr) out_heat[(size_t)plane_id * hw + gy * W + gx] = h;
    double m = h;
    if (gy > 0) m = fmax(m, s_blur[r][c + 1]);
    if (gy < H - 1) m = fmax(m, s_blur[r + 2][c + 1]);
    if (gx > 0) m = fmax(m, s_blur[r + 1][c]);
    if (gx < W - 1) m = fmax(m, s_blur[r + 1][c + 2]);
    if (m == h && h > peak_thresh) {
      const int list = n * n_scales + list_index_base;
      const int slot = atomicAdd(peak_count + list, 1);
      if (slot < max_peaks) {
        peak_conf[(size_t)list * max_peaks + slot] = h;
        peak_key[(size_t)list * max_peaks + slot] = ch * hw + gy * W + gx;
      } else {
        atomicOr(status, 1);
      }
    }
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
  unsigned char* s_id = reinterpret_cast<unsigned char*>(s_py + P);
  unsigned char* s_alive = s_id + P;
  __shared__ double s_best[8];
  __shared__ int s_bestj[8];

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
  for (int e = tid; e < K; e += blockDim.x) {
    const int key = s_key[e];
    const int id = key / hw, rem = key - id * hw;
    const int y = rem / W;
    s_id[e] = (unsigned char)id; s_py[e] = (unsigned short)y; s_px[e] = (unsigned short)(rem - y * W);
    s_alive[e] = 1;
    gconf[e] = s_conf[e];     // export the sorted order (doubles as the caller-visible peak list)
    gkey[e] = key;
  }
  __syncthreads();

  // ---- greedy grouping (postprocessing.py:98-124) ----
  double* skel = skel_g + (size_t)list * P * 15;
  int* skel_xy = skel_xy_g + (size_t)list * P * 5;
  const float* mid = gp.mid[s] + (size_t)n * 40 * hw;
  int nskel = 0;
  for (int i = 0; i < K; ++i) {
    if (!s_alive[i]) continue;                       // consumed earlier (keypoints.pop(matches[0][0]))
    const int sid = s_id[i], sx = s_px[i], sy = s_py[i];
    int hit = 0;
    for (int q = tid; q < nskel; q += blockDim.x) {  // any(norm(kp.xy - s[kp.id,:2]) <= 10) (:100); exact in integers
      const unsigned xy = (unsigned)skel_xy[q * 5 + sid];
      const int ddx = sx - (int)(xy & 0xffffu), ddy = sy - (int)(xy >> 16);
      hit |= (ddx * ddx + ddy * ddy <= 100);
    }
    if (__syncthreads_or(hit)) continue;
    double sk[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) sk[k] = 0.;
    int sk_xy[5] = {0, 0, 0, 0, 0};
    sk[3 * sid] = (double)sx; sk[3 * sid + 1] = (double)sy; sk[3 * sid + 2] = s_conf[i];
    sk_xy[sid] = sx | (sy << 16);
    for (int kk = 0; kk < 4; ++kk) {
      const int t = kk + (kk >= sid ? 1 : 0);         // BFS order over K5: ascending target id (:103)
      const int m = c_mid_index[sid][t];
      const double prx = (double)sx + (double)__ldg(mid + (size_t)(2 * m) * hw + sy * W + sx);       // (:110-112)
      const double pry = (double)sy + (double)__ldg(mid + (size_t)(2 * m + 1) * hw + sy * W + sx);
      double best = DBL_MAX;
      int bestj = INT_MAX;
      for (int j = i + 1 + tid; j < K; j += blockDim.x) {
        if (s_alive[j] && s_id[j] == t) {
          const double ddx = prx - (double)s_px[j], ddy = pry - (double)s_py[j];
          const double d = sqrt(ddx * ddx + ddy * ddy);                                             // np.linalg.norm (:114,117)
          if (d <= 6.0 && d < best) { best = d; bestj = j; }                                        // KP_RADIUS + 1
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
        if (ob < best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
      }
      if (lane == 0) { s_best[warp] = best; s_bestj[warp] = bestj; }
      __syncthreads();
      best = s_best[0]; bestj = s_bestj[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        const double ob = s_best[w];
        const int oj = s_bestj[w];
        if (ob < best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
      }
      if (bestj != INT_MAX) {                                                                       // stable sort by distance -> first minimum
        sk[3 * t] = (double)s_px[bestj]; sk[3 * t + 1] = (double)s_py[bestj]; sk[3 * t + 2] = s_conf[bestj];
        sk_xy[t] = (int)s_px[bestj] | ((int)s_py[bestj] << 16);
        if (tid == 0) s_alive[bestj] = 0;                                                           // keypoints.pop(matches[0][0]) (:120)
      }
      __syncthreads();
    }
    if (tid < 15) {
      double v = 0.;
#pragma unroll
      for (int k = 0; k < 15; ++k) if (k == tid) v = sk[k];
      skel[(size_t)nskel * 15 + tid] = v;
    }
    if (tid >= 32 && tid < 37) {
      int v = 0;
#pragma unroll
      for (int k = 0; k < 5; ++k) if (k == tid - 32) v = sk_xy[k];
      skel_xy[nskel * 5 + (tid - 32)] = v;
    }
    ++nskel;
    __syncthreads();                                  // skel_xy visible to the next seed test
  }
  if (tid == 0) skel_count_g[list] = nskel;
  __syncthreads();
  boxes_from_skeletons(skel, nskel, (double)gp.box_scale[s], true, sbox_g + (size_t)list * P * 5, sbox_count_g + list,
                       skel_keep_g ? skel_keep_g + (size_t)list * P : nullptr);
}

// ------------------------------------------------------------------------------------------------
// K4: per image: gather_skeleton (postprocessing.py:255-261: scale 0..3 concatenated) and
// non_maximum_suppression_numpy (nms.py:4-53).  argsort ties: (conf, index) ascending.
__global__ void __launch_bounds__(256) nms_kernel(const double* __restrict__ sbox_g, const int* __restrict__ sbox_count_g,
                                                  int n_lists, int list_cap, int max_boxes, double nms_thresh,
                                                  double* __restrict__ boxes_g, int* __restrict__ box_count_g,
                                                  double* __restrict__ dets_g, int* __restrict__ det_count_g,
                                                  int* __restrict__ status) {
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
static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

struct DecodeWorkspace {
  unsigned long long* acc[KG_MAX_SCALES];
  double* peak_conf; int* peak_key; int* peak_count;
  double* skel; int* skel_xy; int* skel_count;
  double* sbox; int* sbox_count;
  double* boxes; int* box_count;
  double* dets; int* det_count;
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
  char* z1 = a.take<char>(0);
  w.zero_begin = z0; w.zero_bytes = (size_t)(z1 - z0);
  w.peak_conf = a.take<double>(lists * P);
  w.peak_key = a.take<int>(lists * P);
  w.skel = a.take<double>(lists * P * 15);
  w.skel_xy = a.take<int>(lists * P * 5);
  w.skel_count = a.take<int>(lists);
  w.sbox = a.take<double>(lists * P * 5);
  w.sbox_count = a.take<int>(lists);
  w.boxes = a.take<double>((size_t)cfg.N * cfg.max_boxes * 5);
  w.box_count = a.take<int>(cfg.N);
  w.dets = a.take<double>((size_t)cfg.N * cfg.max_boxes * 5);
  w.det_count = a.take<int>(cfg.N);
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

static size_t group_smem(int P) { return (size_t)P * 18; }
static size_t nms_smem(int B) { return (size_t)B * 13; }

int decode_launch(const kg_decode_config* cfg, const kg_decode_scale* sc, const kg_decode_outputs* out, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream, int* n_launches) {
  KG_TRY(check_config(cfg, sc));
  KG_REQUIRE(out != nullptr && out->d_status != nullptr, "kg_decode: outputs / d_status must be non-null");
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
  double* peak_conf = out->d_peak_conf ? out->d_peak_conf : w.peak_conf;
  int* peak_key = out->d_peak_key ? out->d_peak_key : w.peak_key;
  int* peak_count = w.peak_count;
  for (int s = 0; s < S; ++s) {
    const int H = sc[s].H, W = sc[s].W;
    dim3 g1(ceil_div(H * W, 256), 5, N);
    {
      StageScope t(0, stream);
      vote_kernel<<<g1, 256, 0, stream>>>(sc[s].d_kp, sc[s].d_short, w.acc[s], H, W);
    }
    dim3 g2(ceil_div(W, BT), ceil_div(H, BT), N * 5);
    {
      StageScope t(1, stream);
      blur_peak_kernel<<<g2, 256, 0, stream>>>(w.acc[s], H, W, cfg->peak_thresh, s, S, P, peak_conf, peak_key, peak_count,
                                               out->d_vote[s], out->d_heat[s], out->d_status);
    }
    launches += 2;
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
                                                             out->d_det_count ? out->d_det_count : w.det_count, out->d_status);
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
    KG_CUDA_CHECK(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem(8192)));
    nms_kernel<<<1, 256, nms_smem(cap), st.s>>>(in.as<double>(), d_ints, 1, cap, cap, nms_thresh, boxes.as<double>(), d_ints + 1,
                                                dets.as<double>(), d_ints + 2, d_ints + 3);
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
