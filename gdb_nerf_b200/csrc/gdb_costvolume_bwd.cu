// Backward of K1 (homography warp + variance) and K2 (depth regression ->
// confidence interval) for the training configuration (SURVEY.md 8b "Autograd").
// Reference forward: depth_net.py:399-514; the adjoints below are what autograd
// derives from it (bilinear taps with zero padding, clamps with zero derivative
// where they bite, population variance over views).
#include <algorithm>

#include "gdb_common.cuh"

namespace gdb {

__device__ __forceinline__ float hyp_t(int d, int D) { return linspace01(d, D); }

// ---------------------------------------------------------------------------
// K1 backward.  Thread = (b, d, target pixel); views and channels are looped.
//   var_c = 1/V sum_v (w_vc - mean_c)^2   ->   dL/dw_vc = 2/V (w_vc - mean_c) g_c
//   w_vc = bilinear(feat_v, ix, iy) (zero padding) -> scatter to d_feat, dL/d(ix,iy)
//   ix = X/Z - 0.5, X = rx*depth + P3, Z = max(rz*depth + P11, 1e-6)  -> dL/d(depth)
//   depth = hypothesis(near, far, d) (or its reciprocal)              -> d_range (optional)
// g_var is planar (B,C,D,Ht,Wt); feat / d_feat are channels-last (B,V,Hs,Ws,C).
// ---------------------------------------------------------------------------
template <int C>
__global__ void warp_variance_bwd_kernel(const float* __restrict__ feat, const float* __restrict__ proj,
                                         const float* __restrict__ range, int rh, int rw, int B, int V, int Hs, int Ws, int D,
                                         int Ht, int Wt, int inv_depth, const float* __restrict__ g_var,
                                         float* __restrict__ d_feat, float* __restrict__ d_range) {
  const int HW = Ht * Wt;
  const int64_t total = (int64_t)B * D * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int d = (int)((i / HW) % D);
    const int b = (int)(i / ((int64_t)HW * D));
    const int px = pix % Wt, py = pix / Wt;
    const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    const int ryi = rh == 1 ? 0 : py, rxi = rw == 1 ? 0 : px;
    const float near_ = range[((size_t)(b * 2 + 0) * rh + ryi) * rw + rxi];
    const float far_ = range[((size_t)(b * 2 + 1) * rh + ryi) * rw + rxi];
    float n2 = near_, f2 = far_;
    if (inv_depth) { n2 = fdiv(1.f, near_); f2 = fdiv(1.f, far_); }
    const float tt = hyp_t(d, D);
    const float dv = fadd(n2, fmul(fsub(f2, n2), tt));
    const float depth = inv_depth ? fdiv(1.f, dv) : dv;

    float g[C];
    bool any = false;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      g[c] = g_var[(((size_t)b * C + c) * D + d) * HW + pix];
      any |= g[c] != 0.f;
    }
    if (!any) continue;

    // pass 1: mean over views of the warped features
    float mean[C];
#pragma unroll
    for (int c = 0; c < C; ++c) mean[c] = 0.f;
    for (int v = 0; v < V; ++v) {
      const float* P = proj + ((size_t)b * V + v) * 12;
      float rx = fmaf(P[0], fx, fmaf(P[1], fy, P[2])), ry = fmaf(P[4], fx, fmaf(P[5], fy, P[6])), rz = fmaf(P[8], fx, fmaf(P[9], fy, P[10]));
      float X = fmaf(rx, depth, P[3]), Y = fmaf(ry, depth, P[7]), Z = fmaxf(fmaf(rz, depth, P[11]), 1e-6f);
      float ix = X / Z - 0.5f, iy = Y / Z - 0.5f;
      float x0f = floorf(ix), y0f = floorf(iy);
      float tx = ix - x0f, ty = iy - y0f;
      x0f = fminf(fmaxf(x0f, -2.f), (float)Ws + 1.f);
      y0f = fminf(fmaxf(y0f, -2.f), (float)Hs + 1.f);
      int x0 = (int)x0f, y0 = (int)y0f;
      const float* vb = feat + ((size_t)b * V + v) * Hs * Ws * C;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        if ((unsigned)xx >= (unsigned)Ws || (unsigned)yy >= (unsigned)Hs) continue;
        float w = ((t & 1) ? tx : 1.f - tx) * ((t >> 1) ? ty : 1.f - ty);
        const float* tp = vb + ((size_t)yy * Ws + xx) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) mean[c] = fmaf(__ldg(tp + c), w, mean[c]);
      }
    }
    const float invV = 1.f / (float)V;
#pragma unroll
    for (int c = 0; c < C; ++c) mean[c] *= invV;

    // pass 2: per view adjoint, scatter, coordinate gradient
    float g_depth = 0.f;
    for (int v = 0; v < V; ++v) {
      const float* P = proj + ((size_t)b * V + v) * 12;
      float rx = fmaf(P[0], fx, fmaf(P[1], fy, P[2])), ry = fmaf(P[4], fx, fmaf(P[5], fy, P[6])), rz = fmaf(P[8], fx, fmaf(P[9], fy, P[10]));
      float X = fmaf(rx, depth, P[3]), Y = fmaf(ry, depth, P[7]);
      float Zr = fmaf(rz, depth, P[11]);
      float Z = fmaxf(Zr, 1e-6f);
      float ix = X / Z - 0.5f, iy = Y / Z - 0.5f;
      float x0f = floorf(ix), y0f = floorf(iy);
      float tx = ix - x0f, ty = iy - y0f;
      x0f = fminf(fmaxf(x0f, -2.f), (float)Ws + 1.f);
      y0f = fminf(fmaxf(y0f, -2.f), (float)Hs + 1.f);
      int x0 = (int)x0f, y0 = (int)y0f;
      const float* vb = feat + ((size_t)b * V + v) * Hs * Ws * C;
      float* db = d_feat + ((size_t)b * V + v) * Hs * Ws * C;
      float val[C];
#pragma unroll
      for (int c = 0; c < C; ++c) val[c] = 0.f;
      float tap[4][C];
      bool ok[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        ok[t] = (unsigned)xx < (unsigned)Ws && (unsigned)yy < (unsigned)Hs;
        float w = ((t & 1) ? tx : 1.f - tx) * ((t >> 1) ? ty : 1.f - ty);
        const float* tp = vb + ((ptrdiff_t)yy * Ws + xx) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          tap[t][c] = ok[t] ? __ldg(tp + c) : 0.f;
          val[c] = fmaf(tap[t][c], w, val[c]);
        }
      }
      float gix = 0.f, giy = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float gw = 2.f * invV * (val[c] - mean[c]) * g[c];         // dL/dw_vc (the mean's own dependence cancels: sum_v (w - mean) = 0)
        val[c] = gw;
        gix = fmaf(gw, (tap[1][c] - tap[0][c]) * (1.f - ty) + (tap[3][c] - tap[2][c]) * ty, gix);
        giy = fmaf(gw, (tap[2][c] - tap[0][c]) * (1.f - tx) + (tap[3][c] - tap[1][c]) * tx, giy);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (!ok[t]) continue;
        int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        float w = ((t & 1) ? tx : 1.f - tx) * ((t >> 1) ? ty : 1.f - ty);
        float* tp = db + ((size_t)yy * Ws + xx) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) atomicAdd(tp + c, w * val[c]);
      }
      // ix = X/Z - 0.5: d/d(depth) = (rx - (X/Z) rz') / Z, rz' = 0 where Z is clamped
      float rzz = Zr >= 1e-6f ? rz : 0.f;
      g_depth += gix * (rx - (X / Z) * rzz) / Z + giy * (ry - (Y / Z) * rzz) / Z;
    }
    if (d_range && rh != 1) {
      float g_dv = inv_depth ? -g_depth * depth * depth : g_depth;       // depth = 1 / dv
      float gn = g_dv * (1.f - tt), gf = g_dv * tt;
      if (inv_depth) { gn = -gn * n2 * n2; gf = -gf * f2 * f2; }         // n2 = 1 / near
      atomicAdd(d_range + ((size_t)(b * 2 + 0) * rh + ryi) * rw + rxi, gn);
      atomicAdd(d_range + ((size_t)(b * 2 + 1) * rh + ryi) * rw + rxi, gf);
    }
  }
}

// ---------------------------------------------------------------------------
// K2 backward (depth_net.py:479-514).  Thread = pixel.
// ---------------------------------------------------------------------------
__global__ void depth_range_bwd_kernel(const float* __restrict__ range, int rh, int rw, const float* __restrict__ prob, int B,
                                       int D, int h, int w, float ci_scale, int inv_depth, const float* __restrict__ g_depth,
                                       const float* __restrict__ g_ci, const float* __restrict__ g_vol,
                                       float* __restrict__ d_prob, float* __restrict__ d_range) {
  int hw = h * w;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * hw) return;
  int b = i / hw, p = i % hw;
  int y = p / w, x = p % w;
  int ry = rh == 1 ? 0 : y, rx = rw == 1 ? 0 : x;
  float near_ = range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx];
  float far_ = range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx];
  float n2 = near_, f2 = far_;
  if (inv_depth) { n2 = fdiv(1.f, near_); f2 = fdiv(1.f, far_); }
  auto hyp = [&](int d) { return fadd(n2, fmul(fsub(f2, n2), hyp_t(d, D))); };
  const float* pp = prob + (size_t)b * D * hw + p;
  float mean = 0.f;
  for (int d = 0; d < D; ++d) mean = fadd(mean, fmul(pp[(size_t)d * hw], hyp(d)));
  float var = 0.f, s1 = 0.f;
  for (int d = 0; d < D; ++d) {
    float t = fsub(hyp(d), mean);
    var = fadd(var, fmul(pp[(size_t)d * hw], fmul(t, t)));
    s1 += pp[(size_t)d * hw] * t;
  }
  float stdv = sqrtf(fmaxf(var, 1e-12f));
  float half = ci_scale * stdv;
  float first = hyp(0), last = hyp(D - 1);
  float gd = g_depth ? g_depth[i] : 0.f;
  float glo = g_ci ? g_ci[(size_t)(b * 2 + 0) * hw + p] : 0.f, ghi = g_ci ? g_ci[(size_t)(b * 2 + 1) * hw + p] : 0.f;
  float g_first = g_vol ? g_vol[(size_t)(b * 2 + 0) * hw + p] : 0.f, g_last = g_vol ? g_vol[(size_t)(b * 2 + 1) * hw + p] : 0.f;
  float g_mean = 0.f, g_half = 0.f;
  if (inv_depth) {
    // lo = 1 / min(mean + half, first); hi = 1 / max(mean - half, last); depth = 1 / mean
    float a = mean + half, c = mean - half;
    float lo_in = fminf(a, first), hi_in = fmaxf(c, last);
    float g_lo_in = -glo / (lo_in * lo_in), g_hi_in = -ghi / (hi_in * hi_in);
    if (a <= first) { g_mean += g_lo_in; g_half += g_lo_in; } else g_first += g_lo_in;
    if (c >= last) { g_mean += g_hi_in; g_half -= g_hi_in; } else g_last += g_hi_in;
    g_mean += -gd / (mean * mean);
  } else {
    float a = mean - half, c = mean + half;
    if (a >= first) { g_mean += glo; g_half -= glo; } else g_first += glo;
    if (c <= last) { g_mean += ghi; g_half += ghi; } else g_last += ghi;
    g_mean += gd;
  }
  float g_var = var > 1e-12f ? g_half * ci_scale / (2.f * stdv) : 0.f;
  g_mean += g_var * (-2.f * s1);
  float gn = 0.f, gf = 0.f;
  for (int d = 0; d < D; ++d) {
    float hd = hyp(d), t = hd - mean, pd = pp[(size_t)d * hw];
    d_prob[((size_t)b * D + d) * hw + p] = g_mean * hd + g_var * t * t;
    float gh = g_mean * pd + g_var * 2.f * pd * t;
    if (d == 0) gh += g_first;
    if (d == D - 1) gh += g_last;
    float tt = hyp_t(d, D);
    gn += gh * (1.f - tt);
    gf += gh * tt;
  }
  if (d_range && rh != 1) {
    if (inv_depth) { gn = -gn * n2 * n2; gf = -gf * f2 * f2; }
    d_range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx] = gn;
    d_range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx] = gf;
  }
}

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_warp_variance_bwd(const float* feat_cl, const float* proj, const float* depth_range, int rh, int rw, int B,
                                     int V, int C, int Hs, int Ws, int D, int Ht, int Wt, int inv_depth, const float* g_variance,
                                     float* d_feat_cl, float* d_depth_range, void* stream) {
  GDB_REQUIRE(feat_cl && proj && depth_range && g_variance && d_feat_cl, GDB_E_BADARG, "gdb_warp_variance_bwd: null pointer");
  GDB_REQUIRE(B > 0 && V >= 2 && V <= GDB_MAX_VIEWS && D > 0 && Ht > 0 && Wt > 0 && Hs > 0 && Ws > 0, GDB_E_BADARG,
              "gdb_warp_variance_bwd: bad size");
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == Ht && rw == Wt), GDB_E_BADARG, "gdb_warp_variance_bwd: depth_range must be 1x1 or %dx%d", Ht, Wt);
  int64_t total = (int64_t)B * D * Ht * Wt;
  int blocks = (int)std::min<int64_t>((total + 127) / 128, (int64_t)sm_count() * 32);
  cudaStream_t st = as_stream(stream);
  switch (C) {
    case 8: warp_variance_bwd_kernel<8><<<blocks, 128, 0, st>>>(feat_cl, proj, depth_range, rh, rw, B, V, Hs, Ws, D, Ht, Wt, inv_depth, g_variance, d_feat_cl, d_depth_range); break;
    case 16: warp_variance_bwd_kernel<16><<<blocks, 128, 0, st>>>(feat_cl, proj, depth_range, rh, rw, B, V, Hs, Ws, D, Ht, Wt, inv_depth, g_variance, d_feat_cl, d_depth_range); break;
    case 32: warp_variance_bwd_kernel<32><<<blocks, 128, 0, st>>>(feat_cl, proj, depth_range, rh, rw, B, V, Hs, Ws, D, Ht, Wt, inv_depth, g_variance, d_feat_cl, d_depth_range); break;
    default: return fail(GDB_E_UNSUPPORTED, "gdb_warp_variance_bwd: C=%d not in {8,16,32}", C);
  }
  return cuda_check("gdb_warp_variance_bwd");
}

extern "C" int gdb_depth_range_bwd(const float* depth_range, int rh, int rw, const float* prob, int B, int D, int h, int w,
                                   float ci_scale, int inv_depth, const float* g_depth, const float* g_ci, const float* g_vol_range,
                                   float* d_prob, float* d_depth_range, void* stream) {
  GDB_REQUIRE(depth_range && prob && d_prob && B > 0 && D > 0 && h > 0 && w > 0, GDB_E_BADARG, "gdb_depth_range_bwd: bad argument");
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == h && rw == w), GDB_E_BADARG, "gdb_depth_range_bwd: depth_range must be 1x1 or %dx%d", h, w);
  int n = B * h * w;
  depth_range_bwd_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(depth_range, rh, rw, prob, B, D, h, w, ci_scale, inv_depth,
                                                                        g_depth, g_ci, g_vol_range, d_prob, d_depth_range);
  return cuda_check("gdb_depth_range_bwd");
}
