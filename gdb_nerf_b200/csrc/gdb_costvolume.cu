// Cost-volume side of the path: homography matrices, depth hypotheses,
// fused homography-warp + variance (K1) and depth regression (K2).
// Reference: networks/gdb_nerf/depth_net.py:399-514.
#include <stdlib.h>
#include "gdb_common.cuh"

namespace gdb {

// ---------------------------------------------------------------------------
// small dense inverses in double (device): plumbing for camera matrices
// ---------------------------------------------------------------------------
__device__ inline void inv4x4(const double* a, double* out) {
  double m[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      m[i][j] = a[i * 4 + j];
      m[i][j + 4] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    double best = fabs(m[c][c]);
    for (int r = c + 1; r < 4; ++r)
      if (fabs(m[r][c]) > best) { best = fabs(m[r][c]); piv = r; }
    if (piv != c)
      for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
    double inv = 1.0 / m[c][c];
    for (int j = 0; j < 8; ++j) m[c][j] *= inv;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        double f = m[r][c];
        for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
      }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][j + 4];
}

// ---------------------------------------------------------------------------
// homography matrices  (depth_net.py:453-457 with the scaling of :159-162)
// ---------------------------------------------------------------------------
__global__ void homography_kernel(const float* __restrict__ src_exts, const float* __restrict__ src_ints,
                                  const float* __restrict__ tar_exts, const float* __restrict__ tar_ints,
                                  float src_scale, float tar_scale, int B, int V, float* __restrict__ proj) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * V) return;
  int b = i / V;
  double Ks[9], Kt[9];
  for (int k = 0; k < 9; ++k) {
    // the reference scales rows 0-1 in float32 (in-place multiply) before anything else
    float s = src_ints[i * 9 + k], t = tar_ints[b * 9 + k];
    Ks[k] = (k < 6) ? (double)fmul(s, src_scale) : (double)s;
    Kt[k] = (k < 6) ? (double)fmul(t, tar_scale) : (double)t;
  }
  const float* Es = src_exts + (size_t)i * 16;
  const float* Et = tar_exts + (size_t)b * 16;
  double Ps[12], Pt[16], Pti[16];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double a = 0.0, t = 0.0;
      for (int k = 0; k < 3; ++k) {
        a += Ks[r * 3 + k] * (double)Es[k * 4 + c];
        t += Kt[r * 3 + k] * (double)Et[k * 4 + c];
      }
      Ps[r * 4 + c] = a;
      Pt[r * 4 + c] = t;
    }
  Pt[12] = Pt[13] = Pt[14] = 0.0;
  Pt[15] = 1.0;
  inv4x4(Pt, Pti);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double a = 0.0;
      for (int k = 0; k < 4; ++k) a += Ps[r * 4 + k] * Pti[k * 4 + c];
      proj[(size_t)i * 12 + r * 4 + c] = (float)a;
    }
}

// ---------------------------------------------------------------------------
// depth hypotheses (depth_net.py:399-421)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float hypothesis(float near_, float far_, int d, int D, int inv_depth) {
  if (inv_depth) {
    near_ = fdiv(1.f, near_);
    far_ = fdiv(1.f, far_);
  }
  return fadd(near_, fmul(fsub(far_, near_), linspace01(d, D)));
}

__global__ void depth_values_kernel(const float* __restrict__ range, int rh, int rw, int B, int D, int Ht, int Wt,
                                    int inv_depth, float* __restrict__ out) {
  size_t n = (size_t)B * D * Ht * Wt;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int x = i % Wt;
    int y = (i / Wt) % Ht;
    int d = (i / ((size_t)Wt * Ht)) % D;
    int b = i / ((size_t)Wt * Ht * D);
    int ry = rh == 1 ? 0 : y, rx = rw == 1 ? 0 : x;
    float near_ = range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx];
    float far_ = range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx];
    out[i] = hypothesis(near_, far_, d, D, inv_depth);
  }
}

// ---------------------------------------------------------------------------
// K1: homography warp + variance over views (depth_net.py:459-474)
//
// Thread = (target pixel, 4-channel slice).  C/4 adjacent lanes read one source
// texel (C*4 bytes, channels-last) as consecutive float4s, so every warp-level
// LDG.128 covers whole 128-byte lines.  The projection of a pixel into a view
// is computed by ONE lane of the pixel's lane group (lane q handles view q) and
// broadcast with shuffles: the coordinate arithmetic is issued once per warp
// instead of once per view.  A CTA owns PIX consecutive target pixels and DCH
// depth planes.  Output either channels-last (B,D,Ht,Wt,C) - one coalesced
// STG.128 per thread, what cuDNN's NDHWC kernels consume directly - or planar
// NCDHW (the reference's layout) through a shared-memory transpose.
// ---------------------------------------------------------------------------
struct WarpTap {
  int xx, yy;                 // clamped tap columns x0 | x1 << 16 and rows y0 | y1 << 16 (always inside the map)
  float wx0, wx1, wy0, wy1;   // bilinear weights with the zero padding folded in (0 for a tap outside the map)
};

// grid_sample(bilinear, zeros, align_corners=False) of depth_net.py:469-472 for one (pixel, view, depth): taps outside the
// source map contribute zero, so their weight is zeroed and their address clamped - the loads need no predicate.
__device__ __forceinline__ WarpTap warp_tap(const float* __restrict__ P, float rx, float ry, float rz, float depth, int Ws, int Hs) {
  float X = fmaf(rx, depth, P[3]);
  float Y = fmaf(ry, depth, P[7]);
  float Z = fmaxf(fmaf(rz, depth, P[11]), 1e-6f);
  float iz = 1.f / Z;
  // pixel-space coordinate u = X/Z maps to texel space u - 0.5 (normalise + grid_sample un-normalise cancel)
  float ix = fmaf(X, iz, -0.5f), iy = fmaf(Y, iz, -0.5f);
  float x0f = floorf(ix), y0f = floorf(iy);
  const float tx = ix - x0f, ty = iy - y0f;
  x0f = fminf(fmaxf(x0f, -2.f), (float)Ws + 1.f);    // also maps NaN to -2 (fully outside)
  y0f = fminf(fmaxf(y0f, -2.f), (float)Hs + 1.f);
  const int x0 = (int)x0f, y0 = (int)y0f;
  WarpTap t;
  t.wx0 = (unsigned)x0 < (unsigned)Ws ? 1.f - tx : 0.f;
  t.wx1 = (unsigned)(x0 + 1) < (unsigned)Ws ? tx : 0.f;
  t.wy0 = (unsigned)y0 < (unsigned)Hs ? 1.f - ty : 0.f;
  t.wy1 = (unsigned)(y0 + 1) < (unsigned)Hs ? ty : 0.f;
  t.xx = min(max(x0, 0), Ws - 1) | (min(max(x0 + 1, 0), Ws - 1) << 16);
  t.yy = min(max(y0, 0), Hs - 1) | (min(max(y0 + 1, 0), Hs - 1) << 16);
  return t;
}

template <int C, int V, int OUT_CL>      // 0: (B,C,D,Ht,Wt) planar; 1: (B,D,Ht,Wt,C); 2: (B,Ht,Wt,D,C) depth folded into the channels
__global__ void __launch_bounds__(256)
warp_variance_kernel(const float* __restrict__ feat, const float* __restrict__ proj, const float* __restrict__ range,
                     int rh, int rw, int Hs, int Ws, int D, int Ht, int Wt, int DCH, int inv_depth,
                     float* __restrict__ out) {
  constexpr int LPP = C / 4;          // lanes per pixel
  constexpr int PIX = 256 / LPP;      // pixels per CTA
  constexpr int PAD = (C == 32) ? 1 : (C == 16 ? 2 : 4);
  constexpr int ROW = PIX + PAD;
  constexpr bool SHARE = LPP >= V;    // one lane per (pixel, view) does the projection
  __shared__ float tile[OUT_CL ? 1 : 2][OUT_CL ? 1 : C * ROW];
  __shared__ float sproj[V * 12];

  const int b = blockIdx.z;
  const int HW = Ht * Wt;
  const int q = threadIdx.x % LPP;
  const int pl = threadIdx.x / LPP;
  const int pix = blockIdx.x * PIX + pl;
  const bool live = pix < HW;
  const int px = live ? pix % Wt : 0, py = live ? pix / Wt : 0;
  const int lane = threadIdx.x & 31;
  const int group_base = lane - q;

  if (threadIdx.x < V * 12) sproj[threadIdx.x] = proj[(size_t)b * V * 12 + threadIdx.x];
  __syncthreads();

  // rotation part applied to the pixel centre (for the view(s) this lane projects)
  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
  float rx[SHARE ? 1 : V], ry[SHARE ? 1 : V], rz[SHARE ? 1 : V];
#pragma unroll
  for (int k = 0; k < (SHARE ? 1 : V); ++k) {
    const float* P = sproj + (SHARE ? min(q, V - 1) : k) * 12;
    rx[k] = fmaf(P[0], fx, fmaf(P[1], fy, P[2]));
    ry[k] = fmaf(P[4], fx, fmaf(P[5], fy, P[6]));
    rz[k] = fmaf(P[8], fx, fmaf(P[9], fy, P[10]));
  }
  const int ryi = rh == 1 ? 0 : py, rxi = rw == 1 ? 0 : px;
  const float near_ = range[((size_t)(b * 2 + 0) * rh + ryi) * rw + rxi];
  const float far_ = range[((size_t)(b * 2 + 1) * rh + ryi) * rw + rxi];
  const size_t view_stride = (size_t)Hs * Ws * C;
  const float* fbase = feat + (size_t)b * V * view_stride + q * 4;

  const int d0 = blockIdx.y * DCH;
  const int d1 = min(d0 + DCH, D);
  int buf = 0;
  for (int d = d0; d < d1; ++d, buf ^= 1) {
    float dv = hypothesis(near_, far_, d, D, inv_depth);
    float depth = inv_depth ? fdiv(1.f, dv) : dv;
    WarpTap mine;
    if (SHARE) mine = warp_tap(sproj + min(q, V - 1) * 12, rx[0], ry[0], rz[0], depth, Ws, Hs);
    float4 val[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      WarpTap t;
      if (SHARE) {
        t.xx = __shfl_sync(0xffffffffu, mine.xx, group_base + v);
        t.yy = __shfl_sync(0xffffffffu, mine.yy, group_base + v);
        t.wx0 = __shfl_sync(0xffffffffu, mine.wx0, group_base + v);
        t.wx1 = __shfl_sync(0xffffffffu, mine.wx1, group_base + v);
        t.wy0 = __shfl_sync(0xffffffffu, mine.wy0, group_base + v);
        t.wy1 = __shfl_sync(0xffffffffu, mine.wy1, group_base + v);
      } else {
        t = warp_tap(sproj + v * 12, rx[SHARE ? 0 : v], ry[SHARE ? 0 : v], rz[SHARE ? 0 : v], depth, Ws, Hs);
      }
      const int r0 = (t.yy & 0xffff) * Ws, r1 = (t.yy >> 16) * Ws;
      const int c0 = t.xx & 0xffff, c1 = t.xx >> 16;
      const float4* vb = reinterpret_cast<const float4*>(fbase + v * view_stride);
      constexpr int CQ = C / 4;
      // dead pixels of the last tile read pixel (0, 0)'s taps: in bounds, never stored
      const float4 t00 = __ldg(vb + (r0 + c0) * CQ), t10 = __ldg(vb + (r0 + c1) * CQ);
      const float4 t01 = __ldg(vb + (r1 + c0) * CQ), t11 = __ldg(vb + (r1 + c1) * CQ);
      float4 acc;
      {
        const float w = t.wx0 * t.wy0;
        acc = make_float4(t00.x * w, t00.y * w, t00.z * w, t00.w * w);
      }
      acc = f4_scale_add(acc, t10, t.wx1 * t.wy0);
      acc = f4_scale_add(acc, t01, t.wx0 * t.wy1);
      acc = f4_scale_add(acc, t11, t.wx1 * t.wy1);
      val[v] = acc;
    }
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < V; ++v) { mean.x += val[v].x; mean.y += val[v].y; mean.z += val[v].z; mean.w += val[v].w; }
    const float invV = 1.f / (float)V;
    mean.x *= invV; mean.y *= invV; mean.z *= invV; mean.w *= invV;
    float4 var = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float a = val[v].x - mean.x, bb = val[v].y - mean.y, c = val[v].z - mean.z, e = val[v].w - mean.w;
      var.x = fmaf(a, a, var.x); var.y = fmaf(bb, bb, var.y); var.z = fmaf(c, c, var.z); var.w = fmaf(e, e, var.w);
    }
    var.x *= invV; var.y *= invV; var.z *= invV; var.w *= invV;
    if (OUT_CL) {
      if (live)
        __stcs(reinterpret_cast<float4*>(out + (OUT_CL == 2 ? (((size_t)b * HW + pix) * D + d) : (((size_t)b * D + d) * HW + pix)) * C + q * 4), var);
    } else {
      float* t = tile[buf];
      t[(q * 4 + 0) * ROW + pl] = var.x;
      t[(q * 4 + 1) * ROW + pl] = var.y;
      t[(q * 4 + 2) * ROW + pl] = var.z;
      t[(q * 4 + 3) * ROW + pl] = var.w;
      __syncthreads();
      // coalesced plane writes: C rows of PIX floats
      float* obase = out + (((size_t)b * C) * D + d) * HW + (size_t)blockIdx.x * PIX;
#pragma unroll
      for (int i = 0; i < (C * PIX) / 256; ++i) {
        int idx = i * 256 + threadIdx.x;
        int c = idx / PIX, p = idx % PIX;
        if (blockIdx.x * PIX + p < HW) __stcs(obase + (size_t)c * D * HW + p, t[c * ROW + p]);
      }
      // the other buffer is written next; its readers finished before the barrier above
    }
  }
}

// ---------------------------------------------------------------------------
// K1, second generation (channels-last outputs): thread = (target pixel, 8-channel slice).  ncu showed the 4-channel
// kernel issue-bound (360 warp instructions per depth plane for 12 taps, 80 % issue utilisation, L1 at 53-67 %): the
// per-tap work that does not depend on the channel (projection, clamps, weights, addresses) was repeated by C/4 lanes.
// Here C/8 lanes share a pixel, every tap is ONE 256-bit load (LDG.E.ENL2.256: the lanes of a pixel still cover whole
// 128-byte lines, so the L1 wavefront count per tap is unchanged), the bilinear blend and the variance run as packed
// FFMA2 / FADD2 / FMUL2 on the four 64-bit pairs the load delivers, and the result leaves as one 256-bit store.  The
// (pixel, view) projections of a lane group are spread over its lanes (lane q projects views q, q + LPP, ..) and
// broadcast by shuffles.  Same operation order per channel as the first kernel: bit-identical results.
// ---------------------------------------------------------------------------
typedef unsigned long long u64;
struct U4 { u64 a, b, c, d; };
__device__ __forceinline__ U4 ldg256(const float* p) {
  U4 r;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
  return r;
}
__device__ __forceinline__ void stcs256(float* p, const U4& r) {
  asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(r.a), "l"(r.b), "l"(r.c), "l"(r.d) : "memory");
}
__device__ __forceinline__ u64 shfl_down64(u64 v, int delta) {
  const unsigned lo = __shfl_down_sync(0xffffffffu, (unsigned)v, delta), hi = __shfl_down_sync(0xffffffffu, (unsigned)(v >> 32), delta);
  return (u64)lo | ((u64)hi << 32);
}
__device__ __forceinline__ U4 shfl_down256(const U4& v, int delta) {
  return U4{shfl_down64(v.a, delta), shfl_down64(v.b, delta), shfl_down64(v.c, delta), shfl_down64(v.d, delta)};
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ U4 u4_mul(const U4& v, u64 w) { return U4{mul2(v.a, w), mul2(v.b, w), mul2(v.c, w), mul2(v.d, w)}; }
__device__ __forceinline__ U4 u4_fma(const U4& v, u64 w, const U4& acc) { return U4{fma2(v.a, w, acc.a), fma2(v.b, w, acc.b), fma2(v.c, w, acc.c), fma2(v.d, w, acc.d)}; }
__device__ __forceinline__ U4 u4_add(const U4& x, const U4& y) { return U4{add2(x.a, y.a), add2(x.b, y.b), add2(x.c, y.c), add2(x.d, y.d)}; }
__device__ __forceinline__ U4 u4_sub(const U4& x, const U4& y) { return U4{sub2(x.a, y.a), sub2(x.b, y.b), sub2(x.c, y.c), sub2(x.d, y.d)}; }
__device__ __forceinline__ U4 u4_sqacc(const U4& x, const U4& acc) { return U4{fma2(x.a, x.a, acc.a), fma2(x.b, x.b, acc.b), fma2(x.c, x.c, acc.c), fma2(x.d, x.d, acc.d)}; }

//
// Measured variants, GDB_K1_VARIANT = bit mask (read per call; tools/bench_k1.py):
// bit 1 (NOSHFL): every lane projects its pixel into ALL V views itself instead of receiving the views of the other lanes of its
// pixel by 6 shuffles per view - more issue slots (an issue utilisation of 42-47 % has room), fewer trips through the LSU.
// bit 0 (XS): the x-adjacent pixel of a lane sits LPP lanes further on; where its left
// column is my right column (same clamped rows, neighbour's c0 == my c1 - the rule wherever source and target pitch agree),
// my two right-column taps ARE the neighbour's two left-column taps and arrive by shuffle (8 SHFL per 256-bit tap) instead
// of by load; the other lanes load them as before (predicated).  Same values, same operation order: bit-identical.
template <int C, int V, int OUT_CL, int VAR>      // 1: (B,D,Ht,Wt,C); 2: (B,Ht,Wt,D,C) depth folded into the channels
__global__ void __launch_bounds__(256)
warp_variance8_kernel(const float* __restrict__ feat, const float* __restrict__ proj, const float* __restrict__ range,
                      int rh, int rw, int Hs, int Ws, int D, int Ht, int Wt, int DCH, int inv_depth,
                      float* __restrict__ out) {
  constexpr bool XS = (VAR & 1) != 0, NOSHFL = (VAR & 2) != 0 || C == 8;
  constexpr int LPP = C / 8;                   // lanes per pixel
  constexpr int PIX = 256 / LPP;               // pixels per CTA
  constexpr int VPL = NOSHFL ? V : (V + LPP - 1) / LPP;     // projections per lane
  __shared__ float sproj[V * 12];

  const int b = blockIdx.z;
  const int HW = Ht * Wt;
  const int q = threadIdx.x % LPP;
  const int pix = blockIdx.x * PIX + threadIdx.x / LPP;
  const bool live = pix < HW;
  const int px = live ? pix % Wt : 0, py = live ? pix / Wt : 0;
  const int group_base = (threadIdx.x & 31) - q;

  if (threadIdx.x < V * 12) sproj[threadIdx.x] = proj[(size_t)b * V * 12 + threadIdx.x];
  __syncthreads();

  const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
  float rx[VPL], ry[VPL], rz[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const float* P = sproj + (NOSHFL ? k : min(q + k * LPP, V - 1)) * 12;
    rx[k] = fmaf(P[0], fx, fmaf(P[1], fy, P[2]));
    ry[k] = fmaf(P[4], fx, fmaf(P[5], fy, P[6]));
    rz[k] = fmaf(P[8], fx, fmaf(P[9], fy, P[10]));
  }
  const int ryi = rh == 1 ? 0 : py, rxi = rw == 1 ? 0 : px;
  const float near_ = range[((size_t)(b * 2 + 0) * rh + ryi) * rw + rxi];
  const float far_ = range[((size_t)(b * 2 + 1) * rh + ryi) * rw + rxi];
  const size_t view_stride = (size_t)Hs * Ws * C;
  const float* fbase = feat + (size_t)b * V * view_stride + q * 8;

  const int d0 = blockIdx.y * DCH;
  const int d1 = min(d0 + DCH, D);
  for (int d = d0; d < d1; ++d) {
    const float dv = hypothesis(near_, far_, d, D, inv_depth);
    const float depth = inv_depth ? fdiv(1.f, dv) : dv;
    WarpTap mine[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) mine[k] = warp_tap(sproj + (NOSHFL ? k : min(q + k * LPP, V - 1)) * 12, rx[k], ry[k], rz[k], depth, Ws, Hs);
    U4 val[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int src = group_base + v % LPP;
      const WarpTap m = mine[NOSHFL ? v : v / LPP];
      WarpTap t;
      if (!NOSHFL) {
        t.xx = __shfl_sync(0xffffffffu, m.xx, src);
        t.yy = __shfl_sync(0xffffffffu, m.yy, src);
        t.wx0 = __shfl_sync(0xffffffffu, m.wx0, src);
        t.wx1 = __shfl_sync(0xffffffffu, m.wx1, src);
        t.wy0 = __shfl_sync(0xffffffffu, m.wy0, src);
        t.wy1 = __shfl_sync(0xffffffffu, m.wy1, src);
      } else {
        t = m;
      }
      const int r0 = (t.yy & 0xffff) * Ws, r1 = (t.yy >> 16) * Ws;
      const int c0 = t.xx & 0xffff, c1 = t.xx >> 16;
      const float* vb = fbase + v * view_stride;
      // dead pixels of the last tile read pixel (0, 0)'s taps: in bounds, never stored
      const U4 t00 = ldg256(vb + (size_t)(r0 + c0) * C), t01 = ldg256(vb + (size_t)(r1 + c0) * C);
      U4 t10, t11;
      if constexpr (XS) {
        const int nxx = __shfl_down_sync(0xffffffffu, t.xx, LPP), nyy = __shfl_down_sync(0xffffffffu, t.yy, LPP);
        const bool shared = (int)(threadIdx.x & 31) + LPP < 32 && nyy == t.yy && (nxx & 0xffff) == c1;
        t10 = shfl_down256(t00, LPP);
        t11 = shfl_down256(t01, LPP);
        if (!shared) {
          t10 = ldg256(vb + (size_t)(r0 + c1) * C);
          t11 = ldg256(vb + (size_t)(r1 + c1) * C);
        }
      } else {
        t10 = ldg256(vb + (size_t)(r0 + c1) * C);
        t11 = ldg256(vb + (size_t)(r1 + c1) * C);
      }
      const float w00 = t.wx0 * t.wy0, w10 = t.wx1 * t.wy0, w01 = t.wx0 * t.wy1, w11 = t.wx1 * t.wy1;
      U4 acc = u4_mul(t00, pack2(w00, w00));
      acc = u4_fma(t10, pack2(w10, w10), acc);
      acc = u4_fma(t01, pack2(w01, w01), acc);
      acc = u4_fma(t11, pack2(w11, w11), acc);
      val[v] = acc;
    }
    U4 mean = val[0];                       // 0 + val[0] is exact: same sum order as the first kernel
#pragma unroll
    for (int v = 1; v < V; ++v) mean = u4_add(mean, val[v]);
    const float invV = 1.f / (float)V;
    const u64 iv = pack2(invV, invV);
    mean = u4_mul(mean, iv);
    U4 var{0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int v = 0; v < V; ++v) var = u4_sqacc(u4_sub(val[v], mean), var);
    var = u4_mul(var, iv);
    if (live)
      stcs256(out + (OUT_CL == 2 ? (((size_t)b * HW + pix) * D + d) : (((size_t)b * D + d) * HW + pix)) * C + q * 8, var);
  }
}

// K1, vertical-pair variant (GDB_K1_VARIANT=4, measured): thread = (two vertically adjacent target pixels, 8-channel slice).
// ncu of the kernel above: the L1 data pipe is 88-91 % busy (profiles/r02_ncu_full_k1_variants.json), i.e. only fewer bytes
// through L1 per output can make it faster.  Where source and target pitch agree, the top row of the lower pixel's 2x2
// footprint IS the bottom row of the upper pixel's (same clamped columns, r0 of B == r1 of A): those two taps are taken from
// registers, 6 loads instead of 8 per (pair, view, plane).  The pairing is vertical so that the lanes of a warp still walk
// consecutive x: every load / store instruction covers the same contiguous bytes as in the kernel above.  Same values, same
// operation order per pixel: bit-identical.  Pairs whose footprints do not line up load all 8 taps.
template <int C, int V, int OUT_CL>
__global__ void __launch_bounds__(256, 2)
warp_variance8_vpair_kernel(const float* __restrict__ feat, const float* __restrict__ proj, const float* __restrict__ range,
                            int rh, int rw, int Hs, int Ws, int D, int Ht, int Wt, int DCH, int inv_depth,
                            float* __restrict__ out) {
  constexpr int LPP = C / 8;                   // lanes per pixel pair
  constexpr int PIXP = 256 / LPP;              // pixel pairs per CTA
  constexpr int NK = 2 * V;                    // (pixel, view) projections of a pair: k = s * V + v
  constexpr int VPL = (NK + LPP - 1) / LPP;    // projections per lane
  __shared__ float sproj[V * 12];

  const int b = blockIdx.z;
  const int HW = Ht * Wt;
  const int NP = ((Ht + 1) >> 1) * Wt;         // pairs per target view
  const int q = threadIdx.x % LPP;
  const int pp = blockIdx.x * PIXP + threadIdx.x / LPP;
  const bool liveA = pp < NP;
  const int px = liveA ? pp % Wt : 0, pyp = liveA ? pp / Wt : 0;
  const int pyA = 2 * pyp, pyB = min(2 * pyp + 1, Ht - 1);
  const bool liveB = liveA && 2 * pyp + 1 < Ht;
  const int group_base = (threadIdx.x & 31) - q;

  if (threadIdx.x < V * 12) sproj[threadIdx.x] = proj[(size_t)b * V * 12 + threadIdx.x];
  __syncthreads();

  const float fx = (float)px + 0.5f;
  float rx[VPL], ry[VPL], rz[VPL];
  int po[VPL];
  bool sB[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int k = min(q + j * LPP, NK - 1);
    sB[j] = k >= V;
    po[j] = (k - (sB[j] ? V : 0)) * 12;
    const float* P = sproj + po[j];
    const float fy = (float)(sB[j] ? pyB : pyA) + 0.5f;
    rx[j] = fmaf(P[0], fx, fmaf(P[1], fy, P[2]));
    ry[j] = fmaf(P[4], fx, fmaf(P[5], fy, P[6]));
    rz[j] = fmaf(P[8], fx, fmaf(P[9], fy, P[10]));
  }
  const int rxi = rw == 1 ? 0 : px, ryA = rh == 1 ? 0 : pyA, ryB = rh == 1 ? 0 : pyB;
  const float nearA = range[((size_t)(b * 2 + 0) * rh + ryA) * rw + rxi], farA = range[((size_t)(b * 2 + 1) * rh + ryA) * rw + rxi];
  const float nearB = range[((size_t)(b * 2 + 0) * rh + ryB) * rw + rxi], farB = range[((size_t)(b * 2 + 1) * rh + ryB) * rw + rxi];
  const size_t view_stride = (size_t)Hs * Ws * C;
  const float* fbase = feat + (size_t)b * V * view_stride + q * 8;
  const int pixA = pyA * Wt + px, pixB = pyB * Wt + px;

  const int d0 = blockIdx.y * DCH;
  const int d1 = min(d0 + DCH, D);
  for (int d = d0; d < d1; ++d) {
    const float dvA = hypothesis(nearA, farA, d, D, inv_depth), dvB = hypothesis(nearB, farB, d, D, inv_depth);
    const float depthA = inv_depth ? fdiv(1.f, dvA) : dvA, depthB = inv_depth ? fdiv(1.f, dvB) : dvB;
    WarpTap mine[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) mine[j] = warp_tap(sproj + po[j], rx[j], ry[j], rz[j], sB[j] ? depthB : depthA, Ws, Hs);
    U4 valA[V], valB[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      WarpTap tA, tB;
      {
        constexpr int dummy = 0; (void)dummy;
        const int kA = v, kB = V + v;
        const WarpTap mA = mine[kA / LPP], mB = mine[kB / LPP];
        if (LPP > 1) {
          const int sa = group_base + kA % LPP, sb = group_base + kB % LPP;
          tA.xx = __shfl_sync(0xffffffffu, mA.xx, sa); tA.yy = __shfl_sync(0xffffffffu, mA.yy, sa);
          tA.wx0 = __shfl_sync(0xffffffffu, mA.wx0, sa); tA.wx1 = __shfl_sync(0xffffffffu, mA.wx1, sa);
          tA.wy0 = __shfl_sync(0xffffffffu, mA.wy0, sa); tA.wy1 = __shfl_sync(0xffffffffu, mA.wy1, sa);
          tB.xx = __shfl_sync(0xffffffffu, mB.xx, sb); tB.yy = __shfl_sync(0xffffffffu, mB.yy, sb);
          tB.wx0 = __shfl_sync(0xffffffffu, mB.wx0, sb); tB.wx1 = __shfl_sync(0xffffffffu, mB.wx1, sb);
          tB.wy0 = __shfl_sync(0xffffffffu, mB.wy0, sb); tB.wy1 = __shfl_sync(0xffffffffu, mB.wy1, sb);
        } else {
          tA = mA; tB = mB;
        }
      }
      const float* vb = fbase + v * view_stride;
      const int rA0 = (tA.yy & 0xffff) * Ws, rA1 = (tA.yy >> 16) * Ws, cA0 = tA.xx & 0xffff, cA1 = tA.xx >> 16;
      const int rB0 = (tB.yy & 0xffff) * Ws, rB1 = (tB.yy >> 16) * Ws, cB0 = tB.xx & 0xffff, cB1 = tB.xx >> 16;
      // dead pixels read pixel (0, 0)'s (or the upper pixel's) taps: in bounds, never stored
      const U4 a00 = ldg256(vb + (size_t)(rA0 + cA0) * C), a10 = ldg256(vb + (size_t)(rA0 + cA1) * C);
      const U4 a01 = ldg256(vb + (size_t)(rA1 + cA0) * C), a11 = ldg256(vb + (size_t)(rA1 + cA1) * C);
      const U4 b01 = ldg256(vb + (size_t)(rB1 + cB0) * C), b11 = ldg256(vb + (size_t)(rB1 + cB1) * C);
      const bool shared = rB0 == rA1 && tB.xx == tA.xx;
      {
        const float w00 = tA.wx0 * tA.wy0, w10 = tA.wx1 * tA.wy0, w01 = tA.wx0 * tA.wy1, w11 = tA.wx1 * tA.wy1;
        U4 acc = u4_mul(a00, pack2(w00, w00));
        acc = u4_fma(a10, pack2(w10, w10), acc);
        acc = u4_fma(a01, pack2(w01, w01), acc);
        acc = u4_fma(a11, pack2(w11, w11), acc);
        valA[v] = acc;
      }
      {
        const float w00 = tB.wx0 * tB.wy0, w10 = tB.wx1 * tB.wy0, w01 = tB.wx0 * tB.wy1, w11 = tB.wx1 * tB.wy1;
        U4 acc;
        if (shared) {       // the upper pixel's bottom row is this pixel's top row
          acc = u4_mul(a01, pack2(w00, w00));
          acc = u4_fma(a11, pack2(w10, w10), acc);
        } else {
          const U4 b00 = ldg256(vb + (size_t)(rB0 + cB0) * C), b10 = ldg256(vb + (size_t)(rB0 + cB1) * C);
          acc = u4_mul(b00, pack2(w00, w00));
          acc = u4_fma(b10, pack2(w10, w10), acc);
        }
        acc = u4_fma(b01, pack2(w01, w01), acc);
        acc = u4_fma(b11, pack2(w11, w11), acc);
        valB[v] = acc;
      }
    }
    const float invV = 1.f / (float)V;
    const u64 iv = pack2(invV, invV);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const U4* val = s ? valB : valA;
      U4 mean = val[0];
#pragma unroll
      for (int v = 1; v < V; ++v) mean = u4_add(mean, val[v]);
      mean = u4_mul(mean, iv);
      U4 var{0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int v = 0; v < V; ++v) var = u4_sqacc(u4_sub(val[v], mean), var);
      var = u4_mul(var, iv);
      const int pix = s ? pixB : pixA;
      if (s ? liveB : liveA)
        stcs256(out + (OUT_CL == 2 ? (((size_t)b * HW + pix) * D + d) : (((size_t)b * D + d) * HW + pix)) * C + q * 8, var);
    }
  }
}

// GDB_K1_V1=1 keeps the first-generation kernel (A/B measurements, tools/bench_k1.py)
static bool warp_variance_v1() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GDB_K1_V1"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
// GDB_K1_VARIANT = 0..3 selects the measured variants of the second-generation kernel (VAR above), 4 the vertical-pair kernel
// (read per call)
static int warp_variance_variant() {
  const char* e = getenv("GDB_K1_VARIANT");
  return (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 0;
}

template <int C, int V>
static int launch_warp_variance(const float* feat, const float* proj, const float* range, int rh, int rw, int B, int Hs,
                                int Ws, int D, int Ht, int Wt, int inv_depth, int out_cl, float* out, cudaStream_t st) {
  constexpr int PIX = 256 / (C / 4);
  int tiles = (Ht * Wt + PIX - 1) / PIX;
  // depth chunk: enough CTAs for >= 4 waves of 148 SMs x 4 resident CTAs, but keep planes together for L1 reuse
  int DCH = D;
  while (DCH > 2 && (long)tiles * ((D + DCH - 1) / DCH) * B < 4L * 4 * sm_count()) DCH = (DCH + 1) / 2;
  if (out_cl && C % 8 == 0 && (reinterpret_cast<uintptr_t>(feat) & 31u) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0 && !warp_variance_v1()) {
    constexpr int PIX8 = 256 / (C / 8);
    tiles = (Ht * Wt + PIX8 - 1) / PIX8;
    DCH = D;
    while (DCH > 2 && (long)tiles * ((D + DCH - 1) / DCH) * B < 4L * 4 * sm_count()) DCH = (DCH + 1) / 2;
    dim3 grid8(tiles, (D + DCH - 1) / DCH, B);
    const int var = C >= 16 ? warp_variance_variant() : 0;
    if (var == 4) {
      const int tp = (((Ht + 1) >> 1) * Wt + PIX8 - 1) / PIX8;
      int dch = D;
      while (dch > 2 && (long)tp * ((D + dch - 1) / dch) * B < 4L * 2 * sm_count()) dch = (dch + 1) / 2;
      dim3 gridp(tp, (D + dch - 1) / dch, B);
      if (out_cl == 2)
        warp_variance8_vpair_kernel<C, V, 2><<<gridp, 256, 0, st>>>(feat, proj, range, rh, rw, Hs, Ws, D, Ht, Wt, dch, inv_depth, out);
      else
        warp_variance8_vpair_kernel<C, V, 1><<<gridp, 256, 0, st>>>(feat, proj, range, rh, rw, Hs, Ws, D, Ht, Wt, dch, inv_depth, out);
      return cuda_check("gdb_warp_variance_fwd");
    }
#define GDB_WV8(VAR_)                                                                                                              \
  if (var == VAR_) {                                                                                                               \
    if (out_cl == 2)                                                                                                               \
      warp_variance8_kernel<C, V, 2, (C >= 16 ? VAR_ : 0)><<<grid8, 256, 0, st>>>(feat, proj, range, rh, rw, Hs, Ws, D, Ht, Wt, DCH, inv_depth, out); \
    else                                                                                                                           \
      warp_variance8_kernel<C, V, 1, (C >= 16 ? VAR_ : 0)><<<grid8, 256, 0, st>>>(feat, proj, range, rh, rw, Hs, Ws, D, Ht, Wt, DCH, inv_depth, out); \
  }
    GDB_WV8(0) GDB_WV8(1) GDB_WV8(2) GDB_WV8(3)
#undef GDB_WV8
    return cuda_check("gdb_warp_variance_fwd");
  }
  dim3 grid(tiles, (D + DCH - 1) / DCH, B);
  if (out_cl == 2)
    warp_variance_kernel<C, V, 2><<<grid, 256, 0, st>>>(feat, proj, range, rh, rw, Hs, Ws, D, Ht, Wt, DCH, inv_depth, out);
  else if (out_cl)
    warp_variance_kernel<C, V, 1><<<grid, 256, 0, st>>>(feat, proj, range, rh, rw, Hs, Ws, D, Ht, Wt, DCH, inv_depth, out);
  else
    warp_variance_kernel<C, V, 0><<<grid, 256, 0, st>>>(feat, proj, range, rh, rw, Hs, Ws, D, Ht, Wt, DCH, inv_depth, out);
  return cuda_check("gdb_warp_variance_fwd");
}

// ---------------------------------------------------------------------------
// K2: depth regression -> confidence interval (depth_net.py:479-514)
// ---------------------------------------------------------------------------
__global__ void depth_range_kernel(const float* __restrict__ range, int rh, int rw, const float* __restrict__ prob,
                                   int B, int D, int h, int w, float ci_scale, int inv_depth,
                                   float* __restrict__ depth, float* __restrict__ ci, float* __restrict__ vol_range) {
  int hw = h * w;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * hw) return;
  int b = i / hw, p = i % hw;
  int y = p / w, x = p % w;
  int ry = rh == 1 ? 0 : y, rx = rw == 1 ? 0 : x;
  float near_ = range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx];
  float far_ = range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx];
  const float* pp = prob + (size_t)b * D * hw + p;
  float mean = 0.f;
  for (int d = 0; d < D; ++d) mean = fadd(mean, fmul(pp[(size_t)d * hw], hypothesis(near_, far_, d, D, inv_depth)));
  float var = 0.f;
  for (int d = 0; d < D; ++d) {
    float t = fsub(hypothesis(near_, far_, d, D, inv_depth), mean);
    var = fadd(var, fmul(pp[(size_t)d * hw], fmul(t, t)));
  }
  float half = fmul(ci_scale, sqrtf(fmaxf(var, 1e-12f)));
  float first = hypothesis(near_, far_, 0, D, inv_depth), last = hypothesis(near_, far_, D - 1, D, inv_depth);
  float lo, hi, dep;
  if (inv_depth) {
    lo = fdiv(1.f, fminf(fadd(mean, half), first));
    hi = fdiv(1.f, fmaxf(fsub(mean, half), last));
    dep = fdiv(1.f, mean);
  } else {
    lo = fmaxf(fsub(mean, half), first);
    hi = fminf(fadd(mean, half), last);
    dep = mean;
  }
  depth[i] = dep;
  ci[(size_t)(b * 2 + 0) * hw + p] = lo;
  ci[(size_t)(b * 2 + 1) * hw + p] = hi;
  if (vol_range) {
    vol_range[(size_t)(b * 2 + 0) * hw + p] = first;
    vol_range[(size_t)(b * 2 + 1) * hw + p] = last;
  }
}


// K2 fused with the soft-max of the probability head (cost_reg_net.py:62-63,115-116 -> depth_net.py:172):
// logits are read with arbitrary (batch, depth, pixel) strides so they may be one channel of a channels-last
// multi-head convolution output.  prob_d = exp(l_d - max) / sum (ATen's soft-max order), then K2 unchanged.
template <int DMAX>
__global__ void depth_range_logits_kernel(const float* __restrict__ range, int rh, int rw, const float* __restrict__ logits,
                                          int64_t sB, int64_t sD, int64_t sP, int B, int D, int h, int w, float ci_scale,
                                          int inv_depth, float* __restrict__ depth, float* __restrict__ ci,
                                          float* __restrict__ vol_range, float* __restrict__ prob_out) {
  int hw = h * w;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * hw) return;
  int b = i / hw, p = i % hw;
  int y = p / w, x = p % w;
  int ry = rh == 1 ? 0 : y, rx = rw == 1 ? 0 : x;
  float near_ = range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx];
  float far_ = range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx];
  const float* lp = logits + b * sB + p * sP;
  float pr[DMAX];
  float m = -INFINITY;
#pragma unroll
  for (int d = 0; d < DMAX; ++d) {
    pr[d] = d < D ? __ldg(lp + d * sD) : -INFINITY;
    m = fmaxf(m, pr[d]);
  }
  float sum = 0.f;
#pragma unroll
  for (int d = 0; d < DMAX; ++d) {
    pr[d] = d < D ? expf(pr[d] - m) : 0.f;
    sum += pr[d];
  }
#pragma unroll
  for (int d = 0; d < DMAX; ++d) pr[d] = fdiv(pr[d], sum);
  float mean = 0.f;
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) mean = fadd(mean, fmul(pr[d], hypothesis(near_, far_, d, D, inv_depth)));
  float var = 0.f;
#pragma unroll
  for (int d = 0; d < DMAX; ++d)
    if (d < D) {
      float t = fsub(hypothesis(near_, far_, d, D, inv_depth), mean);
      var = fadd(var, fmul(pr[d], fmul(t, t)));
    }
  if (prob_out) {
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
      if (d < D) prob_out[((size_t)b * D + d) * hw + p] = pr[d];
  }
  float half = fmul(ci_scale, sqrtf(fmaxf(var, 1e-12f)));
  float first = hypothesis(near_, far_, 0, D, inv_depth), last = hypothesis(near_, far_, D - 1, D, inv_depth);
  float lo, hi, dep;
  if (inv_depth) {
    lo = fdiv(1.f, fminf(fadd(mean, half), first));
    hi = fdiv(1.f, fmaxf(fsub(mean, half), last));
    dep = fdiv(1.f, mean);
  } else {
    lo = fmaxf(fsub(mean, half), first);
    hi = fminf(fadd(mean, half), last);
    dep = mean;
  }
  depth[i] = dep;
  ci[(size_t)(b * 2 + 0) * hw + p] = lo;
  ci[(size_t)(b * 2 + 1) * hw + p] = hi;
  if (vol_range) {
    vol_range[(size_t)(b * 2 + 0) * hw + p] = first;
    vol_range[(size_t)(b * 2 + 1) * hw + p] = last;
  }
}

// ---------------------------------------------------------------------------
// K2 fused with the probability head itself: logits = Conv3d(C -> 1, 3x3x3, padding 1, no bias) of the U-Net's last
// feature volume (cost_reg_net.py:62-63), then the soft-max over depth and the depth regression as above.
// y (B,D,H,W,8) channels-last.  A CTA owns a 4 x 32 pixel tile and walks the depth axis with a rolling window of three
// haloed planes in shared memory (each voxel is read from global once per CTA); the D logits of a pixel are parked in
// shared memory and reduced by the pixel's own thread with the K2 arithmetic.  Exact fp32 (the cuDNN head this
// replaces ran a 1-output-channel convolution 20x off the memory roofline).
// ---------------------------------------------------------------------------
constexpr int PH_TY = 4, PH_TX = 16, PH_THREADS = PH_TY * PH_TX, PH_PW = PH_TX + 2, PH_PH = PH_TY + 2, PH_PLANE = 2 * PH_PH * PH_PW;   // float4 per plane

__global__ void __launch_bounds__(PH_THREADS) prob_head_depth_range_kernel(const float* __restrict__ y, const float* __restrict__ wgt,
                                                                    const float* __restrict__ range, int rh, int rw, int B, int D,
                                                                    int H, int W, float ci_scale, int inv_depth,
                                                                    float* __restrict__ depth, float* __restrict__ ci,
                                                                    float* __restrict__ vol_range, float* __restrict__ prob_out) {
  extern __shared__ __align__(16) unsigned char ph_smem[];
  float4* plane = reinterpret_cast<float4*>(ph_smem);                  // [3][2 halves][PH_PH][PH_PW]
  float4* wsm = plane + 3 * PH_PLANE;                                  // [27][2]
  float* lsm = reinterpret_cast<float*>(wsm + 54);                     // [D][PH_THREADS]
  const int tid = threadIdx.x, ty = tid / PH_TX, tx = tid % PH_TX;
  const int b = blockIdx.z, y0 = blockIdx.y * PH_TY, x0 = blockIdx.x * PH_TX;
  const int HW = H * W;
  if (tid < 54) wsm[tid] = __ldg(reinterpret_cast<const float4*>(wgt) + tid);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_plane = [&](int d, int slot) {
    float4* dst = plane + slot * PH_PLANE;
    for (int i = tid; i < PH_PH * PH_PW; i += PH_THREADS) {
      const int py = i / PH_PW, px = i - py * PH_PW;
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      float4 a = zero4, c = zero4;
      if (d >= 0 && d < D && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const float4* src = reinterpret_cast<const float4*>(y + ((((size_t)b * D + d) * H + gy) * W + gx) * 8);
        a = __ldg(src);
        c = __ldg(src + 1);
      }
      dst[i] = a;
      dst[PH_PH * PH_PW + i] = c;
    }
  };
  load_plane(-1, 2);
  load_plane(0, 0);
  for (int d = 0; d < D; ++d) {
    load_plane(d + 1, (d + 1) % 3);                                    // slot (d+1)%3 held plane d-2: no longer read
    __syncthreads();
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float4* pl = plane + ((d + kd + 2) % 3) * PH_PLANE;       // plane d - 1 + kd
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int o = (ty + ky) * PH_PW + tx + kx;
          const float4 a = pl[o], c = pl[PH_PH * PH_PW + o];
          const float4 wa = wsm[((kd * 3 + ky) * 3 + kx) * 2], wc = wsm[((kd * 3 + ky) * 3 + kx) * 2 + 1];
          acc0 = fmaf(a.x, wa.x, acc0); acc1 = fmaf(a.y, wa.y, acc1);
          acc0 = fmaf(a.z, wa.z, acc0); acc1 = fmaf(a.w, wa.w, acc1);
          acc0 = fmaf(c.x, wc.x, acc0); acc1 = fmaf(c.y, wc.y, acc1);
          acc0 = fmaf(c.z, wc.z, acc0); acc1 = fmaf(c.w, wc.w, acc1);
        }
    }
    lsm[d * PH_THREADS + tid] = acc0 + acc1;
    __syncthreads();                                                   // plane (d+2)%3 is overwritten next
  }
  const int gy = y0 + ty, gx = x0 + tx;
  if (gy >= H || gx >= W) return;
  const int pix = gy * W + gx;
  const int ry = rh == 1 ? 0 : gy, rx = rw == 1 ? 0 : gx;
  const float near_ = range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx];
  const float far_ = range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx];
  // soft-max over depth in ATen's order, then depth_regression (depth_net.py:479-514): same arithmetic as depth_range_logits_kernel
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, lsm[d * PH_THREADS + tid]);
  float sum = 0.f;
  for (int d = 0; d < D; ++d) {
    const float e = expf(lsm[d * PH_THREADS + tid] - m);
    lsm[d * PH_THREADS + tid] = e;
    sum += e;
  }
  float mean = 0.f;
  for (int d = 0; d < D; ++d) {
    const float pr = fdiv(lsm[d * PH_THREADS + tid], sum);
    lsm[d * PH_THREADS + tid] = pr;
    mean = fadd(mean, fmul(pr, hypothesis(near_, far_, d, D, inv_depth)));
  }
  float var = 0.f;
  for (int d = 0; d < D; ++d) {
    const float t = fsub(hypothesis(near_, far_, d, D, inv_depth), mean);
    var = fadd(var, fmul(lsm[d * PH_THREADS + tid], fmul(t, t)));
    if (prob_out) prob_out[((size_t)b * D + d) * HW + pix] = lsm[d * PH_THREADS + tid];
  }
  const float half = fmul(ci_scale, sqrtf(fmaxf(var, 1e-12f)));
  const float first = hypothesis(near_, far_, 0, D, inv_depth), last = hypothesis(near_, far_, D - 1, D, inv_depth);
  float lo, hi, dep;
  if (inv_depth) {
    lo = fdiv(1.f, fminf(fadd(mean, half), first));
    hi = fdiv(1.f, fmaxf(fsub(mean, half), last));
    dep = fdiv(1.f, mean);
  } else {
    lo = fmaxf(fsub(mean, half), first);
    hi = fminf(fadd(mean, half), last);
    dep = mean;
  }
  depth[(size_t)b * HW + pix] = dep;
  ci[(size_t)(b * 2 + 0) * HW + pix] = lo;
  ci[(size_t)(b * 2 + 1) * HW + pix] = hi;
  if (vol_range) {
    vol_range[(size_t)(b * 2 + 0) * HW + pix] = first;
    vol_range[(size_t)(b * 2 + 1) * HW + pix] = last;
  }
}

// Depth-split variant.  The kernel above walks all D planes of a 4x16-pixel tile in one CTA: 40 960 pixels of a DTU batch
// are 8.6 warps per SM, each in a chain of 64 dependent plane steps - latency-bound (a variant with 42 % fewer shared-memory
// loads on half the warps ran slower).  Here a CTA walks one depth CHUNK of the tile and keeps the soft-max statistics of
// its chunk on line instead of parking logits: running maximum m, s0 = sum e^(l-m), s1 = sum e^(l-m) x, s2 = sum e^(l-m) x^2
// with x = hypothesis - c about the interval's midpoint c (so that var = s2/s0 - (s1/s0)^2 does not cancel).  The chunk
// partials go to `scratch`; the last CTA of a tile (a self-resetting counter) merges them and writes depth / interval.
// 4 chunks give 35 warps per SM; no logits stash: 10 KB of shared memory per CTA.
__global__ void __launch_bounds__(PH_THREADS) prob_head_split_kernel(const float* __restrict__ y, const float* __restrict__ wgt,
                                                                     const float* __restrict__ range, int rh, int rw, int B, int D,
                                                                     int H, int W, int NCH, float ci_scale, int inv_depth,
                                                                     float4* __restrict__ scratch, int* __restrict__ counters,
                                                                     float* __restrict__ depth, float* __restrict__ ci,
                                                                     float* __restrict__ vol_range) {
  __shared__ __align__(16) float4 plane[3 * PH_PLANE];                 // [3][2 halves][PH_PH][PH_PW]
  __shared__ __align__(16) float4 wsm[54];                             // [27][2]
  __shared__ int s_last;
  const int tid = threadIdx.x, ty = tid / PH_TX, tx = tid % PH_TX;
  const int b = blockIdx.z / NCH, chunk = blockIdx.z - b * NCH;
  const int y0 = blockIdx.y * PH_TY, x0 = blockIdx.x * PH_TX;
  const int HW = H * W;
  const int per = (D + NCH - 1) / NCH;
  const int dlo = chunk * per, dhi = min(dlo + per, D);
  if (tid < 54) wsm[tid] = __ldg(reinterpret_cast<const float4*>(wgt) + tid);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_plane = [&](int d, int slot) {
    float4* dst = plane + slot * PH_PLANE;
    for (int i = tid; i < PH_PH * PH_PW; i += PH_THREADS) {
      const int py = i / PH_PW, px = i - py * PH_PW;
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      float4 a = zero4, c = zero4;
      if (d >= 0 && d < D && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        const float4* src = reinterpret_cast<const float4*>(y + ((((size_t)b * D + d) * H + gy) * W + gx) * 8);
        a = __ldg(src);
        c = __ldg(src + 1);
      }
      dst[i] = a;
      dst[PH_PH * PH_PW + i] = c;
    }
  };
  const int gy = y0 + ty, gx = x0 + tx;
  const bool live = gy < H && gx < W;
  const int ry = rh == 1 ? 0 : min(gy, H - 1), rx = rw == 1 ? 0 : min(gx, W - 1);
  const float near_ = range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx];
  const float far_ = range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx];
  const float first = hypothesis(near_, far_, 0, D, inv_depth), last = hypothesis(near_, far_, D - 1, D, inv_depth);
  const float cmid = 0.5f * (first + last);
  float m = -INFINITY, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  // slot of plane p: (p - dlo + 1) % 3  (plane dlo - 1 -> slot 0)
  load_plane(dlo - 1, 0);
  load_plane(dlo, 1);
  for (int d = dlo; d < dhi; ++d) {
    load_plane(d + 1, (d - dlo + 2) % 3);                              // that slot held plane d - 2: no longer read
    __syncthreads();
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float4* pl = plane + ((d - dlo + kd) % 3) * PH_PLANE;      // plane d - 1 + kd
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int o = (ty + ky) * PH_PW + tx + kx;
          const float4 a = pl[o], c = pl[PH_PH * PH_PW + o];
          const float4 wa = wsm[((kd * 3 + ky) * 3 + kx) * 2], wc = wsm[((kd * 3 + ky) * 3 + kx) * 2 + 1];
          acc0 = fmaf(a.x, wa.x, acc0); acc1 = fmaf(a.y, wa.y, acc1);
          acc0 = fmaf(a.z, wa.z, acc0); acc1 = fmaf(a.w, wa.w, acc1);
          acc0 = fmaf(c.x, wc.x, acc0); acc1 = fmaf(c.y, wc.y, acc1);
          acc0 = fmaf(c.z, wc.z, acc0); acc1 = fmaf(c.w, wc.w, acc1);
        }
    }
    const float l = acc0 + acc1;
    const float x = hypothesis(near_, far_, d, D, inv_depth) - cmid;
    if (l > m) {
      const float sc = expf(m - l);                                    // 0 on the first plane (m = -inf)
      s0 = fmaf(s0, sc, 1.f); s1 = fmaf(s1, sc, x); s2 = fmaf(s2, sc, x * x);
      m = l;
    } else {
      const float e = expf(l - m);
      s0 += e; s1 = fmaf(e, x, s1); s2 = fmaf(e, x * x, s2);
    }
    __syncthreads();                                                   // the slot of plane d - 1 is overwritten next
  }
  const int tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const size_t sbase = ((size_t)tile * NCH) * PH_THREADS;
  scratch[sbase + (size_t)chunk * PH_THREADS + tid] = make_float4(m, s0, s1, s2);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(counters + tile, 1);
    s_last = prev == NCH - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float M = -INFINITY;
  for (int c = 0; c < NCH; ++c) M = fmaxf(M, __ldcg(&scratch[sbase + (size_t)c * PH_THREADS + tid]).x);
  float S0 = 0.f, S1 = 0.f, S2 = 0.f;
  for (int c = 0; c < NCH; ++c) {
    const float4 pt = __ldcg(&scratch[sbase + (size_t)c * PH_THREADS + tid]);
    const float sc = expf(pt.x - M);
    S0 = fmaf(pt.y, sc, S0); S1 = fmaf(pt.z, sc, S1); S2 = fmaf(pt.w, sc, S2);
  }
  if (!live) return;
  const int pix = gy * W + gx;
  const float mx = S1 / S0;
  const float mean = cmid + mx;
  const float var = fmaxf(S2 / S0 - mx * mx, 0.f);
  const float half = fmul(ci_scale, sqrtf(fmaxf(var, 1e-12f)));
  float lo, hi, dep;
  if (inv_depth) {
    lo = fdiv(1.f, fminf(fadd(mean, half), first));
    hi = fdiv(1.f, fmaxf(fsub(mean, half), last));
    dep = fdiv(1.f, mean);
  } else {
    lo = fmaxf(fsub(mean, half), first);
    hi = fminf(fadd(mean, half), last);
    dep = mean;
  }
  depth[(size_t)b * HW + pix] = dep;
  ci[(size_t)(b * 2 + 0) * HW + pix] = lo;
  ci[(size_t)(b * 2 + 1) * HW + pix] = hi;
  if (vol_range) {
    vol_range[(size_t)(b * 2 + 0) * HW + pix] = first;
    vol_range[(size_t)(b * 2 + 1) * HW + pix] = last;
  }
}

// ---------------------------------------------------------------------------
// planar -> channels-last
// ---------------------------------------------------------------------------
template <int CP>
__global__ void planar_to_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t S, int64_t NS) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < NS; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / S, s = i % S;
    const float* sp = src + n * C * S + s;
    float v[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) v[c] = c < C ? __ldg(sp + (int64_t)c * S) : 0.f;
    float4* dp = reinterpret_cast<float4*>(dst + i * CP);
#pragma unroll
    for (int c = 0; c < CP; c += 4) dp[c / 4] = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  }
}

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_planar_to_channels_last(const float* src, float* dst, int N, int C, int64_t S, int Cpad, void* stream) {
  GDB_REQUIRE(src && dst && N > 0 && C > 0 && S > 0, GDB_E_BADARG, "gdb_planar_to_channels_last: bad argument");
  GDB_REQUIRE(Cpad >= C && Cpad % 4 == 0, GDB_E_BADARG, "gdb_planar_to_channels_last: Cpad must be >= C and a multiple of 4");
  GDB_REQUIRE(aligned16(dst), GDB_E_ALIGN, "gdb_planar_to_channels_last: dst not 16-byte aligned");
  int64_t NS = (int64_t)N * S;
  int blocks = (int)std::min<int64_t>((NS + 255) / 256, (int64_t)sm_count() * 16);
  cudaStream_t st = as_stream(stream);
  switch (Cpad) {
    case 4: planar_to_cl_kernel<4><<<blocks, 256, 0, st>>>(src, dst, C, S, NS); break;
    case 8: planar_to_cl_kernel<8><<<blocks, 256, 0, st>>>(src, dst, C, S, NS); break;
    case 16: planar_to_cl_kernel<16><<<blocks, 256, 0, st>>>(src, dst, C, S, NS); break;
    case 32: planar_to_cl_kernel<32><<<blocks, 256, 0, st>>>(src, dst, C, S, NS); break;
    default: return fail(GDB_E_UNSUPPORTED, "gdb_planar_to_channels_last: Cpad=%d not in {4,8,16,32}", Cpad);
  }
  return cuda_check("gdb_planar_to_channels_last");
}

extern "C" int gdb_homography_mats(const float* src_exts, const float* src_ints, const float* tar_exts,
                                   const float* tar_ints, float src_scale, float tar_scale, int B, int V, float* proj,
                                   void* stream) {
  GDB_REQUIRE(src_exts && src_ints && tar_exts && tar_ints && proj && B > 0 && V > 0, GDB_E_BADARG,
              "gdb_homography_mats: bad argument");
  homography_kernel<<<(B * V + 63) / 64, 64, 0, as_stream(stream)>>>(src_exts, src_ints, tar_exts, tar_ints, src_scale,
                                                                    tar_scale, B, V, proj);
  return cuda_check("gdb_homography_mats");
}

extern "C" int gdb_depth_values(const float* depth_range, int rh, int rw, int B, int D, int Ht, int Wt, int inv_depth,
                                float* out, void* stream) {
  GDB_REQUIRE(depth_range && out && B > 0 && D > 0 && Ht > 0 && Wt > 0, GDB_E_BADARG, "gdb_depth_values: bad argument");
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == Ht && rw == Wt), GDB_E_BADARG,
              "gdb_depth_values: depth_range must be 1x1 or %dx%d, got %dx%d", Ht, Wt, rh, rw);
  size_t n = (size_t)B * D * Ht * Wt;
  int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 16);
  depth_values_kernel<<<blocks, 256, 0, as_stream(stream)>>>(depth_range, rh, rw, B, D, Ht, Wt, inv_depth, out);
  return cuda_check("gdb_depth_values");
}

extern "C" int gdb_warp_variance_fwd(const float* feat_cl, const float* proj, const float* depth_range, int rh, int rw,
                                     int B, int V, int C, int Hs, int Ws, int D, int Ht, int Wt, int inv_depth,
                                     int out_channels_last, float* variance, void* stream) {
  GDB_REQUIRE(feat_cl && proj && depth_range && variance, GDB_E_BADARG, "gdb_warp_variance_fwd: null pointer");
  GDB_REQUIRE(B > 0 && Hs > 0 && Ws > 0 && D > 0 && Ht > 0 && Wt > 0, GDB_E_BADARG, "gdb_warp_variance_fwd: bad size");
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == Ht && rw == Wt), GDB_E_BADARG,
              "gdb_warp_variance_fwd: depth_range must be 1x1 or %dx%d, got %dx%d", Ht, Wt, rh, rw);
  GDB_REQUIRE(aligned16(feat_cl) && (!out_channels_last || aligned16(variance)), GDB_E_ALIGN,
              "gdb_warp_variance_fwd: feat_cl / channels-last output not 16-byte aligned");
  GDB_REQUIRE(Hs < 32000 && Ws < 32000, GDB_E_UNSUPPORTED, "gdb_warp_variance_fwd: source map larger than 32000 px");
  GDB_REQUIRE(B <= 65535, GDB_E_UNSUPPORTED, "gdb_warp_variance_fwd: B > 65535");
  cudaStream_t st = as_stream(stream);
#define GDB_WV(CC, VV) \
  if (C == CC && V == VV) return launch_warp_variance<CC, VV>(feat_cl, proj, depth_range, rh, rw, B, Hs, Ws, D, Ht, Wt, inv_depth, out_channels_last, variance, st);
  GDB_WV(32, 2) GDB_WV(32, 3) GDB_WV(32, 4) GDB_WV(16, 2) GDB_WV(16, 3) GDB_WV(16, 4) GDB_WV(8, 2) GDB_WV(8, 3) GDB_WV(8, 4)
#undef GDB_WV
  return fail(GDB_E_UNSUPPORTED, "gdb_warp_variance_fwd: C=%d V=%d not instantiated (C in {8,16,32}, V in {2,3,4})", C, V);
}

extern "C" int gdb_depth_range_fwd(const float* depth_range, int rh, int rw, const float* prob, int B, int D, int h, int w,
                                   float ci_scale, int inv_depth, float* depth, float* ci, float* vol_range, void* stream) {
  GDB_REQUIRE(depth_range && prob && depth && ci && B > 0 && D > 0 && h > 0 && w > 0, GDB_E_BADARG,
              "gdb_depth_range_fwd: bad argument");
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == h && rw == w), GDB_E_BADARG,
              "gdb_depth_range_fwd: depth_range must be 1x1 or %dx%d, got %dx%d", h, w, rh, rw);
  int n = B * h * w;
  depth_range_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(depth_range, rh, rw, prob, B, D, h, w, ci_scale,
                                                                    inv_depth, depth, ci, vol_range);
  return cuda_check("gdb_depth_range_fwd");
}

extern "C" int gdb_depth_range_from_logits_fwd(const float* depth_range, int rh, int rw, const float* logits, int64_t stride_b,
                                               int64_t stride_d, int64_t stride_pix, int B, int D, int h, int w, float ci_scale,
                                               int inv_depth, float* depth, float* ci, float* vol_range, float* prob_out,
                                               void* stream) {
  GDB_REQUIRE(depth_range && logits && depth && ci && B > 0 && D > 0 && h > 0 && w > 0, GDB_E_BADARG,
              "gdb_depth_range_from_logits_fwd: bad argument");
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == h && rw == w), GDB_E_BADARG,
              "gdb_depth_range_from_logits_fwd: depth_range must be 1x1 or %dx%d, got %dx%d", h, w, rh, rw);
  GDB_REQUIRE(D <= 64, GDB_E_UNSUPPORTED, "gdb_depth_range_from_logits_fwd: D=%d > 64 not instantiated", D);
  int n = B * h * w;
  cudaStream_t st = as_stream(stream);
  if (D <= 8)
    depth_range_logits_kernel<8><<<(n + 127) / 128, 128, 0, st>>>(depth_range, rh, rw, logits, stride_b, stride_d, stride_pix, B, D, h, w,
                                                                  ci_scale, inv_depth, depth, ci, vol_range, prob_out);
  else if (D <= 36)
    depth_range_logits_kernel<36><<<(n + 127) / 128, 128, 0, st>>>(depth_range, rh, rw, logits, stride_b, stride_d, stride_pix, B, D, h, w,
                                                                   ci_scale, inv_depth, depth, ci, vol_range, prob_out);
  else
    depth_range_logits_kernel<64><<<(n + 127) / 128, 128, 0, st>>>(depth_range, rh, rw, logits, stride_b, stride_d, stride_pix, B, D, h, w,
                                                                   ci_scale, inv_depth, depth, ci, vol_range, prob_out);
  return cuda_check("gdb_depth_range_from_logits_fwd");
}

extern "C" int64_t gdb_prob_head_split_scratch_floats(int B, int h, int w, int nchunks) {
  const int64_t tiles = (int64_t)B * ((h + PH_TY - 1) / PH_TY) * ((w + PH_TX - 1) / PH_TX);
  return tiles * nchunks * PH_THREADS * 4;
}
extern "C" int64_t gdb_prob_head_split_counters(int B, int h, int w) {
  return (int64_t)B * ((h + PH_TY - 1) / PH_TY) * ((w + PH_TX - 1) / PH_TX);
}

extern "C" int gdb_prob_head_depth_range_split_fwd(const float* y_cl, const float* weight, const float* depth_range, int rh, int rw, int B,
                                                   int C, int D, int h, int w, int nchunks, float ci_scale, int inv_depth, float* scratch,
                                                   int* counters, float* depth, float* ci, float* vol_range, void* stream) {
  GDB_REQUIRE(y_cl && weight && depth_range && depth && ci && scratch && counters && B > 0 && D > 0 && h > 0 && w > 0, GDB_E_BADARG,
              "gdb_prob_head_depth_range_split_fwd: bad argument");
  GDB_REQUIRE(C == 8, GDB_E_UNSUPPORTED, "gdb_prob_head_depth_range_split_fwd: C=%d not instantiated (8)", C);
  GDB_REQUIRE(nchunks >= 1 && nchunks <= D && (long)B * nchunks <= 65535, GDB_E_BADARG,
              "gdb_prob_head_depth_range_split_fwd: nchunks %d outside [1, D] or B * nchunks > 65535", nchunks);
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == h && rw == w), GDB_E_BADARG,
              "gdb_prob_head_depth_range_split_fwd: depth_range must be 1x1 or %dx%d, got %dx%d", h, w, rh, rw);
  GDB_REQUIRE(aligned16(y_cl) && aligned16(weight) && aligned16(scratch), GDB_E_ALIGN,
              "gdb_prob_head_depth_range_split_fwd: y / weight / scratch must be 16-byte aligned");
  // every chunk must hold at least one plane: the last CTA of a tile is found by counting nchunks arrivals
  const int per = (D + nchunks - 1) / nchunks;
  GDB_REQUIRE((nchunks - 1) * per < D, GDB_E_BADARG, "gdb_prob_head_depth_range_split_fwd: %d chunks of %d planes leave one empty (D = %d)",
              nchunks, per, D);
  dim3 grid((w + PH_TX - 1) / PH_TX, (h + PH_TY - 1) / PH_TY, B * nchunks);
  // the arrival counters are zeroed by every call (a memset node, legal inside stream capture): an aborted launch or a
  // recycled buffer can never leave a stale count behind
  {
    cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(int) * (size_t)grid.x * grid.y * B, as_stream(stream));
    if (e != cudaSuccess) return fail((int)e, "gdb_prob_head_depth_range_split_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
  }
  prob_head_split_kernel<<<grid, PH_THREADS, 0, as_stream(stream)>>>(y_cl, weight, depth_range, rh, rw, B, D, h, w, nchunks, ci_scale, inv_depth,
                                                                     reinterpret_cast<float4*>(scratch), counters, depth, ci, vol_range);
  return cuda_check("gdb_prob_head_depth_range_split_fwd");
}

extern "C" int gdb_prob_head_depth_range_fwd(const float* y_cl, const float* weight, const float* depth_range, int rh, int rw, int B,
                                             int C, int D, int h, int w, float ci_scale, int inv_depth, float* depth, float* ci,
                                             float* vol_range, float* prob_out, void* stream) {
  GDB_REQUIRE(y_cl && weight && depth_range && depth && ci && B > 0 && D > 0 && h > 0 && w > 0, GDB_E_BADARG,
              "gdb_prob_head_depth_range_fwd: bad argument");
  GDB_REQUIRE(C == 8, GDB_E_UNSUPPORTED, "gdb_prob_head_depth_range_fwd: C=%d not instantiated (8)", C);
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == h && rw == w), GDB_E_BADARG,
              "gdb_prob_head_depth_range_fwd: depth_range must be 1x1 or %dx%d, got %dx%d", h, w, rh, rw);
  GDB_REQUIRE(aligned16(y_cl) && aligned16(weight), GDB_E_ALIGN, "gdb_prob_head_depth_range_fwd: y / weight must be 16-byte aligned");
  const int smem = (3 * PH_PLANE + 54) * 16 + D * PH_THREADS * 4;
  GDB_REQUIRE(smem <= 227 * 1024, GDB_E_UNSUPPORTED, "gdb_prob_head_depth_range_fwd: D=%d needs %d B of shared memory", D, smem);
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, prob_head_depth_range_kernel, smem);
    if (e != cudaSuccess) return fail((int)e, "gdb_prob_head_depth_range_fwd: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e));
  }
  dim3 grid((w + PH_TX - 1) / PH_TX, (h + PH_TY - 1) / PH_TY, B);
  prob_head_depth_range_kernel<<<grid, PH_THREADS, smem, as_stream(stream)>>>(y_cl, weight, depth_range, rh, rw, B, D, h, w, ci_scale, inv_depth,
                                                                      depth, ci, vol_range, prob_out);
  return cuda_check("gdb_prob_head_depth_range_fwd");
}
