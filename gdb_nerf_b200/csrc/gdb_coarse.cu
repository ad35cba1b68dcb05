// K5: training-only coarse render of the cascade's first stage,
// forward and backward in one kernel template.
//
// Reference: depth_net.py:49-116 (_render_rays), :201-298 (coarse NeRF), :301-341 (build_rays),
// :344-396 (get_img_feat_vectorized).  One ray per cost-volume pixel, S fixed samples per ray
// placed uniformly in depth or disparity inside the stage's confidence interval; per sample a
// trilinear voxel fetch, per view a border-clamped bilinear fetch of the (feature | low-res rgb)
// texture (points behind a camera fetch the corner texel: grid = -99, :369-372) and the four
// direction features; the coarse MLP (same trunk as the fine one, `color` head); compositing
// with T = cumprod(1 - alpha + 1e-10) and NO renormalisation (:109-114).
//
// Lane = (ray, sample): the S samples of a ray sit in adjacent lanes.  BWD = false writes
// rgb (B,3,Hi,Wi); BWD = true recomputes the forward and produces the adjoints exactly like
// gdb_render_bwd.cu (dual numbers seeded on the sample depth, reverse-mode MLP, shared-memory
// accumulated parameter gradients, atomics for the texture / volume taps).
#include <algorithm>

#include "gdb_render_common.cuh"
#include "gdb_autodiff.cuh"

namespace gdb {

struct CoarseParams {
  const float* tex;          // (B*V, Hs, Ws, FP) channels-last feature+rgb texture of the stage (level 0 of gdb_prepare_sources)
  const float* vol;          // (B, D, Hi, Wi, 8)
  const float* ray_range;    // (B, 2, Hi, Wi) confidence interval of the stage (depth units)
  const float* vol_range;    // (B, 2, Hi, Wi)
  const float* cam;          // camera block built with the STAGE intrinsics
  const float* mlp;          // packed like the fine MLP; the `color` head sits in the weight.0 / weight.2 slots
  float* rgb;                // (B, 3, Hi, Wi)                              [forward]
  const float* g_rgb;        // (B, 3, Hi, Wi)                              [backward]
  float* d_mlp;              // accumulated
  float* d_tex;              // (B*V, Hs, Ws, FP) accumulated
  float* d_vol;              // (B, D, Hi, Wi, 8) accumulated
  float* d_ray_range;        // written
  float* d_vol_range;        // written
  int cam_stride, B, Hi, Wi, Hs, Ws, D, S, inv_depth;
};

template <int FEAT_DIM, int V, bool BWD>
__global__ void __launch_bounds__(128, 1) coarse_kernel(const CoarseParams p) {
  using ML = MlpLayout<FEAT_DIM>;
  constexpr int F = ML::F, FP = ML::FP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* wsm = reinterpret_cast<float*>(smem_raw);
  float* gsm = wsm + ((ML::TOTAL + 31) & ~31);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < ML::TOTAL; i += blockDim.x) {
    wsm[i] = p.mlp[i];
    if (BWD) gsm[i] = 0.f;
  }
  __syncthreads();

  const int HW = p.Hi * p.Wi;
  const int NR = p.B * HW;
  const int ns = p.S;
  const int G = 32 / ns;
  const int ngroups = (NR + G - 1) / G;
  const int rl = lane / ns, slot = lane - rl * ns;
  const int seg_base = rl * ns;
  const unsigned full = 0xffffffffu;

  for (int grp = blockIdx.x * nwarps + warp; grp < ngroups; grp += gridDim.x * nwarps) {
    const int ray = grp * G + rl;
    const bool active = rl < G && ray < NR;
    const int ridx = active ? ray : 0;
    const int b = ridx / HW, pix = ridx - b * HW;
    const int py = pix / p.Wi, px = pix - py * p.Wi;
    const float* head = p.cam + (size_t)b * p.cam_stride;

    // ---- sample placement (depth_net.py:79-93)
    const float rn_raw = p.ray_range[(size_t)(b * 2 + 0) * HW + pix], rf_raw = p.ray_range[(size_t)(b * 2 + 1) * HW + pix];
    const float vn_raw = p.vol_range[(size_t)(b * 2 + 0) * HW + pix], vf_raw = p.vol_range[(size_t)(b * 2 + 1) * HW + pix];
    float rn = rn_raw, rf = rf_raw, vn = vn_raw, vf = vf_raw;
    if (p.inv_depth) { rn = 1.f / rf_raw; rf = 1.f / rn_raw; vn = 1.f / vf_raw; vf = 1.f / vn_raw; }
    const float kfrac = ((float)slot + 0.5f) / (float)ns;
    const float step0 = (float)slot / (float)ns, step1 = (float)(slot + 1) / (float)ns;
    const float zs = 0.5f * ((rn + (rf - rn) * step0) + (rn + (rf - rn) * step1));
    const float dnorm = 2.f * (zs - vn) / (vf - vn) - 1.f;
    const float zf = p.inv_depth ? 1.f / zs : zs;
    const Dual z(zf, 1.f);
    const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    const float* M = head + CAM_M;
    const float dx = fmaf(fx, M[0], fmaf(fy, M[1], M[2])), dy = fmaf(fx, M[3], fmaf(fy, M[4], M[5])), dz = fmaf(fx, M[6], fmaf(fy, M[7], M[8]));
    const float ox = head[CAM_O + 0], oy = head[CAM_O + 1], oz = head[CAM_O + 2];
    const Dual wx = Dual(dx) * z + Dual(ox), wy = Dual(dy) * z + Dual(oy), wz = Dual(dz) * z + Dual(oz);

    // ---- voxel feature (grid_sample 3-D, border, align_corners=False)
    float vox[8], t_vox[8], vox_w[8];
    int vox_off[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { vox[k] = 0.f; t_vox[k] = 0.f; vox_w[k] = 0.f; vox_off[k] = 0; }
    if (active) {
      const float u = 2.f * fx / (float)p.Wi - 1.f, v = 2.f * fy / (float)p.Hi - 1.f;
      float ix = fminf(fmaxf(((u + 1.f) * (float)p.Wi - 1.f) * 0.5f, 0.f), (float)(p.Wi - 1));
      float iy = fminf(fmaxf(((v + 1.f) * (float)p.Hi - 1.f) * 0.5f, 0.f), (float)(p.Hi - 1));
      Dual izd = dclamp((Dual(dnorm, 1.f) + Dual(1.f)) * Dual((float)p.D) * Dual(0.5f) - Dual(0.5f), 0.f, (float)(p.D - 1));
      float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(izd.v);
      float tx = ix - x0f, ty = iy - y0f;
      Dual tz(izd.v - z0f, izd.d);
      int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
      int x1 = min(x0 + 1, p.Wi - 1), y1 = min(y0 + 1, p.Hi - 1), z1 = min(z0 + 1, p.D - 1);
      const float* vb = p.vol + (size_t)b * p.D * HW * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int xx = (k & 1) ? x1 : x0, yy = (k & 2) ? y1 : y0, zz = (k & 4) ? z1 : z0;
        float wxy = ((k & 1) ? tx : 1.f - tx) * ((k & 2) ? ty : 1.f - ty);
        Dual wzd = (k & 4) ? tz : Dual(1.f) - tz;
        int off = (zz * p.Hi + yy) * p.Wi + xx;
        vox_off[k] = off;
        vox_w[k] = wxy * wzd.v;
        const float* tp = vb + (size_t)off * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float t = __ldg(tp + c);
          vox[c] = fmaf(t, wxy * wzd.v, vox[c]);
          t_vox[c] = fmaf(t, wxy * wzd.d, t_vox[c]);
        }
      }
    }

    // ---- per view: bilinear texture fetch + direction features (depth_net.py:361-394)
    float fr[V][F], t_fr[V][F], dir[V][4], t_dir[V][4];
    int tap_off[V][4];
    float tap_w[V][4];
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      const float* cv = head + CAM_HEAD + CAM_VIEW * v;
#pragma unroll
      for (int c = 0; c < F; ++c) { fr[v][c] = 0.f; t_fr[v][c] = 0.f; }
#pragma unroll
      for (int c = 0; c < 4; ++c) { dir[v][c] = 0.f; t_dir[v][c] = 0.f; tap_off[v][c] = 0; tap_w[v][c] = 0.f; }
      if (!active) continue;
      Dual cx = wx * Dual(cv[CV_E + 0]) + wy * Dual(cv[CV_E + 1]) + wz * Dual(cv[CV_E + 2]) + Dual(cv[CV_E + 3]);
      Dual cy = wx * Dual(cv[CV_E + 4]) + wy * Dual(cv[CV_E + 5]) + wz * Dual(cv[CV_E + 6]) + Dual(cv[CV_E + 7]);
      Dual cz = wx * Dual(cv[CV_E + 8]) + wy * Dual(cv[CV_E + 9]) + wz * Dual(cv[CV_E + 10]) + Dual(cv[CV_E + 11]);
      Dual ix = cx * Dual(cv[CV_K + 0]) + cy * Dual(cv[CV_K + 1]) + cz * Dual(cv[CV_K + 2]);
      Dual iy = cx * Dual(cv[CV_K + 3]) + cy * Dual(cv[CV_K + 4]) + cz * Dual(cv[CV_K + 5]);
      Dual iz = cx * Dual(cv[CV_K + 6]) + cy * Dual(cv[CV_K + 7]) + cz * Dual(cv[CV_K + 8]);
      Dual gx, gy;
      if (iz.v < 1e-8f) {                                          // behind the camera: constant far-outside coordinate
        gx = Dual(-99.f); gy = Dual(-99.f);
      } else {
        gx = Dual(2.f) * (ix / iz) / Dual((float)p.Ws) - Dual(1.f);
        gy = Dual(2.f) * (iy / iz) / Dual((float)p.Hs) - Dual(1.f);
      }
      Dual pxd = dclamp(((gx + Dual(1.f)) * Dual((float)p.Ws) - Dual(1.f)) * Dual(0.5f), 0.f, (float)(p.Ws - 1));
      Dual pyd = dclamp(((gy + Dual(1.f)) * Dual((float)p.Hs) - Dual(1.f)) * Dual(0.5f), 0.f, (float)(p.Hs - 1));
      float x0f = floorf(pxd.v), y0f = floorf(pyd.v);
      Dual tx(pxd.v - x0f, pxd.d), ty(pyd.v - y0f, pyd.d);
      int x0 = (int)x0f, y0 = (int)y0f;
      int x1 = min(x0 + 1, p.Ws - 1), y1 = min(y0 + 1, p.Hs - 1);
      int o[4] = {y0 * p.Ws + x0, y0 * p.Ws + x1, y1 * p.Ws + x0, y1 * p.Ws + x1};
      Dual wt[4] = {(Dual(1.f) - tx) * (Dual(1.f) - ty), tx * (Dual(1.f) - ty), (Dual(1.f) - tx) * ty, tx * ty};
      const float* base = p.tex + (size_t)(b * V + v) * p.Hs * p.Ws * FP;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        tap_off[v][t] = o[t];
        tap_w[v][t] = wt[t].v;
        const float* tp = base + (size_t)o[t] * FP;
#pragma unroll
        for (int c = 0; c < F; ++c) {
          float tv = __ldg(tp + c);
          fr[v][c] = fmaf(tv, wt[t].v, fr[v][c]);
          t_fr[v][c] = fmaf(tv, wt[t].d, t_fr[v][c]);
        }
      }
      Dual ax = wx - Dual(ox), ay = wy - Dual(oy), az = wz - Dual(oz);
      dunit3(ax, ay, az);
      Dual sx = wx - Dual(cv[CV_C + 0]), sy = wy - Dual(cv[CV_C + 1]), sz = wz - Dual(cv[CV_C + 2]);
      dunit3(sx, sy, sz);
      Dual ddx = ax - sx, ddy = ay - sy, ddz = az - sz;
      dunit3(ddx, ddy, ddz);
      Dual dot = ax * sx + ay * sy + az * sz;
      dir[v][0] = ddx.v; dir[v][1] = ddy.v; dir[v][2] = ddz.v; dir[v][3] = dot.v;
      t_dir[v][0] = ddx.d; t_dir[v][1] = ddy.d; t_dir[v][2] = ddz.d; t_dir[v][3] = dot.d;
    }

    // ---- coarse MLP forward (depth_net.py:248-298), activations kept
    float rpre[V][F], xv[V][F], varc[F], meanc[F];
#pragma unroll 1
    for (int v = 0; v < V; ++v)
#pragma unroll 1
      for (int c = 0; c < F; ++c) {
        float t = wsm[ML::VIEW_B + c];
#pragma unroll
        for (int k = 0; k < 4; ++k) t = fmaf(wsm[ML::VIEW_W + k * FP + c], dir[v][k], t);
        rpre[v][c] = t;
        xv[v][c] = fr[v][c] + fmaxf(t, 0.f);
      }
#pragma unroll 1
    for (int c = 0; c < F; ++c) {
      float m = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) m += xv[v][c];
      m *= (1.f / V);
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { float t = xv[v][c] - m; s = fmaf(t, t, s); }
      meanc[c] = m; varc[c] = s * (1.f / (V - 1));
    }
    float gact[V][32], apre[V], pa[V], im[32], imgpre[16], img[16], hpre[64], h[64], hid[V][64], cpre[V], qv[V];
    {
      float gsh[32];
#pragma unroll 1
      for (int k = 0; k < 32; ++k) {
        float t = wsm[ML::GLOB_B + k];
        for (int c = 0; c < F; ++c) t = fmaf(wsm[ML::GLOB_W + (F + c) * 32 + k], varc[c], fmaf(wsm[ML::GLOB_W + (2 * F + c) * 32 + k], meanc[c], t));
        gsh[k] = t;
      }
      float amax = -1e30f;
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        float s = wsm[ML::AGG_B];
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
          float t = gsh[k];
          for (int c = 0; c < F; ++c) t = fmaf(wsm[ML::GLOB_W + c * 32 + k], xv[v][c], t);
          t = fmaxf(t, 0.f);
          gact[v][k] = t;
          s = fmaf(t, wsm[ML::AGG_W + k], s);
        }
        apre[v] = s;
        amax = fmaxf(amax, fmaxf(s, 0.f));
      }
      float asum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { pa[v] = expf(fmaxf(apre[v], 0.f) - amax); asum += pa[v]; }
#pragma unroll
      for (int v = 0; v < V; ++v) pa[v] /= asum;
#pragma unroll 1
      for (int k = 0; k < 32; ++k) {
        float t = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) t = fmaf(gact[v][k], pa[v], t);
        im[k] = t;
      }
#pragma unroll 1
      for (int k = 0; k < 16; ++k) {
        float t = wsm[ML::FC_B + k];
        for (int c = 0; c < 32; ++c) t = fmaf(wsm[ML::FC_W + c * 16 + k], im[c], t);
        imgpre[k] = t; img[k] = fmaxf(t, 0.f);
      }
#pragma unroll 1
      for (int k = 0; k < 64; ++k) {
        float t = wsm[ML::LR0_B + k];
        for (int c = 0; c < 8; ++c) t = fmaf(wsm[ML::LR0_W + c * 64 + k], vox[c], t);
        for (int c = 0; c < 16; ++c) t = fmaf(wsm[ML::LR0_W + (8 + c) * 64 + k], img[c], t);
        hpre[k] = t; h[k] = fmaxf(t, 0.f);
      }
    }
    float sraw = wsm[ML::SIG_B];
#pragma unroll 1
    for (int k = 0; k < 64; ++k) sraw = fmaf(h[k], wsm[ML::SIG_W + k], sraw);
    const float sigma = sraw > 20.f ? sraw : log1pf(expf(sraw));
    {
      float wmax = -1e30f;
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        float s = wsm[ML::W2_B];
#pragma unroll 1
        for (int k = 0; k < 64; ++k) {
          float t = wsm[ML::W0_B + k];
          for (int c = 0; c < 64; ++c) t = fmaf(wsm[ML::W0_W + c * 64 + k], h[c], t);
          for (int c = 0; c < 8; ++c) t = fmaf(wsm[ML::W0_W + (64 + c) * 64 + k], vox[c], t);
          for (int c = 0; c < 16; ++c) t = fmaf(wsm[ML::W0_W + (72 + c) * 64 + k], img[c], t);
          for (int c = 0; c < F; ++c) t = fmaf(wsm[ML::W0_W + (88 + c) * 64 + k], fr[v][c], t);
          for (int c = 0; c < 4; ++c) t = fmaf(wsm[ML::W0_W + (88 + F + c) * 64 + k], dir[v][c], t);
          t = fmaxf(t, 0.f);
          hid[v][k] = t;
          s = fmaf(t, wsm[ML::W2_W + k], s);
        }
        cpre[v] = s;
        wmax = fmaxf(wmax, fmaxf(s, 0.f));
      }
      float wsum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { qv[v] = expf(fmaxf(cpre[v], 0.f) - wmax); wsum += qv[v]; }
#pragma unroll
      for (int v = 0; v < V; ++v) qv[v] /= wsum;
    }
    float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int v = 0; v < V; ++v) rgb[c] = fmaf(fr[v][F - 3 + c], qv[v], rgb[c]);

    // ---- compositing (depth_net.py:107-114): T = exclusive cumprod of (1 - alpha + 1e-10), no renormalisation
    const float alpha = active ? 1.f - expf(-sigma) : 0.f;
    const float om = 1.f - alpha + 1e-10f;
    float T = 1.f;
    for (int k = 0; k + 1 < ns; ++k) {
      float o = __shfl_sync(full, om, min(seg_base + k, 31));
      if (k < slot) T *= o;
    }
    const float wgt = alpha * T;
    auto seg_sum = [&](float x) {
      float acc = x;
      for (int k = 1; k < ns; ++k) {
        float o = __shfl_down_sync(full, x, k);
        if (slot == 0) acc += o;
      }
      return acc;
    };
    if constexpr (!BWD) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s = seg_sum(wgt * rgb[c]);
        if (active && slot == 0) p.rgb[((size_t)b * 3 + c) * HW + pix] = s;
      }
      continue;
    }

    // =========================================================== backward
    float Gc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) Gc[c] = active ? p.g_rgb[((size_t)b * 3 + c) * HW + pix] : 0.f;
    const float qi = active ? Gc[0] * rgb[0] + Gc[1] * rgb[1] + Gc[2] * rgb[2] : 0.f;     // dL/dw_i
    float g_alpha = qi * T;
    for (int k = 1; k < ns; ++k) {
      float gk = __shfl_sync(full, qi * alpha, min(seg_base + k, 31));
      float prod = 1.f;
      for (int j = 0; j < k; ++j) {
        float o = __shfl_sync(full, om, min(seg_base + j, 31));
        if (j != slot) prod *= o;
      }
      if (k > slot) g_alpha -= gk * prod;
    }
    const float g_sigma = active ? g_alpha * (1.f - alpha) : 0.f;
    const float g_sraw = g_sigma * (sraw > 20.f ? 1.f : 1.f / (1.f + expf(-sraw)));

    float g_h[64], g_vox[8], g_img[16], g_fr[V][F], g_dir[V][4];
#pragma unroll 1
    for (int k = 0; k < 64; ++k) g_h[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) g_vox[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) g_img[k] = 0.f;
    float g_q[V], gq_dot = 0.f;
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float s = 0.f;
      for (int c = 0; c < F; ++c) g_fr[v][c] = 0.f;
      for (int c = 0; c < 3; ++c) { s = fmaf(wgt * Gc[c], fr[v][F - 3 + c], s); g_fr[v][F - 3 + c] = qv[v] * wgt * Gc[c]; }
      g_q[v] = s;
      gq_dot = fmaf(qv[v], s, gq_dot);
    }
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float g_c = qv[v] * (g_q[v] - gq_dot);
      float g_cpre = cpre[v] > 0.f ? g_c : 0.f;
      float g_hid[64];
#pragma unroll 1
      for (int k = 0; k < 64; ++k) g_hid[k] = hid[v][k] > 0.f ? g_cpre * wsm[ML::W2_W + k] : 0.f;
      accum_outer(gsm + ML::W2_W, 64, &g_cpre, 1, hid[v], 64, lane);
      { float s = warp_sum(g_cpre); if (lane == 0) atomicAdd(gsm + ML::W2_B, s); }
      { float one = 1.f; accum_outer(gsm + ML::W0_B, 64, &one, 1, g_hid, 64, lane); }
      accum_outer(gsm + ML::W0_W, 64, h, 64, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + 64 * 64, 64, vox, 8, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + 72 * 64, 64, img, 16, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + 88 * 64, 64, fr[v], F, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + (88 + F) * 64, 64, dir[v], 4, g_hid, 64, lane);
#pragma unroll 1
      for (int c = 0; c < 64; ++c) {
        float s = 0.f;
        for (int k = 0; k < 64; ++k) s = fmaf(wsm[ML::W0_W + c * 64 + k], g_hid[k], s);
        g_h[c] += s;
      }
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float s = 0.f;
        for (int k = 0; k < 64; ++k) s = fmaf(wsm[ML::W0_W + (64 + c) * 64 + k], g_hid[k], s);
        g_vox[c] += s;
      }
#pragma unroll 1
      for (int c = 0; c < 16; ++c) {
        float s = 0.f;
        for (int k = 0; k < 64; ++k) s = fmaf(wsm[ML::W0_W + (72 + c) * 64 + k], g_hid[k], s);
        g_img[c] += s;
      }
#pragma unroll 1
      for (int c = 0; c < F; ++c) {
        float s = 0.f;
        for (int k = 0; k < 64; ++k) s = fmaf(wsm[ML::W0_W + (88 + c) * 64 + k], g_hid[k], s);
        g_fr[v][c] += s;
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float s = 0.f;
        for (int k = 0; k < 64; ++k) s = fmaf(wsm[ML::W0_W + (88 + F + c) * 64 + k], g_hid[k], s);
        g_dir[v][c] = s;
      }
    }
    accum_outer(gsm + ML::SIG_W, 64, &g_sraw, 1, h, 64, lane);
    { float s = warp_sum(g_sraw); if (lane == 0) atomicAdd(gsm + ML::SIG_B, s); }
#pragma unroll 1
    for (int c = 0; c < 64; ++c) g_h[c] = (hpre[c] > 0.f) ? g_h[c] + g_sraw * wsm[ML::SIG_W + c] : 0.f;
    { float one = 1.f; accum_outer(gsm + ML::LR0_B, 64, &one, 1, g_h, 64, lane); }
    accum_outer(gsm + ML::LR0_W, 64, vox, 8, g_h, 64, lane);
    accum_outer(gsm + ML::LR0_W + 8 * 64, 64, img, 16, g_h, 64, lane);
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      float s = 0.f;
      for (int k = 0; k < 64; ++k) s = fmaf(wsm[ML::LR0_W + c * 64 + k], g_h[k], s);
      g_vox[c] += s;
    }
    float g_imgpre[16];
#pragma unroll 1
    for (int c = 0; c < 16; ++c) {
      float s = g_img[c];
      for (int k = 0; k < 64; ++k) s = fmaf(wsm[ML::LR0_W + (8 + c) * 64 + k], g_h[k], s);
      g_imgpre[c] = imgpre[c] > 0.f ? s : 0.f;
    }
    { float one = 1.f; accum_outer(gsm + ML::FC_B, 16, &one, 1, g_imgpre, 16, lane); }
    accum_outer(gsm + ML::FC_W, 16, im, 32, g_imgpre, 16, lane);
    float g_im[32];
#pragma unroll 1
    for (int c = 0; c < 32; ++c) {
      float s = 0.f;
      for (int k = 0; k < 16; ++k) s = fmaf(wsm[ML::FC_W + c * 16 + k], g_imgpre[k], s);
      g_im[c] = s;
    }
    float g_pa[V], gpa_dot = 0.f;
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float s = 0.f;
      for (int k = 0; k < 32; ++k) s = fmaf(g_im[k], gact[v][k], s);
      g_pa[v] = s;
      gpa_dot = fmaf(pa[v], s, gpa_dot);
    }
    float g_var[F], g_mean[F], g_x[V][F];
#pragma unroll 1
    for (int c = 0; c < F; ++c) { g_var[c] = 0.f; g_mean[c] = 0.f; }
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float g_a = pa[v] * (g_pa[v] - gpa_dot);
      float g_apre = apre[v] > 0.f ? g_a : 0.f;
      float g_gpre[32];
#pragma unroll 1
      for (int k = 0; k < 32; ++k) {
        float s = fmaf(pa[v], g_im[k], g_apre * wsm[ML::AGG_W + k]);
        g_gpre[k] = gact[v][k] > 0.f ? s : 0.f;
      }
      accum_outer(gsm + ML::AGG_W, 32, &g_apre, 1, gact[v], 32, lane);
      { float s = warp_sum(g_apre); if (lane == 0) atomicAdd(gsm + ML::AGG_B, s); }
      { float one = 1.f; accum_outer(gsm + ML::GLOB_B, 32, &one, 1, g_gpre, 32, lane); }
      accum_outer(gsm + ML::GLOB_W, 32, xv[v], F, g_gpre, 32, lane);
      accum_outer(gsm + ML::GLOB_W + F * 32, 32, varc, F, g_gpre, 32, lane);
      accum_outer(gsm + ML::GLOB_W + 2 * F * 32, 32, meanc, F, g_gpre, 32, lane);
#pragma unroll 1
      for (int c = 0; c < F; ++c) {
        float sx = 0.f, sv = 0.f, sm = 0.f;
        for (int k = 0; k < 32; ++k) {
          sx = fmaf(wsm[ML::GLOB_W + c * 32 + k], g_gpre[k], sx);
          sv = fmaf(wsm[ML::GLOB_W + (F + c) * 32 + k], g_gpre[k], sv);
          sm = fmaf(wsm[ML::GLOB_W + (2 * F + c) * 32 + k], g_gpre[k], sm);
        }
        g_x[v][c] = sx; g_var[c] += sv; g_mean[c] += sm;
      }
    }
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float g_rpre[F];
#pragma unroll 1
      for (int c = 0; c < F; ++c) {
        float gx = g_x[v][c] + g_mean[c] * (1.f / V) + g_var[c] * (2.f / (V - 1)) * (xv[v][c] - meanc[c]);
        g_fr[v][c] += gx;
        g_rpre[c] = rpre[v][c] > 0.f ? gx : 0.f;
      }
      { float one = 1.f; accum_outer(gsm + ML::VIEW_B, FP, &one, 1, g_rpre, F, lane); }
      accum_outer(gsm + ML::VIEW_W, FP, dir[v], 4, g_rpre, F, lane);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float s = 0.f;
        for (int c = 0; c < F; ++c) s = fmaf(wsm[ML::VIEW_W + k * FP + c], g_rpre[c], s);
        g_dir[v][k] += s;
      }
    }

    // ---- taps, sample position
    float g_dn = 0.f, g_z = 0.f;
    if (active) {
#pragma unroll
      for (int c = 0; c < 8; ++c) g_dn = fmaf(g_vox[c], t_vox[c], g_dn);
      float* dvb = p.d_vol + (size_t)b * p.D * HW * 8;
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {
        if (vox_w[k] == 0.f) continue;
        float* tp = dvb + (size_t)vox_off[k] * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) atomicAdd(tp + c, vox_w[k] * g_vox[c]);
      }
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        for (int c = 0; c < F; ++c) g_z = fmaf(g_fr[v][c], t_fr[v][c], g_z);
        for (int c = 0; c < 4; ++c) g_z = fmaf(g_dir[v][c], t_dir[v][c], g_z);
        float* dbase = p.d_tex + (size_t)(b * V + v) * p.Hs * p.Ws * FP;
#pragma unroll 1
        for (int t = 0; t < 4; ++t) {
          if (tap_w[v][t] == 0.f) continue;
          float* tp = dbase + (size_t)tap_off[v][t] * FP;
          for (int c = 0; c < F; ++c) atomicAdd(tp + c, tap_w[v][t] * g_fr[v][c]);
        }
      }
    }
    float g_zs = p.inv_depth ? -g_z * zf * zf : g_z;
    const float span = vf - vn;
    g_zs += g_dn * 2.f / span;
    float g_vn = g_dn * 2.f * (zs - vf) / (span * span);
    float g_vf = -g_dn * 2.f * (zs - vn) / (span * span);
    float g_rn = g_zs * (1.f - kfrac), g_rf = g_zs * kfrac;
    if (!active) { g_rn = g_rf = g_vn = g_vf = 0.f; }
    g_rn = seg_sum(g_rn); g_rf = seg_sum(g_rf); g_vn = seg_sum(g_vn); g_vf = seg_sum(g_vf);
    if (active && slot == 0) {
      float o_rn = g_rn, o_rf = g_rf, o_vn = g_vn, o_vf = g_vf;
      if (p.inv_depth) {                                             // rn = 1 / rf_raw, rf = 1 / rn_raw (and the same for the volume range)
        o_rf = -g_rn / (rf_raw * rf_raw); o_rn = -g_rf / (rn_raw * rn_raw);
        o_vf = -g_vn / (vf_raw * vf_raw); o_vn = -g_vf / (vn_raw * vn_raw);
      }
      p.d_ray_range[(size_t)(b * 2 + 0) * HW + pix] = o_rn;
      p.d_ray_range[(size_t)(b * 2 + 1) * HW + pix] = o_rf;
      p.d_vol_range[(size_t)(b * 2 + 0) * HW + pix] = o_vn;
      p.d_vol_range[(size_t)(b * 2 + 1) * HW + pix] = o_vf;
    }
    __syncwarp();
  }

  if constexpr (BWD) {
    __syncthreads();
    for (int i = threadIdx.x; i < ML::TOTAL; i += blockDim.x) {
      float gval = gsm[i];
      if (gval != 0.f) atomicAdd(p.d_mlp + i, gval);
    }
  }
}

template <int FEAT_DIM, int V, bool BWD>
static int launch_coarse(const CoarseParams& p, cudaStream_t st, const char* what) {
  using ML = MlpLayout<FEAT_DIM>;
  auto kern = coarse_kernel<FEAT_DIM, V, BWD>;
  const size_t smem = (size_t)2 * ((ML::TOTAL + 31) & ~31) * sizeof(float);
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, kern, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute(%zu B): %s", what, smem, cudaGetErrorString(e));
  }
  const int nwarps = 4;
  const int G = 32 / p.S;
  const long NR = (long)p.B * p.Hi * p.Wi;
  const long ngroups = (NR + G - 1) / G;
  long ctas = (ngroups + nwarps - 1) / nwarps;
  if (ctas > 2L * sm_count()) ctas = 2L * sm_count();
  kern<<<(int)ctas, nwarps * 32, smem, st>>>(p);
  return cuda_check(what);
}

template <bool BWD>
static int dispatch_coarse(const CoarseParams& p, int feat_dim, int V, cudaStream_t st, const char* what) {
#define GDB_C(FD, VV) \
  if (feat_dim == FD && V == VV) return launch_coarse<FD, VV, BWD>(p, st, what);
  GDB_C(32, 2) GDB_C(32, 3) GDB_C(32, 4) GDB_C(16, 2) GDB_C(16, 3) GDB_C(16, 4)
#undef GDB_C
  return fail(GDB_E_UNSUPPORTED, "%s: (feat_dim=%d, V=%d) not instantiated", what, feat_dim, V);
}

static int check_coarse(const char* what, const void* tex, const void* vol, const void* rr, const void* vr, const void* cam,
                        const void* mlp, int cam_stride, int B, int V, int Hi, int Wi, int Hs, int Ws, int D, int S) {
  GDB_REQUIRE(tex && vol && rr && vr && cam && mlp, GDB_E_BADARG, "%s: null pointer", what);
  GDB_REQUIRE(B > 0 && Hi > 0 && Wi > 0 && Hs > 0 && Ws > 0 && D > 0, GDB_E_BADARG, "%s: bad size", what);
  GDB_REQUIRE(S >= 1 && S <= 32, GDB_E_BADARG, "%s: num_samples must be 1..32", what);
  GDB_REQUIRE(cam_stride == CAM_HEAD + CAM_VIEW * V, GDB_E_BADARG, "%s: cam_stride %d != %d", what, cam_stride, CAM_HEAD + CAM_VIEW * V);
  return 0;
}

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_coarse_render_fwd(const float* tex, const float* vol_cl, const float* ray_range, const float* vol_range,
                                     const float* cam, int cam_stride, const float* mlp, int B, int V, int Hi, int Wi, int Hs,
                                     int Ws, int feat_dim, int D, int num_samples, int inv_depth, float* rgb, void* stream) {
  int rc = check_coarse("gdb_coarse_render_fwd", tex, vol_cl, ray_range, vol_range, cam, mlp, cam_stride, B, V, Hi, Wi, Hs, Ws, D, num_samples);
  if (rc) return rc;
  GDB_REQUIRE(rgb, GDB_E_BADARG, "gdb_coarse_render_fwd: null output");
  CoarseParams p{};
  p.tex = tex; p.vol = vol_cl; p.ray_range = ray_range; p.vol_range = vol_range; p.cam = cam; p.mlp = mlp; p.rgb = rgb;
  p.cam_stride = cam_stride; p.B = B; p.Hi = Hi; p.Wi = Wi; p.Hs = Hs; p.Ws = Ws; p.D = D; p.S = num_samples; p.inv_depth = inv_depth;
  return dispatch_coarse<false>(p, feat_dim, V, as_stream(stream), "gdb_coarse_render_fwd");
}

extern "C" int gdb_coarse_render_bwd(const float* tex, const float* vol_cl, const float* ray_range, const float* vol_range,
                                     const float* cam, int cam_stride, const float* mlp, int B, int V, int Hi, int Wi, int Hs,
                                     int Ws, int feat_dim, int D, int num_samples, int inv_depth, const float* g_rgb, float* d_mlp,
                                     float* d_tex, float* d_vol, float* d_ray_range, float* d_vol_range, void* stream) {
  int rc = check_coarse("gdb_coarse_render_bwd", tex, vol_cl, ray_range, vol_range, cam, mlp, cam_stride, B, V, Hi, Wi, Hs, Ws, D, num_samples);
  if (rc) return rc;
  GDB_REQUIRE(g_rgb && d_mlp && d_tex && d_vol && d_ray_range && d_vol_range, GDB_E_BADARG, "gdb_coarse_render_bwd: null pointer");
  CoarseParams p{};
  p.tex = tex; p.vol = vol_cl; p.ray_range = ray_range; p.vol_range = vol_range; p.cam = cam; p.mlp = mlp;
  p.g_rgb = g_rgb; p.d_mlp = d_mlp; p.d_tex = d_tex; p.d_vol = d_vol; p.d_ray_range = d_ray_range; p.d_vol_range = d_vol_range;
  p.cam_stride = cam_stride; p.B = B; p.Hi = Hi; p.Wi = Wi; p.Hs = Hs; p.Ws = Ws; p.D = D; p.S = num_samples; p.inv_depth = inv_depth;
  return dispatch_coarse<true>(p, feat_dim, V, as_stream(stream), "gdb_coarse_render_bwd");
}
