// K4: backward of the fused per-bundle render (training config, SURVEY.md 8b "Autograd").
//
// Adjoint of gdb_render_fused_fwd with respect to everything the reference
// differentiates through (bundle_sampler.py:193-371, nerf.py:58-115,
// utils.py:19-43,88-121, network.py:54-91):
//   - the MLP parameters                      -> d_mlp   (packed layout of the forward)
//   - the feature+rgb texture, all mip levels -> d_tex   (pulled down to level 0 by gdb_prepare_sources_bwd)
//   - the regularised feature volume          -> d_vol
//   - the per-bundle depth interval / volume range (through the sample positions:
//     tap coordinates, mip level, direction features, voxel depth coordinate,
//     composited depth)                        -> d_depth_range, d_vol_range
// Source images and camera parameters are inputs of the model and get no gradient.
//
// Method.  Same work decomposition as the forward (lane = (bundle, sample slot)),
// everything recomputed from the forward's inputs - nothing is saved by the
// forward.  Per lane:
//   1. encode with DUAL numbers seeded on the sample depth z (forward mode):
//      every gathered quantity carries its derivative d/dz (d/d(dnorm) for the
//      voxel feature), including the derivative through bilinear weights, the
//      mip interpolation fraction, the projected footprint and the unit
//      direction vectors; clamps have zero derivative where they bite, as autograd.
//   2. MLP forward keeping the activations in (thread-local) memory.
//   3. compositing forward + backward with segmented warp shuffles.
//   4. MLP backward (reverse mode); weight gradients are reduced over the warp
//      with a 31-shuffle transpose-reduce per 32 weights, accumulated in a
//      per-CTA shared-memory copy of the gradient block and flushed once.
//   5. scatter of the tap adjoints (atomics) and dL/dz = <adjoint, tangent>,
//      chained to the bundle's near/far and the volume range, segment-summed.
#include <algorithm>

#include "gdb_render_common.cuh"
#include "gdb_autodiff.cuh"

namespace gdb {

struct RenderBwdParams {
  RenderParams f;              // the forward's arguments (outputs unused)
  const float* g_feat;         // (B, CT, Hb, Wb) planar upstream gradient of the bundle features
  const float* g_depth;        // (B, Hb, Wb) or null
  const float* g_opacity;      // (B, Hb, Wb) or null
  float* d_mlp;                // packed, accumulated (caller zeroes)
  float* d_tex;                // mip chain, accumulated (caller zeroes)
  float* d_vol;                // (B, D, Hb, Wb, 8) accumulated (caller zeroes)
  float* d_depth_range;        // (B, 2, Hb, Wb) written
  float* d_vol_range;          // (B, 2, Hb, Wb) written
};

template <int BS, int FEAT_DIM, int V>
__global__ void __launch_bounds__(128, 1) render_bwd_kernel(const RenderBwdParams q) {
  using ML = MlpLayout<FEAT_DIM>;
  constexpr int BB = BS * BS, F = ML::F, FP = ML::FP, R = 3 * BB, CT = R + F + 8;
  const RenderParams& p = q.f;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* wsm = reinterpret_cast<float*>(smem_raw);                 // weights
  float* gsm = wsm + ((ML::TOTAL + 31) & ~31);                     // gradient accumulator
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < ML::TOTAL; i += blockDim.x) { wsm[i] = p.mlp[i]; gsm[i] = 0.f; }
  __syncthreads();

  const int HW = p.Hb * p.Wb;
  const int NB = p.B * HW;
  const int ns = p.max_samples;
  const int G = 32 / ns;
  const int ngroups = (NB + G - 1) / G;
  const int bl = lane / ns, slot = lane - bl * ns;
  const int seg_base = bl * ns;
  const unsigned full = 0xffffffffu;

  for (int grp = blockIdx.x * nwarps + warp; grp < ngroups; grp += gridDim.x * nwarps) {
    const int bundle = grp * G + bl;
    const bool has_bundle = bl < G && bundle < NB;
    const int bidx = has_bundle ? bundle : 0;
    const int b = bidx / HW, pix = bidx - b * HW;
    const int yb = pix / p.Wb, xb = pix - yb * p.Wb;
    const float* head = p.cam + (size_t)b * p.cam_stride;

    const float nr_raw = p.depth_range[(size_t)(b * 2 + 0) * HW + pix], fr_raw = p.depth_range[(size_t)(b * 2 + 1) * HW + pix];
    const float vn_raw = p.vol_range[(size_t)(b * 2 + 0) * HW + pix], vf_raw = p.vol_range[(size_t)(b * 2 + 1) * HW + pix];
    float nr = nr_raw, fr_ = fr_raw, vn = vn_raw, vf = vf_raw;
    const int n = bundle_sample_count(nr, fr_, head[CAM_MINIV], ns, p.inv_depth, p.adaptive);
    if (p.inv_depth) { nr = fdiv(1.f, nr); fr_ = fdiv(1.f, fr_); vn = fdiv(1.f, vn); vf = fdiv(1.f, vf); }
    const bool active = has_bundle && slot < n;
    float zf, dnorm;
    sample_depth(nr, fr_, vn, vf, n, slot, p.inv_depth, zf, dnorm);
    const float zs = p.inv_depth ? fdiv(1.f, zf) : zf;            // sample position in the sampling domain
    BundleGeom<BS> geo;
    geo.init(head, yb, xb, p.H, p.W);
    const float ox = head[CAM_O + 0], oy = head[CAM_O + 1], oz = head[CAM_O + 2];
    const Dual z(zf, 1.f);                                        // seed: d/dz

    // ---------------------------------------------------------------- 1. encode with duals
    Dual cwx(0.f), cwy(0.f), cwz(0.f);
#pragma unroll
    for (int j = 0; j < BB; ++j) {
      float dx, dy, dz;
      geo.ray_dir(head, j, dx, dy, dz);
      cwx = cwx + (Dual(dx) * z + Dual(ox)); cwy = cwy + (Dual(dy) * z + Dual(oy)); cwz = cwz + (Dual(dz) * z + Dual(oz));
    }
    cwx = cwx * Dual(1.f / BB); cwy = cwy * Dual(1.f / BB); cwz = cwz * Dual(1.f / BB);
    Dual ball;
    {
      Dual ex = cwx - Dual(ox), ey = cwy - Dual(oy), ez = cwz - Dual(oz);
      ball = dsqrt(ex * ex + ey * ey + ez * ez) * Dual(geo.unit_ball);
    }

    // voxel feature: value, tangent with respect to dnorm, taps
    float vox[8], t_vox[8];
    int vox_off[8];
    float vox_w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { vox[k] = 0.f; t_vox[k] = 0.f; vox_off[k] = 0; vox_w[k] = 0.f; }
    if (active) {
      float ix = fminf(fmaxf(((geo.u + 1.f) * (float)p.Wb - 1.f) * 0.5f, 0.f), (float)(p.Wb - 1));
      float iy = fminf(fmaxf(((geo.v + 1.f) * (float)p.Hb - 1.f) * 0.5f, 0.f), (float)(p.Hb - 1));
      Dual izd = dclamp((Dual(dnorm, 1.f) + Dual(1.f)) * Dual((float)p.D) * Dual(0.5f) - Dual(0.5f), 0.f, (float)(p.D - 1));
      float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(izd.v);
      float tx = ix - x0f, ty = iy - y0f;
      Dual tz(izd.v - z0f, izd.d);
      int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
      int x1 = min(x0 + 1, p.Wb - 1), y1 = min(y0 + 1, p.Hb - 1), z1 = min(z0 + 1, p.D - 1);
      const float* vb = p.vol + (size_t)b * p.D * HW * p.vol_stride;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int xx = (k & 1) ? x1 : x0, yy = (k & 2) ? y1 : y0, zz = (k & 4) ? z1 : z0;
        float wxy = ((k & 1) ? tx : 1.f - tx) * ((k & 2) ? ty : 1.f - ty);
        Dual wz = (k & 4) ? tz : Dual(1.f) - tz;
        int off = ((zz * p.Hb + yy) * p.Wb + xx);
        vox_off[k] = off;
        vox_w[k] = wxy * wz.v;
        const float* tp = vb + (size_t)off * p.vol_stride;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float t = __ldg(tp + c);
          vox[c] = fmaf(t, wxy * wz.v, vox[c]);
          t_vox[c] = fmaf(t, wxy * wz.d, t_vox[c]);
        }
      }
    }

    float fr[V][F], t_fr[V][F], dir[V][4], t_dir[V][4], col[V][R], t_col[V][R];
    int tap_off[V][8];        // 4 taps of level l0, 4 of level l1 (texel offsets inside the level's view slab)
    float tap_w[V][8];
    int64_t tap_base[V][2];   // float offsets of the (level, view) slabs inside tex
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      const float* cv = head + CAM_HEAD + CAM_VIEW * v;
      Dual ccx(0.f), ccy(0.f), ccz(0.f);
#pragma unroll 1
      for (int j = 0; j < BB; ++j) {
        float dx, dy, dz;
        geo.ray_dir(head, j, dx, dy, dz);
        Dual wx = Dual(dx) * z + Dual(ox), wy = Dual(dy) * z + Dual(oy), wz = Dual(dz) * z + Dual(oz);
        Dual cx = wx * Dual(cv[CV_E + 0]) + wy * Dual(cv[CV_E + 1]) + wz * Dual(cv[CV_E + 2]) + Dual(cv[CV_E + 3]);
        Dual cy = wx * Dual(cv[CV_E + 4]) + wy * Dual(cv[CV_E + 5]) + wz * Dual(cv[CV_E + 6]) + Dual(cv[CV_E + 7]);
        Dual cz = wx * Dual(cv[CV_E + 8]) + wy * Dual(cv[CV_E + 9]) + wz * Dual(cv[CV_E + 10]) + Dual(cv[CV_E + 11]);
        ccx = ccx + cx; ccy = ccy + cy; ccz = ccz + cz;
        // fine colour of ray j (bundle_sampler.py:327-337)
        Dual ix = cx * Dual(cv[CV_K + 0]) + cy * Dual(cv[CV_K + 1]) + cz * Dual(cv[CV_K + 2]);
        Dual iy = cx * Dual(cv[CV_K + 3]) + cy * Dual(cv[CV_K + 4]) + cz * Dual(cv[CV_K + 5]);
        Dual iz = dmax(cx * Dual(cv[CV_K + 6]) + cy * Dual(cv[CV_K + 7]) + cz * Dual(cv[CV_K + 8]), 1e-6f);
        Dual gx = Dual(2.f) * (ix / iz) / Dual((float)p.W) - Dual(1.f), gy = Dual(2.f) * (iy / iz) / Dual((float)p.H) - Dual(1.f);
        Dual px = dclamp(((gx + Dual(1.f)) * Dual((float)p.W) - Dual(1.f)) * Dual(0.5f), 0.f, (float)(p.W - 1));
        Dual py = dclamp(((gy + Dual(1.f)) * Dual((float)p.H) - Dual(1.f)) * Dual(0.5f), 0.f, (float)(p.H - 1));
        float x0f = floorf(px.v), y0f = floorf(py.v);
        Dual tx(px.v - x0f, px.d), ty(py.v - y0f, py.d);
        int x0 = (int)x0f, y0 = (int)y0f;
        int x1 = min(x0 + 1, p.W - 1), y1 = min(y0 + 1, p.H - 1);
        Dual w00 = (Dual(1.f) - tx) * (Dual(1.f) - ty), w10 = tx * (Dual(1.f) - ty), w01 = (Dual(1.f) - tx) * ty, w11 = tx * ty;
        float cval[3] = {0.f, 0.f, 0.f}, ctan[3] = {0.f, 0.f, 0.f};
        if (active) {
          const float* ib = p.rgba + (size_t)(b * V + v) * p.H * p.W * 4;
          float4 a00 = ldg4(ib + (size_t)(y0 * p.W + x0) * 4), a10 = ldg4(ib + (size_t)(y0 * p.W + x1) * 4);
          float4 a01 = ldg4(ib + (size_t)(y1 * p.W + x0) * 4), a11 = ldg4(ib + (size_t)(y1 * p.W + x1) * 4);
          cval[0] = a00.x * w00.v + a10.x * w10.v + a01.x * w01.v + a11.x * w11.v;
          cval[1] = a00.y * w00.v + a10.y * w10.v + a01.y * w01.v + a11.y * w11.v;
          cval[2] = a00.z * w00.v + a10.z * w10.v + a01.z * w01.v + a11.z * w11.v;
          ctan[0] = a00.x * w00.d + a10.x * w10.d + a01.x * w01.d + a11.x * w11.d;
          ctan[1] = a00.y * w00.d + a10.y * w10.d + a01.y * w01.d + a11.y * w11.d;
          ctan[2] = a00.z * w00.d + a10.z * w10.d + a01.z * w01.d + a11.z * w11.d;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { col[v][c * BB + j] = cval[c]; t_col[v][c * BB + j] = ctan[c]; }
      }
      ccx = ccx * Dual(1.f / BB); ccy = ccy * Dual(1.f / BB); ccz = ccz * Dual(1.f / BB);
      Dual dist = dsqrt(ccx * ccx + ccy * ccy + ccz * ccz);
      Dual sec = dist / ccz;
      Dual sec_sq = sec * sec;
      Dual rb = dist / ball;
      Dual foot = sec_sq / (dsqrt(dmax(rb * rb - Dual(1.f), 1e-12f)) + dsqrt(dmax(sec_sq - Dual(1.f), 1e-12f)));
      Dual lod = dlog2(foot / Dual(cv[CV_PIXR]));
      const float fb = (float)BS;
      Dual pxc = ccx * Dual(cv[CV_K + 0] / fb) + ccy * Dual(cv[CV_K + 1] / fb) + ccz * Dual(cv[CV_K + 2] / fb);
      Dual pyc = ccx * Dual(cv[CV_K + 3] / fb) + ccy * Dual(cv[CV_K + 4] / fb) + ccz * Dual(cv[CV_K + 5] / fb);
      Dual pzc = dmax(ccx * Dual(cv[CV_K + 6]) + ccy * Dual(cv[CV_K + 7]) + ccz * Dual(cv[CV_K + 8]), 1e-6f);
      Dual u01 = pxc / pzc / Dual((float)p.Wb), v01 = pyc / pzc / Dual((float)p.Hb);
#pragma unroll
      for (int c = 0; c < F; ++c) { fr[v][c] = 0.f; t_fr[v][c] = 0.f; }
#pragma unroll
      for (int c = 0; c < 4; ++c) { dir[v][c] = 0.f; t_dir[v][c] = 0.f; }
#pragma unroll
      for (int k = 0; k < 8; ++k) { tap_off[v][k] = 0; tap_w[v][k] = 0.f; }
      tap_base[v][0] = tap_base[v][1] = 0;
      if (active) {
        Dual flod = dclamp(lod, 0.f, (float)p.L);
        if (!(flod.v >= 0.f)) flod = Dual(0.f, 0.f);
        int l0 = (int)floorf(flod.v);
        int l1 = min(l0 + 1, p.L);
        Dual frac(flod.v - (float)l0, flod.d);
        const bool tri = flod.v > 0.f;
        int lw[2] = {p.Wb >> l0, p.Wb >> l1}, lh[2] = {p.Hb >> l0, p.Hb >> l1};
        tap_base[v][0] = p.tex_level[l0] + (int64_t)(b * V + v) * lh[0] * lw[0] * FP;
        tap_base[v][1] = p.tex_level[l1] + (int64_t)(b * V + v) * lh[1] * lw[1] * FP;
        Dual lev[2][F];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s == 1 && !tri) break;
          // nvdiffrast indexTextureLinear, boundary 'clamp'
          const float wd = (float)lw[s], hd = (float)lh[s];
          Dual u = dclamp(u01 * Dual(wd) - Dual(0.5f), 0.f, wd - 1.f), w = dclamp(v01 * Dual(hd) - Dual(0.5f), 0.f, hd - 1.f);
          bool cu = (u.v == 0.f) || (u.v == wd - 1.f), cvv = (w.v == 0.f) || (w.v == hd - 1.f);
          int iu0 = (int)floorf(u.v), iv0 = (int)floorf(w.v);
          int iu1 = iu0 + (cu ? 0 : 1), iv1 = iv0 + (cvv ? 0 : 1);
          Dual fu(u.v - (float)iu0, u.d), fv(w.v - (float)iv0, w.d);
          int o[4] = {iv0 * lw[s] + iu0, iv0 * lw[s] + iu1, iv1 * lw[s] + iu0, iv1 * lw[s] + iu1};
          Dual wt[4] = {(Dual(1.f) - fu) * (Dual(1.f) - fv), fu * (Dual(1.f) - fv), (Dual(1.f) - fu) * fv, fu * fv};
          const float* base = p.tex + tap_base[v][s];
#pragma unroll
          for (int c = 0; c < F; ++c) lev[s][c] = Dual(0.f);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            tap_off[v][s * 4 + t] = o[t];
            tap_w[v][s * 4 + t] = wt[t].v;
            const float* tp = base + (size_t)o[t] * FP;
#pragma unroll
            for (int c = 0; c < F; ++c) {
              float tv = __ldg(tp + c);
              lev[s][c].v = fmaf(tv, wt[t].v, lev[s][c].v);
              lev[s][c].d = fmaf(tv, wt[t].d, lev[s][c].d);
            }
          }
        }
        if (tri) {
#pragma unroll
          for (int c = 0; c < F; ++c) {
            Dual r = lev[0][c] + frac * (lev[1][c] - lev[0][c]);
            fr[v][c] = r.v; t_fr[v][c] = r.d;
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) { tap_w[v][t] *= (1.f - frac.v); tap_w[v][4 + t] *= frac.v; }
        } else {
#pragma unroll
          for (int c = 0; c < F; ++c) { fr[v][c] = lev[0][c].v; t_fr[v][c] = lev[0][c].d; }
        }
        Dual tx = cwx - Dual(ox), ty = cwy - Dual(oy), tz = cwz - Dual(oz);
        dunit3(tx, ty, tz);
        Dual sx = cwx - Dual(cv[CV_C + 0]), sy = cwy - Dual(cv[CV_C + 1]), sz = cwz - Dual(cv[CV_C + 2]);
        dunit3(sx, sy, sz);
        Dual ddx = tx - sx, ddy = ty - sy, ddz = tz - sz;
        dunit3(ddx, ddy, ddz);
        Dual dot = tx * sx + ty * sy + tz * sz;
        dir[v][0] = ddx.v; dir[v][1] = ddy.v; dir[v][2] = ddz.v; dir[v][3] = dot.v;
        t_dir[v][0] = ddx.d; t_dir[v][1] = ddy.d; t_dir[v][2] = ddz.d; t_dir[v][3] = dot.d;
      }
    }

    // ---------------------------------------------------------------- 2. MLP forward, activations kept
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
    float gact[V][32], apre[V], pa[V], im[32], imgpre[16], img[16], hpre[64], h[64], hid[V][64], cpre[V], qv[V], fhpre[8];
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
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      float t = wsm[ML::FH_B + k];
      for (int c = 0; c < 64; ++c) t = fmaf(wsm[ML::FH_W + c * 8 + k], h[c], t);
      fhpre[k] = t;
    }
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

    // ---------------------------------------------------------------- 3. compositing forward + backward
    const float alpha = active ? 1.f - expf(-sigma) : 0.f;
    float T = 1.f;
    for (int k = 0; k + 1 < ns; ++k) {
      float o = __shfl_sync(full, 1.f - alpha, min(seg_base + k, 31));
      if (k < slot) T *= o;
    }
    const float uw = alpha * T;                                   // unnormalised weight
    float wtot = 0.f;
    for (int k = 0; k < ns; ++k) {
      float o = __shfl_sync(full, uw, min(seg_base + k, 31));
      if (k < n) wtot += o;
    }
    const float wden = fmaxf(wtot, 1e-6f);
    const float wgt = active ? uw / wden : 0.f;

    // upstream gradients of my bundle
    const float* gf = q.g_feat + (size_t)b * CT * HW + pix;
    const float gD = (q.g_depth && has_bundle) ? q.g_depth[(size_t)b * HW + pix] : 0.f;
    const float gO = (q.g_opacity && has_bundle) ? q.g_opacity[(size_t)b * HW + pix] : 0.f;
    float g_out[R + F], g_fh[8];
    float qi = 0.f;                                               // dL/d(normalised weight)
#pragma unroll 1
    for (int c = 0; c < R + F; ++c) {
      float Gc = has_bundle ? gf[(size_t)c * HW] : 0.f;
      float val = 0.f;
      if (c < R) {
#pragma unroll
        for (int v = 0; v < V; ++v) val = fmaf(col[v][c], qv[v], val);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) val = fmaf(fr[v][c - R], qv[v], val);
      }
      qi = fmaf(Gc, val, qi);
      g_out[c] = wgt * Gc;
    }
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      float Gc = has_bundle ? gf[(size_t)(R + F + k) * HW] : 0.f;
      float val = fmaxf(fhpre[k], 0.f);
      qi = fmaf(Gc, val, qi);
      g_fh[k] = (fhpre[k] > 0.f) ? wgt * Gc : 0.f;
    }
    float g_z = 0.f;                                              // direct dL/dz (through the composited depth)
    {
      float g_zz;                                                 // dL/d(zz), zz = z or 1/z
      if (p.inv_depth) {
        float zz = fdiv(1.f, zf);
        float sd = 0.f;
        for (int k = 0; k < ns; ++k) {
          float o = __shfl_sync(full, wgt * zz, min(seg_base + k, 31));
          if (k < n) sd += o;
        }
        float g_sd = -gD / (sd * sd);
        qi = fmaf(g_sd, zz, qi);
        g_zz = g_sd * wgt;
        g_z = -g_zz * zz * zz;
      } else {
        qi = fmaf(gD, zf, qi);
        g_zz = gD * wgt;
        g_z = g_zz;
      }
      qi += gO;
    }
    if (!active) qi = 0.f;
    // w = u / max(sum u, 1e-6)
    float qw = 0.f;
    for (int k = 0; k < ns; ++k) {
      float o = __shfl_sync(full, qi * wgt, min(seg_base + k, 31));
      if (k < n) qw += o;
    }
    const float g_u = active ? ((wtot > 1e-6f) ? (qi - qw) / wden : qi / wden) : 0.f;
    // u_i = alpha_i * prod_{j<i} (1 - alpha_j):  dL/dalpha_i = g_u_i T_i - sum_{k>i} g_u_k alpha_k prod_{j<k, j!=i} (1 - alpha_j)
    float g_alpha = g_u * T;
    for (int k = 1; k < ns; ++k) {                                // later sample k (absolute slot)
      float gk = __shfl_sync(full, g_u * alpha, min(seg_base + k, 31));
      float prod = 1.f;
      for (int j = 0; j < k; ++j) {
        float om = __shfl_sync(full, 1.f - alpha, min(seg_base + j, 31));
        if (j != slot) prod *= om;
      }
      if (k > slot && k < n) g_alpha -= gk * prod;
    }
    const float g_sigma = active ? g_alpha * (1.f - alpha) : 0.f;  // alpha = 1 - exp(-sigma)
    const float g_sraw = g_sigma * (sraw > 20.f ? 1.f : 1.f / (1.f + expf(-sraw)));

    // ---------------------------------------------------------------- 4. MLP backward
    float g_h[64], g_vox[8], g_img[16], g_fr[V][F], g_dir[V][4], g_col_dot = 0.f;
#pragma unroll 1
    for (int k = 0; k < 64; ++k) g_h[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) g_vox[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) g_img[k] = 0.f;
    // blend: out = sum_v q_v [col_v | fr_v]
    float g_q[V], gq_dot = 0.f;
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float s = 0.f;
      for (int c = 0; c < R; ++c) { s = fmaf(g_out[c], col[v][c], s); g_col_dot = fmaf(qv[v] * g_out[c], t_col[v][c], g_col_dot); }
      for (int c = 0; c < F; ++c) { s = fmaf(g_out[R + c], fr[v][c], s); g_fr[v][c] = qv[v] * g_out[R + c]; }
      g_q[v] = s;
      gq_dot = fmaf(qv[v], s, gq_dot);
    }
    // weight.2 / weight.0 per view
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float g_c = qv[v] * (g_q[v] - gq_dot);
      float g_cpre = cpre[v] > 0.f ? g_c : 0.f;
      float g_hid[64];
#pragma unroll 1
      for (int k = 0; k < 64; ++k) g_hid[k] = hid[v][k] > 0.f ? g_cpre * wsm[ML::W2_W + k] : 0.f;
      // parameter gradients: weight.2 (vector + bias), weight.0 (bias, rows h|vox|img|fr_v|dir_v)
      accum_outer(gsm + ML::W2_W, 64, &g_cpre, 1, hid[v], 64, lane);
      { float s = warp_sum(g_cpre); if (lane == 0) atomicAdd(gsm + ML::W2_B, s); }
      { float one = 1.f; accum_outer(gsm + ML::W0_B, 64, &one, 1, g_hid, 64, lane); }
      accum_outer(gsm + ML::W0_W, 64, h, 64, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + 64 * 64, 64, vox, 8, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + 72 * 64, 64, img, 16, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + 88 * 64, 64, fr[v], F, g_hid, 64, lane);
      accum_outer(gsm + ML::W0_W + (88 + F) * 64, 64, dir[v], 4, g_hid, 64, lane);
      // input gradients
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
    // feat_head, sigma -> h
    accum_outer(gsm + ML::FH_W, 8, h, 64, g_fh, 8, lane);
    { float one = 1.f; accum_outer(gsm + ML::FH_B, 8, &one, 1, g_fh, 8, lane); }
    accum_outer(gsm + ML::SIG_W, 64, &g_sraw, 1, h, 64, lane);
    { float s = warp_sum(g_sraw); if (lane == 0) atomicAdd(gsm + ML::SIG_B, s); }
#pragma unroll 1
    for (int c = 0; c < 64; ++c) {
      float s = g_sraw * wsm[ML::SIG_W + c];
      for (int k = 0; k < 8; ++k) s = fmaf(wsm[ML::FH_W + c * 8 + k], g_fh[k], s);
      g_h[c] = (hpre[c] > 0.f) ? g_h[c] + s : 0.f;                // through relu: g_h is now dL/d(h_pre)
    }
    // lr0
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
    // fc
    { float one = 1.f; accum_outer(gsm + ML::FC_B, 16, &one, 1, g_imgpre, 16, lane); }
    accum_outer(gsm + ML::FC_W, 16, im, 32, g_imgpre, 16, lane);
    float g_im[32];
#pragma unroll 1
    for (int c = 0; c < 32; ++c) {
      float s = 0.f;
      for (int k = 0; k < 16; ++k) s = fmaf(wsm[ML::FC_W + c * 16 + k], g_imgpre[k], s);
      g_im[c] = s;
    }
    // im = sum_v pa_v g_v ; pa = softmax(relu(apre))
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
    // var/mean, residual, view_fc
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

    // ---------------------------------------------------------------- 5. taps and sample position
    float g_dn = 0.f;                                              // dL/d(dnorm)
    if (active) {
#pragma unroll
      for (int c = 0; c < 8; ++c) g_dn = fmaf(g_vox[c], t_vox[c], g_dn);
      float* dvb = q.d_vol + (size_t)b * p.D * HW * 8;
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {
        if (vox_w[k] == 0.f) continue;
        float* tp = dvb + (size_t)vox_off[k] * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) atomicAdd(tp + c, vox_w[k] * g_vox[c]);
      }
      g_z += g_col_dot;
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        for (int c = 0; c < F; ++c) g_z = fmaf(g_fr[v][c], t_fr[v][c], g_z);
        for (int c = 0; c < 4; ++c) g_z = fmaf(g_dir[v][c], t_dir[v][c], g_z);
#pragma unroll 1
        for (int t = 0; t < 8; ++t) {
          if (tap_w[v][t] == 0.f) continue;
          float* tp = q.d_tex + tap_base[v][t >> 2] + (size_t)tap_off[v][t] * FP;
          for (int c = 0; c < F; ++c) atomicAdd(tp + c, tap_w[v][t] * g_fr[v][c]);
        }
      }
    } else {
      g_z = 0.f;
    }
    // z -> sampling-domain position zs -> (near, far); dnorm -> (zs, vol near, vol far)
    float g_zs = p.inv_depth ? -g_z * zf * zf : g_z;               // z = 1 / zs
    const float span = vf - vn;
    g_zs += g_dn * 2.f / span;
    float g_vn = g_dn * 2.f * (zs - vf) / (span * span);
    float g_vf = -g_dn * 2.f * (zs - vn) / (span * span);
    const float kfrac = ((float)slot + 0.5f) / (float)n;
    float g_nr = g_zs * (1.f - kfrac), g_fr_ = g_zs * kfrac;
    if (!active) { g_nr = g_fr_ = g_vn = g_vf = 0.f; }
    auto seg_sum = [&](float x) {
      float acc = x;
      for (int k = 1; k < ns; ++k) {
        float o = __shfl_down_sync(full, x, k);
        if (slot == 0 && k < n) acc += o;
      }
      return acc;
    };
    g_nr = seg_sum(g_nr); g_fr_ = seg_sum(g_fr_); g_vn = seg_sum(g_vn); g_vf = seg_sum(g_vf);
    if (has_bundle && slot == 0) {
      if (p.inv_depth) {                                           // the kernel's nr = 1 / nr_raw etc.
        g_nr = -g_nr / (nr_raw * nr_raw); g_fr_ = -g_fr_ / (fr_raw * fr_raw);
        g_vn = -g_vn / (vn_raw * vn_raw); g_vf = -g_vf / (vf_raw * vf_raw);
      }
      q.d_depth_range[(size_t)(b * 2 + 0) * HW + pix] = g_nr;
      q.d_depth_range[(size_t)(b * 2 + 1) * HW + pix] = g_fr_;
      q.d_vol_range[(size_t)(b * 2 + 0) * HW + pix] = g_vn;
      q.d_vol_range[(size_t)(b * 2 + 1) * HW + pix] = g_vf;
    }
    __syncwarp();
  }

  __syncthreads();
  for (int i = threadIdx.x; i < ML::TOTAL; i += blockDim.x) {
    float gval = gsm[i];
    if (gval != 0.f) atomicAdd(q.d_mlp + i, gval);
  }
}

template <int BS, int FEAT_DIM, int V>
static int launch_render_bwd(const RenderBwdParams& q, cudaStream_t st) {
  using ML = MlpLayout<FEAT_DIM>;
  auto kern = render_bwd_kernel<BS, FEAT_DIM, V>;
  const size_t smem = (size_t)2 * ((ML::TOTAL + 31) & ~31) * sizeof(float);
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, kern, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "gdb_render_fused_bwd: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
  }
  const int nwarps = 4;
  const int G = 32 / q.f.max_samples;
  const long NB = (long)q.f.B * q.f.Hb * q.f.Wb;
  const long ngroups = (NB + G - 1) / G;
  long ctas = (ngroups + nwarps - 1) / nwarps;
  if (ctas > 2L * sm_count()) ctas = 2L * sm_count();
  kern<<<(int)ctas, nwarps * 32, smem, st>>>(q);
  return cuda_check("gdb_render_fused_bwd");
}

// -------- texture gradient: pull the mip levels down to level 0 and split off the feature channels
// (adjoint of the box-filter chain of gdb_prepare_sources; the rgb channels end at the input images: dropped)
__global__ void tex_pull_kernel(float* __restrict__ d_tex, int64_t fine_off, int64_t coarse_off, int BV, int hf, int wf, int FP) {
  // fine level (hf x wf) += 0.25 * coarse level (hf/2 x wf/2)
  int64_t n = (int64_t)BV * hf * wf * FP;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % FP);
    int64_t t = i / FP;
    int x = (int)(t % wf);
    t /= wf;
    int y = (int)(t % hf);
    int64_t bv = t / hf;
    d_tex[fine_off + i] += 0.25f * d_tex[coarse_off + ((bv * (hf / 2) + (y >> 1)) * (wf / 2) + (x >> 1)) * FP + c];
  }
}
__global__ void tex_split_kernel(const float* __restrict__ d_tex0, int BV, int Cf, int HW, int FP, float* __restrict__ d_feat) {
  // (BV, HW, FP) level-0 gradient -> (BV, Cf, HW) planar feature gradient
  int64_t n = (int64_t)BV * Cf * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int s = (int)(i % HW);
    int64_t t = i / HW;
    int c = (int)(t % Cf);
    int64_t bv = t / Cf;
    d_feat[i] = d_tex0[(bv * HW + s) * FP + c];
  }
}

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_render_fused_bwd(const float* rgba, const float* tex, const float* vol_cl, const float* depth_range,
                                    const float* vol_range, const float* cam, int cam_stride, const float* mlp, int B, int V,
                                    int H, int W, int bundle_size, int feat_dim, int D, int vol_stride, int max_samples,
                                    int max_mip_level, int inv_depth, int adaptive, const float* g_feat, const float* g_depth,
                                    const float* g_opacity, float* d_mlp, float* d_tex, float* d_vol, float* d_depth_range,
                                    float* d_vol_range, void* stream) {
  GDB_REQUIRE(rgba && tex && vol_cl && depth_range && vol_range && cam && mlp && g_feat && d_mlp && d_tex && d_vol &&
                  d_depth_range && d_vol_range, GDB_E_BADARG, "gdb_render_fused_bwd: null pointer");
  GDB_REQUIRE(B > 0 && H > 0 && W > 0 && D > 0 && bundle_size > 0 && H % bundle_size == 0 && W % bundle_size == 0, GDB_E_BADARG,
              "gdb_render_fused_bwd: bad size");
  GDB_REQUIRE(max_samples >= 1 && max_samples <= 32 && max_mip_level >= 0 && max_mip_level <= 3, GDB_E_BADARG,
              "gdb_render_fused_bwd: max_samples must be 1..32 and max_mip_level 0..3");
  GDB_REQUIRE(aligned16(rgba), GDB_E_ALIGN, "gdb_render_fused_bwd: rgba must be 16-byte aligned");
  GDB_REQUIRE(cam_stride == CAM_HEAD + CAM_VIEW * V, GDB_E_BADARG, "gdb_render_fused_bwd: cam_stride %d != %d", cam_stride,
              CAM_HEAD + CAM_VIEW * V);
  GDB_REQUIRE(vol_stride >= 8, GDB_E_BADARG, "gdb_render_fused_bwd: vol_stride %d < 8", vol_stride);
  RenderBwdParams q{};
  RenderParams& p = q.f;
  p.rgba = rgba; p.tex = tex; p.vol = vol_cl; p.depth_range = depth_range; p.vol_range = vol_range; p.cam = cam; p.mlp = mlp;
  p.cam_stride = cam_stride; p.vol_stride = vol_stride;
  p.B = B; p.H = H; p.W = W; p.Hb = H / bundle_size; p.Wb = W / bundle_size; p.D = D; p.max_samples = max_samples;
  p.L = max_mip_level; p.inv_depth = inv_depth; p.adaptive = adaptive;
  const int FPad = (feat_dim + 3 + 3) & ~3;
  p.tex_level[0] = 0;
  for (int k = 1; k <= 3; ++k) p.tex_level[k] = p.tex_level[k - 1] + (int64_t)B * V * (p.Hb >> (k - 1)) * (p.Wb >> (k - 1)) * FPad;
  q.g_feat = g_feat; q.g_depth = g_depth; q.g_opacity = g_opacity;
  q.d_mlp = d_mlp; q.d_tex = d_tex; q.d_vol = d_vol; q.d_depth_range = d_depth_range; q.d_vol_range = d_vol_range;
  cudaStream_t st = as_stream(stream);
#define GDB_R(BSZ, FD, VV) \
  if (bundle_size == BSZ && feat_dim == FD && V == VV) return launch_render_bwd<BSZ, FD, VV>(q, st);
  GDB_R(2, 16, 2) GDB_R(2, 16, 3) GDB_R(2, 16, 4) GDB_R(4, 32, 2) GDB_R(4, 32, 3) GDB_R(4, 32, 4)
#undef GDB_R
  return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_bwd: (bundle_size=%d, feat_dim=%d, V=%d) not instantiated", bundle_size, feat_dim, V);
}

extern "C" int gdb_prepare_sources_bwd(float* d_tex, int BV, int Cf, int Hb, int Wb, int max_mip_level, float* d_feat, void* stream) {
  GDB_REQUIRE(d_tex && d_feat && BV > 0 && Cf > 0 && Hb > 0 && Wb > 0 && max_mip_level >= 0 && max_mip_level <= 3, GDB_E_BADARG,
              "gdb_prepare_sources_bwd: bad argument");
  const int FP = (Cf + 3 + 3) & ~3;
  int64_t lvl[5] = {0, 0, 0, 0, 0};
  for (int k = 1; k <= 4; ++k) lvl[k] = lvl[k - 1] + (int64_t)BV * (Hb >> (k - 1)) * (Wb >> (k - 1)) * FP;
  cudaStream_t st = as_stream(stream);
  for (int k = max_mip_level; k >= 1; --k) {
    int hf = Hb >> (k - 1), wf = Wb >> (k - 1);
    int64_t n = (int64_t)BV * hf * wf * FP;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
    tex_pull_kernel<<<blocks, 256, 0, st>>>(d_tex, lvl[k - 1], lvl[k], BV, hf, wf, FP);
  }
  int64_t n = (int64_t)BV * Cf * Hb * Wb;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
  tex_split_kernel<<<blocks, 256, 0, st>>>(d_tex, BV, Cf, Hb * Wb, FP, d_feat);
  return cuda_check("gdb_prepare_sources_bwd");
}
