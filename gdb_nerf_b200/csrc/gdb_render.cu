// K3: fused per-bundle render - sample placement, multi-view projection,
// bilinear / trilinear / mip-mapped gathers, aggregation + radiance MLP and
// alpha compositing in one kernel (fp32 SIMT variant).
//
// Reference: BundleSampler.sample/encode (bundle_sampler.py:193-371),
// NeRF.forward (nerf.py:58-115), render_weight_from_density /
// accumulate_value_along_rays (utils.py:19-43,88-121),
// Network.render_bundles (network.py:54-91).
//
// Work decomposition: a warp owns G = 32/max_samples consecutive bundles; lane
// = (bundle-in-group, sample slot), so the <= max_samples samples of a bundle
// sit in adjacent lanes and compositing is a segmented warp shuffle.  Warps are
// independent (no CTA barrier after the weights are staged): while one warp
// waits on its gathers another runs its MLP.  The MLP weights (packed [K][N])
// live in shared memory once per CTA; each lane keeps its activations in a
// private shared-memory column ([row][lane], bank = lane: conflict-free).
#include <mutex>
#include "gdb_render_common.cuh"

namespace gdb {

template <int BS, int FEAT_DIM, int V>
struct RenderCfg {
  using ML = MlpLayout<FEAT_DIM>;
  static constexpr int BB = BS * BS;
  static constexpr int F = ML::F;
  static constexpr int FP = ML::FP;
  static constexpr int R = 3 * BB;
  static constexpr int CT = R + F + 8;            // bundle feature channels
  static constexpr int RFD = R + F + 4;           // row width of rgbs_feat_dir
  // per-warp scratch rows ([row][32 lanes])
  static constexpr int R_FR = 0;                  // V*F   gathered feature+rgb per view
  static constexpr int R_DIR = V * F;             // V*4   direction features per view
  static constexpr int R_VOX = V * (F + 4);       // 8
  static constexpr int R_IMG = R_VOX + 8;         // 16    (x_v overlays IMG|H until global_fc is done)
  static constexpr int R_H = R_IMG + 16;          // 64
  static constexpr int R_XV = R_IMG;
  static constexpr int XV_ROWS = (V * F > 80) ? V * F : 80;
  static constexpr int ROWS = R_IMG + XV_ROWS;
  static constexpr int MLP_BYTES = ((ML::TOTAL * 4 + 127) / 128) * 128;
  static constexpr int WARP_BYTES = ROWS * 32 * 4;
};

template <int BS, int FEAT_DIM, int V>
__global__ void __launch_bounds__(256, 1) render_fused_kernel(const RenderParams p) {
  using C = RenderCfg<BS, FEAT_DIM, V>;
  using ML = typename C::ML;
  constexpr int BB = C::BB, F = C::F, FP = C::FP, R = C::R, CT = C::CT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* wsm = reinterpret_cast<float*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* col = reinterpret_cast<float*>(smem_raw + C::MLP_BYTES + (size_t)warp * C::WARP_BYTES) + lane;
#define SM(row) col[(row) * 32]

  for (int i = threadIdx.x * 4; i < ML::TOTAL; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(wsm + i) = ldg4(p.mlp + i);
  __syncthreads();

  const int HW = p.Hb * p.Wb;
  const int NB = p.B * HW;
  const int ns = p.max_samples;
  const int G = 32 / ns;
  const int ngroups = (NB + G - 1) / G;
  const int bl = lane / ns, slot = lane - bl * ns;
  const int seg_base = bl * ns;                     // first lane of my bundle

  for (int grp = blockIdx.x * nwarps + warp; grp < ngroups; grp += gridDim.x * nwarps) {
    const int bundle = grp * G + bl;
    const bool has_bundle = bl < G && bundle < NB;
    const int bidx = has_bundle ? bundle : 0;
    const int b = bidx / HW, pix = bidx - b * HW;
    const int yb = pix / p.Wb, xb = pix - yb * p.Wb;
    const float* head = p.cam + (size_t)b * p.cam_stride;

    float nr = p.depth_range[(size_t)(b * 2 + 0) * HW + pix], fr = p.depth_range[(size_t)(b * 2 + 1) * HW + pix];
    float vn = p.vol_range[(size_t)(b * 2 + 0) * HW + pix], vf = p.vol_range[(size_t)(b * 2 + 1) * HW + pix];
    const int n = bundle_sample_count(nr, fr, head[CAM_MINIV], ns, p.inv_depth, p.adaptive);
    if (p.inv_depth) { nr = fdiv(1.f, nr); fr = fdiv(1.f, fr); vn = fdiv(1.f, vn); vf = fdiv(1.f, vf); }
    const bool active = has_bundle && slot < n;
    float z, dnorm;
    sample_depth(nr, fr, vn, vf, n, slot, p.inv_depth, z, dnorm);
    BundleGeom<BS> g;
    g.init(head, yb, xb, p.H, p.W);
    const float ox = head[CAM_O + 0], oy = head[CAM_O + 1], oz = head[CAM_O + 2];
    const int64_t srow = (p.offsets && active) ? (int64_t)p.offsets[bundle] + slot : -1;

    // ---- world-space sphere centre and radius (bundle_sampler.py:255-263)
    float cwx = 0.f, cwy = 0.f, cwz = 0.f;
#pragma unroll
    for (int j = 0; j < BB; ++j) {
      float dx, dy, dz;
      g.ray_dir(head, j, dx, dy, dz);
      cwx += fmaf(dx, z, ox); cwy += fmaf(dy, z, oy); cwz += fmaf(dz, z, oz);
    }
    cwx *= (1.f / BB); cwy *= (1.f / BB); cwz *= (1.f / BB);
    float ball;
    {
      float ex = cwx - ox, ey = cwy - oy, ez = cwz - oz;
      ball = sqrtf(ex * ex + ey * ey + ez * ez) * g.unit_ball;
    }

    // ---- voxel feature: trilinear, border, align_corners=False (bundle_sampler.py:322-324)
    if (active) {
      float ix = fminf(fmaxf(((g.u + 1.f) * (float)p.Wb - 1.f) * 0.5f, 0.f), (float)(p.Wb - 1));
      float iy = fminf(fmaxf(((g.v + 1.f) * (float)p.Hb - 1.f) * 0.5f, 0.f), (float)(p.Hb - 1));
      float iz = fminf(fmaxf(((dnorm + 1.f) * (float)p.D - 1.f) * 0.5f, 0.f), (float)(p.D - 1));
      float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
      float tx = ix - x0f, ty = iy - y0f, tz = iz - z0f;
      int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
      int x1 = min(x0 + 1, p.Wb - 1), y1 = min(y0 + 1, p.Hb - 1), z1 = min(z0 + 1, p.D - 1);
      const float* vb = p.vol + (size_t)b * p.vol_sb;
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int xx = (k & 1) ? x1 : x0, yy = (k & 2) ? y1 : y0, zz = (k & 4) ? z1 : z0;
        float w = ((k & 1) ? tx : 1.f - tx) * ((k & 2) ? ty : 1.f - ty) * ((k & 4) ? tz : 1.f - tz);
        const float* tp = vb + zz * p.vol_sz + yy * p.vol_sy + xx * p.vol_sx;
        lo = f4_scale_add(lo, ldg4(tp), w);
        hi = f4_scale_add(hi, ldg4(tp + 4), w);
      }
      SM(C::R_VOX + 0) = lo.x; SM(C::R_VOX + 1) = lo.y; SM(C::R_VOX + 2) = lo.z; SM(C::R_VOX + 3) = lo.w;
      SM(C::R_VOX + 4) = hi.x; SM(C::R_VOX + 5) = hi.y; SM(C::R_VOX + 6) = hi.z; SM(C::R_VOX + 7) = hi.w;
      if (p.tap_vox) {
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[0] = lo;
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[1] = hi;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) SM(C::R_VOX + k) = 0.f;
    }

    // ---- per view: sphere centre in camera space, level of detail, mip-mapped feature fetch, direction features
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      const float* cv = head + CAM_HEAD + CAM_VIEW * v;
      float ccx = 0.f, ccy = 0.f, ccz = 0.f;
#pragma unroll
      for (int j = 0; j < BB; ++j) {
        float dx, dy, dz;
        g.ray_dir(head, j, dx, dy, dz);
        float wx = fmaf(dx, z, ox), wy = fmaf(dy, z, oy), wz = fmaf(dz, z, oz);
        ccx += fmaf(wx, cv[CV_E + 0], fmaf(wy, cv[CV_E + 1], fmaf(wz, cv[CV_E + 2], cv[CV_E + 3])));
        ccy += fmaf(wx, cv[CV_E + 4], fmaf(wy, cv[CV_E + 5], fmaf(wz, cv[CV_E + 6], cv[CV_E + 7])));
        ccz += fmaf(wx, cv[CV_E + 8], fmaf(wy, cv[CV_E + 9], fmaf(wz, cv[CV_E + 10], cv[CV_E + 11])));
      }
      ccx *= (1.f / BB); ccy *= (1.f / BB); ccz *= (1.f / BB);
      float dist = sqrtf(ccx * ccx + ccy * ccy + ccz * ccz);
      float sec = dist / ccz;
      float sec_sq = sec * sec;
      float rb = dist / ball;
      float foot = sec_sq / (sqrtf(fmaxf(rb * rb - 1.f, 1e-12f)) + sqrtf(fmaxf(sec_sq - 1.f, 1e-12f)));
      float lod = log2f(foot / cv[CV_PIXR]);
      // centre projected with intrinsics / b, normalised to [0,1] (bundle_sampler.py:351-353)
      const float fb = (float)BS;
      float k0 = cv[CV_K + 0] / fb, k1 = cv[CV_K + 1] / fb, k2 = cv[CV_K + 2] / fb;
      float k3 = cv[CV_K + 3] / fb, k4 = cv[CV_K + 4] / fb, k5 = cv[CV_K + 5] / fb;
      float pxc = fmaf(ccx, k0, fmaf(ccy, k1, ccz * k2));
      float pyc = fmaf(ccx, k3, fmaf(ccy, k4, ccz * k5));
      float pzc = fmaxf(fmaf(ccx, cv[CV_K + 6], fmaf(ccy, cv[CV_K + 7], ccz * cv[CV_K + 8])), 1e-6f);
      float u01 = pxc / pzc / (float)p.Wb, v01 = pyc / pzc / (float)p.Hb;

      if (active) {
        float flod = fminf(fmaxf(lod, 0.f), (float)p.L);
        if (!(flod >= 0.f)) flod = 0.f;
        int l0 = (int)floorf(flod);
        int l1 = min(l0 + 1, p.L);
        float frac = flod - (float)l0;
        const bool tri = flod > 0.f;
        int w0 = p.Wb >> l0, h0 = p.Hb >> l0, w1 = p.Wb >> l1, h1 = p.Hb >> l1;
        TexTap ta = tex_tap(u01, v01, w0, h0);
        TexTap tb = tex_tap(u01, v01, w1, h1);
        const float* base0 = p.tex + p.tex_level[l0] + (size_t)(b * V + v) * h0 * w0 * FP;
        const float* base1 = p.tex + p.tex_level[l1] + (size_t)(b * V + v) * h1 * w1 * FP;
#pragma unroll
        for (int q = 0; q < FP / 4; ++q) {
          float4 a = bilerp4(ldg4(base0 + (size_t)ta.o00 * FP + q * 4), ldg4(base0 + (size_t)ta.o10 * FP + q * 4),
                             ldg4(base0 + (size_t)ta.o01 * FP + q * 4), ldg4(base0 + (size_t)ta.o11 * FP + q * 4), ta.fu, ta.fv);
          if (tri) {
            float4 bq = bilerp4(ldg4(base1 + (size_t)tb.o00 * FP + q * 4), ldg4(base1 + (size_t)tb.o10 * FP + q * 4),
                                ldg4(base1 + (size_t)tb.o01 * FP + q * 4), ldg4(base1 + (size_t)tb.o11 * FP + q * 4), tb.fu, tb.fv);
            a.x = lerpf(a.x, bq.x, frac); a.y = lerpf(a.y, bq.y, frac); a.z = lerpf(a.z, bq.z, frac); a.w = lerpf(a.w, bq.w, frac);
          }
          if (q * 4 + 0 < F) SM(C::R_FR + v * F + q * 4 + 0) = a.x;
          if (q * 4 + 1 < F) SM(C::R_FR + v * F + q * 4 + 1) = a.y;
          if (q * 4 + 2 < F) SM(C::R_FR + v * F + q * 4 + 2) = a.z;
          if (q * 4 + 3 < F) SM(C::R_FR + v * F + q * 4 + 3) = a.w;
        }
        // direction features (bundle_sampler.py:362-367)
        float tx = cwx - ox, ty = cwy - oy, tz = cwz - oz;
        unit3(tx, ty, tz);
        float sx = cwx - cv[CV_C + 0], sy = cwy - cv[CV_C + 1], sz = cwz - cv[CV_C + 2];
        unit3(sx, sy, sz);
        float ddx = tx - sx, ddy = ty - sy, ddz = tz - sz;
        unit3(ddx, ddy, ddz);
        float dot = tx * sx + ty * sy + tz * sz;
        SM(C::R_DIR + v * 4 + 0) = ddx; SM(C::R_DIR + v * 4 + 1) = ddy; SM(C::R_DIR + v * 4 + 2) = ddz; SM(C::R_DIR + v * 4 + 3) = dot;
        if (p.tap_rfd) {
          float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow) * C::RFD + R;
          for (int c = 0; c < F; ++c) tp[c] = SM(C::R_FR + v * F + c);
          tp[F + 0] = ddx; tp[F + 1] = ddy; tp[F + 2] = ddz; tp[F + 3] = dot;
        }
      } else {
        for (int c = 0; c < F; ++c) SM(C::R_FR + v * F + c) = 0.f;
        for (int c = 0; c < 4; ++c) SM(C::R_DIR + v * 4 + c) = 0.f;
      }
    }

    // ======================= MLP (nerf.py:58-115) =======================
    // -- view_fc + residual (nerf.py:69-71): x_v = featrgb_v + relu(W_view dir_v + b)
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
      float d0 = SM(C::R_DIR + v * 4 + 0), d1 = SM(C::R_DIR + v * 4 + 1), d2 = SM(C::R_DIR + v * 4 + 2), d3 = SM(C::R_DIR + v * 4 + 3);
#pragma unroll
      for (int q = 0; q < FP / 4; ++q) {
        float t[4];
        load_row<4>(t, wsm + ML::VIEW_B + q * 4);
        axpy_row<4>(t, wsm + ML::VIEW_W + 0 * FP + q * 4, d0);
        axpy_row<4>(t, wsm + ML::VIEW_W + 1 * FP + q * 4, d1);
        axpy_row<4>(t, wsm + ML::VIEW_W + 2 * FP + q * 4, d2);
        axpy_row<4>(t, wsm + ML::VIEW_W + 3 * FP + q * 4, d3);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (q * 4 + e < F) SM(C::R_XV + v * F + q * 4 + e) = SM(C::R_FR + v * F + q * 4 + e) + fmaxf(t[e], 0.f);
      }
    }
    // -- var/mean over views + global_fc (nerf.py:73-78); the [var|mean] slice is view independent
    float img[16];
    {
      float gsh[32], gv[V][32];
      load_row<32>(gsh, wsm + ML::GLOB_B);
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int k = 0; k < 32; ++k) gv[v][k] = 0.f;
#pragma unroll 1
      for (int c = 0; c < F; ++c) {
        float xs[V], mean = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) { xs[v] = SM(C::R_XV + v * F + c); mean += xs[v]; }
        mean *= (1.f / V);
        float var = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) { float t = xs[v] - mean; var = fmaf(t, t, var); }
        var *= (1.f / (V - 1));                                   // torch.var_mean is unbiased
        axpy_row<32>(gsh, wsm + ML::GLOB_W + (F + c) * 32, var);
        axpy_row<32>(gsh, wsm + ML::GLOB_W + (2 * F + c) * 32, mean);
#pragma unroll
        for (int v = 0; v < V; ++v) axpy_row<32>(gv[v], wsm + ML::GLOB_W + c * 32, xs[v]);
      }
      // -- agg_w_fc + softmax over views + weighted sum (nerf.py:79-80)
      float aw[V], amax = -1e30f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
#pragma unroll
        for (int k = 0; k < 32; ++k) gv[v][k] = fmaxf(gv[v][k] + gsh[k], 0.f);
        aw[v] = fmaxf(dot_row<32>(gv[v], wsm + ML::AGG_W) + wsm[ML::AGG_B], 0.f);
        amax = fmaxf(amax, aw[v]);
      }
      float asum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { aw[v] = expf(aw[v] - amax); asum += aw[v]; }
      float im[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) im[k] = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float a = aw[v] / asum;
#pragma unroll
        for (int k = 0; k < 32; ++k) im[k] = fmaf(gv[v][k], a, im[k]);
      }
      // -- fc (nerf.py:82)
      load_row<16>(img, wsm + ML::FC_B);
#pragma unroll
      for (int k = 0; k < 32; ++k) axpy_row<16>(img, wsm + ML::FC_W + k * 16, im[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) { img[k] = fmaxf(img[k], 0.f); SM(C::R_IMG + k) = img[k]; }

    // -- lr0, sigma, feat_head (nerf.py:100-102,112)
    float sigma, fh[8];
    {
      float h[64];
      load_row<64>(h, wsm + ML::LR0_B);
#pragma unroll 2
      for (int k = 0; k < 8; ++k) axpy_row<64>(h, wsm + ML::LR0_W + k * 64, SM(C::R_VOX + k));
#pragma unroll 2
      for (int k = 0; k < 16; ++k) axpy_row<64>(h, wsm + ML::LR0_W + (8 + k) * 64, SM(C::R_IMG + k));
      load_row<8>(fh, wsm + ML::FH_B);
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        h[k] = fmaxf(h[k], 0.f);
        SM(C::R_H + k) = h[k];
      }
      float s = dot_row<64>(h, wsm + ML::SIG_W) + wsm[ML::SIG_B];
      sigma = s > 20.f ? s : log1pf(expf(s));                     // nn.Softplus(beta=1, threshold=20)
#pragma unroll 4
      for (int k = 0; k < 64; ++k) axpy_row<8>(fh, wsm + ML::FH_W + k * 8, SM(C::R_H + k));
#pragma unroll
      for (int k = 0; k < 8; ++k) fh[k] = fmaxf(fh[k], 0.f);
    }

    // -- blending weights (nerf.py:106-109): weight.0 on [h|vox|img | featrgb_v|dir_v]; shared slice once
    float wv[V];
    {
      float U[64];
      load_row<64>(U, wsm + ML::W0_B);
#pragma unroll 2
      for (int k = 0; k < 64; ++k) axpy_row<64>(U, wsm + ML::W0_W + k * 64, SM(C::R_H + k));
#pragma unroll 2
      for (int k = 0; k < 8; ++k) axpy_row<64>(U, wsm + ML::W0_W + (64 + k) * 64, SM(C::R_VOX + k));
#pragma unroll 2
      for (int k = 0; k < 16; ++k) axpy_row<64>(U, wsm + ML::W0_W + (72 + k) * 64, SM(C::R_IMG + k));
      float wmax = -1e30f;
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        float hid[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) hid[k] = U[k];
#pragma unroll 2
        for (int k = 0; k < F + 4; ++k) axpy_row<64>(hid, wsm + ML::W0_W + (88 + k) * 64, SM(C::R_FR + (k < F ? v * F + k : V * F + v * 4 + (k - F))));
#pragma unroll
        for (int k = 0; k < 64; ++k) hid[k] = fmaxf(hid[k], 0.f);
        float s = fmaxf(dot_row<64>(hid, wsm + ML::W2_W) + wsm[ML::W2_B], 0.f);
        wv[v] = s;
        wmax = fmaxf(wmax, s);
      }
      float wsum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { wv[v] = expf(wv[v] - wmax); wsum += wv[v]; }
#pragma unroll
      for (int v = 0; v < V; ++v) wv[v] /= wsum;
    }

    // ======================= compositing weights (utils.py:19-43) =======================
    const unsigned full = 0xffffffffu;
    float alpha = active ? 1.f - expf(-sigma) : 0.f;
    float one_minus = 1.f - alpha;
    float T = 1.f;
    for (int k = 0; k + 1 < ns; ++k) {
      float o = __shfl_sync(full, one_minus, min(seg_base + k, 31));
      if (k < slot) T *= o;                                        // exclusive product in sample order
    }
    float wgt = alpha * T;
    float wtot = 0.f;
    for (int k = 0; k < ns; ++k) {
      float o = __shfl_sync(full, wgt, min(seg_base + k, 31));
      if (k < n) wtot += o;
    }
    wgt = active ? wgt / fmaxf(wtot, 1e-6f) : 0.f;
    if (p.tap_sigma && active) p.tap_sigma[srow] = sigma;
    if (p.tap_w && active) p.tap_w[srow] = wgt;

    // segmented sum over the lanes of my bundle, result valid in the slot-0 lane
    auto seg_sum = [&](float x) {
      float acc = x;
      for (int k = 1; k < ns; ++k) {
        float o = __shfl_down_sync(full, x, k);
        if (slot == 0 && k < n) acc += o;
      }
      return acc;
    };
    const bool writer = has_bundle && slot == 0;
    // planar: channel c at of[c * ostr]; channels-last: fine colours in out_feat (R per bundle), the rest in out_dec
    const size_t ostr = p.out_cl ? 1 : (size_t)HW;
    float* of = p.out_cl ? p.out_feat + (size_t)bidx * R : p.out_feat + (size_t)b * CT * HW + pix;
    float* od = p.out_cl ? p.out_dec + (size_t)bidx * p.dec_stride - R : of;
    if (writer && p.out_cl)
      for (int k = F + 8; k < p.dec_stride; ++k) od[R + k] = k == F + 8 ? p.dec_pad0 : 0.f;                                // pad channels of the decoder input
    float* tf = (p.tap_feat && active) ? p.tap_feat + srow * CT : nullptr;

    // -- fine colours: project every ray of the bundle into every view, bilinear on the full-res image,
    //    blend with the view weights (bundle_sampler.py:327-337, nerf.py:110) and composite
#pragma unroll 1
    for (int j = 0; j < BB; ++j) {
      float dx, dy, dz;
      g.ray_dir(head, j, dx, dy, dz);
      float wx = fmaf(dx, z, ox), wy = fmaf(dy, z, oy), wz = fmaf(dz, z, oz);
      float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float* cv = head + CAM_HEAD + CAM_VIEW * v;
        float cx = fmaf(wx, cv[CV_E + 0], fmaf(wy, cv[CV_E + 1], fmaf(wz, cv[CV_E + 2], cv[CV_E + 3])));
        float cy = fmaf(wx, cv[CV_E + 4], fmaf(wy, cv[CV_E + 5], fmaf(wz, cv[CV_E + 6], cv[CV_E + 7])));
        float cz = fmaf(wx, cv[CV_E + 8], fmaf(wy, cv[CV_E + 9], fmaf(wz, cv[CV_E + 10], cv[CV_E + 11])));
        float ix = fmaf(cx, cv[CV_K + 0], fmaf(cy, cv[CV_K + 1], cz * cv[CV_K + 2]));
        float iy = fmaf(cx, cv[CV_K + 3], fmaf(cy, cv[CV_K + 4], cz * cv[CV_K + 5]));
        float iz = fmaxf(fmaf(cx, cv[CV_K + 6], fmaf(cy, cv[CV_K + 7], cz * cv[CV_K + 8])), 1e-6f);
        float gx = 2.f * (ix / iz) / (float)p.W - 1.f, gy = 2.f * (iy / iz) / (float)p.H - 1.f;
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
          Bilin bl4 = bilin_border(gx, gy, p.W, p.H);
          const float* ib = p.rgba + (size_t)(b * V + v) * p.H * p.W * 4;
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o00 * 4), bl4.w00);
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o10 * 4), bl4.w10);
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o01 * 4), bl4.w01);
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o11 * 4), bl4.w11);
          if (p.tap_rfd) {
            float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow) * C::RFD;
            tp[0 * BB + j] = c4.x; tp[1 * BB + j] = c4.y; tp[2 * BB + j] = c4.z;
          }
        }
        cr = fmaf(c4.x, wv[v], cr); cg = fmaf(c4.y, wv[v], cg); cb = fmaf(c4.z, wv[v], cb);
      }
      if (tf) { tf[0 * BB + j] = cr; tf[1 * BB + j] = cg; tf[2 * BB + j] = cb; }
      float sr = seg_sum(wgt * cr), sg = seg_sum(wgt * cg), sb = seg_sum(wgt * cb);
      if (writer) {
        of[(size_t)(0 * BB + j) * ostr] = sr; of[(size_t)(1 * BB + j) * ostr] = sg; of[(size_t)(2 * BB + j) * ostr] = sb;
      }
    }
    // -- blended feature+rgb channels
#pragma unroll 1
    for (int c = 0; c < F; ++c) {
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) a = fmaf(SM(C::R_FR + v * F + c), wv[v], a);
      if (tf) tf[R + c] = a;
      float s = seg_sum(wgt * a);
      if (writer) od[(size_t)(R + c) * ostr] = s;
    }
    // -- geometry head channels
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (tf) tf[R + F + k] = fh[k];
      float s = seg_sum(wgt * fh[k]);
      if (writer) od[(size_t)(R + F + k) * ostr] = s;
    }
    // -- depth and opacity (network.py:83-89)
    {
      float zz = p.inv_depth ? fdiv(1.f, z) : z;
      float sd = seg_sum(wgt * zz), so = seg_sum(wgt);
      if (writer) {
        p.out_depth[(size_t)b * HW + pix] = p.inv_depth ? fdiv(1.f, sd) : sd;
        p.out_opacity[(size_t)b * HW + pix] = so;
      }
    }
    __syncwarp();
  }
#undef SM
}

template <int BS, int FEAT_DIM, int V>
static int launch_render(const RenderParams& p, cudaStream_t st) {
  using C = RenderCfg<BS, FEAT_DIM, V>;
  static int nwarps = 0;
  auto kern = render_fused_kernel<BS, FEAT_DIM, V>;
  if (nwarps == 0) {
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int nw = (max_smem - C::MLP_BYTES) / C::WARP_BYTES;
    nw = nw > 8 ? 8 : nw;
    if (nw < 1) return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_fwd: shared memory too small (%d B)", max_smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::MLP_BYTES + nw * C::WARP_BYTES);
    if (e != cudaSuccess) return fail((int)e, "gdb_render_fused_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    nwarps = nw;
  }
  const int G = 32 / p.max_samples;
  const long NB = (long)p.B * p.Hb * p.Wb;
  const long ngroups = (NB + G - 1) / G;
  long ctas = (ngroups + nwarps - 1) / nwarps;
  if (ctas > sm_count()) ctas = sm_count();
  size_t smem = C::MLP_BYTES + (size_t)nwarps * C::WARP_BYTES;
  kern<<<(int)ctas, nwarps * 32, smem, st>>>(p);
  return cuda_check("gdb_render_fused_fwd");
}

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_mlp_param_floats(int feat_dim) {
  switch (feat_dim) {
    case 8: return MlpLayout<8>::TOTAL;
    case 16: return MlpLayout<16>::TOTAL;
    case 32: return MlpLayout<32>::TOTAL;
    default: return GDB_E_UNSUPPORTED;
  }
}

extern "C" int gdb_render_fused_fwd(const float* rgba, const float* tex, const float* vol_cl, const float* depth_range,
                                    const float* vol_range, const float* cam, int cam_stride, const float* mlp, int B,
                                    int V, int H, int W, int bundle_size, int feat_dim, int D, int vol_stride, int vol_layout, int max_samples,
                                    int max_mip_level, int inv_depth, int adaptive, int precision, int out_channels_last,
                                    int dec_stride, float* out_feat, float* out_dec, float* out_depth, float* out_opacity,
                                    int row_lo, int row_hi, const gdb_render_taps* taps, void* stream) {
  GDB_REQUIRE(rgba && tex && vol_cl && depth_range && vol_range && cam && mlp && out_feat && out_depth && out_opacity,
              GDB_E_BADARG, "gdb_render_fused_fwd: null pointer");
  GDB_REQUIRE(B > 0 && H > 0 && W > 0 && D > 0, GDB_E_BADARG, "gdb_render_fused_fwd: bad size");
  GDB_REQUIRE(bundle_size > 0 && H % bundle_size == 0 && W % bundle_size == 0, GDB_E_BADARG,
              "gdb_render_fused_fwd: image %dx%d not divisible by bundle size %d", H, W, bundle_size);
  GDB_REQUIRE(max_samples >= 1 && max_samples <= 32, GDB_E_BADARG, "gdb_render_fused_fwd: max_samples must be 1..32");
  GDB_REQUIRE(max_mip_level >= 0 && max_mip_level <= 3, GDB_E_UNSUPPORTED, "gdb_render_fused_fwd: max_mip_level must be 0..3");
  GDB_REQUIRE(precision >= 0 && precision <= 6, GDB_E_UNSUPPORTED,
              "gdb_render_fused_fwd: precision %d not built (0 = fp32 SIMT, 1 = fp16-operand tcgen05 MLP, 2 = split-fp16 tcgen05 MLP, "
              "3 / 4 / 6 = first / second (round 1) / fourth generation of the fp16-operand kernel, kept for A/B; 5 = 1)", precision);
  GDB_REQUIRE(row_lo >= 0 && row_hi <= H / bundle_size && row_lo < row_hi, GDB_E_BADARG,
              "gdb_render_fused_fwd: row range [%d, %d) outside the %d bundle rows", row_lo, row_hi, H / bundle_size);
  GDB_REQUIRE((row_lo == 0 && row_hi == H / bundle_size) || precision == 1 || (precision >= 4 && precision <= 6), GDB_E_UNSUPPORTED,
              "gdb_render_fused_fwd: a partial row range needs precision 1, 4, 5 or 6 (got %d)", precision);
  GDB_REQUIRE(!((precision == 1 || precision == 2 || precision == 5 || precision == 6) && out_channels_last) || (aligned16(out_feat) && aligned16(out_dec)), GDB_E_ALIGN,
              "gdb_render_fused_fwd: channels-last outputs must be 16-byte aligned");
  GDB_REQUIRE(aligned16(rgba) && aligned16(tex) && aligned16(vol_cl) && aligned16(mlp), GDB_E_ALIGN,
              "gdb_render_fused_fwd: rgba/tex/vol/mlp must be 16-byte aligned");
  GDB_REQUIRE(cam_stride == CAM_HEAD + CAM_VIEW * V, GDB_E_BADARG, "gdb_render_fused_fwd: cam_stride %d != %d", cam_stride,
              CAM_HEAD + CAM_VIEW * V);
  RenderParams p{};
  p.rgba = rgba; p.tex = tex; p.vol = vol_cl; p.depth_range = depth_range; p.vol_range = vol_range; p.cam = cam; p.mlp = mlp;
  GDB_REQUIRE(!out_channels_last || out_dec, GDB_E_BADARG, "gdb_render_fused_fwd: channels-last output needs out_dec");
  const int dec_min = feat_dim + 3 + 8;
  GDB_REQUIRE(dec_stride == 0 || (dec_stride >= dec_min && dec_stride <= dec_min + 3), GDB_E_BADARG,
              "gdb_render_fused_fwd: dec_stride %d outside [%d, %d]", dec_stride, dec_min, dec_min + 3);
  p.dec_stride = dec_stride ? dec_stride : dec_min;
  GDB_REQUIRE(out_channels_last >= 0 && out_channels_last <= 2, GDB_E_BADARG, "gdb_render_fused_fwd: out_channels_last must be 0, 1 or 2");
  GDB_REQUIRE(out_channels_last != 2 || p.dec_stride > dec_min, GDB_E_BADARG,
              "gdb_render_fused_fwd: out_channels_last = 2 (constant-one channel) needs a pad channel (dec_stride > %d)", dec_min);
  p.dec_pad0 = out_channels_last == 2 ? 1.f : 0.f;
  p.out_feat = out_feat; p.out_dec = out_dec; p.out_depth = out_depth; p.out_opacity = out_opacity; p.out_cl = out_channels_last ? 1 : 0;
  if (taps && (taps->rgbs_feat_dir || taps->vox_feat || taps->sigma || taps->feat || taps->weights)) {
    GDB_REQUIRE(taps->offsets && taps->S_total > 0, GDB_E_BADARG, "gdb_render_fused_fwd: taps need offsets and S_total");
    GDB_REQUIRE(!taps->vox_feat || aligned16(taps->vox_feat), GDB_E_ALIGN, "gdb_render_fused_fwd: vox tap not aligned");
    p.offsets = taps->offsets; p.S_total = taps->S_total; p.tap_rfd = taps->rgbs_feat_dir; p.tap_vox = taps->vox_feat;
    p.tap_sigma = taps->sigma; p.tap_feat = taps->feat; p.tap_w = taps->weights;
  }
  GDB_REQUIRE(vol_stride >= 8 && vol_stride % 4 == 0, GDB_E_BADARG, "gdb_render_fused_fwd: vol_stride %d must be >= 8 and a multiple of 4", vol_stride);
  p.cam_stride = cam_stride;
  p.vol_stride = vol_stride;
  GDB_REQUIRE(vol_layout == 0 || vol_layout == 1, GDB_E_BADARG, "gdb_render_fused_fwd: vol_layout must be 0 (B,D,Hb,Wb,.) or 1 (B,Hb,Wb,D,.)");
  p.B = B; p.H = H; p.W = W; p.Hb = H / bundle_size; p.Wb = W / bundle_size; p.D = D; p.max_samples = max_samples;
  p.L = max_mip_level; p.inv_depth = inv_depth; p.adaptive = adaptive;
  p.pix_lo = row_lo * p.Wb; p.pix_hi = row_hi * p.Wb;
  if (vol_layout == 0) {
    p.vol_sx = vol_stride; p.vol_sy = (int64_t)vol_stride * p.Wb; p.vol_sz = p.vol_sy * p.Hb; p.vol_sb = p.vol_sz * D;
  } else {
    p.vol_sz = vol_stride; p.vol_sx = (int64_t)vol_stride * D; p.vol_sy = p.vol_sx * p.Wb; p.vol_sb = p.vol_sy * p.Hb;
  }
  const int m = 1 << max_mip_level;
  GDB_REQUIRE(p.Hb % m == 0 && p.Wb % m == 0, GDB_E_BADARG, "gdb_render_fused_fwd: bundle map %dx%d not divisible by %d", p.Hb, p.Wb, m);
  const int FPad = (feat_dim + 3 + 3) & ~3;
  p.tex_level[0] = 0;
  for (int k = 1; k <= 3; ++k) p.tex_level[k] = p.tex_level[k - 1] + (int64_t)B * V * (p.Hb >> (k - 1)) * (p.Wb >> (k - 1)) * FPad;
  cudaStream_t st = as_stream(stream);
  if (precision == 6) return render_tc3_dispatch(p, bundle_size, feat_dim, V, st);
  if (precision == 1 || precision == 4 || precision == 5) return render_tc2_dispatch(p, bundle_size, feat_dim, V, precision == 4 ? 2 : 3, st);
  if (precision == 2 && render_tc4_covers(p, bundle_size, feat_dim, V)) return render_tc4_dispatch(p, bundle_size, feat_dim, V, st);
  if (precision >= 2) return render_tc_dispatch(p, bundle_size, feat_dim, V, precision == 2, st);
#define GDB_R(BSZ, FD, VV) \
  if (bundle_size == BSZ && feat_dim == FD && V == VV) return launch_render<BSZ, FD, VV>(p, st);
  GDB_R(2, 16, 2) GDB_R(2, 16, 3) GDB_R(2, 16, 4) GDB_R(4, 32, 2) GDB_R(4, 32, 3) GDB_R(4, 32, 4)
#undef GDB_R
  return fail(GDB_E_UNSUPPORTED,
              "gdb_render_fused_fwd: (bundle_size=%d, feat_dim=%d, V=%d) not instantiated; built: (2,16,2..4), (4,32,2..4)",
              bundle_size, feat_dim, V);
}

// ---------------------------------------------------------------- tile counters --
// The persistent tensor-core kernels hand their tiles out dynamically: SMs do not run at one speed (measured: a static
// stride leaves 8 % of the SM cycles idle at the end of the DTU launch), an atomic counter evens the finish line out.
namespace gdb {
constexpr int TILE_SLOTS_STREAM = 64, TILE_SLOTS_CAPTURE = 960;
__device__ unsigned int g_tile_counters[TILE_SLOTS_STREAM + TILE_SLOTS_CAPTURE];

unsigned int* acquire_tile_counter(cudaStream_t st) {
  struct DevState {
    unsigned int* base = nullptr;
    cudaStream_t streams[TILE_SLOTS_STREAM];
    int n_streams = 0, n_capture = 0;
  };
  static std::mutex mu;
  static DevState devs[GDB_MAX_DEVICES];
  static int off = -1;
  if (off < 0) { const char* e = getenv("GDB_K3_STATIC"); off = (e && e[0] == '1') ? 1 : 0; }     // A/B: static tile assignment
  if (off || st == cudaStreamPerThread) return nullptr;     // the per-thread handle names a different stream in every host thread
  std::lock_guard<std::mutex> lock(mu);
  DevState& d = devs[current_device()];
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (!d.base) {
    // the symbol is resolved outside captures only (a first-use module lookup is not something to do inside a global-mode capture);
    // a launch captured before any eager one simply keeps the static stride
    if (cs != cudaStreamCaptureStatusNone) return nullptr;
    if (cudaGetSymbolAddress(reinterpret_cast<void**>(&d.base), g_tile_counters) != cudaSuccess) {
      cudaGetLastError();
      d.base = nullptr;
      return nullptr;
    }
  }
  int slot = -1;
  if (cs == cudaStreamCaptureStatusActive) {
    if (d.n_capture >= TILE_SLOTS_CAPTURE) return nullptr;
    slot = TILE_SLOTS_STREAM + d.n_capture++;
  } else {
    for (int i = 0; i < d.n_streams && slot < 0; ++i)
      if (d.streams[i] == st) slot = i;
    if (slot < 0) {
      if (d.n_streams >= TILE_SLOTS_STREAM) return nullptr;
      slot = d.n_streams;
      d.streams[d.n_streams++] = st;
    }
  }
  if (cudaMemsetAsync(d.base + slot, 0, sizeof(unsigned int), st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return d.base + slot;
}
}  // namespace gdb

extern "C" int gdb_abi_version(void) { return GDB_ABI_VERSION; }
extern "C" const char* gdb_last_error_string(void) { return err_buf(); }
