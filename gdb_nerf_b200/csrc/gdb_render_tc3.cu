// K3, fourth-generation tensor-core kernel (fp16 MLP operands, fp32 accumulation in TMEM): `precision = 6` of the C ABI, a
// MEASURED VARIANT - the third generation (gdb_render_tc2.cu) stays the default: this kernel issues 7 % fewer instructions and
// is 2-3 % faster at DTU 512x640 but 4-7 % slower at LLFF 640x960 and NeRF 800x800 4x4 (profiles/r02_k3_gen3_vs_gen4.log).
//
// Reference: bundle_sampler.py:193-371, nerf.py:58-115, utils.py:19-43,88-121.
//
// Same tile plan, operand aliasing, GEMM schedule and compositing as the third generation (gdb_render_tc2.cu, GEN = 3; see
// the header comment there and gdb_render_tc2.cuh).  What changed is the instruction diet of the two gather passes, which the
// ncu source page of generation 3 (profiles/r02_k3_phase_profile_gen3_dtu.txt, 605 M warp instructions per 8 DTU views)
// showed to be 52 % of all issued instructions, a third of them address arithmetic, predicate handling and the zero
// initialisation (CS2R) of predicated load destinations:
//
//  * P1 (thread = row) hands P2 (lane = (row, 16-byte quad of a texel)) the EIGHT FINAL TAP WEIGHTS and the eight 32-bit tap
//    indices of a (row, view) - the bilinear weights of both mip levels already multiplied by the tri-linear fraction -
//    through five 16-byte slots of the row's own X_v / FD_v operand rows.  A fetch lane then runs eight unconditional
//    LDG.128 and eight packed FMAs: no per-lane unpacking, no nested lerps, no branches on the tri-linear flag, no predicates.
//    Rows without a sample (and the level-1 taps of a bi-linear sample) carry zero weights and the address of a tap that is
//    fetched anyway, so they cost an L1 hit and nothing else.  (Weighted-sum form instead of nested lerps: results differ from
//    generation 3 in the last fp32 bit; both are within 1e-6 of the fp64 oracle on the operands, which are then rounded to fp16.)
//  * the colour pass and the voxel taps select a harmless address instead of predicating the load,
//  * reciprocals use the correctly rounded MUFU.RCP sequence (__frcp_rn) instead of the IEEE division (bit-identical),
//  * inactive rows are not masked to zero anywhere in the MLP operands: rows are independent in every GEMM, an inactive row's
//    operands are finite by construction and its compositing weight is exactly 0.
//
//  * operand chunks are CHB = 2048 + 64 / 80 bytes apart instead of 2048 (the UMMA leading-dimension offset is free): the quad lanes
//    of a fetch iteration write the same rows of consecutive chunks, which with a 2048-byte stride fall on the same banks
//    (ncu: 14 M of 75 M shared-memory wavefronts were bank-conflict replays, all of them these stores).
//
// Staging (north star: "TMA / shared-memory staging of source-feature tiles") was settled by measurement, not built: the
// MEMSRC = 1 instantiation serves EVERY gather of the kernel from shared memory (no global latency, no L1 tag stage, no DRAM,
// no staging cost - a bound no TMA scheme can beat) and runs 0.90 ms against 0.98 ms per 8 DTU views (NeRF 4x4: 1.94 vs 2.43,
// LLFF 1.67 vs 1.84; profiles/r02_k3_variants.log, profiles/r02_ncu_full_k3_g4_ldsbound_dtu.json): the L1 data pipe stays at
// 73 % because a shared-memory tap costs the same wavefronts as a global one, and the operand plan leaves 7.8 KB of the SM's
// 227 KB free where one tile's colour footprint alone is 3 x 17 KB.
#include <string>

#include "gdb_render_tc2.cuh"

namespace gdb {

// first term of a weighted tap sum
__device__ __forceinline__ float4 f4_scale(float4 v, float w) {
  const unsigned long long ww = pack2(w, w);
  unsigned long long lo, hi;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(lo) : "l"(pack2(v.x, v.y)), "l"(ww));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hi) : "l"(pack2(v.z, v.w)), "l"(ww));
  float4 r;
  unpack2(lo, r.x, r.y);
  unpack2(hi, r.z, r.w);
  return r;
}

// operand plan of generation 4: Tc2Cfg's regions with a padded chunk stride
template <int BS, int FEAT_DIM, int V, int NG>
struct Tc3Cfg : Tc2Cfg<BS, FEAT_DIM, V, NG> {
  using Base = Tc2Cfg<BS, FEAT_DIM, V, NG>;
  static constexpr int CHB = 2048 + (NG >= 4 ? 64 : 80);   // bytes between consecutive operand chunks (what the 227 KB allow)
  static constexpr int A_X = 0;
  static constexpr int A_FD = A_X + Base::CH_X * CHB;
  static constexpr int A_S = A_FD + Base::CH_FD * CHB;
  static constexpr int A_END = ((A_S + Base::CH_S * CHB + 127) / 128) * 128;
  static constexpr int CAM_OFF = A_END + 128;          // mbarrier at A_END
  static constexpr int GROUP_BYTES = CAM_OFF + ((CAM_HEAD + CAM_VIEW * V) * 4 + 127) / 128 * 128;
  static constexpr int ZERO_OFF = Base::GROUP_OFF + NG * GROUP_BYTES;
  static constexpr int ONE_OFF = ZERO_OFF + 2048;
  static constexpr int SMEM = ONE_OFF + 2048;
  static_assert(32 * Base::NCP * 4 <= 512 * (Base::CH_X + Base::CH_FD), "compositing stash must fit in the warp's rows of X + FD");
};

// `nch` consecutive chunks (CHB apart) starting at `a` (odd counts pair the last chunk with the zero chunk)
template <int CHB>
__device__ __forceinline__ void mma_chunks_p(uint32_t d_tmem, uint32_t a, int nch, uint32_t zero_chunk, uint32_t b_addr, int N,
                                             uint32_t accumulate) {
  for (int ks = 0; 2 * ks < nch; ++ks) {
    const uint32_t a0 = a + ks * 2 * CHB;
    const uint32_t a1 = (2 * ks + 1 < nch) ? a0 + CHB : zero_chunk;
    mma_step(d_tmem, a0, a1, b_addr + ks * 2 * (N * 16), N, (ks > 0 || accumulate) ? 1u : 0u);
  }
}

template <int BS, int FEAT_DIM, int V, int NG, bool TAPS, int PB, int MEMSRC = 0>
__global__ void __launch_bounds__(128 * NG, 1) render_tc3_kernel(const RenderParams p) {
  using C = Tc3Cfg<BS, FEAT_DIM, V, NG>;
  constexpr int BB = C::BB, F = C::F, FP = C::FP, R = C::R, CT = C::CT, QL = C::QL, IPW = C::IPW;
  constexpr int CHB = C::CHB, GROUP_BYTES = C::GROUP_BYTES, ZERO_OFF = C::ZERO_OFF, ONE_OFF = C::ONE_OFF;
  static_assert(C::XCH >= 3 && C::FDCH >= 2, "five 16-byte descriptor slots per (row, view)");
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  float* vec = reinterpret_cast<float*>(smem + C::VEC_OFF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 7;                      // group = tile slot
  const int row = tid & 127;                   // sample row == TMEM lane
  const int wq = warp & 3;
  unsigned char* gsm = smem + C::GROUP_OFF + (size_t)g * GROUP_BYTES;
  const uint32_t mbar = smem_u32(gsm + C::A_END);
  const unsigned full = 0xffffffffu;
  // MEMSRC = 1 (GDB_K3_VARIANT=ldsbound, TIMING ONLY - the results are meaningless): every gather is served from the first 16 KB
  // of shared memory instead of global memory.  It bounds from below what ANY staging scheme (TMA boxes or otherwise) that
  // turns the taps into shared-memory reads could reach: no global latency, no L1 tag stage, no DRAM traffic, and no staging cost.
  auto gather = [&](const float4* ptr) -> float4 {
    if constexpr (MEMSRC == 1) return *reinterpret_cast<const float4*>(smem + ((reinterpret_cast<uintptr_t>(ptr)) & 0x3FF0));
    else return __ldg(ptr);
  };

  // ---- one-time setup: weights -> smem (fp16 B operands + fp32 vectors), constant chunks, mbarriers, TMEM
  {
    tc2_stage_weights<C>(smem, vec, p.mlp, tid, blockDim.x);       // (its constant chunks land at this plan's ZERO_OFF / ONE_OFF)
    if (row == 0) mbar_init(mbar, 1);
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                   "r"(C::TALLOC)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_group = tmem_base_s + g * C::TC;                      // column offset of my group
  const uint32_t tmem_row = tmem_group + ((uint32_t)(wq * 32) << 16);       // my warp's lane quarter
  uint32_t parity = 0;

  const uint32_t w_base = smem_u32(smem);
  const uint32_t zero_chunk = w_base + ZERO_OFF, one_chunk = w_base + ONE_OFF;
  const uint32_t aX = smem_u32(gsm) + C::A_X, aFD = smem_u32(gsm) + C::A_FD, aS = smem_u32(gsm) + C::A_S;
  unsigned char* const sX = gsm + C::A_X;
  unsigned char* const sFD = gsm + C::A_FD;
  unsigned char* const sS = gsm + C::A_S;

  const int HW = p.Hb * p.Wb;
  const int ns = p.max_samples;
  const int G = 32 / ns;                                   // bundles per warp
  const int pix_lo = p.pix_lo, pix_hi = p.pix_hi;          // bundle range of every view this launch renders (image-tile split)
  const int tiles_pv = (pix_hi - pix_lo + 4 * G - 1) / (4 * G);   // tiles per target view (a tile = 4 warps x G bundles, one view)
  const int tiles = p.B * tiles_pv;
  // slot-major rows: consecutive lanes hold the SAME slot of ADJACENT bundles, so the rows a gather instruction covers read
  // neighbouring texels
  auto rowof = [&](int bb, int k) { return k * G + bb; };
  const int slot = lane / G;
  const int bl = lane - slot * G;
  const bool lane_ok = slot < ns;
  float* scam = reinterpret_cast<float*>(gsm + C::CAM_OFF);
  const float* head = scam;
  int cur_b = -1;

  // lane roles of the feature fetch: lane = (row-in-iteration, quad)
  const int gr = lane / QL, gq = lane - gr * QL;
  const bool glane = lane < IPW * QL;
  const bool last_quad = gq == QL - 1;
  const float4* tex4 = reinterpret_cast<const float4*>(p.tex);
  const float inv_Wb = 1.f / (float)p.Wb, inv_Hb = 1.f / (float)p.Hb, two_W = 2.f / (float)p.W, two_H = 2.f / (float)p.H;

  // (near, far, vol_near, vol_far) of my row's bundle in tile `t`
  auto load_ranges = [&](int t) -> float4 {
    const int tb = t / tiles_pv;
    const int praw = pix_lo + ((t - tb * tiles_pv) * 4 + wq) * G + bl;
    const int px = (lane_ok && praw < pix_hi) ? praw : pix_lo;
    const float* dr = p.depth_range + (size_t)(tb * 2) * HW + px;
    const float* vr = p.vol_range + (size_t)(tb * 2) * HW + px;
    return make_float4(__ldg(dr), __ldg(dr + HW), __ldg(vr), __ldg(vr + HW));
  };
  float4 rng_next = make_float4(1.f, 2.f, 1.f, 2.f);
  if ((int)(blockIdx.x * NG + g) < tiles) rng_next = load_ranges(blockIdx.x * NG + g);

#pragma unroll 1
  for (int tile = blockIdx.x * NG + g; tile < tiles; tile += gridDim.x * NG) {
    const int b = tile / tiles_pv;                         // uniform over the group
    if (b != cur_b) {                                      // stage this view's camera block
      group_sync(g);
      for (int i = row; i < CAM_HEAD + CAM_VIEW * V; i += 128) scam[i] = p.cam[(size_t)b * p.cam_stride + i];
      group_sync(g);
      cur_b = b;
    }
    const int pix_warp0 = pix_lo + ((tile - b * tiles_pv) * 4 + wq) * G;       // first bundle of my warp
    const int pix_raw = pix_warp0 + bl;
    const bool has_bundle = lane_ok && pix_raw < pix_hi;
    const int pix = has_bundle ? pix_raw : pix_lo;
    const int bidx = b * HW + pix;
    const int yb = pix / p.Wb, xb = pix - yb * p.Wb;

    // =========================== P0: sample placement (thread = row) ===========================
    const float4 rng = rng_next;
    float nr = rng.x, fr_ = rng.y, vn = rng.z, vf = rng.w;
    const int n = bundle_sample_count(nr, fr_, head[CAM_MINIV], ns, p.inv_depth, p.adaptive);
    if (p.inv_depth) { nr = fdiv(1.f, nr); fr_ = fdiv(1.f, fr_); vn = fdiv(1.f, vn); vf = fdiv(1.f, vf); }
    const bool active = has_bundle && slot < n;
    float z, dnorm;
    sample_depth(nr, fr_, vn, vf, n, slot, p.inv_depth, z, dnorm);
    BundleGeom<BS> geo;
    geo.init(head, yb, xb, p.H, p.W);
    const float ox = head[CAM_O + 0], oy = head[CAM_O + 1], oz = head[CAM_O + 2];
    const int64_t srow = (TAPS && p.offsets && active) ? (int64_t)p.offsets[bidx] + slot : -1;

    float cwx = 0.f, cwy = 0.f, cwz = 0.f;
#pragma unroll
    for (int j = 0; j < BB; ++j) {
      float dx, dy, dz;
      geo.ray_dir(head, j, dx, dy, dz);
      cwx += fmaf(dx, z, ox); cwy += fmaf(dy, z, oy); cwz += fmaf(dz, z, oz);
    }
    cwx *= (1.f / BB); cwy *= (1.f / BB); cwz *= (1.f / BB);
    float ball;
    {
      float ex = cwx - ox, ey = cwy - oy, ez = cwz - oz;
      ball = sqrtf(ex * ex + ey * ey + ez * ez) * geo.unit_ball;
    }

    // ---- voxel feature (bundle_sampler.py:322-324), kept as one packed fp16 chunk until region X is free
    uint4 voxh = make_uint4(0, 0, 0, 0);
    if (active) {
      float ix = fminf(fmaxf(((geo.u + 1.f) * (float)p.Wb - 1.f) * 0.5f, 0.f), (float)(p.Wb - 1));
      float iy = fminf(fmaxf(((geo.v + 1.f) * (float)p.Hb - 1.f) * 0.5f, 0.f), (float)(p.Hb - 1));
      float iz = fminf(fmaxf(((dnorm + 1.f) * (float)p.D - 1.f) * 0.5f, 0.f), (float)(p.D - 1));
      float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
      float tx = ix - x0f, ty = iy - y0f, tz = iz - z0f;
      int x0 = (int)x0f, y0 = (int)y0f, z0 = (int)z0f;
      int x1 = min(x0 + 1, p.Wb - 1), y1 = min(y0 + 1, p.Hb - 1), z1 = min(z0 + 1, p.D - 1);
      const float* vb_ = p.vol + (size_t)b * p.vol_sb;
      // a bundle centre sits on a voxel centre whenever the cost volume has the bundle map's resolution: the x / y fractions are
      // then exactly zero for most bundles and six of the eight taps carry weight 0; those read the address of tap 0 (an L1 hit)
      const float4* tp0 = reinterpret_cast<const float4*>(vb_ + z0 * p.vol_sz + y0 * p.vol_sy + x0 * p.vol_sx);
      float4 tl[8], th[8];
      float tw[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int xx = (k & 1) ? x1 : x0, yy = (k & 2) ? y1 : y0, zz = (k & 4) ? z1 : z0;
        tw[k] = ((k & 1) ? tx : 1.f - tx) * ((k & 2) ? ty : 1.f - ty) * ((k & 4) ? tz : 1.f - tz);
        const float4* tp = reinterpret_cast<const float4*>(vb_ + zz * p.vol_sz + yy * p.vol_sy + xx * p.vol_sx);
        tp = tw[k] != 0.f ? tp : tp0;
        tl[k] = gather(tp);
        th[k] = gather(tp + 1);
      }
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        lo = f4_scale_add(lo, tl[k], tw[k]);
        hi = f4_scale_add(hi, th[k], tw[k]);
      }
      voxh = make_uint4(pack_h2(lo.x, lo.y), pack_h2(lo.z, lo.w), pack_h2(hi.x, hi.y), pack_h2(hi.z, hi.w));
      if (TAPS && p.tap_vox) {
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[0] = lo;
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[1] = hi;
      }
    }

    // ====================== P1: per-view fetch descriptors (thread = row) ======================
    // eight float4 indices into the mip chain (tap (0,0), (1,0), (0,1), (1,1) of level l0, then of level l1) and their eight final
    // weights (bilinear weight x tri-linear share of the level); the direction features
    float tdx = cwx - ox, tdy = cwy - oy, tdz = cwz - oz;      // unit vector target camera -> sample (view independent)
    unit3_fast(tdx, tdy, tdz);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float* cv = head + CAM_HEAD + CAM_VIEW * v;
      // centre of the bundle's points in the source camera frame (bundle_sampler.py:340; the mean commutes with the rigid map)
      const float ccx = fmaf(cwx, cv[CV_E + 0], fmaf(cwy, cv[CV_E + 1], fmaf(cwz, cv[CV_E + 2], cv[CV_E + 3])));
      const float ccy = fmaf(cwx, cv[CV_E + 4], fmaf(cwy, cv[CV_E + 5], fmaf(cwz, cv[CV_E + 6], cv[CV_E + 7])));
      const float ccz = fmaf(cwx, cv[CV_E + 8], fmaf(cwy, cv[CV_E + 9], fmaf(cwz, cv[CV_E + 10], cv[CV_E + 11])));
      // mip level (:343-348): only its fractional part reaches the output, approximate division is ample
      const float dist = sqrtf(ccx * ccx + ccy * ccy + ccz * ccz);
      const float sec = __fdividef(dist, ccz);
      const float sec_sq = sec * sec;
      const float rb = __fdividef(dist, ball);
      const float foot = __fdividef(sec_sq, sqrtf(fmaxf(rb * rb - 1.f, 1e-12f)) + sqrtf(fmaxf(sec_sq - 1.f, 1e-12f)));
      const float lod = log2f(__fdividef(foot, cv[CV_PIXR]));
      constexpr float ifb = 1.f / (float)BS;                  // power of two: exact
      const float pxc = fmaf(ccx, cv[CV_K + 0] * ifb, fmaf(ccy, cv[CV_K + 1] * ifb, ccz * (cv[CV_K + 2] * ifb)));
      const float pyc = fmaf(ccx, cv[CV_K + 3] * ifb, fmaf(ccy, cv[CV_K + 4] * ifb, ccz * (cv[CV_K + 5] * ifb)));
      const float pzc = fmaxf(fmaf(ccx, cv[CV_K + 6], fmaf(ccy, cv[CV_K + 7], ccz * cv[CV_K + 8])), 1e-6f);
      const float rz = __frcp_rn(pzc);
      const float u01 = pxc * rz * inv_Wb, v01 = pyc * rz * inv_Hb;
      uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
      float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0, dr = w0;
      if (active) {
        float flod = fminf(fmaxf(lod, 0.f), (float)p.L);
        if (!(flod >= 0.f)) flod = 0.f;
        const int l0 = (int)floorf(flod);
        const bool tri = flod > 0.f;
        const int wl0 = p.Wb >> l0, hl0 = p.Hb >> l0;
        const TexTap ta = tex_tap(u01, v01, wl0, hl0);
        const uint32_t base0 = (uint32_t)(p.tex_level[l0] >> 2) + (uint32_t)((b * V + v) * hl0 * wl0) * QL;
        o0 = make_uint4(base0 + ta.o00 * QL, base0 + ta.o10 * QL, base0 + ta.o01 * QL, base0 + ta.o11 * QL);
        const float frac = tri ? flod - (float)l0 : 0.f;
        {
          const float au = 1.f - ta.fu, av = (1.f - ta.fv) * (1.f - frac), bv = ta.fv * (1.f - frac);
          w0 = make_float4(au * av, ta.fu * av, au * bv, ta.fu * bv);
        }
        o1 = o0;
        if (tri) {
          const int l1 = min(l0 + 1, p.L);
          const int wl1 = p.Wb >> l1, hl1 = p.Hb >> l1;
          const TexTap tb = tex_tap(u01, v01, wl1, hl1);
          const uint32_t base1 = (uint32_t)(p.tex_level[l1] >> 2) + (uint32_t)((b * V + v) * hl1 * wl1) * QL;
          o1 = make_uint4(base1 + tb.o00 * QL, base1 + tb.o10 * QL, base1 + tb.o01 * QL, base1 + tb.o11 * QL);
          const float au = 1.f - tb.fu, av = (1.f - tb.fv) * frac, bv = tb.fv * frac;
          w1 = make_float4(au * av, tb.fu * av, au * bv, tb.fu * bv);
        }
        float sx = cwx - cv[CV_C + 0], sy = cwy - cv[CV_C + 1], sz = cwz - cv[CV_C + 2];
        unit3_fast(sx, sy, sz);
        float ddx = tdx - sx, ddy = tdy - sy, ddz = tdz - sz;
        unit3_fast(ddx, ddy, ddz);
        dr = make_float4(ddx, ddy, ddz, tdx * sx + tdy * sy + tdz * sz);
        if (TAPS && p.tap_rfd) {
          float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow) * C::RFD + R + F;
          tp[0] = dr.x; tp[1] = dr.y; tp[2] = dr.z; tp[3] = dr.w;
        }
      }
      // the descriptor of (row, view) travels through the row's own 16-byte slots of the X_v (3) and FD_v (2) operand regions:
      // the fetch lanes of the row read it there and overwrite it with the operands afterwards
      unsigned char* dx_ = sX + (v * C::XCH) * CHB + row * 16;
      unsigned char* df_ = sFD + (v * C::FDCH) * CHB + row * 16;
      *reinterpret_cast<uint4*>(dx_) = o0;
      *reinterpret_cast<uint4*>(dx_ + CHB) = o1;
      *reinterpret_cast<float4*>(dx_ + 2 * CHB) = w0;
      *reinterpret_cast<float4*>(df_) = w1;
      *reinterpret_cast<float4*>(df_ + CHB) = dr;
    }
    {
      // the depth ranges of my next tile, in flight underneath this tile
      const int tn = tile + gridDim.x * NG;
      if (tn < tiles) rng_next = load_ranges(tn);
    }

    // ================= P2: mip-mapped feature fetch, lane = (row, quad) =================
    // writes FD_v = [featrgb_v | dir_v], X_v = [x_v | 1] (nerf.py:69-71) and S = [var | mean] over views (nerf.py:73)
    {
      __syncwarp();                       // the descriptors of my warp's 32 rows are in shared memory
      // view_fc weights of my quad's four channels, loaded once per tile
      float4 vw0, vw1, vw2, vw3, vbq;
      {
        const float* vq = vec + (glane ? gq : 0) * 4;
        vw0 = *reinterpret_cast<const float4*>(vq + C::X_VIEW_W + 0 * FP);
        vw1 = *reinterpret_cast<const float4*>(vq + C::X_VIEW_W + 1 * FP);
        vw2 = *reinterpret_cast<const float4*>(vq + C::X_VIEW_W + 2 * FP);
        vw3 = *reinterpret_cast<const float4*>(vq + C::X_VIEW_W + 3 * FP);
        vbq = *reinterpret_cast<const float4*>(vq + C::X_VIEW_B);
      }
      const float vw[4][4] = {{vw0.x, vw0.y, vw0.z, vw0.w}, {vw1.x, vw1.y, vw1.z, vw1.w}, {vw2.x, vw2.y, vw2.z, vw2.w}, {vw3.x, vw3.y, vw3.z, vw3.w}};
      const float vb[4] = {vbq.x, vbq.y, vbq.z, vbq.w};
      const float4* tq = tex4 + (glane ? gq : 0);
      // rows >= ns * G of a warp never hold a sample: the fetch iterations that would cover only such rows are skipped (see
      // gdb_render_tc2.cu)
      const int nit = min(C::NIT, (ns * G + IPW - 1) / IPW);
#pragma unroll 1
      for (int it = 0; it < nit; ++it) {
        const int src_raw = it * IPW + gr;
        const bool ok = glane && src_raw < 32;
        const int src = min(src_raw, 31);
        const int orow16 = (wq * 32 + src) * 16;
        const int64_t srow_g = TAPS ? __shfl_sync(full, srow, src) : 0;
        const bool act_g = TAPS ? (__shfl_sync(full, active ? 1 : 0, src) != 0) : true;
        const unsigned char* dsx = sX + orow16;
        const unsigned char* dsf = sFD + orow16;
        float xq[V][4];
        // PB = 1: the eight taps of one view in flight at a time; PB = V: the 8 V taps of the iteration in flight at once
        float4 fb[V];
        if constexpr (PB > 1) {
          float4 t[V][8];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const uint4 o0 = *reinterpret_cast<const uint4*>(dsx + (v * C::XCH) * CHB);
            const uint4 o1 = *reinterpret_cast<const uint4*>(dsx + (v * C::XCH + 1) * CHB);
            t[v][0] = gather(tq + o0.x); t[v][1] = gather(tq + o0.y); t[v][2] = gather(tq + o0.z); t[v][3] = gather(tq + o0.w);
            t[v][4] = gather(tq + o1.x); t[v][5] = gather(tq + o1.y); t[v][6] = gather(tq + o1.z); t[v][7] = gather(tq + o1.w);
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 w0 = *reinterpret_cast<const float4*>(dsx + (v * C::XCH + 2) * CHB);
            const float4 w1 = *reinterpret_cast<const float4*>(dsf + (v * C::FDCH) * CHB);
            float4 f = f4_scale(t[v][0], w0.x);
            f = f4_scale_add(f, t[v][1], w0.y);
            f = f4_scale_add(f, t[v][2], w0.z);
            f = f4_scale_add(f, t[v][3], w0.w);
            f = f4_scale_add(f, t[v][4], w1.x);
            f = f4_scale_add(f, t[v][5], w1.y);
            f = f4_scale_add(f, t[v][6], w1.z);
            fb[v] = f4_scale_add(f, t[v][7], w1.w);
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4 f;
          if constexpr (PB > 1) {
            f = fb[v];
          } else {
            const uint4 o0 = *reinterpret_cast<const uint4*>(dsx + (v * C::XCH) * CHB);
            const uint4 o1 = *reinterpret_cast<const uint4*>(dsx + (v * C::XCH + 1) * CHB);
            // unconditional taps in flight.  The second mip level is skipped when NO row of this instruction blends two levels
            // (a warp-uniform branch; P1 gave such rows o1 = o0 and zero weights, and the level is a smooth function of depth and
            // position, so neighbouring bundles mostly agree)
            const bool two_levels = __any_sync(full, o1.x != o0.x);
            const float4 t0 = gather(tq + o0.x), t1 = gather(tq + o0.y), t2 = gather(tq + o0.z), t3 = gather(tq + o0.w);
            const float4 w0 = *reinterpret_cast<const float4*>(dsx + (v * C::XCH + 2) * CHB);
            if (two_levels) {
              const float4 t4 = gather(tq + o1.x), t5 = gather(tq + o1.y), t6 = gather(tq + o1.z), t7 = gather(tq + o1.w);
              const float4 w1 = *reinterpret_cast<const float4*>(dsf + (v * C::FDCH) * CHB);
              f = f4_scale(t0, w0.x);
              f = f4_scale_add(f, t1, w0.y);
              f = f4_scale_add(f, t2, w0.z);
              f = f4_scale_add(f, t3, w0.w);
              f = f4_scale_add(f, t4, w1.x);
              f = f4_scale_add(f, t5, w1.y);
              f = f4_scale_add(f, t6, w1.z);
              f = f4_scale_add(f, t7, w1.w);
            } else {
              f = f4_scale(t0, w0.x);
              f = f4_scale_add(f, t1, w0.y);
              f = f4_scale_add(f, t2, w0.z);
              f = f4_scale_add(f, t3, w0.w);
            }
          }
          const float4 q2 = *reinterpret_cast<const float4*>(dsf + (v * C::FDCH + 1) * CHB);
          const float dir[4] = {q2.x, q2.y, q2.z, q2.w};
          if (TAPS && p.tap_rfd && ok && act_g) {
            float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow_g) * C::RFD + R + gq * 4;
            tp[0] = f.x; tp[1] = f.y; tp[2] = f.z;
            if (!last_quad) tp[3] = f.w;
          }
          const float fe[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float t = vb[e];
            t = fmaf(vw[0][e], dir[0], t);
            t = fmaf(vw[1][e], dir[1], t);
            t = fmaf(vw[2][e], dir[2], t);
            t = fmaf(vw[3][e], dir[3], t);
            xq[v][e] = fe[e] + fmaxf(t, 0.f);
          }
          __syncwarp();                     // every lane of the row has read the descriptor that the operands now replace
          if (ok) {
            unsigned char* fdp = sFD + (v * C::FDCH + (gq >> 1)) * CHB + orow16;
            unsigned char* xp = sX + (v * C::XCH + (gq >> 1)) * CHB + orow16;
            if (last_quad) {
              // featrgb's pad channel is K slot F: dir_v follows in FD, the constant one in X
              *reinterpret_cast<uint4*>(fdp) = make_uint4(pack_h2(fe[0], fe[1]), pack_h2(fe[2], dir[0]), pack_h2(dir[1], dir[2]), pack_h2(dir[3], 0.f));
              *reinterpret_cast<uint4*>(xp) = make_uint4(pack_h2(xq[v][0], xq[v][1]), pack_h2(xq[v][2], 1.f), 0u, 0u);
              xq[v][3] = 0.f;
            } else {
              *reinterpret_cast<uint2*>(fdp + (gq & 1) * 8) = make_uint2(pack_h2(fe[0], fe[1]), pack_h2(fe[2], fe[3]));
              *reinterpret_cast<uint2*>(xp + (gq & 1) * 8) = make_uint2(pack_h2(xq[v][0], xq[v][1]), pack_h2(xq[v][2], xq[v][3]));
            }
          }
        }
        if (ok) {
          float var[4], mean[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float mu = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) mu += xq[v][e];
            mu *= (1.f / V);
            float s2 = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) { float t = xq[v][e] - mu; s2 = fmaf(t, t, s2); }
            var[e] = s2 * (1.f / (V - 1));
            mean[e] = mu;
          }
          const int kv = gq, km = QL + gq;      // quad positions of var / mean inside S
          *reinterpret_cast<uint2*>(sS + (kv >> 1) * CHB + orow16 + (kv & 1) * 8) = make_uint2(pack_h2(var[0], var[1]), pack_h2(var[2], var[3]));
          *reinterpret_cast<uint2*>(sS + (km >> 1) * CHB + orow16 + (km & 1) * 8) = make_uint2(pack_h2(mean[0], mean[1]), pack_h2(mean[2], mean[3]));
        }
      }
    }

    // ================= GEMM 1: global_fc, G_v = [var|mean] W_gs + [x_v|1] W_gx =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        mma_chunks_p<CHB>(tmem_group + v * 32, aS, C::SCH, zero_chunk, w_base + C::W_GS, 32, 0);
        mma_chunks_p<CHB>(tmem_group + v * 32, aX + v * C::XCH * CHB, C::XCH, zero_chunk, w_base + C::W_GX, 32, 1);
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
    {
      // pass 1: aggregation logits (re-reading TMEM keeps only one view's 32 columns live)
      float aw[V];
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        float gv[32];
        tmem_ld32(tmem_row + v * 32, gv);
        float s0 = vec[C::X_SCAL + 0], s1 = 0.f, s2_ = 0.f, s3 = 0.f;     // four independent chains (FMA latency)
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          s0 = fmaf(fmaxf(gv[k + 0], 0.f), vec[C::X_AGG_W + k + 0], s0);
          s1 = fmaf(fmaxf(gv[k + 1], 0.f), vec[C::X_AGG_W + k + 1], s1);
          s2_ = fmaf(fmaxf(gv[k + 2], 0.f), vec[C::X_AGG_W + k + 2], s2_);
          s3 = fmaf(fmaxf(gv[k + 3], 0.f), vec[C::X_AGG_W + k + 3], s3);
        }
        float s = fmaxf((s0 + s1) + (s2_ + s3), 0.f);
#pragma unroll
        for (int u = 0; u < V; ++u)
          if (u == v) aw[u] = s;
      }
      float amax = aw[0];
#pragma unroll
      for (int v = 1; v < V; ++v) amax = fmaxf(amax, aw[v]);
      float asum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { aw[v] = expf(aw[v] - amax); asum += aw[v]; }
      const float rsum = 1.f / asum;
      // pass 2: softmax-weighted sum over views
      float im[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) im[k] = 0.f;
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        float gv[32];
        tmem_ld32(tmem_row + v * 32, gv);
        float a = aw[0];
#pragma unroll
        for (int u = 1; u < V; ++u)
          if (u == v) a = aw[u];
        a *= rsum;
#pragma unroll
        for (int k = 0; k < 32; ++k) im[k] = fmaf(fmaxf(gv[k], 0.f), a, im[k]);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(sS + ch * CHB + row * 16) = make_uint4(pack_h2(im[ch * 8 + 0], im[ch * 8 + 1]), pack_h2(im[ch * 8 + 2], im[ch * 8 + 3]),
                                                                          pack_h2(im[ch * 8 + 4], im[ch * 8 + 5]), pack_h2(im[ch * 8 + 6], im[ch * 8 + 7]));
    }
    // ================= GEMM 2: fc =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      mma_chunks_p<CHB>(tmem_group, aS, 4, zero_chunk, w_base + C::W_FC, 16, 0);
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
    {
      float img[16];
      tmem_ld16(tmem_row, img);
#pragma unroll
      for (int k = 0; k < 16; ++k) img[k] = fmaxf(img[k] + vec[C::X_FC_B + k], 0.f);
      // X[8] <- vox, S[0..1] <- img (their previous contents were consumed by GEMMs 1 and 2)
      *reinterpret_cast<uint4*>(sX + 8 * CHB + row * 16) = voxh;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
        *reinterpret_cast<uint4*>(sS + ch * CHB + row * 16) = make_uint4(pack_h2(img[ch * 8 + 0], img[ch * 8 + 1]), pack_h2(img[ch * 8 + 2], img[ch * 8 + 3]),
                                                                          pack_h2(img[ch * 8 + 4], img[ch * 8 + 5]), pack_h2(img[ch * 8 + 6], img[ch * 8 + 7]));
    }
    // ================= GEMM 3: lr0 on [vox | img | 1] =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      mma_step(tmem_group, aX + 8 * CHB, aS, w_base + C::W_LR0, 64, 0);
      mma_step(tmem_group, aS + CHB, one_chunk, w_base + C::W_LR0 + 2 * 64 * 16, 64, 1);
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float h[32];
      tmem_ld32(tmem_row + half * 32, h);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(sX + (half * 4 + ch) * CHB + row * 16) =
            make_uint4(pack_h2(fmaxf(h[ch * 8 + 0], 0.f), fmaxf(h[ch * 8 + 1], 0.f)), pack_h2(fmaxf(h[ch * 8 + 2], 0.f), fmaxf(h[ch * 8 + 3], 0.f)),
                       pack_h2(fmaxf(h[ch * 8 + 4], 0.f), fmaxf(h[ch * 8 + 5], 0.f)), pack_h2(fmaxf(h[ch * 8 + 6], 0.f), fmaxf(h[ch * 8 + 7], 0.f)));
    }
    // ================= GEMM 4: [sigma | feat_head] and weight.0, view by view through NB 64-column buffers =================
    float sigma = 0.f, fh[8], wv[V];
#pragma unroll
    for (int r = 0; r < C::ROUNDS; ++r) {
      const int v0 = C::round_start(r), nv = C::round_n(r);
      tc_fence_before();
      fence_async_smem();
      group_sync(g);
      if (row == 0) {
        tc_fence_after();
        if (r == 0) mma_chunks_p<CHB>(tmem_group + (C::NB - 1) * 64, aX, 8, zero_chunk, w_base + C::W_SH, 16, 0);
#pragma unroll 1
        for (int i = 0; i < nv; ++i) {
          const uint32_t d = tmem_group + i * 64;
          mma_chunks_p<CHB>(d, aX, 8, zero_chunk, w_base + C::W_0S, 64, 0);                                   // h
          mma_step(d, aX + 8 * CHB, aS, w_base + C::W_0S + 8 * 64 * 16, 64, 1);                        // vox | img[0:8]
          mma_step(d, aS + CHB, one_chunk, w_base + C::W_0S + 10 * 64 * 16, 64, 1);                    // img[8:16] | 1
          mma_chunks_p<CHB>(d, aFD + (v0 + i) * C::FDCH * CHB, C::FDCH, zero_chunk, w_base + C::W_0V, 64, 1); // featrgb_v | dir_v
        }
        umma_commit(mbar);
      }
      mbar_wait(mbar, parity); parity ^= 1;
      tc_fence_after();
      if (r == 0) {
        float sh[16];
        tmem_ld16(tmem_row + (C::NB - 1) * 64, sh);
        float s = sh[0] + vec[C::X_SCAL + 1];
        sigma = s > 20.f ? s : log1pf(expf(s));
#pragma unroll
        for (int k = 0; k < 8; ++k) fh[k] = fmaxf(sh[1 + k] + vec[C::X_FH_B + k], 0.f);
      }
#pragma unroll 1
      for (int i = 0; i < nv; ++i) {
        float q0 = vec[C::X_SCAL + 2], q1 = 0.f, q2 = 0.f, q3 = 0.f;        // four independent chains (FMA latency)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float hid[32];
          tmem_ld32(tmem_row + i * 64 + half * 32, hid);
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            q0 = fmaf(fmaxf(hid[k + 0], 0.f), vec[C::X_W2_W + half * 32 + k + 0], q0);
            q1 = fmaf(fmaxf(hid[k + 1], 0.f), vec[C::X_W2_W + half * 32 + k + 1], q1);
            q2 = fmaf(fmaxf(hid[k + 2], 0.f), vec[C::X_W2_W + half * 32 + k + 2], q2);
            q3 = fmaf(fmaxf(hid[k + 3], 0.f), vec[C::X_W2_W + half * 32 + k + 3], q3);
          }
        }
        const float s2 = fmaxf((q0 + q1) + (q2 + q3), 0.f);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (v == v0 + i) wv[v] = s2;
      }
    }
    {
      float wmax = -1e30f;
#pragma unroll
      for (int v = 0; v < V; ++v) wmax = fmaxf(wmax, wv[v]);
      float wsum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { wv[v] = expf(wv[v] - wmax); wsum += wv[v]; }
      const float rws = 1.f / wsum;
#pragma unroll
      for (int v = 0; v < V; ++v) wv[v] *= rws;
    }
    tc_fence_before();      // TMEM reads of this tile are ordered before the barrier that precedes the next tile's MMAs

    // ======================= compositing weights (utils.py:19-43) =======================
    float alpha = active ? 1.f - expf(-sigma) : 0.f;
    float one_minus = 1.f - alpha;
    float T = 1.f;
    for (int k = 0; k + 1 < ns; ++k) {
      float o = __shfl_sync(full, one_minus, min(rowof(bl, k), 31));
      if (k < slot) T *= o;
    }
    float wgt = alpha * T;
    float wtot = 0.f;
    for (int k = 0; k < ns; ++k) {
      float o = __shfl_sync(full, wgt, min(rowof(bl, k), 31));
      if (k < n) wtot += o;
    }
    wgt = active ? wgt / fmaxf(wtot, 1e-6f) : 0.f;
    if (TAPS && p.tap_sigma && active) p.tap_sigma[srow] = sigma;
    if (TAPS && p.tap_w && active) p.tap_w[srow] = wgt;

    // compositing / colour stash: the warp's own rows of the dead operand regions X and FD
    constexpr int STASH0 = C::A_X;
    auto stash_f = [&](int t) { return reinterpret_cast<float4*>(gsm + STASH0 + ((t * 4) >> 9) * CHB + wq * 512 + ((t * 4) & 511)); };
    float* tf = (TAPS && p.tap_feat && active) ? p.tap_feat + srow * CT : nullptr;

    // ---- blended features sum_v w_v featrgb_v (featrgb read back from the FD operand), geometry head, depth, opacity:
    //      weighted by the compositing weight and transposed through shared memory (the warp's own rows),
    //      then summed over a bundle's samples in slot order with lane = (bundle, channel quad) and stored coalesced
    {
      float vals[C::NCP];
#pragma unroll
      for (int c = 0; c < C::NCP; ++c) vals[c] = 0.f;
#pragma unroll
      for (int ch = 0; ch < C::FDCH; ++ch) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const uint4 q = *reinterpret_cast<const uint4*>(sFD + (v * C::FDCH + ch) * CHB + row * 16);
          const float2 f0 = h2_to_f2(q.x), f1 = h2_to_f2(q.y), f2 = h2_to_f2(q.z), f3 = h2_to_f2(q.w);
          acc[0] = fmaf(f0.x, wv[v], acc[0]); acc[1] = fmaf(f0.y, wv[v], acc[1]);
          acc[2] = fmaf(f1.x, wv[v], acc[2]); acc[3] = fmaf(f1.y, wv[v], acc[3]);
          acc[4] = fmaf(f2.x, wv[v], acc[4]); acc[5] = fmaf(f2.y, wv[v], acc[5]);
          acc[6] = fmaf(f3.x, wv[v], acc[6]); acc[7] = fmaf(f3.y, wv[v], acc[7]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (ch * 8 + e < F) vals[ch * 8 + e] = acc[e];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) vals[F + k] = fh[k];
      if (tf) {
#pragma unroll
        for (int c = 0; c < F + 8; ++c) tf[R + c] = vals[c];
      }
      vals[F + 8] = p.inv_depth ? fdiv(1.f, z) : z;
      vals[F + 9] = 1.f;
      __syncwarp();          // every lane has read its FD rows
#pragma unroll
      for (int q4 = 0; q4 < C::NCP / 4; ++q4)
        *stash_f(lane * C::NCP + ((q4 ^ (lane & 7)) << 2)) =
            make_float4(wgt * vals[q4 * 4 + 0], wgt * vals[q4 * 4 + 1], wgt * vals[q4 * 4 + 2], wgt * vals[q4 * 4 + 3]);
      __syncwarp();
      if (p.out_cl && p.dec_stride == F + 9) {
        // lane = (bundle, channel quad): float4 sums over the bundle's samples, 16-byte stores into the channels-last decoder
        // input.  F + 8 = 3 (mod 4): the depth is the last lane of the last decoder quad (whose slot in memory is the pad
        // channel), the opacity the first lane of the quad after it.
        constexpr int NQ = (F + 9) / 4 + 1;
        static_assert((F + 8) % 4 == 3 && NQ * 4 <= C::NCP, "quad plan of the compositing stash");
#pragma unroll 1
        for (int base = 0; base < G * NQ; base += 32) {
          const int item = base + lane;
          const int bb = min(item / NQ, G - 1), q = item - (item / NQ) * NQ;
          const int nb = __shfl_sync(full, n, rowof(bb, 0));
          const int pixb = pix_warp0 + bb;
          if (item < G * NQ && pixb < pix_hi) {
            const int r0 = rowof(bb, 0);
            float4 a = *stash_f(r0 * C::NCP + ((q ^ (r0 & 7)) << 2));
            for (int k = 1; k < nb; ++k) {
              const int rl = rowof(bb, k);
              const float4 o = *stash_f(rl * C::NCP + ((q ^ (rl & 7)) << 2));
              a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            }
            const size_t ob = (size_t)b * HW + pixb;
            if (q < NQ - 1) {
              if (q == NQ - 2) {
                p.out_depth[ob] = p.inv_depth ? fdiv(1.f, a.w) : a.w;
                a.w = p.dec_pad0;
              }
              *reinterpret_cast<float4*>(p.out_dec + ob * (F + 9) + q * 4) = a;
            } else {
              p.out_opacity[ob] = a.x;
            }
          }
        }
      } else
#pragma unroll 1
      for (int base = 0; base < G * C::NC; base += 32) {
        const int item = base + lane;
        const int bb = min(item / C::NC, G - 1), c = item - (item / C::NC) * C::NC;
        const int nb = __shfl_sync(full, n, rowof(bb, 0));          // every lane of a bundle holds its count
        const int pixb = pix_warp0 + bb;
        if (item < G * C::NC && pixb < pix_hi) {
          float a = 0.f;
          for (int k = 0; k < nb; ++k) {
            const int rl = rowof(bb, k);
            const float o = *reinterpret_cast<const float*>(stash_f(rl * C::NCP + (((c >> 2) ^ (rl & 7)) << 2) + (c & 3)));
            a = k == 0 ? o : a + o;
          }
          if (c < F + 8) {
            float* odb = p.out_cl ? p.out_dec + (size_t)(b * HW + pixb) * p.dec_stride + c : p.out_feat + ((size_t)b * CT + R + c) * HW + pixb;
            *odb = a;
            if (p.out_cl && c == F + 7)
              for (int k = F + 8; k < p.dec_stride; ++k) odb[k - c] = k == F + 8 ? p.dec_pad0 : 0.f;                   // pad channels of the decoder input
          } else if (c == F + 8) {
            p.out_depth[(size_t)b * HW + pixb] = p.inv_depth ? fdiv(1.f, a) : a;
          } else {
            p.out_opacity[(size_t)b * HW + pixb] = a;
          }
        }
      }
    }

    // ============== P6: fine colours, lane = (row, ray) (bundle_sampler.py:327-337) ==============
    // every lane of the warp is done with its stash rows
    __syncwarp();
    {
      // per-row parameters through the row's 16-byte slots of S[0..2] (free since GEMM 4): (z, x0, y0, w), the view weights,
      // (active, packed-sample row)
      static_assert(C::CH_S >= 3 && V <= 4, "row-parameter slots of the colour pass");
      *reinterpret_cast<float4*>(sS + row * 16) = make_float4(z, geo.x0, geo.y0, wgt);
      {
        float w4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < V; ++v) w4[v] = wv[v];
        *reinterpret_cast<float4*>(sS + CHB + row * 16) = make_float4(w4[0], w4[1], w4[2], w4[3]);
      }
      if (TAPS)
        *reinterpret_cast<uint4*>(sS + 2 * CHB + row * 16) = make_uint4(active ? 1u : 0u, (uint32_t)(srow & 0xffffffff), (uint32_t)((uint64_t)srow >> 32), 0u);
      __syncwarp();
      // component-wise stash of the weighted colours: float index row * R + c * BB + j, in the warp's rows of FD (+ X)
      static_assert(32 * R * 4 <= 512 * (C::CH_X + C::CH_FD), "colour stash must fit in the warp's rows");
      auto cst = [&](int t) { return reinterpret_cast<float*>(gsm + STASH0 + ((t * 4) >> 9) * CHB + wq * 512 + ((t * 4) & 511)); };
      const int nit6 = min(BB, (ns * G * BB + 31) / 32);      // (row, ray) items of the rows that can hold a sample
#pragma unroll 1
      for (int it = 0; it < nit6; ++it) {
        const int item = it * 32 + lane;
        const int r = item / BB, j = item - r * BB;
        const float4 ra = *reinterpret_cast<const float4*>(sS + (wq * 32 + r) * 16);
        const float4 rb = *reinterpret_cast<const float4*>(sS + CHB + (wq * 32 + r) * 16);
        uint4 rc = make_uint4(0u, 0u, 0u, 0u);
        if (TAPS) rc = *reinterpret_cast<const uint4*>(sS + 2 * CHB + (wq * 32 + r) * 16);
        const float zr = ra.x, wr = ra.w;
        // production build: a row without compositing weight contributes w * colour = 0 whatever it gathers: it reads pixel 0
        const bool actr = TAPS ? rc.x != 0 : wr != 0.f;
        const int64_t srow_r = TAPS ? (int64_t)(((uint64_t)rc.z << 32) | rc.y) : 0;
        const float wvr[4] = {rb.x, rb.y, rb.z, rb.w};
        const float x = ra.y + (float)(j % BS), y = ra.z + (float)(j / BS);
        const float* M = head + CAM_M;
        const float dx = fmaf(x, M[0], fmaf(y, M[1], M[2]));
        const float dy = fmaf(x, M[3], fmaf(y, M[4], M[5]));
        const float dz = fmaf(x, M[6], fmaf(y, M[7], M[8]));
        const float wx = fmaf(dx, zr, ox), wy = fmaf(dy, zr, oy), wz = fmaf(dz, zr, oz);
        // all 4 V taps of the (row, ray) in flight at once, unconditional
        float4 t[V][4];
        float tw[V][4];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float* cv = head + CAM_HEAD + CAM_VIEW * v;
          float cx = fmaf(wx, cv[CV_E + 0], fmaf(wy, cv[CV_E + 1], fmaf(wz, cv[CV_E + 2], cv[CV_E + 3])));
          float cy = fmaf(wx, cv[CV_E + 4], fmaf(wy, cv[CV_E + 5], fmaf(wz, cv[CV_E + 6], cv[CV_E + 7])));
          float cz = fmaf(wx, cv[CV_E + 8], fmaf(wy, cv[CV_E + 9], fmaf(wz, cv[CV_E + 10], cv[CV_E + 11])));
          float ix = fmaf(cx, cv[CV_K + 0], fmaf(cy, cv[CV_K + 1], cz * cv[CV_K + 2]));
          float iy = fmaf(cx, cv[CV_K + 3], fmaf(cy, cv[CV_K + 4], cz * cv[CV_K + 5]));
          float iz = fmaxf(fmaf(cx, cv[CV_K + 6], fmaf(cy, cv[CV_K + 7], cz * cv[CV_K + 8])), 1e-6f);
          const float rz = __frcp_rn(iz);
          float gx = (ix * rz) * two_W - 1.f, gy = (iy * rz) * two_H - 1.f;
          gx = actr ? gx : -1.f;
          gy = actr ? gy : -1.f;
          const Bilin bl4 = bilin_border(gx, gy, p.W, p.H);
          const float4* ib = reinterpret_cast<const float4*>(p.rgba) + (size_t)(b * V + v) * p.H * p.W;
          t[v][0] = gather(ib + bl4.o00);
          t[v][1] = gather(ib + bl4.o10);
          t[v][2] = gather(ib + bl4.o01);
          t[v][3] = gather(ib + bl4.o11);
          tw[v][0] = bl4.w00; tw[v][1] = bl4.w10; tw[v][2] = bl4.w01; tw[v][3] = bl4.w11;
        }
        float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) c4 = f4_scale_add(c4, t[v][k], tw[v][k]);
          if (TAPS && p.tap_rfd && actr) {
            float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow_r) * C::RFD;
            tp[0 * BB + j] = c4.x; tp[1 * BB + j] = c4.y; tp[2 * BB + j] = c4.z;
          }
          cr = fmaf(c4.x, wvr[v], cr); cg = fmaf(c4.y, wvr[v], cg); cb = fmaf(c4.z, wvr[v], cb);
        }
        if (TAPS && p.tap_feat && actr) {
          float* tfr = p.tap_feat + srow_r * CT;
          tfr[0 * BB + j] = cr; tfr[1 * BB + j] = cg; tfr[2 * BB + j] = cb;
        }
        *cst(r * R + 0 * BB + j) = wr * cr;
        *cst(r * R + 1 * BB + j) = wr * cg;
        *cst(r * R + 2 * BB + j) = wr * cb;
      }
      __syncwarp();
      // sum over the samples of a bundle in slot order, lane = (bundle, quad of the 3 b^2 fine-colour channels)
      constexpr int R4 = R / 4;
#pragma unroll 1
      for (int base = 0; base < G * R4; base += 32) {
        const int item = base + lane;
        const int bb = min(item / R4, G - 1), q = item - (item / R4) * R4;
        const int nb = __shfl_sync(full, n, rowof(bb, 0));
        const int pixb = pix_warp0 + bb;
        if (item < G * R4 && pixb < pix_hi) {
          float4 a = *reinterpret_cast<const float4*>(cst(rowof(bb, 0) * R + q * 4));
          for (int k = 1; k < nb; ++k) {
            const float4 o = *reinterpret_cast<const float4*>(cst(rowof(bb, k) * R + q * 4));
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
          }
          if (p.out_cl) {
            *reinterpret_cast<float4*>(p.out_feat + ((size_t)b * HW + pixb) * R + q * 4) = a;
          } else {
            float* ofb = p.out_feat + ((size_t)b * CT + q * 4) * HW + pixb;
            ofb[0] = a.x; ofb[(size_t)HW] = a.y; ofb[2 * (size_t)HW] = a.z; ofb[3 * (size_t)HW] = a.w;
          }
        }
      }
      __syncwarp();           // stash reads complete before the next tile's descriptors overwrite X / FD
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(C::TALLOC) : "memory");
  }
}

template <int BS, int FEAT_DIM, int V, int NG, bool TAPS, int PB, int MEMSRC = 0>
static int launch_render_tc3_t(const RenderParams& p, cudaStream_t st) {
  using C = Tc3Cfg<BS, FEAT_DIM, V, NG>;
  constexpr int SMEM = C::SMEM;
  static_assert(SMEM + 1024 <= 227 * 1024, "shared memory plan (dynamic + the kernel's 1 KB static section)");
  auto kern = render_tc3_kernel<BS, FEAT_DIM, V, NG, TAPS, PB, MEMSRC>;
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, kern, SMEM);
    if (e != cudaSuccess) return fail((int)e, "gdb_render_fused_fwd(tc3): cudaFuncSetAttribute(%d B): %s", SMEM, cudaGetErrorString(e));
  }
  const int G = 32 / p.max_samples;
  const long tiles = (long)p.B * ((p.pix_hi - p.pix_lo + 4 * G - 1) / (4 * G));
  long ctas = (tiles + NG - 1) / NG;
  if (ctas > sm_count()) ctas = sm_count();
  kern<<<(int)ctas, 128 * NG, SMEM, st>>>(p);
  return cuda_check("gdb_render_fused_fwd(tc3)");
}
template <int BS, int FEAT_DIM, int V, int NG, int PB = 1>
static int launch_render_tc3(const RenderParams& p, cudaStream_t st) {
  const bool taps = p.tap_rfd || p.tap_vox || p.tap_sigma || p.tap_feat || p.tap_w;
  return taps ? launch_render_tc3_t<BS, FEAT_DIM, V, NG, true, PB>(p, st) : launch_render_tc3_t<BS, FEAT_DIM, V, NG, false, PB>(p, st);
}

// the fourth-generation kernel (`precision = 6` of the C ABI).
// GDB_K3_VARIANT (development A/B, V = 3 only): "pb" = all 8 V taps of a fetch iteration in flight; "ng3" / "ng1" = one tile slot
// less per SM (more registers per thread, a larger L1); "ng3pb" / "ng1pb" = both.
int render_tc3_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, cudaStream_t st) {
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("GDB_K3_VARIANT");
    const std::string s = e ? e : "";
    variant = s == "pb" ? 1 : s == "ng3" || s == "ng1" ? 2 : s == "ng3pb" || s == "ng1pb" ? 3 : s == "ldsbound" ? 4 : 0;
  }
  if (V == 3 && variant == 4) {      // timing only, see MEMSRC
    if (bundle_size == 2 && feat_dim == 16) return launch_render_tc3_t<2, 16, 3, 4, false, 1, 1>(p, st);
    if (bundle_size == 4 && feat_dim == 32) return launch_render_tc3_t<4, 32, 3, 2, false, 1, 1>(p, st);
  }
  if (V == 3 && variant) {
    if (bundle_size == 2 && feat_dim == 16)
      return variant == 1 ? launch_render_tc3<2, 16, 3, 4, 3>(p, st) : variant == 2 ? launch_render_tc3<2, 16, 3, 3, 1>(p, st) : launch_render_tc3<2, 16, 3, 3, 3>(p, st);
    if (bundle_size == 4 && feat_dim == 32)
      return variant == 1 ? launch_render_tc3<4, 32, 3, 2, 3>(p, st) : variant == 2 ? launch_render_tc3<4, 32, 3, 1, 1>(p, st) : launch_render_tc3<4, 32, 3, 1, 3>(p, st);
  }
  if (bundle_size == 2 && feat_dim == 16 && V == 2) return launch_render_tc3<2, 16, 2, 4>(p, st);
  if (bundle_size == 2 && feat_dim == 16 && V == 3) return launch_render_tc3<2, 16, 3, 4>(p, st);
  if (bundle_size == 2 && feat_dim == 16 && V == 4) return launch_render_tc3<2, 16, 4, 2>(p, st);
  if (bundle_size == 4 && feat_dim == 32 && V == 2) return launch_render_tc3<4, 32, 2, 2>(p, st);
  if (bundle_size == 4 && feat_dim == 32 && V == 3) return launch_render_tc3<4, 32, 3, 2>(p, st);
  if (bundle_size == 4 && feat_dim == 32 && V == 4) return launch_render_tc3<4, 32, 4, 1>(p, st);
  return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_fwd(tc3): (bundle_size=%d, feat_dim=%d, V=%d) not instantiated", bundle_size, feat_dim, V);
}

}  // namespace gdb
