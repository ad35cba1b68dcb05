// K3, second-generation tensor-core variant (precision 1: fp16 operands, fp32 accumulation in TMEM).
//
// Reference: bundle_sampler.py:193-371, nerf.py:58-115, utils.py:19-43,88-121.
//
// Same arithmetic contract as gdb_render_tc.cu (SPLIT = false); what changed is how the work is laid out on the SM:
//
//  * NG = 4 tiles of 128 sample rows are in flight per SM (16 warps, <= 128 registers) instead of 2 x 4 warps at 255
//    registers.  A tile needs 23 operand chunks (46 KB): the regions are aliased by lifetime
//        X  [x_v | 1] per view      -> after GEMM 1: [h (8 chunks) | vox]
//        S  [var | mean]            -> after GEMM 1: aggregated 32-vector -> after GEMM 2: img
//        FD [featrgb_v | dir_v]     -> read again for the final blend, then (with X) the per-ray colour stash
//    and a tile owns only 128 TMEM columns (weight.0 runs view by view through two 64-column buffers).
//  * Every warp owns 32 rows end to end (geometry, gathers, epilogues, compositing), so the only cross-warp
//    synchronisation is the barrier in front of each MMA issue.
//  * Geometry is computed once per (row, view) with thread = row; the feature fetch then runs with
//    lane = (row, 16-byte quad of the texel): tap addresses and weights are redistributed by warp shuffles and each
//    LDG.128 of a warp covers whole texels (80 / 144 contiguous bytes) instead of 32 unrelated lines.  The quad lanes
//    write the fp16 operand chunks directly.
//  * Fine colours are fetched after the view weights are known with lane = (row, ray): the four rays of a bundle read
//    neighbouring source pixels, the blend over views and the compositing weight are applied in registers and the
//    sum over a bundle's samples runs per (bundle, ray) from a small shared-memory stash (no shuffles).
//  * The biases of global_fc, lr0 and weight.0 ride in the GEMMs (a constant-one K slot fed from a shared constant chunk),
//    odd chunk counts pair with a shared zero chunk, the per-channel compositing goes through a swizzled shared-memory
//    transpose (lane = (bundle, channel), coalesced stores), zero-weight voxel taps are skipped, and the loops over
//    gather iterations / rays / views of the epilogues are rolled: 7.3 k SASS instructions instead of 13.3 k in the first
//    tcgen05 variant, whose top stall reason was instruction fetch.  The production instantiation (TAPS = false)
//    carries none of the per-sample parity outputs.
// Round 2, last session (profiles/r02_k3_rcpa.log): the correctly rounded reciprocal of the colour pass (__frcp_rn: MUFU.RCP, two
// Newton steps, a range check and a CALL to a slow path) sat between the projection of a (row, ray) into a view and that view's
// four gathers, so the twelve "branch-free" gathers of an item were in fact three dependent groups.  Defaults now:
//   GDB_X_RCPA    colour pass: MUFU.RCP alone (1 ulp of a pixel coordinate, 6e-5 px at 800 px)
//   GDB_X_P1A     P0 / P1 (per row and view): single-MUFU square roots / reciprocal (ball radius, mip level, texture coordinate)
//   GDB_X_LINPROJ colour pass: pixel -> source-image projection as ONE 3x3 matrix + offset per (target view, source view),
//                 composed in double precision when the camera block is staged: image = z * Q (x, y, 1) + c, 9 FMAs per
//                 (row, ray, view) instead of 12 + 18
// 0.885 -> 0.856 ms per 8 DTU views, 2.22 -> 1.93 ms at NeRF-synthetic 4x4, 1.66 -> 1.59 ms at LLFF; errors against the oracle
// unchanged (fine rgb max 1.2e-5 / 1.7e-5).  -DGDB_X_EXACT restores the exact chain (A/B).  Both generations (precision 1 / 4)
// follow the same flags and stay bit-identical to each other.
#ifndef GDB_X_EXACT
#define GDB_X_RCPA
#define GDB_X_P1A
#define GDB_X_LINPROJ
#endif
#include "gdb_render_tc2.cuh"

namespace gdb {

// TAPS: the per-sample parity outputs (tests); the production instantiation carries none of that code or state
//
// GEN = 2: the round-1 kernel as measured in profiles/r01_* (kept for A/B, `precision = 4`).
// GEN = 3: the same arithmetic (bit-identical results) restructured against the long-scoreboard stalls that ncu showed to be
//   the top stall reason (37 % of all warp samples, profiles/r02_k3_phase_profile_gen2.txt):
//    * per-(row, view) fetch descriptors travel through the row's own 48 bytes of the X_v operand region (three LDS.128 of
//      a fetch lane instead of thirteen warp shuffles; 36 registers freed for loads in flight),
//    * the gathers are issued in batches as predicated loads (no branches between them): level-0 taps of ALL views, then
//      level-1 taps of all views (FB = 1), or all 8 V taps at once (FB = 2); twelve colour taps per (row, ray) at once;
//      the voxel taps at once; the depth ranges of the NEXT tile are prefetched under the current tile,
//    * compositing and the output stores run on float4 quads (lane = (bundle, channel quad), STG.128 into the channels-last
//      decoder input), the fine colours are stashed component-wise so that a lane sums and stores one float4 of a bundle,
//    * the colour pass reads its per-row parameters from shared memory (two LDS.128) instead of 5 + V shuffles.
// sum_k relu(g[k]) w[k] over 32 accumulator columns in the four chains k mod 4, two chains per packed FFMA2 (bit-identical
// to the scalar chains); w is 16-byte aligned in shared memory
__device__ __forceinline__ void relu_dot32(const float (&g)[32], const float* w, unsigned long long& a01, unsigned long long& a23) {
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const ulonglong2 ww = *reinterpret_cast<const ulonglong2*>(w + k);
    a01 = fma2(pack2(fmaxf(g[k + 0], 0.f), fmaxf(g[k + 1], 0.f)), ww.x, a01);
    a23 = fma2(pack2(fmaxf(g[k + 2], 0.f), fmaxf(g[k + 3], 0.f)), ww.y, a23);
  }
}

#ifdef GDB_X_P1A
#define GDB_P1_SQRT sqrt_approx
#define GDB_P1_RCP rcp_approx
#else
#define GDB_P1_SQRT sqrtf
#define GDB_P1_RCP __frcp_rn
#endif

//
// Measured variants kept behind compile-time flags (tools/build_variant.sh; profiles/r02_k3_experiments_s3.log; ms per 8 views,
// DTU / LLFF / NeRF-synthetic against 0.882 / 1.655 / 2.27 of the default build in the same call):
//   -DGDB_X_SPINHINT=ns  suspend-time hint on the mbarrier probe (the probe loop is 5.6 % of the issued instructions): no effect
//   -DGDB_X_FMA2         packed FFMA2 in the ReLU-dot epilogues: 124 bytes of spills at 128 registers, 0.998 at DTU (+3.7 %), -0.7 % at NeRF
//   -DGDB_K3_CHPAD=64    operand chunks 2048 + 64 bytes apart (removes the two-way conflicts of the fetch lanes' 8-byte stores):
//                        0.894 / 1.676 (+1.3 %), 2.229 (-0.3 %)
//   -DGDB_X_UNCOND       plain gathers instead of predicated ones with zero-initialised destinations: 0.906 / 1.708 (+2.5 / +3 %)
//   -DGDB_X_ROLLR        the two rounds of GEMM 4 as one rolled loop: 0.936 / 1.774 (+6 %)
//   -DGDB_X_EXACT        the exact reciprocal / square roots and the step-by-step projection chain of the colour pass (the default
//                        until the last session of round 2): 0.885 / 1.66 / 2.22 against 0.856 / 1.59 / 1.93
template <int BS, int FEAT_DIM, int V, int NG, bool TAPS, int GEN, int FB>
__global__ void __launch_bounds__(128 * NG, 1) render_tc2_kernel(const RenderParams p) {
  using C = Tc2Cfg<BS, FEAT_DIM, V, NG>;
  using ML = typename C::ML;
  constexpr int BB = C::BB, F = C::F, FP = C::FP, R = C::R, CT = C::CT, QL = C::QL, IPW = C::IPW;
  constexpr int CHB = C::CHB;                  // bytes between consecutive operand chunks
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  float* vec = reinterpret_cast<float*>(smem + C::VEC_OFF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 7;                      // group = tile slot
  const int row = tid & 127;                   // sample row == TMEM lane
  const int wq = warp & 3;
  unsigned char* gsm = smem + C::GROUP_OFF + (size_t)g * C::GROUP_BYTES;
  const uint32_t mbar = smem_u32(gsm + C::A_END);
  const unsigned full = 0xffffffffu;

  // ---- one-time setup: weights -> smem (fp16 B operands + fp32 vectors), constant chunks, mbarriers, TMEM
  {
    tc2_stage_weights<C>(smem, vec, p.mlp, tid, blockDim.x);
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
  const uint32_t zero_chunk = w_base + C::ZERO_OFF, one_chunk = w_base + C::ONE_OFF;
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
  // row (= lane) of sample `k` of the warp's bundle `bb`.  GEN 2: bundle-major (the packed sample order).  GEN 3: slot-major,
  // consecutive lanes hold the SAME slot of ADJACENT bundles, so the rows a gather instruction covers read neighbouring
  // texels (one run of contiguous bytes per image row instead of one 128-byte line per lane)
  auto rowof = [&](int bb, int k) { return GEN == 3 ? k * G + bb : bb * ns + k; };
  const int slot = GEN == 3 ? lane / G : lane % ns;
  const int bl = GEN == 3 ? lane - slot * G : lane / ns;
  const bool lane_ok = GEN == 3 ? slot < ns : bl < G;
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
  if (GEN == 3 && (int)(blockIdx.x * NG + g) < tiles) rng_next = load_ranges(blockIdx.x * NG + g);

  // Tile assignment.  Every tile slot starts with tile blockIdx.x * NG + g; the following ones come from an atomic counter when
  // the launch carries one (p.tile_counter, zero at launch: SMs do not run at one speed - with a static stride 8 % of the SM
  // cycles of a DTU launch were idle at the end, 0.96 -> 0.89 ms), else from the static stride.  Row 0 of the group draws the
  // next tile at the top of the current one; the group reads it behind the barrier in front of GEMM 1.
  __shared__ int next_tile_s[NG];
  const bool dyn = GEN == 3 && p.tile_counter != nullptr;
  int tn = 0;
#pragma unroll 1
  for (int tile = blockIdx.x * NG + g; tile < tiles; tile = tn) {
    unsigned int drawn = 0;                                // the atomic's result is not touched before GEMM 1: its latency hides under P0-P2
    if (dyn) {
      if (row == 0) drawn = atomicAdd(p.tile_counter, 1u);
    } else {
      tn = tile + gridDim.x * NG;
    }
    const int b = tile / tiles_pv;                         // uniform over the group
    if (b != cur_b) {                                      // stage this view's camera block
      group_sync(g);
      for (int i = row; i < CAM_HEAD + CAM_VIEW * V; i += 128) scam[i] = p.cam[(size_t)b * p.cam_stride + i];
      group_sync(g);
#ifdef GDB_X_LINPROJ
      if (row < V) {
        // Q = K_v R_v M, c = K_v (R_v o + t_v): image coordinates of pixel (x, y) at ray depth z are z * Q (x, y, 1) + c
        const float* cv = scam + CAM_HEAD + CAM_VIEW * row;
        double A[9], Q[9], c3[3];
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += (double)cv[CV_K + i * 3 + k] * (double)cv[CV_E + k * 4 + j];
            A[i * 3 + j] = a;
          }
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) {
            double a = 0.0;
            for (int k = 0; k < 3; ++k) a += A[i * 3 + k] * (double)scam[CAM_M + k * 3 + j];
            Q[i * 3 + j] = a;
          }
          double a = 0.0;
          for (int k = 0; k < 3; ++k) a += A[i * 3 + k] * (double)scam[CAM_O + k] + (double)cv[CV_K + i * 3 + k] * (double)cv[CV_E + k * 4 + 3];
          c3[i] = a;
        }
        float* L = scam + C::LIN_FLOATS_OFF + row * 12;
        for (int i = 0; i < 9; ++i) L[i] = (float)Q[i];
        for (int i = 0; i < 3; ++i) L[9 + i] = (float)c3[i];
      }
      group_sync(g);
#endif
      cur_b = b;
    }
    const int pix_warp0 = pix_lo + ((tile - b * tiles_pv) * 4 + wq) * G;       // first bundle of my warp
    const int pix_raw = pix_warp0 + bl;
    const bool has_bundle = lane_ok && pix_raw < pix_hi;
    const int pix = has_bundle ? pix_raw : pix_lo;
    const int bidx = b * HW + pix;
    const int yb = pix / p.Wb, xb = pix - yb * p.Wb;

    // =========================== P0: sample placement (thread = row) ===========================
    const float4 rng = GEN == 3 ? rng_next : load_ranges(tile);
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
      ball = GDB_P1_SQRT(ex * ex + ey * ey + ez * ez) * geo.unit_ball;
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
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
      if constexpr (GEN == 3) {
        // all (non-zero-weight) taps in flight at once; same accumulation order as below
        float4 tl[8], th[8];
        float tw[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          int xx = (k & 1) ? x1 : x0, yy = (k & 2) ? y1 : y0, zz = (k & 4) ? z1 : z0;
          tw[k] = ((k & 1) ? tx : 1.f - tx) * ((k & 2) ? ty : 1.f - ty) * ((k & 4) ? tz : 1.f - tz);
          const float4* tp = reinterpret_cast<const float4*>(vb_ + zz * p.vol_sz + yy * p.vol_sy + xx * p.vol_sx);
          tl[k] = ldg4_if(tp, tw[k] != 0.f);
          th[k] = ldg4_if(tp + 1, tw[k] != 0.f);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (tw[k] != 0.f) {
            lo = f4_scale_add(lo, tl[k], tw[k]);
            hi = f4_scale_add(hi, th[k], tw[k]);
          }
      } else
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int xx = (k & 1) ? x1 : x0, yy = (k & 2) ? y1 : y0, zz = (k & 4) ? z1 : z0;
        float w = ((k & 1) ? tx : 1.f - tx) * ((k & 2) ? ty : 1.f - ty) * ((k & 4) ? tz : 1.f - tz);
        // a bundle centre sits on a voxel centre whenever the cost volume has the bundle map's resolution: the x / y
        // fractions are then exactly zero for most bundles and six of the eight taps carry weight 0 - not fetched
        if (w != 0.f) {
          const float* tp = vb_ + zz * p.vol_sz + yy * p.vol_sy + xx * p.vol_sx;
          lo = f4_scale_add(lo, ldg4(tp), w);
          hi = f4_scale_add(hi, ldg4(tp + 4), w);
        }
      }
      voxh = make_uint4(pack_h2(lo.x, lo.y), pack_h2(lo.z, lo.w), pack_h2(hi.x, hi.y), pack_h2(hi.z, hi.w));
      if (TAPS && p.tap_vox) {
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[0] = lo;
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[1] = hi;
      }
    }

    // ====================== P1: per-view fetch descriptors (thread = row) ======================
    // a0/a1: float4 index of tap (0,0) at mip levels l0/l1; pk: row strides and flags; bilinear fractions; direction features
    int d_a0[V], d_a1[V];
    uint32_t d_pk[V];
    float d_fu0[V], d_fv0[V], d_fu1[V], d_fv1[V], d_fr[V], d_dir[V][4];
    float tdx = cwx - ox, tdy = cwy - oy, tdz = cwz - oz;      // unit vector target camera -> sample (view independent)
    unit3_fast(tdx, tdy, tdz);
    // GEN 3 hands the descriptors over through shared memory at once: the loop stays rolled (530 fewer instructions of
    // straight-line code per tile for the instruction cache: 0.962 -> 0.938 ms at DTU)
#pragma unroll(GEN == 3 ? 1 : V)
    for (int v = 0; v < V; ++v) {
      const float* cv = head + CAM_HEAD + CAM_VIEW * v;
      // centre of the bundle's points in the source camera frame (bundle_sampler.py:340; the mean commutes with the rigid map)
      const float ccx = fmaf(cwx, cv[CV_E + 0], fmaf(cwy, cv[CV_E + 1], fmaf(cwz, cv[CV_E + 2], cv[CV_E + 3])));
      const float ccy = fmaf(cwx, cv[CV_E + 4], fmaf(cwy, cv[CV_E + 5], fmaf(cwz, cv[CV_E + 6], cv[CV_E + 7])));
      const float ccz = fmaf(cwx, cv[CV_E + 8], fmaf(cwy, cv[CV_E + 9], fmaf(cwz, cv[CV_E + 10], cv[CV_E + 11])));
      // mip level (:343-348): only its fractional part reaches the output, approximate division is ample
      const float dist = GDB_P1_SQRT(ccx * ccx + ccy * ccy + ccz * ccz);
      const float sec = __fdividef(dist, ccz);
      const float sec_sq = sec * sec;
      const float rb = __fdividef(dist, ball);
      const float foot = __fdividef(sec_sq, GDB_P1_SQRT(fmaxf(rb * rb - 1.f, 1e-12f)) + GDB_P1_SQRT(fmaxf(sec_sq - 1.f, 1e-12f)));
      const float lod = log2f(__fdividef(foot, cv[CV_PIXR]));
      constexpr float ifb = 1.f / (float)BS;                  // power of two: exact
      const float pxc = fmaf(ccx, cv[CV_K + 0] * ifb, fmaf(ccy, cv[CV_K + 1] * ifb, ccz * (cv[CV_K + 2] * ifb)));
      const float pyc = fmaf(ccx, cv[CV_K + 3] * ifb, fmaf(ccy, cv[CV_K + 4] * ifb, ccz * (cv[CV_K + 5] * ifb)));
      const float pzc = fmaxf(fmaf(ccx, cv[CV_K + 6], fmaf(ccy, cv[CV_K + 7], ccz * cv[CV_K + 8])), 1e-6f);
      const float rz = GDB_P1_RCP(pzc);
      const float u01 = pxc * rz * inv_Wb, v01 = pyc * rz * inv_Hb;
      d_a0[v] = 0; d_a1[v] = 0; d_pk[v] = 0;
      d_fu0[v] = d_fv0[v] = d_fu1[v] = d_fv1[v] = d_fr[v] = 0.f;
      d_dir[v][0] = d_dir[v][1] = d_dir[v][2] = d_dir[v][3] = 0.f;
      if (active) {
        float flod = fminf(fmaxf(lod, 0.f), (float)p.L);
        if (!(flod >= 0.f)) flod = 0.f;
        int l0 = (int)floorf(flod);
        int l1 = min(l0 + 1, p.L);
        const bool tri = flod > 0.f;
        int w0 = p.Wb >> l0, h0 = p.Hb >> l0, w1 = p.Wb >> l1, h1 = p.Hb >> l1;
        TexTap ta = tex_tap(u01, v01, w0, h0);
        TexTap tb = tex_tap(u01, v01, w1, h1);
        d_a0[v] = (int)(p.tex_level[l0] >> 2) + ((b * V + v) * h0 * w0 + ta.o00) * QL;
        d_a1[v] = (int)(p.tex_level[l1] >> 2) + ((b * V + v) * h1 * w1 + tb.o00) * QL;
        d_pk[v] = (uint32_t)((ta.o01 - ta.o00) * QL) | ((uint32_t)((tb.o01 - tb.o00) * QL) << 14) |
                  ((uint32_t)(ta.o10 - ta.o00) << 28) | ((uint32_t)(tb.o10 - tb.o00) << 29) | ((tri ? 1u : 0u) << 30) | (1u << 31);
        d_fu0[v] = ta.fu; d_fv0[v] = ta.fv; d_fu1[v] = tb.fu; d_fv1[v] = tb.fv; d_fr[v] = flod - (float)l0;
        float sx = cwx - cv[CV_C + 0], sy = cwy - cv[CV_C + 1], sz = cwz - cv[CV_C + 2];
        unit3_fast(sx, sy, sz);
        float ddx = tdx - sx, ddy = tdy - sy, ddz = tdz - sz;
        unit3_fast(ddx, ddy, ddz);
        d_dir[v][0] = ddx; d_dir[v][1] = ddy; d_dir[v][2] = ddz; d_dir[v][3] = tdx * sx + tdy * sy + tdz * sz;
        if (TAPS && p.tap_rfd) {
          float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow) * C::RFD + R + F;
          tp[0] = ddx; tp[1] = ddy; tp[2] = ddz; tp[3] = d_dir[v][3];
        }
      }
      if constexpr (GEN == 3) {
        // the descriptor of (row, view) travels through the row's own 3 x 16 bytes of the X_v operand region: the fetch
        // lanes of the row read it there and overwrite it with the operand afterwards
        static_assert(C::XCH >= 3, "three 16-byte slots per (row, view)");
        unsigned char* dp = sX + (v * C::XCH) * CHB + row * 16;
        *reinterpret_cast<uint4*>(dp) = make_uint4((uint32_t)d_a0[v], (uint32_t)d_a1[v], d_pk[v], __float_as_uint(d_fr[v]));
        *reinterpret_cast<float4*>(dp + CHB) = make_float4(d_fu0[v], d_fv0[v], d_fu1[v], d_fv1[v]);
        *reinterpret_cast<float4*>(dp + 2 * CHB) = make_float4(d_dir[v][0], d_dir[v][1], d_dir[v][2], d_dir[v][3]);
      }
    }
    if constexpr (GEN == 3) {
      // the depth ranges of my next tile, in flight underneath this tile (dynamic assignment: behind GEMM 1's barrier)
      if (!dyn && tn < tiles) rng_next = load_ranges(tn);
    }

    // ================= P2: mip-mapped feature fetch, lane = (row, quad) =================
    // writes FD_v = [featrgb_v | dir_v], X_v = [x_v | 1] (nerf.py:69-71) and S = [var | mean] over views (nerf.py:73)
    if constexpr (GEN == 3) {
      __syncwarp();                       // the descriptors of my warp's 32 rows are in shared memory
      // view_fc weights of my quad's four channels, loaded once per tile (the descriptors no longer occupy registers)
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
      // rows >= ns * G of a warp never hold a sample (slot-major rows: their slot index is >= max_samples; 30 of 32 rows are in
      // use with 3 or 6 samples per bundle): the fetch iterations that would cover only such rows are skipped.  Their operand rows
      // keep whatever the previous tile left there - rows are independent in every GEMM, nothing downstream reads them (the
      // compositing sums run over rows < ns * G only) and their compositing weight is exactly 0.
      const int nit = min(C::NIT, (ns * G + IPW - 1) / IPW);
#pragma unroll 1
      for (int it = 0; it < nit; ++it) {
        const int src_raw = it * IPW + gr;
        const bool ok = glane && src_raw < 32;
        const int src = min(src_raw, 31);
        const int orow16 = (wq * 32 + src) * 16;
        const int64_t srow_g = TAPS ? __shfl_sync(full, srow, src) : 0;
        const unsigned char* dsc = sX + orow16;
        const float4* tq = tex4 + (glane ? gq : 0);
        // four taps of one mip level of one view as predicated loads
        auto taps4 = [&](const uint4 q0, int level, float4(&t)[4]) {
          const uint32_t pk = q0.z;
          const int a = level ? (int)q0.y : (int)q0.x;
          const int dy = level ? (int)((pk >> 14) & 0x3FFF) : (int)(pk & 0x3FFF);
          const int dx = (int)((pk >> (28 + level)) & 1) * QL;
          const bool pr = ok && (pk >> 31) && (level == 0 || ((pk >> 30) & 1));
          const float4* b0 = tq + a;
#ifdef GDB_X_UNCOND
          // experiment: every address is valid whatever the predicate (zero descriptors of rows without a sample point at texel 0,
          // level 1 exists even when unused) and every consumer selects on the same flags: plain loads, no zero initialisation
          (void)pr;
          t[0] = __ldg(b0); t[1] = __ldg(b0 + dx); t[2] = __ldg(b0 + dy); t[3] = __ldg(b0 + (dy + dx));
#else
          t[0] = ldg4_if(b0, pr);
          t[1] = ldg4_if(b0 + dx, pr);
          t[2] = ldg4_if(b0 + dy, pr);
          t[3] = ldg4_if(b0 + (dy + dx), pr);
#endif
        };
        auto mix = [&](float4& f, const float4(&t)[4], const uint4 q0, const float4 q1) {      // level-1 blend (tri-linear part)
          if ((q0.z >> 30) & 1) {
            const float4 bq = bilerp4(t[0], t[1], t[2], t[3], q1.z, q1.w);
            const float frac = __uint_as_float(q0.w);
            f.x = lerpf(f.x, bq.x, frac); f.y = lerpf(f.y, bq.y, frac); f.z = lerpf(f.z, bq.z, frac); f.w = lerpf(f.w, bq.w, frac);
          }
        };
        float4 f[V];
        if constexpr (FB == 2) {            // all 8 V taps of the iteration in flight
          float4 t0[V][4], t1[V][4];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * CHB);
            taps4(q0, 0, t0[v]);
            taps4(q0, 1, t1[v]);
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * CHB);
            const float4 q1 = *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 1) * CHB);
            f[v] = bilerp4(t0[v][0], t0[v][1], t0[v][2], t0[v][3], q1.x, q1.y);
            mix(f[v], t1[v], q0, q1);
          }
        } else if constexpr (FB == 1) {     // level 0 of all views, then level 1 of all views
          float4 t[V][4];
#pragma unroll
          for (int v = 0; v < V; ++v) taps4(*reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * CHB), 0, t[v]);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 q1 = *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 1) * CHB);
            f[v] = bilerp4(t[v][0], t[v][1], t[v][2], t[v][3], q1.x, q1.y);
          }
#pragma unroll
          for (int v = 0; v < V; ++v) taps4(*reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * CHB), 1, t[v]);
#pragma unroll
          for (int v = 0; v < V; ++v)
            mix(f[v], t[v], *reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * CHB),
                *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 1) * CHB));
        } else {                             // both levels of one view at a time
#pragma unroll
          for (int v = 0; v < V; ++v) {
            float4 t0[4], t1[4];
            const uint4 q0 = *reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * CHB);
            taps4(q0, 0, t0);
            taps4(q0, 1, t1);
            const float4 q1 = *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 1) * CHB);
            f[v] = bilerp4(t0[0], t0[1], t0[2], t0[3], q1.x, q1.y);
            mix(f[v], t1, q0, q1);
          }
        }
        float xq[V][4];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const uint32_t pk = *reinterpret_cast<const uint32_t*>(dsc + (v * C::XCH) * CHB + 8);
          const float4 q2 = *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 2) * CHB);
          const float dir[4] = {q2.x, q2.y, q2.z, q2.w};
          const bool act = ok && (pk >> 31);
          if (TAPS && p.tap_rfd && act) {
            float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow_g) * C::RFD + R + gq * 4;
            tp[0] = f[v].x; tp[1] = f[v].y; tp[2] = f[v].z;
            if (!last_quad) tp[3] = f[v].w;
          }
          const float fe[4] = {f[v].x, f[v].y, f[v].z, f[v].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float t = vb[e];
            t = fmaf(vw[0][e], dir[0], t);
            t = fmaf(vw[1][e], dir[1], t);
            t = fmaf(vw[2][e], dir[2], t);
            t = fmaf(vw[3][e], dir[3], t);
            xq[v][e] = act ? fe[e] + fmaxf(t, 0.f) : 0.f;
          }
          __syncwarp();                     // every lane of the row has read the descriptor that the operand now replaces
          if (ok) {
            unsigned char* fdp = sFD + (v * C::FDCH + (gq >> 1)) * CHB + orow16;
            unsigned char* xp = sX + (v * C::XCH + (gq >> 1)) * CHB + orow16;
            if (last_quad) {
              *reinterpret_cast<uint4*>(fdp) = make_uint4(pack_h2(fe[0], fe[1]), pack_h2(fe[2], dir[0]), pack_h2(dir[1], dir[2]), pack_h2(dir[3], 0.f));
              *reinterpret_cast<uint4*>(xp) = make_uint4(pack_h2(xq[v][0], xq[v][1]), pack_h2(xq[v][2], act ? 1.f : 0.f), 0u, 0u);
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
    } else
#pragma unroll 1
    for (int it = 0; it < C::NIT; ++it) {
      const int src_raw = it * IPW + gr;
      const bool ok = glane && src_raw < 32;
      const int src = min(src_raw, 31);
      const int orow16 = (wq * 32 + src) * 16;
      const int64_t srow_g = TAPS ? __shfl_sync(full, srow, src) : 0;
      float xq[V][4];
      // view_fc weights of my quad's four channels (re-read per iteration: keeping them live across the tile costs spills)
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
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int a0 = __shfl_sync(full, d_a0[v], src), a1 = __shfl_sync(full, d_a1[v], src);
        const uint32_t pk = __shfl_sync(full, d_pk[v], src);
        const float fu0 = __shfl_sync(full, d_fu0[v], src), fv0 = __shfl_sync(full, d_fv0[v], src);
        const float fu1 = __shfl_sync(full, d_fu1[v], src), fv1 = __shfl_sync(full, d_fv1[v], src);
        const float frac = __shfl_sync(full, d_fr[v], src);
        float dir[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) dir[j] = __shfl_sync(full, d_dir[v][j], src);
        const bool act = ok && (pk >> 31);
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) {
          const int dy0 = pk & 0x3FFF, dy1 = (pk >> 14) & 0x3FFF;
          const int dx0 = ((pk >> 28) & 1) * QL, dx1 = ((pk >> 29) & 1) * QL;
          const int i0 = a0 + gq;
          f = bilerp4(__ldg(tex4 + i0), __ldg(tex4 + (i0 + dx0)), __ldg(tex4 + (i0 + dy0)), __ldg(tex4 + (i0 + dy0 + dx0)), fu0, fv0);
          if ((pk >> 30) & 1) {
            const int i1 = a1 + gq;
            float4 bq = bilerp4(__ldg(tex4 + i1), __ldg(tex4 + (i1 + dx1)), __ldg(tex4 + (i1 + dy1)), __ldg(tex4 + (i1 + dy1 + dx1)), fu1, fv1);
            f.x = lerpf(f.x, bq.x, frac); f.y = lerpf(f.y, bq.y, frac); f.z = lerpf(f.z, bq.z, frac); f.w = lerpf(f.w, bq.w, frac);
          }
          if (TAPS && p.tap_rfd) {
            float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow_g) * C::RFD + R + gq * 4;
            tp[0] = f.x; tp[1] = f.y; tp[2] = f.z;
            if (!last_quad) tp[3] = f.w;
          }
        }
        // view_fc + residual, my four channels
        float fe[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float t = vb[e];
          t = fmaf(vw[0][e], dir[0], t);
          t = fmaf(vw[1][e], dir[1], t);
          t = fmaf(vw[2][e], dir[2], t);
          t = fmaf(vw[3][e], dir[3], t);
          xq[v][e] = act ? fe[e] + fmaxf(t, 0.f) : 0.f;
        }
        if (ok) {
          unsigned char* fdp = sFD + (v * C::FDCH + (gq >> 1)) * CHB + orow16;
          unsigned char* xp = sX + (v * C::XCH + (gq >> 1)) * CHB + orow16;
          if (last_quad) {
            // featrgb's pad channel is K slot F: dir_v follows in FD, the constant one in X
            *reinterpret_cast<uint4*>(fdp) = make_uint4(pack_h2(f.x, f.y), pack_h2(f.z, dir[0]), pack_h2(dir[1], dir[2]), pack_h2(dir[3], 0.f));
            *reinterpret_cast<uint4*>(xp) = make_uint4(pack_h2(xq[v][0], xq[v][1]), pack_h2(xq[v][2], act ? 1.f : 0.f), 0u, 0u);
            xq[v][3] = 0.f;
          } else {
            *reinterpret_cast<uint2*>(fdp + (gq & 1) * 8) = make_uint2(pack_h2(f.x, f.y), pack_h2(f.z, f.w));
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

    // ================= GEMM 1: global_fc, G_v = [var|mean] W_gs + [x_v|1] W_gx =================
    if (dyn && row == 0) next_tile_s[g] = (int)(gridDim.x * NG + drawn);
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (dyn) {
      tn = next_tile_s[g];
      if (tn < tiles) rng_next = load_ranges(tn);
    }
    if (row == 0) {
      tc_fence_after();
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        mma_chunks(tmem_group + v * 32, aS, C::SCH, zero_chunk, w_base + C::W_GS, 32, 0, CHB);
        mma_chunks(tmem_group + v * 32, aX + v * C::XCH * CHB, C::XCH, zero_chunk, w_base + C::W_GX, 32, 1, CHB);
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
#ifdef GDB_X_FMA2
        {
          unsigned long long a01 = pack2(s0, 0.f), a23 = pack2(0.f, 0.f);
          relu_dot32(gv, vec + C::X_AGG_W, a01, a23);
          unpack2(a01, s0, s1); unpack2(a23, s2_, s3);
        }
#else
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          s0 = fmaf(fmaxf(gv[k + 0], 0.f), vec[C::X_AGG_W + k + 0], s0);
          s1 = fmaf(fmaxf(gv[k + 1], 0.f), vec[C::X_AGG_W + k + 1], s1);
          s2_ = fmaf(fmaxf(gv[k + 2], 0.f), vec[C::X_AGG_W + k + 2], s2_);
          s3 = fmaf(fmaxf(gv[k + 3], 0.f), vec[C::X_AGG_W + k + 3], s3);
        }
#endif
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
#ifdef GDB_X_FMA2
        const unsigned long long aa = pack2(a, a);
#pragma unroll
        for (int k = 0; k < 32; k += 2) fma2_acc(im[k], im[k + 1], fmaxf(gv[k], 0.f), fmaxf(gv[k + 1], 0.f), aa);
#else
#pragma unroll
        for (int k = 0; k < 32; ++k) im[k] = fmaf(fmaxf(gv[k], 0.f), a, im[k]);
#endif
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
      mma_chunks(tmem_group, aS, 4, zero_chunk, w_base + C::W_FC, 16, 0, CHB);
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
#ifdef GDB_X_ROLLR
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int r = 0; r < C::ROUNDS; ++r) {
      const int v0 = C::round_start(r), nv = C::round_n(r);
      tc_fence_before();
      fence_async_smem();
      group_sync(g);
      if (row == 0) {
        tc_fence_after();
        if (r == 0) mma_chunks(tmem_group + (C::NB - 1) * 64, aX, 8, zero_chunk, w_base + C::W_SH, 16, 0, CHB);
#pragma unroll 1
        for (int i = 0; i < nv; ++i) {
          const uint32_t d = tmem_group + i * 64;
          mma_chunks(d, aX, 8, zero_chunk, w_base + C::W_0S, 64, 0, CHB);                                   // h
          mma_step(d, aX + 8 * CHB, aS, w_base + C::W_0S + 8 * 64 * 16, 64, 1);                        // vox | img[0:8]
          mma_step(d, aS + CHB, one_chunk, w_base + C::W_0S + 10 * 64 * 16, 64, 1);                    // img[8:16] | 1
          mma_chunks(d, aFD + (v0 + i) * C::FDCH * CHB, C::FDCH, zero_chunk, w_base + C::W_0V, 64, 1, CHB); // featrgb_v | dir_v
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
#ifdef GDB_X_FMA2
          {
            unsigned long long a01 = pack2(q0, q1), a23 = pack2(q2, q3);
            relu_dot32(hid, vec + C::X_W2_W + half * 32, a01, a23);
            unpack2(a01, q0, q1); unpack2(a23, q2, q3);
          }
#else
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            q0 = fmaf(fmaxf(hid[k + 0], 0.f), vec[C::X_W2_W + half * 32 + k + 0], q0);
            q1 = fmaf(fmaxf(hid[k + 1], 0.f), vec[C::X_W2_W + half * 32 + k + 1], q1);
            q2 = fmaf(fmaxf(hid[k + 2], 0.f), vec[C::X_W2_W + half * 32 + k + 2], q2);
            q3 = fmaf(fmaxf(hid[k + 3], 0.f), vec[C::X_W2_W + half * 32 + k + 3], q3);
          }
#endif
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
#pragma unroll
      for (int v = 0; v < V; ++v) wv[v] /= wsum;
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

    auto stash_f = [&](int t) { return reinterpret_cast<float4*>(gsm + ((t * 4) >> 9) * CHB + wq * 512 + ((t * 4) & 511)); };
    const size_t ostr = p.out_cl ? 1 : (size_t)HW;
    float* tf = (TAPS && p.tap_feat && active) ? p.tap_feat + srow * CT : nullptr;

    // ---- blended features sum_v w_v featrgb_v (featrgb read back from the FD operand), geometry head, depth, opacity:
    //      weighted by the compositing weight and transposed through shared memory (the warp's own rows of region X),
    //      then summed over a bundle's samples in slot order with lane = (bundle, channel) and stored coalesced
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
      if (GEN == 3 && p.out_cl && p.dec_stride == F + 9) {
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
    // every lane of the warp is done with its FD rows (the stash aliases X + FD, rows of this warp only)
    __syncwarp();
    if constexpr (GEN == 3) {
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
      // component-wise stash of the weighted colours: float index row * R + c * BB + j, in the warp's rows of X + FD
      static_assert(32 * R * 4 <= 512 * (C::CH_X + C::CH_FD), "colour stash must fit in the warp's rows of X + FD");
      auto cst = [&](int t) { return reinterpret_cast<float*>(gsm + ((t * 4) >> 9) * CHB + wq * 512 + ((t * 4) & 511)); };
      const int nit6 = min(BB, (ns * G * BB + 31) / 32);      // (row, ray) items of the rows that can hold a sample
#pragma unroll 1      // (2 or 4 items per trip change nothing at 4x4 bundles: profiles/r02_k3_experiments_s3.log)
      for (int it = 0; it < nit6; ++it) {
        const int item = it * 32 + lane;
        const int r = item / BB, j = item - r * BB;
        const float4 ra = *reinterpret_cast<const float4*>(sS + (wq * 32 + r) * 16);
        const float4 rb = *reinterpret_cast<const float4*>(sS + CHB + (wq * 32 + r) * 16);
        uint4 rc = make_uint4(0u, 0u, 0u, 0u);
        if (TAPS) rc = *reinterpret_cast<const uint4*>(sS + 2 * CHB + (wq * 32 + r) * 16);
        const float zr = ra.x, wr = ra.w;
        // production build: a row without compositing weight contributes w * colour = 0 whatever it gathers - not fetched
        const bool actr = TAPS ? rc.x != 0 : wr != 0.f;
        const int64_t srow_r = TAPS ? (int64_t)(((uint64_t)rc.z << 32) | rc.y) : 0;
        const float wvr[4] = {rb.x, rb.y, rb.z, rb.w};
        const float x = ra.y + (float)(j % BS), y = ra.z + (float)(j / BS);
#ifndef GDB_X_LINPROJ
        const float* M = head + CAM_M;
        const float dx = fmaf(x, M[0], fmaf(y, M[1], M[2]));
        const float dy = fmaf(x, M[3], fmaf(y, M[4], M[5]));
        const float dz = fmaf(x, M[6], fmaf(y, M[7], M[8]));
        const float wx = fmaf(dx, zr, ox), wy = fmaf(dy, zr, oy), wz = fmaf(dz, zr, oz);
#endif
        // all 4 V taps of the (row, ray) in flight at once
        float4 t[V][4];
        float tw[V][4];
#pragma unroll
        for (int v = 0; v < V; ++v) {
#ifdef GDB_X_LINPROJ
          const float* L = scam + C::LIN_FLOATS_OFF + v * 12;
          const float ix = fmaf(zr, fmaf(x, L[0], fmaf(y, L[1], L[2])), L[9]);
          const float iy = fmaf(zr, fmaf(x, L[3], fmaf(y, L[4], L[5])), L[10]);
          const float iz = fmaxf(fmaf(zr, fmaf(x, L[6], fmaf(y, L[7], L[8])), L[11]), 1e-6f);
#else
          const float* cv = head + CAM_HEAD + CAM_VIEW * v;
          float cx = fmaf(wx, cv[CV_E + 0], fmaf(wy, cv[CV_E + 1], fmaf(wz, cv[CV_E + 2], cv[CV_E + 3])));
          float cy = fmaf(wx, cv[CV_E + 4], fmaf(wy, cv[CV_E + 5], fmaf(wz, cv[CV_E + 6], cv[CV_E + 7])));
          float cz = fmaf(wx, cv[CV_E + 8], fmaf(wy, cv[CV_E + 9], fmaf(wz, cv[CV_E + 10], cv[CV_E + 11])));
          float ix = fmaf(cx, cv[CV_K + 0], fmaf(cy, cv[CV_K + 1], cz * cv[CV_K + 2]));
          float iy = fmaf(cx, cv[CV_K + 3], fmaf(cy, cv[CV_K + 4], cz * cv[CV_K + 5]));
          float iz = fmaxf(fmaf(cx, cv[CV_K + 6], fmaf(cy, cv[CV_K + 7], cz * cv[CV_K + 8])), 1e-6f);
#endif
#ifdef GDB_X_RCPA
          const float rz = rcp_approx(iz);
#else
          const float rz = __frcp_rn(iz);
#endif
          float gx = (ix * rz) * two_W - 1.f, gy = (iy * rz) * two_H - 1.f;
          const Bilin bl4 = bilin_border(gx, gy, p.W, p.H);
          const float4* ib = reinterpret_cast<const float4*>(p.rgba) + (size_t)(b * V + v) * p.H * p.W;
#ifdef GDB_X_UNCOND
          t[v][0] = __ldg(ib + bl4.o00); t[v][1] = __ldg(ib + bl4.o10); t[v][2] = __ldg(ib + bl4.o01); t[v][3] = __ldg(ib + bl4.o11);
#else
          t[v][0] = ldg4_if(ib + bl4.o00, actr);
          t[v][1] = ldg4_if(ib + bl4.o10, actr);
          t[v][2] = ldg4_if(ib + bl4.o01, actr);
          t[v][3] = ldg4_if(ib + bl4.o11, actr);
#endif
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
      __syncwarp();           // stash reads complete before the next tile's descriptors overwrite X
    } else {
    auto stash = [&](int t) { return reinterpret_cast<float4*>(gsm + (t >> 5) * CHB + wq * 512 + (t & 31) * 16); };
#pragma unroll 1
    for (int it = 0; it < BB; ++it) {
      const int item = it * 32 + lane;
      const int r = item / BB, j = item - r * BB;
      const float zr = __shfl_sync(full, z, r);
      const float x0r = __shfl_sync(full, geo.x0, r), y0r = __shfl_sync(full, geo.y0, r);
      const float wr = __shfl_sync(full, wgt, r);
      const bool actr = __shfl_sync(full, active ? 1 : 0, r) != 0;
      const int64_t srow_r = TAPS ? __shfl_sync(full, srow, r) : 0;
      float wvr[V];
#pragma unroll
      for (int v = 0; v < V; ++v) wvr[v] = __shfl_sync(full, wv[v], r);
      const float x = x0r + (float)(j % BS), y = y0r + (float)(j / BS);
#ifndef GDB_X_LINPROJ
      const float* M = head + CAM_M;
      const float dx = fmaf(x, M[0], fmaf(y, M[1], M[2]));
      const float dy = fmaf(x, M[3], fmaf(y, M[4], M[5]));
      const float dz = fmaf(x, M[6], fmaf(y, M[7], M[8]));
      const float wx = fmaf(dx, zr, ox), wy = fmaf(dy, zr, oy), wz = fmaf(dz, zr, oz);
#endif
      float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
#ifdef GDB_X_LINPROJ
        const float* L = scam + C::LIN_FLOATS_OFF + v * 12;
        const float ix = fmaf(zr, fmaf(x, L[0], fmaf(y, L[1], L[2])), L[9]);
        const float iy = fmaf(zr, fmaf(x, L[3], fmaf(y, L[4], L[5])), L[10]);
        const float iz = fmaxf(fmaf(zr, fmaf(x, L[6], fmaf(y, L[7], L[8])), L[11]), 1e-6f);
#else
        const float* cv = head + CAM_HEAD + CAM_VIEW * v;
        float cx = fmaf(wx, cv[CV_E + 0], fmaf(wy, cv[CV_E + 1], fmaf(wz, cv[CV_E + 2], cv[CV_E + 3])));
        float cy = fmaf(wx, cv[CV_E + 4], fmaf(wy, cv[CV_E + 5], fmaf(wz, cv[CV_E + 6], cv[CV_E + 7])));
        float cz = fmaf(wx, cv[CV_E + 8], fmaf(wy, cv[CV_E + 9], fmaf(wz, cv[CV_E + 10], cv[CV_E + 11])));
        float ix = fmaf(cx, cv[CV_K + 0], fmaf(cy, cv[CV_K + 1], cz * cv[CV_K + 2]));
        float iy = fmaf(cx, cv[CV_K + 3], fmaf(cy, cv[CV_K + 4], cz * cv[CV_K + 5]));
        float iz = fmaxf(fmaf(cx, cv[CV_K + 6], fmaf(cy, cv[CV_K + 7], cz * cv[CV_K + 8])), 1e-6f);
#endif
#ifdef GDB_X_RCPA
        const float rz = rcp_approx(iz);
#else
        const float rz = fdiv(1.f, iz);
#endif
        float gx = (ix * rz) * two_W - 1.f, gy = (iy * rz) * two_H - 1.f;
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (actr) {
          Bilin bl4 = bilin_border(gx, gy, p.W, p.H);
          const float* ib = p.rgba + (size_t)(b * V + v) * p.H * p.W * 4;
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o00 * 4), bl4.w00);
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o10 * 4), bl4.w10);
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o01 * 4), bl4.w01);
          c4 = f4_scale_add(c4, ldg4(ib + (size_t)bl4.o11 * 4), bl4.w11);
          if (TAPS && p.tap_rfd) {
            float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow_r) * C::RFD;
            tp[0 * BB + j] = c4.x; tp[1 * BB + j] = c4.y; tp[2 * BB + j] = c4.z;
          }
        }
        cr = fmaf(c4.x, wvr[v], cr); cg = fmaf(c4.y, wvr[v], cg); cb = fmaf(c4.z, wvr[v], cb);
      }
      if (TAPS && p.tap_feat && actr) {
        float* tfr = p.tap_feat + srow_r * CT;
        tfr[0 * BB + j] = cr; tfr[1 * BB + j] = cg; tfr[2 * BB + j] = cb;
      }
      *stash(r * BB + j) = make_float4(wr * cr, wr * cg, wr * cb, 0.f);
    }
    __syncwarp();
    // sum over the samples of a bundle in slot order, thread = (bundle, ray)
#pragma unroll 1
    for (int base = 0; base < G * BB; base += 32) {
      const int item = base + lane;
      const int bb = min(item / BB, G - 1), j = item % BB;
      const int nb = __shfl_sync(full, n, bb * ns);                 // every lane of a bundle holds its count
      const int pixb = pix_warp0 + bb;
      if (item < G * BB && pixb < pix_hi) {
        float4 a = *stash((bb * ns) * BB + j);
        for (int k = 1; k < nb; ++k) {
          const float4 o = *stash((bb * ns + k) * BB + j);
          a.x += o.x; a.y += o.y; a.z += o.z;
        }
        float* ofb = p.out_cl ? p.out_feat + (size_t)(b * HW + pixb) * R : p.out_feat + (size_t)b * CT * HW + pixb;
        ofb[(size_t)(0 * BB + j) * ostr] = a.x;
        ofb[(size_t)(1 * BB + j) * ostr] = a.y;
        ofb[(size_t)(2 * BB + j) * ostr] = a.z;
      }
    }
    __syncwarp();           // stash reads complete before the next tile's gathers overwrite X / FD
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

template <int BS, int FEAT_DIM, int V, int NG, bool TAPS, int GEN, int FB>
static int launch_render_tc2_t(const RenderParams& p, cudaStream_t st) {
  using C = Tc2Cfg<BS, FEAT_DIM, V, NG>;
  static_assert(C::SMEM <= 227 * 1024, "shared memory plan");
  auto kern = render_tc2_kernel<BS, FEAT_DIM, V, NG, TAPS, GEN, FB>;
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, kern, C::SMEM);
    if (e != cudaSuccess) return fail((int)e, "gdb_render_fused_fwd(tc2): cudaFuncSetAttribute(%d B): %s", C::SMEM, cudaGetErrorString(e));
  }
  if ((long)p.Wb * C::QL >= (1 << 14))
    return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_fwd(tc2): bundle map width %d too large for the packed tap stride", p.Wb);
  const int G = 32 / p.max_samples;
  const long NB = (long)p.B * (p.pix_hi - p.pix_lo);
  const long tiles = (long)p.B * ((p.pix_hi - p.pix_lo + 4 * G - 1) / (4 * G));
  (void)NB;
  long ctas = (tiles + NG - 1) / NG;
  if (ctas > sm_count()) ctas = sm_count();
  // dynamic tile assignment where it was measured to pay: 2x2 bundles (DTU 0.96 -> 0.89 ms; the 4x4 tiles of NeRF-synthetic,
  // twice as long and in 54 rounds instead of 28, lose 0.4 % to the extra barrier-side traffic)
  RenderParams q = p;
  static int dyn4 = -1;
  if (dyn4 < 0) { const char* e = getenv("GDB_K3_DYN4"); dyn4 = (e && e[0] == '1') ? 1 : 0; }     // A/B: the counter for 4x4 bundles too
  q.tile_counter = (GEN == 3 && (BS == 2 || dyn4) && ctas == sm_count()) ? acquire_tile_counter(st) : nullptr;
  kern<<<(int)ctas, 128 * NG, C::SMEM, st>>>(q);
  return cuda_check("gdb_render_fused_fwd(tc2)");
}
template <int BS, int FEAT_DIM, int V, int NG, int GEN, int FB>
static int launch_render_tc2(const RenderParams& p, cudaStream_t st) {
  const bool taps = p.tap_rfd || p.tap_vox || p.tap_sigma || p.tap_feat || p.tap_w;
  return taps ? launch_render_tc2_t<BS, FEAT_DIM, V, NG, true, GEN, FB>(p, st) : launch_render_tc2_t<BS, FEAT_DIM, V, NG, false, GEN, FB>(p, st);
}

// gen = 2: the round-1 kernel (A/B reference, `precision = 4` of the C ABI); gen = 3: the batched-gather kernel (default).
// The batching mode of the feature fetch (FB, see the kernel's header comment): measured best is one view at a time (FB = 0) for
// 2x2 bundles (16 warps per SM at 128 registers: 0.95 / 1.07 ms at DTU with FB = 0 / 2) and all 8 V taps of an iteration in flight
// (FB = 2) for 4x4 bundles (8 warps per SM at 255 registers: 2.36 / 2.24 ms at NeRF 800x800); GDB_K3_FB = 0 / 1 / 2 overrides
// it for V = 3.
int render_tc2_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, int gen, cudaStream_t st) {
  static int fb_env = -2;
  if (fb_env == -2) {
    const char* e = getenv("GDB_K3_FB");
    fb_env = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : -1;
  }
  if (gen == 2) {
    if (bundle_size == 2 && feat_dim == 16 && V == 2) return launch_render_tc2<2, 16, 2, 4, 2, 0>(p, st);
    if (bundle_size == 2 && feat_dim == 16 && V == 3) return launch_render_tc2<2, 16, 3, 4, 2, 0>(p, st);
    if (bundle_size == 2 && feat_dim == 16 && V == 4) return launch_render_tc2<2, 16, 4, 2, 2, 0>(p, st);
    if (bundle_size == 4 && feat_dim == 32 && V == 2) return launch_render_tc2<4, 32, 2, 2, 2, 0>(p, st);
    if (bundle_size == 4 && feat_dim == 32 && V == 3) return launch_render_tc2<4, 32, 3, 2, 2, 0>(p, st);
    if (bundle_size == 4 && feat_dim == 32 && V == 4) return launch_render_tc2<4, 32, 4, 1, 2, 0>(p, st);
  } else {
    if (V == 3) {
      const int fb = fb_env >= 0 ? fb_env : (bundle_size == 4 ? 2 : 0);
      if (bundle_size == 2 && feat_dim == 16 && fb) return fb == 1 ? launch_render_tc2<2, 16, 3, 4, 3, 1>(p, st) : launch_render_tc2<2, 16, 3, 4, 3, 2>(p, st);
      if (bundle_size == 4 && feat_dim == 32 && fb) return fb == 1 ? launch_render_tc2<4, 32, 3, 2, 3, 1>(p, st) : launch_render_tc2<4, 32, 3, 2, 3, 2>(p, st);
    }
    if (bundle_size == 2 && feat_dim == 16 && V == 2) return launch_render_tc2<2, 16, 2, 4, 3, 0>(p, st);
    if (bundle_size == 2 && feat_dim == 16 && V == 3) return launch_render_tc2<2, 16, 3, 4, 3, 0>(p, st);
    if (bundle_size == 2 && feat_dim == 16 && V == 4) {
      static int ng_env = -1;
      if (ng_env < 0) { const char* e = getenv("GDB_K3_NG_V4"); ng_env = (e && e[0] == '2') ? 2 : 3; }
      return ng_env == 2 ? launch_render_tc2<2, 16, 4, 2, 3, 0>(p, st) : launch_render_tc2<2, 16, 4, 3, 3, 0>(p, st);
    }
    if (bundle_size == 4 && feat_dim == 32 && V == 2) return launch_render_tc2<4, 32, 2, 2, 3, 0>(p, st);
    if (bundle_size == 4 && feat_dim == 32 && V == 3) return launch_render_tc2<4, 32, 3, 2, 3, 0>(p, st);
    if (bundle_size == 4 && feat_dim == 32 && V == 4) return launch_render_tc2<4, 32, 4, 1, 3, 0>(p, st);
  }
  return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_fwd(tc2): (bundle_size=%d, feat_dim=%d, V=%d) not instantiated", bundle_size, feat_dim, V);
}

}  // namespace gdb
