// K3 (tensor-core variant): the fused per-bundle render with the MLP GEMMs on the
// 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in
// TMEM).  Everything that is not a GEMM - sample placement, projections,
// gathers, softmax over views, compositing - stays SIMT, as in gdb_render.cu.
//
// Reference: bundle_sampler.py:193-371, nerf.py:58-115, utils.py:19-43,88-121.
//
// Structure.  One CTA per SM, NG independent groups of 128 threads.  A group
// owns a tile of 128 sample rows (thread = row = TMEM lane): 4 warps x
// floor(32/max_samples) bundles.  Per tile the group alternates SIMT phases that
// write the next A operand (fp16, UMMA "interleaved" K-major core-matrix layout:
// [K/8 chunks][128 rows][8 halves], one conflict-free STS.128 per thread per
// chunk) with GEMM phases issued by the group's thread 0:
//
//   G_v   = [var|mean] W_gs + x_v W_gx          (3 views, N=32)   -> agg softmax -> im
//   FC    = im W_fc                              (N=16)            -> img
//   LR0   = [vox|img] W_lr0                      (N=64)            -> h
//   SH    = h [w_sigma | W_feat_head]            (N=16)            -> sigma, geometry head
//   W0_v  = [h|vox|img] W_0s + [featrgb_v|dir_v] W_0v  (3 views, N=64) -> view weights
//
// Completion is signalled through tcgen05.commit -> mbarrier; accumulators are
// read back with tcgen05.ld (32x32b: thread i of warp w reads lane 32*(w%4)+i).
// Two groups ping-pong on one SM (their phases interleave freely), sharing one
// copy of the weights in shared memory and splitting the 512 TMEM columns.
//
// SPLIT = true is the fp32-class variant of the same kernel: every operand is
// stored as TWO fp16 planes, x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 significant bits), and every K step issues three MMAs, hi*hi + hi*lo +
// lo*hi, into the same fp32 TMEM accumulator (the dropped lo*lo term is 2^-22
// relative).  To keep two groups resident with twice the operand bytes the
// operand regions are aliased by lifetime and weight.0 runs in two passes
// ([h|vox|img] slice, then the per-view slice, same accumulator).
#include <cuda_fp16.h>

#include "gdb_render_common.cuh"

#include "gdb_tcgen05.cuh"

namespace gdb {

// ------------------------------------------------------------ configuration --
template <int BS, int FEAT_DIM, int V, bool SPLIT>
struct TcCfg {
  using ML = MlpLayout<FEAT_DIM>;
  static constexpr int BB = BS * BS;
  static constexpr int F = ML::F;
  static constexpr int FP = ML::FP;
  static constexpr int R = 3 * BB;
  static constexpr int CT = R + F + 8;
  static constexpr int RFD = R + F + 4;
  static constexpr int P = SPLIT ? 2 : 1;                    // fp16 planes per operand
  // K extents (multiples of 16) of the A operands
  static constexpr int K_GS = ((2 * F + 15) / 16) * 16;      // [var | mean]
  static constexpr int K_GX = ((F + 15) / 16) * 16;          // x_v
  static constexpr int K_IM = 32;
  static constexpr int K_HVI = 96;                           // [h(64) | vox(8) | img(16) | 0(8)]
  static constexpr int K_FD = ((F + 4 + 15) / 16) * 16;      // [featrgb_v | dir_v]
  // fp16 weight matrices in shared memory (bytes), UMMA B layout [K/8][N][8]; the lo plane follows at +W_PLANE
  static constexpr int W_GS = 0;
  static constexpr int W_GX = W_GS + 32 * K_GS * 2;
  static constexpr int W_FC = W_GX + 32 * K_GX * 2;
  static constexpr int W_LR0 = W_FC + 16 * K_IM * 2;
  static constexpr int W_SH = W_LR0 + 64 * 32 * 2;
  static constexpr int W_0S = W_SH + 16 * 64 * 2;
  static constexpr int W_0V = W_0S + 64 * K_HVI * 2;
  static constexpr int W_PLANE = W_0V + 64 * K_FD * 2;
  static constexpr int W_END = W_PLANE * P;
  // fp32 vectors (floats, after the matrices)
  static constexpr int X_VIEW_W = 0;                 // [4][FP]
  static constexpr int X_VIEW_B = X_VIEW_W + 4 * FP;
  static constexpr int X_GLOB_B = X_VIEW_B + FP;     // 32
  static constexpr int X_AGG_W = X_GLOB_B + 32;      // 32
  static constexpr int X_FC_B = X_AGG_W + 32;        // 16
  static constexpr int X_LR0_B = X_FC_B + 16;        // 64
  static constexpr int X_W0_B = X_LR0_B + 64;        // 64
  static constexpr int X_W2_W = X_W0_B + 64;         // 64
  static constexpr int X_FH_B = X_W2_W + 64;         // 8
  static constexpr int X_SCAL = X_FH_B + 8;          // agg_b, sig_b, w2_b, pad
  static constexpr int X_END = X_SCAL + 4;
  static constexpr int VEC_OFF = ((W_END + 127) / 128) * 128;
  static constexpr int GROUP_OFF = ((VEC_OFF + X_END * 4 + 127) / 128) * 128;
  // per-group A operands (bytes); a chunk is 128 rows x 16 B = 2 KB; the lo plane of a region follows its hi plane.
  //   region S : [var|mean] (GEMM 1), then the aggregated 32-vector (GEMM 2)
  //   region X : x_v of all views (GEMM 1), then [h|vox|img|0] (GEMM 3, 4) and - two-pass mode - [featrgb_v|dir_v] (GEMM 4b)
  //   region D : [featrgb_v|dir_v], one-pass mode only (written with the gathers, read by GEMM 4)
  static constexpr bool TWO_PASS = SPLIT;
  static constexpr int CH_S = K_GS / 8;
  static constexpr int CH_X0 = V * (K_GX / 8) > K_HVI / 8 ? V * (K_GX / 8) : K_HVI / 8;
  static constexpr int CH_X = (TWO_PASS && V * (K_FD / 8) > CH_X0) ? V * (K_FD / 8) : CH_X0;
  static constexpr int CH_D = TWO_PASS ? 0 : V * (K_FD / 8);
  static constexpr int LO_S = CH_S * 2048, LO_X = CH_X * 2048, LO_D = CH_D * 2048;       // plane sizes = lo offsets
  static constexpr int A_GS = 0;
  static constexpr int A_X = A_GS + P * LO_S;
  static constexpr int A_GX = A_X;
  static constexpr int A_HVI = A_X;
  static constexpr int A_FD = TWO_PASS ? A_X : A_X + P * LO_X;
  static constexpr int LO_FD = TWO_PASS ? LO_X : LO_D;
  static constexpr int A_END = A_X + P * LO_X + P * LO_D;
  static constexpr int CAM_OFF = A_END + 128;                  // after the mbarrier: camera block of the tile's view
  static constexpr int GROUP_BYTES = CAM_OFF + ((CAM_HEAD + CAM_VIEW * V) * 4 + 127) / 128 * 128;
  // fine colours are gathered with the other taps (before the MLP) and kept in registers when they fit
  static constexpr bool EARLY_RGB = (3 * BB * V <= 48) && !SPLIT;
  // TMEM columns per group
  static constexpr int T_W0 = 0;      // V x 64
  static constexpr int T_LR0 = 0;     // 64 (consumed before W0 is issued)
  static constexpr int T_G = 64;      // V x 32 (consumed before W0 is issued)
  static constexpr int T_SH = 64 * V; // 16
  static constexpr int T_FC = 64 * V + 16;
  static constexpr int T_COLS = (64 * V + 32 <= 256) ? 256 : 512;
  static constexpr int NG = (GROUP_OFF + 2 * GROUP_BYTES <= 227 * 1024 && T_COLS == 256) ? 2 : 1;
  static constexpr int SMEM = GROUP_OFF + NG * GROUP_BYTES;
  static_assert(64 * V + 32 <= T_COLS && 64 + 32 * V <= 64 * V, "TMEM column plan");
  static_assert(SMEM <= 227 * 1024, "shared memory plan");
  static_assert(K_IM / 8 <= CH_S, "A_IM aliases A_GS");
};

// weights: fp32 packed [K][N] block (global) -> fp16 UMMA B operand [K/8][N][8] in shared memory
template <bool SPLIT>
__device__ __forceinline__ void stage_b(unsigned char* dst, int lo_off, const float* __restrict__ src, int src_row0, int rows_valid, int N,
                                        int Kpad, int tid, int nthreads) {
  for (int i = tid; i < (Kpad / 8) * N; i += nthreads) {
    int c = i / N, n = i % N;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = c * 8 + j;
      v[j] = k < rows_valid ? __ldg(src + (size_t)(src_row0 + k) * N + n) : 0.f;
    }
    if constexpr (SPLIT) {
      uint4 hi, lo;
      split8(v, hi, lo);
      *reinterpret_cast<uint4*>(dst + (size_t)i * 16) = hi;
      *reinterpret_cast<uint4*>(dst + lo_off + (size_t)i * 16) = lo;
    } else {
      uint4 q = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
      *reinterpret_cast<uint4*>(dst + (size_t)i * 16) = q;
    }
  }
}

// D (+)= A B over `ksteps` K steps of 16.  SPLIT: hi*hi + hi*lo + lo*hi per step (lo planes at a_addr + a_lo, b_addr + b_lo).
template <bool SPLIT>
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_addr, uint32_t a_lo, uint32_t b_addr, uint32_t b_lo, int N, int ksteps,
                                           uint32_t accumulate) {
  const uint32_t idesc = umma_idesc_f16(N);
  for (int ks = 0; ks < ksteps; ++ks) {
    uint64_t ad = umma_desc(a_addr + ks * 2 * 2048, 2048, 128);
    uint64_t bd = umma_desc(b_addr + ks * 2 * (N * 16), N * 16, 128);
    umma_f16(d_tmem, ad, bd, idesc, (ks > 0 || accumulate) ? 1u : 0u);
    if constexpr (SPLIT) {
      uint64_t adl = umma_desc(a_addr + a_lo + ks * 2 * 2048, 2048, 128);
      uint64_t bdl = umma_desc(b_addr + b_lo + ks * 2 * (N * 16), N * 16, 128);
      umma_f16(d_tmem, ad, bdl, idesc, 1u);
      umma_f16(d_tmem, adl, bd, idesc, 1u);
    }
  }
}

template <int BS, int FEAT_DIM, int V, bool SPLIT>
__global__ void __launch_bounds__(256, 1) render_tc_kernel(const RenderParams p) {
  using C = TcCfg<BS, FEAT_DIM, V, SPLIT>;
  using ML = typename C::ML;
  constexpr int BB = C::BB, F = C::F, FP = C::FP, R = C::R, CT = C::CT;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  float* vec = reinterpret_cast<float*>(smem + C::VEC_OFF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 7;                      // group
  const int row = tid & 127;                   // sample row == TMEM lane
  const int ngroups_cta = blockDim.x >> 7;
  unsigned char* gsm = smem + C::GROUP_OFF + (size_t)g * C::GROUP_BYTES;
  const uint32_t mbar = smem_u32(gsm + C::A_END);

  // ---- one-time setup: weights -> smem (fp16 B operands + fp32 vectors), mbarriers, TMEM
  {
    const float* m = p.mlp;
    stage_b<SPLIT>(smem + C::W_GS, C::W_PLANE, m + ML::GLOB_W, F, 2 * F, 32, C::K_GS, tid, blockDim.x);
    stage_b<SPLIT>(smem + C::W_GX, C::W_PLANE, m + ML::GLOB_W, 0, F, 32, C::K_GX, tid, blockDim.x);
    stage_b<SPLIT>(smem + C::W_FC, C::W_PLANE, m + ML::FC_W, 0, 32, 16, C::K_IM, tid, blockDim.x);
    stage_b<SPLIT>(smem + C::W_LR0, C::W_PLANE, m + ML::LR0_W, 0, 24, 64, 32, tid, blockDim.x);
    stage_b<SPLIT>(smem + C::W_0S, C::W_PLANE, m + ML::W0_W, 0, 88, 64, C::K_HVI, tid, blockDim.x);
    stage_b<SPLIT>(smem + C::W_0V, C::W_PLANE, m + ML::W0_W, 88, F + 4, 64, C::K_FD, tid, blockDim.x);
    // [sigma | feat_head] as one N=16 operand: n=0 sigma, n=1..8 geometry head
    for (int i = tid; i < 8 * 16; i += blockDim.x) {
      int c = i / 16, n = i % 16;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int k = c * 8 + j;
        v[j] = n == 0 ? __ldg(m + ML::SIG_W + k) : (n <= 8 ? __ldg(m + ML::FH_W + k * 8 + (n - 1)) : 0.f);
      }
      if constexpr (SPLIT) {
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(smem + C::W_SH + (size_t)i * 16) = hi;
        *reinterpret_cast<uint4*>(smem + C::W_SH + C::W_PLANE + (size_t)i * 16) = lo;
      } else {
        uint4 q = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        *reinterpret_cast<uint4*>(smem + C::W_SH + (size_t)i * 16) = q;
      }
    }
    for (int i = tid; i < 5 * FP; i += blockDim.x) vec[C::X_VIEW_W + i] = m[ML::VIEW_W + i];   // W [4][FP] + b [FP]
    for (int i = tid; i < 32; i += blockDim.x) { vec[C::X_GLOB_B + i] = m[ML::GLOB_B + i]; vec[C::X_AGG_W + i] = m[ML::AGG_W + i]; }
    for (int i = tid; i < 16; i += blockDim.x) vec[C::X_FC_B + i] = m[ML::FC_B + i];
    for (int i = tid; i < 64; i += blockDim.x) {
      vec[C::X_LR0_B + i] = m[ML::LR0_B + i]; vec[C::X_W0_B + i] = m[ML::W0_B + i]; vec[C::X_W2_W + i] = m[ML::W2_W + i];
    }
    for (int i = tid; i < 8; i += blockDim.x) vec[C::X_FH_B + i] = m[ML::FH_B + i];
    if (tid == 0) { vec[C::X_SCAL + 0] = m[ML::AGG_B]; vec[C::X_SCAL + 1] = m[ML::SIG_B]; vec[C::X_SCAL + 2] = m[ML::W2_B]; }
    if (row == 0) mbar_init(mbar, 1);
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                   "r"(C::T_COLS * C::NG)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_group = tmem_base_s + g * C::T_COLS;                  // column offset of my group
  const uint32_t tmem_row = tmem_group + ((uint32_t)((warp & 3) * 32) << 16);  // my warp's lane quarter
  uint32_t parity = 0;

  const int HW = p.Hb * p.Wb;
  const int ns = p.max_samples;
  const int G = 32 / ns;                                   // bundles per warp
  const int tiles_pv = (HW + 4 * G - 1) / (4 * G);         // tiles per target view (a tile = 4 warps x G bundles, one view)
  const int tiles = p.B * tiles_pv;
  const int bl = lane / ns, slot = lane - bl * ns;
  const int seg_base = bl * ns;
  const int wq = warp & 3;
  float* scam = reinterpret_cast<float*>(gsm + C::CAM_OFF);
  int cur_b = -1;

  for (int tile = blockIdx.x * ngroups_cta + g; tile < tiles; tile += gridDim.x * ngroups_cta) {
    const int b = tile / tiles_pv;                         // uniform over the group
    if (b != cur_b) {                                      // stage this view's camera block (128 floats for V = 3)
      group_sync(g);
      for (int i = row; i < CAM_HEAD + CAM_VIEW * V; i += 128) scam[i] = p.cam[(size_t)b * p.cam_stride + i];
      group_sync(g);
      cur_b = b;
    }
    const int pix_raw = ((tile - b * tiles_pv) * 4 + wq) * G + bl;
    const bool has_bundle = bl < G && pix_raw < HW;
    const int pix = has_bundle ? pix_raw : 0;
    const int bidx = b * HW + pix;
    const int bundle = bidx;
    const int yb = pix / p.Wb, xb = pix - yb * p.Wb;
    const float* head = scam;

    float nr = p.depth_range[(size_t)(b * 2 + 0) * HW + pix], fr_ = p.depth_range[(size_t)(b * 2 + 1) * HW + pix];
    float vn = p.vol_range[(size_t)(b * 2 + 0) * HW + pix], vf = p.vol_range[(size_t)(b * 2 + 1) * HW + pix];
    const int n = bundle_sample_count(nr, fr_, head[CAM_MINIV], ns, p.inv_depth, p.adaptive);
    if (p.inv_depth) { nr = fdiv(1.f, nr); fr_ = fdiv(1.f, fr_); vn = fdiv(1.f, vn); vf = fdiv(1.f, vf); }
    const bool active = has_bundle && slot < n;
    float z, dnorm;
    sample_depth(nr, fr_, vn, vf, n, slot, p.inv_depth, z, dnorm);
    BundleGeom<BS> geo;
    geo.init(head, yb, xb, p.H, p.W);
    const float ox = head[CAM_O + 0], oy = head[CAM_O + 1], oz = head[CAM_O + 2];
    const int64_t srow = (p.offsets && active) ? (int64_t)p.offsets[bundle] + slot : -1;

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

    // ---- voxel feature (bundle_sampler.py:322-324) -> chunk 8 of [h|vox|img|0]
    float vox[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) vox[k] = 0.f;
    if (active) {
      float ix = fminf(fmaxf(((geo.u + 1.f) * (float)p.Wb - 1.f) * 0.5f, 0.f), (float)(p.Wb - 1));
      float iy = fminf(fmaxf(((geo.v + 1.f) * (float)p.Hb - 1.f) * 0.5f, 0.f), (float)(p.Hb - 1));
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
      vox[0] = lo.x; vox[1] = lo.y; vox[2] = lo.z; vox[3] = lo.w; vox[4] = hi.x; vox[5] = hi.y; vox[6] = hi.z; vox[7] = hi.w;
      if (p.tap_vox) {
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[0] = lo;
        reinterpret_cast<float4*>(p.tap_vox + srow * 8)[1] = hi;
      }
    }

    // ---- per view: level of detail, mip-mapped feature fetch, direction features; registers keep fp32 copies
    float fr[V][FP];      // feature + rgb per view (fp32, for the final blend)
    float xv[V][FP];      // after view_fc residual
#pragma unroll
    for (int v = 0; v < V; ++v) {
#pragma unroll
      for (int c = 0; c < FP; ++c) { fr[v][c] = 0.f; xv[v][c] = 0.f; }
    }
    float col[C::EARLY_RGB ? V : 1][C::EARLY_RGB ? R : 1];     // fine colours per view (channel-major, then ray)
    float dirs[C::TWO_PASS ? V : 1][4];                        // direction features, kept for the second weight.0 pass
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float* cv = head + CAM_HEAD + CAM_VIEW * v;
      float ccx = 0.f, ccy = 0.f, ccz = 0.f;
#pragma unroll
      for (int j = 0; j < BB; ++j) {
        float dx, dy, dz;
        geo.ray_dir(head, j, dx, dy, dz);
        float wx = fmaf(dx, z, ox), wy = fmaf(dy, z, oy), wz = fmaf(dz, z, oz);
        float cx = fmaf(wx, cv[CV_E + 0], fmaf(wy, cv[CV_E + 1], fmaf(wz, cv[CV_E + 2], cv[CV_E + 3])));
        float cy = fmaf(wx, cv[CV_E + 4], fmaf(wy, cv[CV_E + 5], fmaf(wz, cv[CV_E + 6], cv[CV_E + 7])));
        float cz = fmaf(wx, cv[CV_E + 8], fmaf(wy, cv[CV_E + 9], fmaf(wz, cv[CV_E + 10], cv[CV_E + 11])));
        ccx += cx; ccy += cy; ccz += cz;
        if constexpr (C::EARLY_RGB) {
          // fine colour of ray j in view v (bundle_sampler.py:327-337): issued together with the other gathers
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
          col[v][0 * BB + j] = c4.x; col[v][1 * BB + j] = c4.y; col[v][2 * BB + j] = c4.z;
        }
      }
      ccx *= (1.f / BB); ccy *= (1.f / BB); ccz *= (1.f / BB);
      float dist = sqrtf(ccx * ccx + ccy * ccy + ccz * ccz);
      float sec = dist / ccz;
      float sec_sq = sec * sec;
      float rb = dist / ball;
      float foot = sec_sq / (sqrtf(fmaxf(rb * rb - 1.f, 1e-12f)) + sqrtf(fmaxf(sec_sq - 1.f, 1e-12f)));
      float lod = log2f(foot / cv[CV_PIXR]);
      const float fb = (float)BS;
      float k0 = cv[CV_K + 0] / fb, k1 = cv[CV_K + 1] / fb, k2 = cv[CV_K + 2] / fb;
      float k3 = cv[CV_K + 3] / fb, k4 = cv[CV_K + 4] / fb, k5 = cv[CV_K + 5] / fb;
      float pxc = fmaf(ccx, k0, fmaf(ccy, k1, ccz * k2));
      float pyc = fmaf(ccx, k3, fmaf(ccy, k4, ccz * k5));
      float pzc = fmaxf(fmaf(ccx, cv[CV_K + 6], fmaf(ccy, cv[CV_K + 7], ccz * cv[CV_K + 8])), 1e-6f);
      float u01 = pxc / pzc / (float)p.Wb, v01 = pyc / pzc / (float)p.Hb;
      float dir[4] = {0.f, 0.f, 0.f, 0.f};
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
          fr[v][q * 4 + 0] = a.x; fr[v][q * 4 + 1] = a.y; fr[v][q * 4 + 2] = a.z; fr[v][q * 4 + 3] = a.w;
        }
        float tx = cwx - ox, ty = cwy - oy, tz = cwz - oz;
        unit3(tx, ty, tz);
        float sx = cwx - cv[CV_C + 0], sy = cwy - cv[CV_C + 1], sz = cwz - cv[CV_C + 2];
        unit3(sx, sy, sz);
        float ddx = tx - sx, ddy = ty - sy, ddz = tz - sz;
        unit3(ddx, ddy, ddz);
        dir[0] = ddx; dir[1] = ddy; dir[2] = ddz; dir[3] = tx * sx + ty * sy + tz * sz;
        if (p.tap_rfd) {
          float* tp = p.tap_rfd + ((size_t)v * p.S_total + srow) * C::RFD + R;
#pragma unroll
          for (int c = 0; c < F; ++c) tp[c] = fr[v][c];
          tp[F + 0] = dir[0]; tp[F + 1] = dir[1]; tp[F + 2] = dir[2]; tp[F + 3] = dir[3];
        }
      }
      // view_fc + residual (nerf.py:69-71), fp32 SIMT (K = 4)
#pragma unroll
      for (int c = 0; c < F; ++c) {
        float t = vec[C::X_VIEW_B + c];
        t = fmaf(vec[C::X_VIEW_W + 0 * FP + c], dir[0], t);
        t = fmaf(vec[C::X_VIEW_W + 1 * FP + c], dir[1], t);
        t = fmaf(vec[C::X_VIEW_W + 2 * FP + c], dir[2], t);
        t = fmaf(vec[C::X_VIEW_W + 3 * FP + c], dir[3], t);
        xv[v][c] = active ? fr[v][c] + fmaxf(t, 0.f) : 0.f;
      }
      // A operands of this view: x_v and (one-pass mode) [featrgb_v | dir_v]
#pragma unroll
      for (int ch = 0; ch < C::K_GX / 8; ++ch) {
        float t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t8[j] = (ch * 8 + j < F) ? xv[v][ch * 8 + j] : 0.f;
        store_chunk_p<SPLIT>(gsm + C::A_GX + v * (C::K_GX / 8) * 2048, C::LO_X, ch, row, t8);
      }
      if constexpr (C::TWO_PASS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dirs[v][j] = dir[j];
      } else {
#pragma unroll
        for (int ch = 0; ch < C::K_FD / 8; ++ch) {
          float t8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = ch * 8 + j;
            t8[j] = k < F ? fr[v][k < F ? k : 0] : (k < F + 4 ? dir[(k - F) & 3] : 0.f);
          }
          store_chunk(gsm + C::A_FD + v * (C::K_FD / 8) * 2048, ch, row, t8);
        }
      }
    }
    // [var | mean] over views (nerf.py:73)
    {
      float vm[C::K_GS];
#pragma unroll
      for (int k = 0; k < C::K_GS; ++k) vm[k] = 0.f;
#pragma unroll
      for (int c = 0; c < F; ++c) {
        float mean = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) mean += xv[v][c];
        mean *= (1.f / V);
        float var = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) { float t = xv[v][c] - mean; var = fmaf(t, t, var); }
        vm[c] = var * (1.f / (V - 1));
        vm[F + c] = mean;
      }
#pragma unroll
      for (int ch = 0; ch < C::K_GS / 8; ++ch) {
        float t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t8[j] = vm[ch * 8 + j];
        store_chunk_p<SPLIT>(gsm + C::A_GS, C::LO_S, ch, row, t8);
      }
    }

    const uint32_t a_base = smem_u32(gsm), w_base = smem_u32(smem);
    // ================= GEMM 1: global_fc =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
#pragma unroll
      for (int v = 0; v < V; ++v) {
        issue_gemm<SPLIT>(tmem_group + C::T_G + v * 32, a_base + C::A_GS, C::LO_S, w_base + C::W_GS, C::W_PLANE, 32, C::K_GS / 16, 0);
        issue_gemm<SPLIT>(tmem_group + C::T_G + v * 32, a_base + C::A_GX + v * (C::K_GX / 8) * 2048, C::LO_X, w_base + C::W_GX, C::W_PLANE, 32,
                          C::K_GX / 16, 1);
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
    {
      // pass 1: aggregation logits (TMEM reads are cheap; re-reading keeps only one view's 32 columns live)
      float aw[V];
      float amax = -1e30f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float gv[32];
        tmem_ld32(tmem_row + C::T_G + v * 32, gv);
        float s = vec[C::X_SCAL + 0];
#pragma unroll
        for (int k = 0; k < 32; ++k) s = fmaf(fmaxf(gv[k] + vec[C::X_GLOB_B + k], 0.f), vec[C::X_AGG_W + k], s);
        aw[v] = fmaxf(s, 0.f);
        amax = fmaxf(amax, aw[v]);
      }
      float asum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { aw[v] = expf(aw[v] - amax); asum += aw[v]; }
      // pass 2: softmax-weighted sum over views
      float im[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) im[k] = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float gv[32];
        tmem_ld32(tmem_row + C::T_G + v * 32, gv);
        float a = aw[v] / asum;
#pragma unroll
        for (int k = 0; k < 32; ++k) im[k] = fmaf(fmaxf(gv[k] + vec[C::X_GLOB_B + k], 0.f), a, im[k]);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t8[j] = im[ch * 8 + j];
        store_chunk_p<SPLIT>(gsm + C::A_GS, C::LO_S, ch, row, t8);   // A_IM aliases A_GS (its GEMM has completed)
      }
    }
    // ================= GEMM 2: fc =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      issue_gemm<SPLIT>(tmem_group + C::T_FC, a_base + C::A_GS, C::LO_S, w_base + C::W_FC, C::W_PLANE, 16, C::K_IM / 16, 0);
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
    {
      float img[16];
      tmem_ld16(tmem_row + C::T_FC, img);
#pragma unroll
      for (int k = 0; k < 16; ++k) img[k] = fmaxf(img[k] + vec[C::X_FC_B + k], 0.f);
      // region X held x_v until GEMM 1 completed; from here on it is [h | vox | img | 0]
      float t8[8];
      store_chunk_p<SPLIT>(gsm + C::A_HVI, C::LO_X, 8, row, vox);
#pragma unroll
      for (int j = 0; j < 8; ++j) t8[j] = img[j];
      store_chunk_p<SPLIT>(gsm + C::A_HVI, C::LO_X, 9, row, t8);
#pragma unroll
      for (int j = 0; j < 8; ++j) t8[j] = img[8 + j];
      store_chunk_p<SPLIT>(gsm + C::A_HVI, C::LO_X, 10, row, t8);
      *reinterpret_cast<uint4*>(gsm + C::A_HVI + 11 * 2048 + row * 16) = make_uint4(0, 0, 0, 0);
      if constexpr (SPLIT) *reinterpret_cast<uint4*>(gsm + C::A_HVI + C::LO_X + 11 * 2048 + row * 16) = make_uint4(0, 0, 0, 0);
    }
    // ================= GEMM 3: lr0 on [vox | img | 0] =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      issue_gemm<SPLIT>(tmem_group + C::T_LR0, a_base + C::A_HVI + 8 * 2048, C::LO_X, w_base + C::W_LR0, C::W_PLANE, 64, 2, 0);
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float h[32];
      tmem_ld32(tmem_row + C::T_LR0 + half * 32, h);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float t8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t8[j] = fmaxf(h[ch * 8 + j] + vec[C::X_LR0_B + half * 32 + ch * 8 + j], 0.f);
        store_chunk_p<SPLIT>(gsm + C::A_HVI, C::LO_X, half * 4 + ch, row, t8);
      }
    }
    // ================= GEMM 4: [sigma | feat_head] and weight.0 =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      issue_gemm<SPLIT>(tmem_group + C::T_SH, a_base + C::A_HVI, C::LO_X, w_base + C::W_SH, C::W_PLANE, 16, 4, 0);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        issue_gemm<SPLIT>(tmem_group + C::T_W0 + v * 64, a_base + C::A_HVI, C::LO_X, w_base + C::W_0S, C::W_PLANE, 64, C::K_HVI / 16, 0);
        if constexpr (!C::TWO_PASS)
          issue_gemm<SPLIT>(tmem_group + C::T_W0 + v * 64, a_base + C::A_FD + v * (C::K_FD / 8) * 2048, C::LO_FD, w_base + C::W_0V, C::W_PLANE,
                            64, C::K_FD / 16, 1);
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
    if constexpr (C::TWO_PASS) {
      // second pass of weight.0: [featrgb_v | dir_v] replaces [h|vox|img] in region X (its MMAs have completed)
#pragma unroll
      for (int v = 0; v < V; ++v) {
#pragma unroll
        for (int ch = 0; ch < C::K_FD / 8; ++ch) {
          float t8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = ch * 8 + j;
            t8[j] = k < F ? fr[v][k < F ? k : 0] : (k < F + 4 ? dirs[v][(k - F) & 3] : 0.f);
          }
          store_chunk_p<SPLIT>(gsm + C::A_FD + v * (C::K_FD / 8) * 2048, C::LO_FD, ch, row, t8);
        }
      }
      tc_fence_before();
      fence_async_smem();
      group_sync(g);
      if (row == 0) {
        tc_fence_after();
#pragma unroll
        for (int v = 0; v < V; ++v)
          issue_gemm<SPLIT>(tmem_group + C::T_W0 + v * 64, a_base + C::A_FD + v * (C::K_FD / 8) * 2048, C::LO_FD, w_base + C::W_0V, C::W_PLANE,
                            64, C::K_FD / 16, 1);
        umma_commit(mbar);
      }
      mbar_wait(mbar, parity); parity ^= 1;
      tc_fence_after();
    }
    float sigma, fh[8], wv[V];
    {
      float sh[16];
      tmem_ld16(tmem_row + C::T_SH, sh);
      float s = sh[0] + vec[C::X_SCAL + 1];
      sigma = s > 20.f ? s : log1pf(expf(s));
#pragma unroll
      for (int k = 0; k < 8; ++k) fh[k] = fmaxf(sh[1 + k] + vec[C::X_FH_B + k], 0.f);
      float wmax = -1e30f;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float s2 = vec[C::X_SCAL + 2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float hid[32];
          tmem_ld32(tmem_row + C::T_W0 + v * 64 + half * 32, hid);
#pragma unroll
          for (int k = 0; k < 32; ++k)
            s2 = fmaf(fmaxf(hid[k] + vec[C::X_W0_B + half * 32 + k], 0.f), vec[C::X_W2_W + half * 32 + k], s2);
        }
        wv[v] = fmaxf(s2, 0.f);
        wmax = fmaxf(wmax, wv[v]);
      }
      float wsum = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) { wv[v] = expf(wv[v] - wmax); wsum += wv[v]; }
#pragma unroll
      for (int v = 0; v < V; ++v) wv[v] /= wsum;
    }
    tc_fence_before();      // TMEM reads of this tile are ordered before the barrier that precedes the next tile's MMAs

    // ======================= compositing weights (utils.py:19-43) =======================
    const unsigned full = 0xffffffffu;
    float alpha = active ? 1.f - expf(-sigma) : 0.f;
    float one_minus = 1.f - alpha;
    float T = 1.f;
    for (int k = 0; k + 1 < ns; ++k) {
      float o = __shfl_sync(full, one_minus, min(seg_base + k, 31));
      if (k < slot) T *= o;
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

    auto seg_sum = [&](float x) {
      float acc = x;
      for (int k = 1; k < ns; ++k) {
        float o = __shfl_down_sync(full, x, k);
        if (slot == 0 && k < n) acc += o;
      }
      return acc;
    };
    const bool writer = has_bundle && slot == 0;
    const size_t ostr = p.out_cl ? 1 : (size_t)HW;
    float* of = p.out_cl ? p.out_feat + (size_t)bidx * R : p.out_feat + (size_t)b * CT * HW + pix;
    float* od = p.out_cl ? p.out_dec + (size_t)bidx * p.dec_stride - R : of;
    if (writer && p.out_cl)
      for (int k = F + 8; k < p.dec_stride; ++k) od[R + k] = k == F + 8 ? p.dec_pad0 : 0.f;                                // pad channels of the decoder input
    float* tf = (p.tap_feat && active) ? p.tap_feat + srow * CT : nullptr;

    if constexpr (C::EARLY_RGB) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        float a = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) a = fmaf(col[v][c], wv[v], a);
        if (tf) tf[c] = a;
        float s = seg_sum(wgt * a);
        if (writer) of[(size_t)c * ostr] = s;
      }
    } else {
#pragma unroll 1
    for (int j = 0; j < BB; ++j) {
      float dx, dy, dz;
      geo.ray_dir(head, j, dx, dy, dz);
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
    }
#pragma unroll
    for (int c = 0; c < F; ++c) {
      float a = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v) a = fmaf(fr[v][c], wv[v], a);
      if (tf) tf[R + c] = a;
      float s = seg_sum(wgt * a);
      if (writer) od[(size_t)(R + c) * ostr] = s;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (tf) tf[R + F + k] = fh[k];
      float s = seg_sum(wgt * fh[k]);
      if (writer) od[(size_t)(R + F + k) * ostr] = s;
    }
    {
      float zz = p.inv_depth ? fdiv(1.f, z) : z;
      float sd = seg_sum(wgt * zz), so = seg_sum(wgt);
      if (writer) {
        p.out_depth[(size_t)b * HW + pix] = p.inv_depth ? fdiv(1.f, sd) : sd;
        p.out_opacity[(size_t)b * HW + pix] = so;
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(C::T_COLS * C::NG) : "memory");
  }
}

template <int BS, int FEAT_DIM, int V, bool SPLIT>
static int launch_render_tc(const RenderParams& p, cudaStream_t st) {
  using C = TcCfg<BS, FEAT_DIM, V, SPLIT>;
  auto kern = render_tc_kernel<BS, FEAT_DIM, V, SPLIT>;
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, kern, C::SMEM);
    if (e != cudaSuccess) return fail((int)e, "gdb_render_fused_fwd(tc): cudaFuncSetAttribute(%d B): %s", C::SMEM, cudaGetErrorString(e));
  }
  const int G = 32 / p.max_samples;
  const long NB = (long)p.B * p.Hb * p.Wb;
  const long tiles = (NB + 4 * G - 1) / (4 * G);
  long ctas = (tiles + C::NG - 1) / C::NG;
  if (ctas > sm_count()) ctas = sm_count();
  kern<<<(int)ctas, 128 * C::NG, C::SMEM, st>>>(p);
  return cuda_check("gdb_render_fused_fwd(tc)");
}

int render_tc_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, int split, cudaStream_t st) {
#define GDB_R(BSZ, FD, VV)                                                                   \
  if (bundle_size == BSZ && feat_dim == FD && V == VV)                                        \
    return split ? launch_render_tc<BSZ, FD, VV, true>(p, st) : launch_render_tc<BSZ, FD, VV, false>(p, st);
  GDB_R(2, 16, 2) GDB_R(2, 16, 3) GDB_R(2, 16, 4) GDB_R(4, 32, 2) GDB_R(4, 32, 3) GDB_R(4, 32, 4)
#undef GDB_R
  return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_fwd(tc): (bundle_size=%d, feat_dim=%d, V=%d) not instantiated", bundle_size, feat_dim, V);
}

}  // namespace gdb
