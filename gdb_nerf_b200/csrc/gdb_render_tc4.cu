// K3, fp32-class arithmetic (`precision = 2`) in the third-generation layout: the production kernel of the split-fp16 MLP for
// 2x2 bundles and three source views (DTU / LLFF evaluation).
//
// Reference: bundle_sampler.py:193-371, nerf.py:58-115, utils.py:19-43,88-121.
//
// Arithmetic: every MMA operand is stored as TWO fp16 planes, x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 significant
// bits), and every K step issues three MMAs, hi*hi + hi*lo + lo*hi, into the same fp32 TMEM accumulator - exactly the scheme of
// the first-generation split kernel (gdb_render_tc.cu, SPLIT = true, which stays for 4x4 bundles, V != 3 and the parity taps).
// Everything around the GEMMs is the third-generation kernel (gdb_render_tc2.cu, GEN = 3): slot-major rows, descriptors through
// shared memory, coalesced lane = (row, quad) feature fetch with batched predicated gathers, fine colours with lane = (row, ray),
// float4 compositing, never-used rows skipped.
//
// Two planes double the operand bytes, so two tiles of 128 rows are resident per SM (8 warps, up to 255 registers) and the
// operand plan is cut from 23 to 20 chunks per plane:
//   * global_fc's `mean` input disappears: mean . W_m = sum_u x_u . (W_m / V), i.e. V extra accumulating MMAs over the X_u
//     operands that are in shared memory anyway (the tensor pipe is idle 93 % of the time);
//   * var[16:19] rides in the spare K slots 20..22 of every [x_v | 1] operand (the matching rows of that operand's weight
//     matrix hold global_fc's var rows 16..18, and zeros in the matrix of the mean term), so S = var[0:16] is 2 chunks;
//   * the aggregated 32-vector (GEMM 2's input) goes to region X, dead after GEMM 1.
//        X  [x_v | 1 | var16..18] per view   -> after GEMM 1: aggregated vector (4) -> after GEMM 3: [h (8) | vox]
//        S  var[0:16]                         -> after GEMM 2: img (2)
//        FD [featrgb_v | dir_v]               -> read again for the final blend (hi + lo: fp32-class decoder features)
#include "gdb_render_tc2.cuh"

namespace gdb {

template <int BS, int FEAT_DIM, int V>
struct Tc4Cfg {
  using ML = MlpLayout<FEAT_DIM>;
  static constexpr int NG = 2;
  static constexpr int BB = BS * BS;
  static constexpr int F = ML::F;
  static constexpr int FP = ML::FP;
  static constexpr int R = 3 * BB;
  static constexpr int CT = R + F + 8;
  static constexpr int QL = FP / 4;
  static constexpr int IPW = 32 / QL;
  static constexpr int NIT = (32 + IPW - 1) / IPW;
  static constexpr int VT = F - 16;                    // var channels that ride in the X_v operands
  static_assert(F == FP - 1 && QL % 2 == 1, "texel = F channels + one pad channel, the last quad starts a chunk");
  static_assert(VT >= 0 && VT <= 3 && F + 1 + VT <= 24, "operand plan written for 16 < F <= 19");
  static constexpr int XCH = 3;                        // [x_v (F) | 1 | var tail (VT)]
  static constexpr int FDCH = (F + 4 + 7) / 8;         // [featrgb_v (F) | dir_v (4)]
  static constexpr int SCH = 2;                        // var[0:16]
  static_assert(FDCH == 3, "operand plan");
  static constexpr int CH_X = cmax(V * XCH, 9);
  static constexpr int CH_FD = V * FDCH;
  static constexpr int CH_S = SCH;
  // fp16 weight matrices (bytes within one plane), UMMA B layout [K/8][N][8]; the lo plane follows at + W_PLANE
  static constexpr int W_GS = 0;                       // K = 16: var[0:16]
  static constexpr int W_GX = W_GS + 16 * 32 * 2;      // K = 32: x | bias | var tail
  static constexpr int W_GM = W_GX + 32 * 32 * 2;      // K = 32: mean rows / V
  static constexpr int W_FC = W_GM + 32 * 32 * 2;
  static constexpr int W_LR0 = W_FC + 32 * 16 * 2;
  static constexpr int W_SH = W_LR0 + 32 * 64 * 2;
  static constexpr int W_0S = W_SH + 64 * 16 * 2;
  static constexpr int W_0V = W_0S + 96 * 64 * 2;
  static constexpr int W_PLANE = W_0V + 32 * 64 * 2;
  // fp32 vectors (floats)
  static constexpr int X_VIEW_W = 0;
  static constexpr int X_VIEW_B = X_VIEW_W + 4 * FP;
  static constexpr int X_AGG_W = X_VIEW_B + FP;
  static constexpr int X_FC_B = X_AGG_W + 32;
  static constexpr int X_W2_W = X_FC_B + 16;
  static constexpr int X_FH_B = X_W2_W + 64;
  static constexpr int X_SCAL = X_FH_B + 8;
  static constexpr int X_END = X_SCAL + 4;
  static constexpr int VEC_OFF = 2 * W_PLANE;
  static constexpr int GROUP_OFF = ((VEC_OFF + X_END * 4 + 127) / 128) * 128;
  // per-group operand regions (bytes from the plane base); the lo plane follows at + PL
  static constexpr int A_X = 0;
  static constexpr int A_FD = A_X + CH_X * 2048;
  static constexpr int A_S = A_FD + CH_FD * 2048;
  static constexpr int PL = A_S + CH_S * 2048;
  static constexpr int MBAR_OFF = 2 * PL;
  static constexpr int CAM_OFF = MBAR_OFF + 128;
  static constexpr int GROUP_BYTES = CAM_OFF + ((CAM_HEAD + CAM_VIEW * V) * 4 + 127) / 128 * 128;
  static constexpr int ZERO_OFF = GROUP_OFF + NG * GROUP_BYTES;
  static constexpr int ONE_OFF = ZERO_OFF + 2048;
  static constexpr int SMEM = ONE_OFF + 2048;
  static constexpr int NC = F + 10;
  static constexpr int NCP = NC <= 32 ? 32 : 64;
  static_assert(32 * NCP * 4 <= 512 * (CH_X + CH_FD), "compositing stash must fit in the warp's rows of X + FD");
  static_assert(32 * R * 4 <= 512 * (CH_X + CH_FD), "colour stash must fit in the warp's rows of X + FD");
  static constexpr int TC = 256;                       // TMEM columns per group
  static_assert(V * 64 + 16 <= TC && V * 32 <= TC, "TMEM column plan");
};

// hi / lo fp16 planes of four fp32 values
__device__ __forceinline__ void split4(float a, float b, float c, float d, uint2& hi, uint2& lo) {
  const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
  const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(a - b0.x, b - b0.y), l1 = __floats2half2_rn(c - b1.x, d - b1.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}
// one 16-byte chunk row (8 values) into both planes
__device__ __forceinline__ void store8_split(unsigned char* hi_ptr, int lo_off, const float (&v)[8]) {
  uint4 hi, lo;
  split8(v, hi, lo);
  *reinterpret_cast<uint4*>(hi_ptr) = hi;
  *reinterpret_cast<uint4*>(hi_ptr + lo_off) = lo;
}

// weights: fp32 packed block (global) -> two fp16 planes of a UMMA B operand [Kpad/8][N][8]
template <class Fn>
__device__ __forceinline__ void stage_b_split(unsigned char* dst, int lo_off, int N, int Kpad, int tid, int nthreads, Fn value) {
  for (int i = tid; i < (Kpad / 8) * N; i += nthreads) {
    const int c = i / N, n = i - c * N;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = value(c * 8 + j, n);
    store8_split(dst + (size_t)i * 16, lo_off, v);
  }
}

// one K step (16) of D (+)= A B with split operands: hi*hi + hi*lo + lo*hi.  (a0, a1) / (a0l, a1l): the two K chunks of the hi /
// lo plane of A (a1 > a0); b / bl: the K step of the hi / lo plane of B.
__device__ __forceinline__ void mma3(uint32_t d_tmem, uint32_t a0, uint32_t a1, uint32_t a0l, uint32_t a1l, uint32_t b, uint32_t bl, int N,
                                     uint32_t accumulate) {
  const uint32_t idesc = umma_idesc_f16(N);
  const uint64_t ad = umma_desc(a0, a1 - a0, 128), adl = umma_desc(a0l, a1l - a0l, 128);
  const uint64_t bd = umma_desc(b, N * 16, 128), bdl = umma_desc(bl, N * 16, 128);
  umma_f16(d_tmem, ad, bd, idesc, accumulate);
  umma_f16(d_tmem, ad, bdl, idesc, 1u);
  umma_f16(d_tmem, adl, bd, idesc, 1u);
}

// FB = 1: the taps of all views of a fetch iteration are in flight at once (8 warps per SM leave 255 registers per thread)
template <int BS, int FEAT_DIM, int V, int FB>
__global__ void __launch_bounds__(256, 1) render_tc4_kernel(const RenderParams p) {
  using C = Tc4Cfg<BS, FEAT_DIM, V>;
  using ML = typename C::ML;
  constexpr int NG = C::NG, BB = C::BB, F = C::F, FP = C::FP, R = C::R, CT = C::CT, QL = C::QL, IPW = C::IPW, PL = C::PL, WP = C::W_PLANE;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  float* vec = reinterpret_cast<float*>(smem + C::VEC_OFF);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = tid >> 7;
  const int row = tid & 127;
  const int wq = warp & 3;
  unsigned char* gsm = smem + C::GROUP_OFF + (size_t)g * C::GROUP_BYTES;
  const uint32_t mbar = smem_u32(gsm + C::MBAR_OFF);
  const unsigned full = 0xffffffffu;

  // ---- one-time setup: weights -> smem (two fp16 planes + fp32 vectors), constant chunks, mbarriers, TMEM
  {
    const float* __restrict__ m = p.mlp;
    const int nt = blockDim.x;
    const float invV = 1.f / (float)V;
    stage_b_split(smem + C::W_GS, WP, 32, 16, tid, nt, [&](int k, int n) { return __ldg(m + ML::GLOB_W + (size_t)(F + k) * 32 + n); });
    stage_b_split(smem + C::W_GX, WP, 32, 32, tid, nt, [&](int k, int n) {
      return k < F ? __ldg(m + ML::GLOB_W + (size_t)k * 32 + n)
                   : (k == F ? __ldg(m + ML::GLOB_B + n)
                             : (k < F + 1 + C::VT ? __ldg(m + ML::GLOB_W + (size_t)(F + 16 + (k - F - 1)) * 32 + n) : 0.f));
    });
    stage_b_split(smem + C::W_GM, WP, 32, 32, tid, nt, [&](int k, int n) { return k < F ? __ldg(m + ML::GLOB_W + (size_t)(2 * F + k) * 32 + n) * invV : 0.f; });
    stage_b_split(smem + C::W_FC, WP, 16, 32, tid, nt, [&](int k, int n) { return __ldg(m + ML::FC_W + k * 16 + n); });
    stage_b_split(smem + C::W_LR0, WP, 64, 32, tid, nt, [&](int k, int n) {
      return k < 24 ? __ldg(m + ML::LR0_W + k * 64 + n) : (k == 24 ? __ldg(m + ML::LR0_B + n) : 0.f);
    });
    stage_b_split(smem + C::W_SH, WP, 16, 64, tid, nt, [&](int k, int n) {
      return n == 0 ? __ldg(m + ML::SIG_W + k) : (n <= 8 ? __ldg(m + ML::FH_W + k * 8 + (n - 1)) : 0.f);
    });
    stage_b_split(smem + C::W_0S, WP, 64, 96, tid, nt, [&](int k, int n) {
      return k < 88 ? __ldg(m + ML::W0_W + (size_t)k * 64 + n) : (k == 88 ? __ldg(m + ML::W0_B + n) : 0.f);
    });
    stage_b_split(smem + C::W_0V, WP, 64, 32, tid, nt, [&](int k, int n) { return k < F + 4 ? __ldg(m + ML::W0_W + (size_t)(88 + k) * 64 + n) : 0.f; });
    for (int i = tid; i < 5 * FP; i += nt) vec[C::X_VIEW_W + i] = m[ML::VIEW_W + i];
    for (int i = tid; i < 32; i += nt) vec[C::X_AGG_W + i] = m[ML::AGG_W + i];
    for (int i = tid; i < 16; i += nt) vec[C::X_FC_B + i] = m[ML::FC_B + i];
    for (int i = tid; i < 64; i += nt) vec[C::X_W2_W + i] = m[ML::W2_W + i];
    for (int i = tid; i < 8; i += nt) vec[C::X_FH_B + i] = m[ML::FH_B + i];
    if (tid == 0) { vec[C::X_SCAL + 0] = m[ML::AGG_B]; vec[C::X_SCAL + 1] = m[ML::SIG_B]; vec[C::X_SCAL + 2] = m[ML::W2_B]; }
    for (int i = tid; i < 128; i += nt) {
      *reinterpret_cast<uint4*>(smem + C::ZERO_OFF + i * 16) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(smem + C::ONE_OFF + i * 16) = make_uint4(0x3C00u, 0, 0, 0);    // fp16 1.0 in K slot 0 (its lo plane is the zero chunk)
    }
    if (row == 0) mbar_init(mbar, 1);
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_group = tmem_base_s + g * C::TC;
  const uint32_t tmem_row = tmem_group + ((uint32_t)(wq * 32) << 16);
  uint32_t parity = 0;

  const uint32_t w_base = smem_u32(smem);
  const uint32_t zero_chunk = w_base + C::ZERO_OFF, one_chunk = w_base + C::ONE_OFF;
  const uint32_t aX = smem_u32(gsm) + C::A_X, aFD = smem_u32(gsm) + C::A_FD, aS = smem_u32(gsm) + C::A_S;
  unsigned char* const sX = gsm + C::A_X;
  unsigned char* const sFD = gsm + C::A_FD;
  unsigned char* const sS = gsm + C::A_S;
  // K steps over `nch` consecutive chunks of this group's planes starting at `a` (odd counts pair with the zero chunk)
  auto mma3_chunks = [&](uint32_t d, uint32_t a, int nch, int b_off, int N, uint32_t accumulate) {
    for (int ks = 0; 2 * ks < nch; ++ks) {
      const uint32_t a0 = a + ks * 4096;
      const bool pair = 2 * ks + 1 < nch;
      const uint32_t a1 = pair ? a0 + 2048 : zero_chunk, a1l = pair ? a0 + PL + 2048 : zero_chunk;
      const uint32_t b = w_base + b_off + ks * 2 * (N * 16);
      mma3(d, a0, a1, a0 + PL, a1l, b, b + WP, N, (ks > 0 || accumulate) ? 1u : 0u);
    }
  };

  const int HW = p.Hb * p.Wb;
  const int ns = p.max_samples;
  const int G = 32 / ns;
  const int pix_lo = p.pix_lo, pix_hi = p.pix_hi;
  const int tiles_pv = (pix_hi - pix_lo + 4 * G - 1) / (4 * G);
  const int tiles = p.B * tiles_pv;
  auto rowof = [&](int bb, int k) { return k * G + bb; };     // slot-major rows
  const int slot = lane / G;
  const int bl = lane - slot * G;
  const bool lane_ok = slot < ns;
  float* scam = reinterpret_cast<float*>(gsm + C::CAM_OFF);
  const float* head = scam;
  int cur_b = -1;

  const int gr = lane / QL, gq = lane - gr * QL;
  const bool glane = lane < IPW * QL;
  const bool last_quad = gq == QL - 1;
  const float4* tex4 = reinterpret_cast<const float4*>(p.tex);
  const float two_W = 2.f / (float)p.W, two_H = 2.f / (float)p.H;

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

  // tile assignment as in gdb_render_tc2.cu: the first tile of a slot is static, the following ones come from the launch's atomic
  // counter (p.tile_counter) when there is one
  __shared__ int next_tile_s[NG];
  const bool dyn = p.tile_counter != nullptr;
  int tn = 0;
#pragma unroll 1
  for (int tile = blockIdx.x * NG + g; tile < tiles; tile = tn) {
    unsigned int drawn = 0;                                // the atomic's result is not touched before GEMM 1: its latency hides under P0-P2
    if (dyn) {
      if (row == 0) drawn = atomicAdd(p.tile_counter, 1u);
    } else {
      tn = tile + gridDim.x * NG;
    }
    const int b = tile / tiles_pv;
    if (b != cur_b) {
      group_sync(g);
      for (int i = row; i < CAM_HEAD + CAM_VIEW * V; i += 128) scam[i] = p.cam[(size_t)b * p.cam_stride + i];
      group_sync(g);
      cur_b = b;
    }
    const int pix_warp0 = pix_lo + ((tile - b * tiles_pv) * 4 + wq) * G;
    const int pix_raw = pix_warp0 + bl;
    const bool has_bundle = lane_ok && pix_raw < pix_hi;
    const int pix = has_bundle ? pix_raw : pix_lo;
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

    // ---- voxel feature (bundle_sampler.py:322-324), kept in registers until region X is free
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
      const float* vb_ = p.vol + (size_t)b * p.vol_sb;
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
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
      vox[0] = lo.x; vox[1] = lo.y; vox[2] = lo.z; vox[3] = lo.w; vox[4] = hi.x; vox[5] = hi.y; vox[6] = hi.z; vox[7] = hi.w;
    }

    // ====================== P1: per-view fetch descriptors (thread = row) ======================
    float tdx = cwx - ox, tdy = cwy - oy, tdz = cwz - oz;
    unit3_fast(tdx, tdy, tdz);
#pragma unroll 1      // rolled: less straight-line code per tile for the instruction cache (1.707 -> 1.692 ms at DTU)
    for (int v = 0; v < V; ++v) {
      const float* cv = head + CAM_HEAD + CAM_VIEW * v;
      const float ccx = fmaf(cwx, cv[CV_E + 0], fmaf(cwy, cv[CV_E + 1], fmaf(cwz, cv[CV_E + 2], cv[CV_E + 3])));
      const float ccy = fmaf(cwx, cv[CV_E + 4], fmaf(cwy, cv[CV_E + 5], fmaf(cwz, cv[CV_E + 6], cv[CV_E + 7])));
      const float ccz = fmaf(cwx, cv[CV_E + 8], fmaf(cwy, cv[CV_E + 9], fmaf(cwz, cv[CV_E + 10], cv[CV_E + 11])));
      const float dist = sqrtf(ccx * ccx + ccy * ccy + ccz * ccz);
      const float sec = __fdividef(dist, ccz);
      const float sec_sq = sec * sec;
      const float rb = __fdividef(dist, ball);
      const float foot = __fdividef(sec_sq, sqrtf(fmaxf(rb * rb - 1.f, 1e-12f)) + sqrtf(fmaxf(sec_sq - 1.f, 1e-12f)));
      const float lod = log2f(__fdividef(foot, cv[CV_PIXR]));
      constexpr float ifb = 1.f / (float)BS;
      const float pxc = fmaf(ccx, cv[CV_K + 0] * ifb, fmaf(ccy, cv[CV_K + 1] * ifb, ccz * (cv[CV_K + 2] * ifb)));
      const float pyc = fmaf(ccx, cv[CV_K + 3] * ifb, fmaf(ccy, cv[CV_K + 4] * ifb, ccz * (cv[CV_K + 5] * ifb)));
      const float pzc = fmaxf(fmaf(ccx, cv[CV_K + 6], fmaf(ccy, cv[CV_K + 7], ccz * cv[CV_K + 8])), 1e-6f);
      // the texture coordinate with the reference's two true divisions (the fp16-operand kernel multiplies by reciprocals: one ulp
      // of u is 3e-5 texels at LLFF, which white-noise features turn into 2e-5 of a fetched feature - noise for that class,
      // a fifth of the budget of this one)
      const float u01 = pxc / pzc / (float)p.Wb, v01 = pyc / pzc / (float)p.Hb;
      int d_a0 = 0, d_a1 = 0;
      uint32_t d_pk = 0;
      float d_fu0 = 0.f, d_fv0 = 0.f, d_fu1 = 0.f, d_fv1 = 0.f, d_fr = 0.f, dd0 = 0.f, dd1 = 0.f, dd2 = 0.f, dd3 = 0.f;
      if (active) {
        float flod = fminf(fmaxf(lod, 0.f), (float)p.L);
        if (!(flod >= 0.f)) flod = 0.f;
        int l0 = (int)floorf(flod);
        int l1 = min(l0 + 1, p.L);
        const bool tri = flod > 0.f;
        int w0 = p.Wb >> l0, h0 = p.Hb >> l0, w1 = p.Wb >> l1, h1 = p.Hb >> l1;
        TexTap ta = tex_tap(u01, v01, w0, h0);
        TexTap tb = tex_tap(u01, v01, w1, h1);
        d_a0 = (int)(p.tex_level[l0] >> 2) + ((b * V + v) * h0 * w0 + ta.o00) * QL;
        d_a1 = (int)(p.tex_level[l1] >> 2) + ((b * V + v) * h1 * w1 + tb.o00) * QL;
        d_pk = (uint32_t)((ta.o01 - ta.o00) * QL) | ((uint32_t)((tb.o01 - tb.o00) * QL) << 14) |
               ((uint32_t)(ta.o10 - ta.o00) << 28) | ((uint32_t)(tb.o10 - tb.o00) << 29) | ((tri ? 1u : 0u) << 30) | (1u << 31);
        d_fu0 = ta.fu; d_fv0 = ta.fv; d_fu1 = tb.fu; d_fv1 = tb.fv; d_fr = flod - (float)l0;
        float sx = cwx - cv[CV_C + 0], sy = cwy - cv[CV_C + 1], sz = cwz - cv[CV_C + 2];
        unit3_fast(sx, sy, sz);
        float ddx = tdx - sx, ddy = tdy - sy, ddz = tdz - sz;
        unit3_fast(ddx, ddy, ddz);
        dd0 = ddx; dd1 = ddy; dd2 = ddz; dd3 = tdx * sx + tdy * sy + tdz * sz;
      }
      // the descriptor of (row, view) travels through the row's own 3 x 16 bytes of the hi plane of X_v
      unsigned char* dp = sX + (v * C::XCH) * 2048 + row * 16;
      *reinterpret_cast<uint4*>(dp) = make_uint4((uint32_t)d_a0, (uint32_t)d_a1, d_pk, __float_as_uint(d_fr));
      *reinterpret_cast<float4*>(dp + 2048) = make_float4(d_fu0, d_fv0, d_fu1, d_fv1);
      *reinterpret_cast<float4*>(dp + 4096) = make_float4(dd0, dd1, dd2, dd3);
    }
    if (!dyn && tn < tiles) rng_next = load_ranges(tn);      // the depth ranges of my next tile, in flight underneath this one

    // ================= P2: mip-mapped feature fetch, lane = (row, quad) =================
    // writes (both planes) FD_v = [featrgb_v | dir_v], X_v = [x_v | 1 | var16..18] and S = var[0:16]
    {
      __syncwarp();
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
      const int nit = min(C::NIT, (ns * G + IPW - 1) / IPW);          // rows >= ns * G never hold a sample
#pragma unroll 1
      for (int it = 0; it < nit; ++it) {
        const int src_raw = it * IPW + gr;
        const bool ok = glane && src_raw < 32;
        const int src = min(src_raw, 31);
        const int orow16 = (wq * 32 + src) * 16;
        const unsigned char* dsc = sX + orow16;
        const float4* tq = tex4 + (glane ? gq : 0);
        auto taps4 = [&](const uint4 q0, int level, float4(&t)[4]) {
          const uint32_t pk = q0.z;
          const int a = level ? (int)q0.y : (int)q0.x;
          const int dy = level ? (int)((pk >> 14) & 0x3FFF) : (int)(pk & 0x3FFF);
          const int dx = (int)((pk >> (28 + level)) & 1) * QL;
          const bool pr = ok && (pk >> 31) && (level == 0 || ((pk >> 30) & 1));
          const float4* b0 = tq + a;
          t[0] = ldg4_if(b0, pr);
          t[1] = ldg4_if(b0 + dx, pr);
          t[2] = ldg4_if(b0 + dy, pr);
          t[3] = ldg4_if(b0 + (dy + dx), pr);
        };
        float xq[V][4];
        auto blend = [&](const float4(&t0)[4], const float4(&t1)[4], const uint4 q0, const float4 q1) -> float4 {
          float4 f = bilerp4(t0[0], t0[1], t0[2], t0[3], q1.x, q1.y);
          if ((q0.z >> 30) & 1) {
            const float4 bq = bilerp4(t1[0], t1[1], t1[2], t1[3], q1.z, q1.w);
            const float frac = __uint_as_float(q0.w);
            f.x = lerpf(f.x, bq.x, frac); f.y = lerpf(f.y, bq.y, frac); f.z = lerpf(f.z, bq.z, frac); f.w = lerpf(f.w, bq.w, frac);
          }
          return f;
        };
        float4 fb[V];
        if constexpr (FB == 1) {
          float4 t0[V][4], t1[V][4];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * 2048);
            taps4(q0, 0, t0[v]);
            taps4(q0, 1, t1[v]);
          }
#pragma unroll
          for (int v = 0; v < V; ++v)
            fb[v] = blend(t0[v], t1[v], *reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * 2048),
                          *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 1) * 2048));
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const uint4 q0 = *reinterpret_cast<const uint4*>(dsc + (v * C::XCH) * 2048);
          float4 f;
          if constexpr (FB == 1) {
            f = fb[v];
          } else {
            float4 t0[4], t1[4];
            taps4(q0, 0, t0);
            taps4(q0, 1, t1);
            f = blend(t0, t1, q0, *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 1) * 2048));
          }
          const float4 q2 = *reinterpret_cast<const float4*>(dsc + (v * C::XCH + 2) * 2048);
          const float dir[4] = {q2.x, q2.y, q2.z, q2.w};
          const bool act = ok && (q0.z >> 31);
          const float fe[4] = {f.x, f.y, f.z, f.w};
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
            unsigned char* fdp = sFD + (v * C::FDCH + (gq >> 1)) * 2048 + orow16;
            unsigned char* xp = sX + (v * C::XCH + (gq >> 1)) * 2048 + orow16;
            if (last_quad) {
              // featrgb's pad channel is K slot F: dir_v follows in FD; in X the constant one (the var tail follows after the view loop)
              const float fd8[8] = {fe[0], fe[1], fe[2], dir[0], dir[1], dir[2], dir[3], 0.f};
              store8_split(fdp, PL, fd8);
              uint2 hi, lo;
              split4(xq[v][0], xq[v][1], xq[v][2], act ? 1.f : 0.f, hi, lo);
              *reinterpret_cast<uint2*>(xp) = hi;
              *reinterpret_cast<uint2*>(xp + PL) = lo;
              xq[v][3] = 0.f;
            } else {
              uint2 hi, lo;
              split4(fe[0], fe[1], fe[2], fe[3], hi, lo);
              *reinterpret_cast<uint2*>(fdp + (gq & 1) * 8) = hi;
              *reinterpret_cast<uint2*>(fdp + PL + (gq & 1) * 8) = lo;
              split4(xq[v][0], xq[v][1], xq[v][2], xq[v][3], hi, lo);
              *reinterpret_cast<uint2*>(xp + (gq & 1) * 8) = hi;
              *reinterpret_cast<uint2*>(xp + PL + (gq & 1) * 8) = lo;
            }
          }
        }
        if (ok) {
          float var[4];
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
          }
          uint2 hi, lo;
          split4(var[0], var[1], var[2], last_quad ? 0.f : var[3], hi, lo);
          if (last_quad) {
            // var[16..18] into K slots 20..22 of every [x_v | 1] operand (the mean is not stored at all: see the header)
#pragma unroll
            for (int v = 0; v < V; ++v) {
              unsigned char* xp = sX + (v * C::XCH + 2) * 2048 + orow16 + 8;
              *reinterpret_cast<uint2*>(xp) = hi;
              *reinterpret_cast<uint2*>(xp + PL) = lo;
            }
          } else {
            unsigned char* sp = sS + (gq >> 1) * 2048 + orow16 + (gq & 1) * 8;
            *reinterpret_cast<uint2*>(sp) = hi;
            *reinterpret_cast<uint2*>(sp + PL) = lo;
          }
        }
      }
    }

    // ================= GEMM 1: global_fc, G_v = var W_var + [x_v|1|var tail] W_gx + sum_u x_u (W_mean / V) =================
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
        const uint32_t d = tmem_group + v * 32;
        mma3_chunks(d, aS, C::SCH, C::W_GS, 32, 0);
        mma3_chunks(d, aX + v * C::XCH * 2048, C::XCH, C::W_GX, 32, 1);
#pragma unroll 1
        for (int u = 0; u < V; ++u) mma3_chunks(d, aX + u * C::XCH * 2048, C::XCH, C::W_GM, 32, 1);
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
    {
      float aw[V];
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        float gv[32];
        tmem_ld32(tmem_row + v * 32, gv);
        float s0 = vec[C::X_SCAL + 0], s1 = 0.f, s2_ = 0.f, s3 = 0.f;
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
      // X[0..3] <- aggregated vector (region X was consumed by GEMM 1)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const float t8[8] = {im[ch * 8 + 0], im[ch * 8 + 1], im[ch * 8 + 2], im[ch * 8 + 3], im[ch * 8 + 4], im[ch * 8 + 5], im[ch * 8 + 6], im[ch * 8 + 7]};
        store8_split(sX + ch * 2048 + row * 16, PL, t8);
      }
    }
    // ================= GEMM 2: fc =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      mma3_chunks(tmem_group, aX, 4, C::W_FC, 16, 0);
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
      store8_split(sX + 8 * 2048 + row * 16, PL, vox);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const float t8[8] = {img[ch * 8 + 0], img[ch * 8 + 1], img[ch * 8 + 2], img[ch * 8 + 3], img[ch * 8 + 4], img[ch * 8 + 5], img[ch * 8 + 6], img[ch * 8 + 7]};
        store8_split(sS + ch * 2048 + row * 16, PL, t8);
      }
    }
    // ================= GEMM 3: lr0 on [vox | img | 1] =================
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      mma3(tmem_group, aX + 8 * 2048, aS, aX + PL + 8 * 2048, aS + PL, w_base + C::W_LR0, w_base + C::W_LR0 + WP, 64, 0);
      mma3(tmem_group, aS + 2048, one_chunk, aS + PL + 2048, zero_chunk, w_base + C::W_LR0 + 2 * 64 * 16, w_base + C::W_LR0 + WP + 2 * 64 * 16, 64, 1);
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float h[32];
      tmem_ld32(tmem_row + half * 32, h);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const float t8[8] = {fmaxf(h[ch * 8 + 0], 0.f), fmaxf(h[ch * 8 + 1], 0.f), fmaxf(h[ch * 8 + 2], 0.f), fmaxf(h[ch * 8 + 3], 0.f),
                             fmaxf(h[ch * 8 + 4], 0.f), fmaxf(h[ch * 8 + 5], 0.f), fmaxf(h[ch * 8 + 6], 0.f), fmaxf(h[ch * 8 + 7], 0.f)};
        store8_split(sX + (half * 4 + ch) * 2048 + row * 16, PL, t8);
      }
    }
    // ================= GEMM 4: [sigma | feat_head] and weight.0 of every view (V x 64 + 16 TMEM columns) =================
    float sigma = 0.f, fh[8], wv[V];
    tc_fence_before();
    fence_async_smem();
    group_sync(g);
    if (row == 0) {
      tc_fence_after();
      mma3_chunks(tmem_group + V * 64, aX, 8, C::W_SH, 16, 0);
#pragma unroll 1
      for (int i = 0; i < V; ++i) {
        const uint32_t d = tmem_group + i * 64;
        mma3_chunks(d, aX, 8, C::W_0S, 64, 0);                                                          // h
        mma3(d, aX + 8 * 2048, aS, aX + PL + 8 * 2048, aS + PL, w_base + C::W_0S + 8 * 64 * 16, w_base + C::W_0S + WP + 8 * 64 * 16, 64, 1);   // vox | img[0:8]
        mma3(d, aS + 2048, one_chunk, aS + PL + 2048, zero_chunk, w_base + C::W_0S + 10 * 64 * 16, w_base + C::W_0S + WP + 10 * 64 * 16, 64, 1);   // img[8:16] | 1
        mma3_chunks(d, aFD + i * C::FDCH * 2048, C::FDCH, C::W_0V, 64, 1);                                // featrgb_v | dir_v
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, parity); parity ^= 1;
    tc_fence_after();
    {
      float sh[16];
      tmem_ld16(tmem_row + V * 64, sh);
      float s = sh[0] + vec[C::X_SCAL + 1];
      sigma = s > 20.f ? s : log1pf(expf(s));
#pragma unroll
      for (int k = 0; k < 8; ++k) fh[k] = fmaxf(sh[1 + k] + vec[C::X_FH_B + k], 0.f);
    }
#pragma unroll 1
    for (int i = 0; i < V; ++i) {
      float q0 = vec[C::X_SCAL + 2], q1 = 0.f, q2 = 0.f, q3 = 0.f;
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
        if (v == i) wv[v] = s2;
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
    tc_fence_before();

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

    auto stash_f = [&](int t) { return reinterpret_cast<float4*>(gsm + ((t * 4) >> 9) * 2048 + wq * 512 + ((t * 4) & 511)); };

    // ---- blended features sum_v w_v featrgb_v (featrgb = hi + lo read back from the FD operand), geometry head, depth, opacity
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
          const uint4 q = *reinterpret_cast<const uint4*>(sFD + (v * C::FDCH + ch) * 2048 + row * 16);
          const uint4 ql = *reinterpret_cast<const uint4*>(sFD + PL + (v * C::FDCH + ch) * 2048 + row * 16);
          const float2 f0 = h2_to_f2(q.x), f1 = h2_to_f2(q.y), f2 = h2_to_f2(q.z), f3 = h2_to_f2(q.w);
          const float2 l0 = h2_to_f2(ql.x), l1 = h2_to_f2(ql.y), l2 = h2_to_f2(ql.z), l3 = h2_to_f2(ql.w);
          acc[0] = fmaf(f0.x + l0.x, wv[v], acc[0]); acc[1] = fmaf(f0.y + l0.y, wv[v], acc[1]);
          acc[2] = fmaf(f1.x + l1.x, wv[v], acc[2]); acc[3] = fmaf(f1.y + l1.y, wv[v], acc[3]);
          acc[4] = fmaf(f2.x + l2.x, wv[v], acc[4]); acc[5] = fmaf(f2.y + l2.y, wv[v], acc[5]);
          acc[6] = fmaf(f3.x + l3.x, wv[v], acc[6]); acc[7] = fmaf(f3.y + l3.y, wv[v], acc[7]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (ch * 8 + e < F) vals[ch * 8 + e] = acc[e];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) vals[F + k] = fh[k];
      vals[F + 8] = p.inv_depth ? fdiv(1.f, z) : z;
      vals[F + 9] = 1.f;
      __syncwarp();          // every lane has read its FD rows
#pragma unroll
      for (int q4 = 0; q4 < C::NCP / 4; ++q4)
        *stash_f(lane * C::NCP + ((q4 ^ (lane & 7)) << 2)) =
            make_float4(wgt * vals[q4 * 4 + 0], wgt * vals[q4 * 4 + 1], wgt * vals[q4 * 4 + 2], wgt * vals[q4 * 4 + 3]);
      __syncwarp();
      if (p.out_cl && p.dec_stride == F + 9) {
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
        const int nb = __shfl_sync(full, n, rowof(bb, 0));
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
              for (int k = F + 8; k < p.dec_stride; ++k) odb[k - c] = k == F + 8 ? p.dec_pad0 : 0.f;
          } else if (c == F + 8) {
            p.out_depth[(size_t)b * HW + pixb] = p.inv_depth ? fdiv(1.f, a) : a;
          } else {
            p.out_opacity[(size_t)b * HW + pixb] = a;
          }
        }
      }
    }

    // ============== P6: fine colours, lane = (row, ray) (bundle_sampler.py:327-337) ==============
    __syncwarp();
    {
      // per-row parameters through the row's 16-byte slots of S (both chunks of the hi plane; free since GEMM 4)
      *reinterpret_cast<float4*>(sS + row * 16) = make_float4(z, geo.x0, geo.y0, wgt);
      {
        float w4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < V; ++v) w4[v] = wv[v];
        *reinterpret_cast<float4*>(sS + 2048 + row * 16) = make_float4(w4[0], w4[1], w4[2], w4[3]);
      }
      __syncwarp();
      auto cst = [&](int t) { return reinterpret_cast<float*>(gsm + ((t * 4) >> 9) * 2048 + wq * 512 + ((t * 4) & 511)); };
      const int nit6 = min(BB, (ns * G * BB + 31) / 32);
#pragma unroll 1
      for (int it = 0; it < nit6; ++it) {
        const int item = it * 32 + lane;
        const int r = item / BB, j = item - r * BB;
        const float4 ra = *reinterpret_cast<const float4*>(sS + (wq * 32 + r) * 16);
        const float4 rb = *reinterpret_cast<const float4*>(sS + 2048 + (wq * 32 + r) * 16);
        const float zr = ra.x, wr = ra.w;
        const bool actr = wr != 0.f;          // a row without compositing weight contributes w * colour = 0 whatever it gathers
        const float wvr[4] = {rb.x, rb.y, rb.z, rb.w};
        const float x = ra.y + (float)(j % BS), y = ra.z + (float)(j / BS);
        const float* M = head + CAM_M;
        const float dx = fmaf(x, M[0], fmaf(y, M[1], M[2]));
        const float dy = fmaf(x, M[3], fmaf(y, M[4], M[5]));
        const float dz = fmaf(x, M[6], fmaf(y, M[7], M[8]));
        const float wx = fmaf(dx, zr, ox), wy = fmaf(dy, zr, oy), wz = fmaf(dz, zr, oz);
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
#ifdef GDB_X_EXACT
          const float rz = __frcp_rn(iz);
#else
          // MUFU.RCP alone (1 ulp of a PIXEL coordinate of the colour tap; the texture coordinate of P1 keeps its true divisions):
          // no range check / slow-path call between a view's projection and its gathers (gdb_render_tc2.cu, head of the file)
          const float rz = rcp_approx(iz);
#endif
          float gx = (ix * rz) * two_W - 1.f, gy = (iy * rz) * two_H - 1.f;
          const Bilin bl4 = bilin_border(gx, gy, p.W, p.H);
          const float4* ib = reinterpret_cast<const float4*>(p.rgba) + (size_t)(b * V + v) * p.H * p.W;
          t[v][0] = ldg4_if(ib + bl4.o00, actr);
          t[v][1] = ldg4_if(ib + bl4.o10, actr);
          t[v][2] = ldg4_if(ib + bl4.o01, actr);
          t[v][3] = ldg4_if(ib + bl4.o11, actr);
          tw[v][0] = bl4.w00; tw[v][1] = bl4.w10; tw[v][2] = bl4.w01; tw[v][3] = bl4.w11;
        }
        float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) c4 = f4_scale_add(c4, t[v][k], tw[v][k]);
          cr = fmaf(c4.x, wvr[v], cr); cg = fmaf(c4.y, wvr[v], cg); cb = fmaf(c4.z, wvr[v], cb);
        }
        *cst(r * R + 0 * BB + j) = wr * cr;
        *cst(r * R + 1 * BB + j) = wr * cg;
        *cst(r * R + 2 * BB + j) = wr * cb;
      }
      __syncwarp();
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
      __syncwarp();
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
  }
}

template <int BS, int FEAT_DIM, int V, int FB>
static int launch_render_tc4(const RenderParams& p, cudaStream_t st) {
  using C = Tc4Cfg<BS, FEAT_DIM, V>;
  static_assert(C::SMEM + 1024 <= 227 * 1024, "shared memory plan (dynamic + the kernel's 1 KB static section)");
  auto kern = render_tc4_kernel<BS, FEAT_DIM, V, FB>;
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, kern, C::SMEM);
    if (e != cudaSuccess) return fail((int)e, "gdb_render_fused_fwd(tc4): cudaFuncSetAttribute(%d B): %s", C::SMEM, cudaGetErrorString(e));
  }
  if ((long)p.Wb * C::QL >= (1 << 14))
    return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_fwd(tc4): bundle map width %d too large for the packed tap stride", p.Wb);
  const int G = 32 / p.max_samples;
  const long tiles = (long)p.B * ((p.pix_hi - p.pix_lo + 4 * G - 1) / (4 * G));
  long ctas = (tiles + C::NG - 1) / C::NG;
  if (ctas > sm_count()) ctas = sm_count();
  RenderParams q = p;
  q.tile_counter = ctas == sm_count() ? acquire_tile_counter(st) : nullptr;
  kern<<<(int)ctas, 128 * C::NG, C::SMEM, st>>>(q);
  return cuda_check("gdb_render_fused_fwd(tc4)");
}

// true when this kernel covers the call: 2x2 bundles, feat_dim 16, three source views, no parity taps
bool render_tc4_covers(const RenderParams& p, int bundle_size, int feat_dim, int V) {
  const bool taps = p.tap_rfd || p.tap_vox || p.tap_sigma || p.tap_feat || p.tap_w;
  static int off = -1;
  if (off < 0) { const char* e = getenv("GDB_K3_SPLIT_GEN1"); off = (e && e[0] == '1') ? 1 : 0; }     // A/B: the first-generation split kernel
  return !off && !taps && bundle_size == 2 && feat_dim == 16 && V == 3;
}
int render_tc4_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, cudaStream_t st) {
  static int fb = -1;
  // all taps of a fetch iteration in flight is the default (measured: 1.85 -> 1.77 ms per 8 DTU views, 3.42 -> 3.26 ms at LLFF); GDB_K3_SPLIT_FB=0 is the A/B
  if (fb < 0) { const char* e = getenv("GDB_K3_SPLIT_FB"); fb = (e && e[0] == '0') ? 0 : 1; }
  if (bundle_size == 2 && feat_dim == 16 && V == 3) return fb ? launch_render_tc4<2, 16, 3, 1>(p, st) : launch_render_tc4<2, 16, 3, 0>(p, st);
  return fail(GDB_E_UNSUPPORTED, "gdb_render_fused_fwd(tc4): (bundle_size=%d, feat_dim=%d, V=%d) not instantiated", bundle_size, feat_dim, V);
}

}  // namespace gdb
