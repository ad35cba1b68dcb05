// Configuration (operand / TMEM plan) and small device helpers shared by the second- and third-generation tcgen05 render
// kernels (gdb_render_tc2.cu, gdb_render_tc3.cu).
#pragma once
#include <cuda_fp16.h>

#include "gdb_render_common.cuh"

#include "gdb_tcgen05.cuh"

namespace gdb {

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int cmin(int a, int b) { return a < b ? a : b; }

template <int BS, int FEAT_DIM, int V, int NG>
struct Tc2Cfg {
  using ML = MlpLayout<FEAT_DIM>;
  static constexpr int BB = BS * BS;
  static constexpr int F = ML::F;
  static constexpr int FP = ML::FP;
  static constexpr int R = 3 * BB;
  static constexpr int CT = R + F + 8;
  static constexpr int RFD = R + F + 4;
  static constexpr int QL = FP / 4;                    // 16-byte quads per texel = lanes per row in the feature fetch
  static constexpr int IPW = 32 / QL;                  // rows per warp iteration of the feature fetch
  static constexpr int NIT = (32 + IPW - 1) / IPW;
  static_assert(F == FP - 1, "the texel has exactly one pad channel (F = feat_dim + 3, feat_dim a multiple of 4)");
  static_assert(QL % 2 == 1, "the last quad starts a chunk");
  // operand chunks (a chunk = 8 K values x 128 rows x fp16 = 2 KB)
  static constexpr int XCH = (F + 1 + 7) / 8;          // [x_v (F) | 1]
  static constexpr int FDCH = (F + 4 + 7) / 8;         // [featrgb_v (F) | dir_v (4)]
  static constexpr int SCH = (2 * FP + 7) / 8;         // [var (FP) | mean (FP)]
  static constexpr int KS_X = (XCH + 1) / 2, KS_FD = (FDCH + 1) / 2, KS_S = (SCH + 1) / 2;   // K steps of 16
  static constexpr int CH_X = cmax(V * XCH, 9);        // later [h (8) | vox]
  static constexpr int CH_FD = V * FDCH;
  static constexpr int CH_S = cmax(SCH, 4);            // later the aggregated vector (4), then img (2)
  // fp16 weight matrices (bytes), UMMA B layout [K/8][N][8]
  static constexpr int W_GS = 0;
  static constexpr int W_GX = W_GS + 32 * KS_S * 32;
  static constexpr int W_FC = W_GX + 32 * KS_X * 32;
  static constexpr int W_LR0 = W_FC + 16 * 32 * 2;
  static constexpr int W_SH = W_LR0 + 64 * 32 * 2;
  static constexpr int W_0S = W_SH + 16 * 64 * 2;
  static constexpr int W_0V = W_0S + 64 * 96 * 2;
  static constexpr int W_END = W_0V + 64 * KS_FD * 32;
  // fp32 vectors (floats)
  static constexpr int X_VIEW_W = 0;                   // [4][FP]
  static constexpr int X_VIEW_B = X_VIEW_W + 4 * FP;   // [FP]
  static constexpr int X_AGG_W = X_VIEW_B + FP;        // 32
  static constexpr int X_FC_B = X_AGG_W + 32;          // 16
  static constexpr int X_W2_W = X_FC_B + 16;           // 64
  static constexpr int X_FH_B = X_W2_W + 64;           // 8
  static constexpr int X_SCAL = X_FH_B + 8;            // agg_b, sig_b, w2_b, pad
  static constexpr int X_END = X_SCAL + 4;
  static constexpr int VEC_OFF = ((W_END + 127) / 128) * 128;
  static constexpr int GROUP_OFF = ((VEC_OFF + X_END * 4 + 127) / 128) * 128;
  // per-group regions (bytes from the group base); S lies after X so that chunk pairs (X[8], S[0]) have a positive stride
#ifndef GDB_K3_CHPAD
#define GDB_K3_CHPAD 0
#endif
  static constexpr int CHB = 2048 + GDB_K3_CHPAD;      // bytes between consecutive operand chunks (the UMMA leading-dimension offset is free)
  static constexpr int A_X = 0;
  static constexpr int A_FD = A_X + CH_X * CHB;
  static constexpr int A_S = A_FD + CH_FD * CHB;
  static constexpr int A_END = ((A_S + CH_S * CHB + 127) / 128) * 128;
  static constexpr int CAM_OFF = A_END + 128;          // mbarrier at A_END
#ifdef GDB_X_LINPROJ
  static constexpr int LIN_FLOATS_OFF = ((CAM_HEAD + CAM_VIEW * V) * 4 + 127) / 128 * 32;   // composed projections, floats from the camera block
  static constexpr int GROUP_BYTES = CAM_OFF + LIN_FLOATS_OFF * 4 + (12 * V * 4 + 127) / 128 * 128;
#else
  static constexpr int GROUP_BYTES = CAM_OFF + ((CAM_HEAD + CAM_VIEW * V) * 4 + 127) / 128 * 128;
#endif
  static constexpr int ZERO_OFF = GROUP_OFF + NG * GROUP_BYTES;   // constant chunks after every group
  static constexpr int ONE_OFF = ZERO_OFF + 2048;
  static constexpr int SMEM = ONE_OFF + 2048;
  // per-(row, ray) weighted colours: a warp uses its own rows' 512 B of the first BB chunks of X + FD
  static_assert(BB <= CH_X + CH_FD, "colour stash must fit in X + FD");
  static constexpr int NC = F + 10;                    // composited channels: featrgb (F), geometry head (8), depth, opacity
  static constexpr int NCP = NC <= 32 ? 32 : 64;       // padded row of the transposition stash (swizzled by float4)
  static_assert(32 * NCP * 4 <= 512 * (CH_X + CH_FD), "compositing stash must fit in the warp's rows of X + FD");
  // TMEM columns per group
  static constexpr int TC = cmin((512 / NG) & ~31, 256);
  static constexpr int TALLOC = NG * TC <= 256 ? 256 : 512;      // tcgen05.alloc takes a power of two
  static constexpr int NB = TC / 64;                   // 64-column buffers for weight.0
  static_assert(V * 32 <= TC && NB >= 2, "TMEM column plan");
  static constexpr int ROUNDS = 1 + (cmax(V - (NB - 1), 0) + NB - 1) / NB;
  __host__ __device__ static constexpr int round_start(int r) { return r == 0 ? 0 : (NB - 1) + (r - 1) * NB; }
  __host__ __device__ static constexpr int round_n(int r) { return cmax(0, cmin(r == 0 ? NB - 1 : NB, V - round_start(r))); }
};

// weights: fp32 packed block (global) -> fp16 UMMA B operand [Kpad/8][N][8]; value(k, n) supplied by the caller
template <class Fn>
__device__ __forceinline__ void stage_b2(unsigned char* dst, int N, int Kpad, int tid, int nthreads, Fn value) {
  for (int i = tid; i < (Kpad / 8) * N; i += nthreads) {
    const int c = i / N, n = i - c * N;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = value(c * 8 + j, n);
    *reinterpret_cast<uint4*>(dst + (size_t)i * 16) =
        make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  }
}

// one-time per CTA: weights (fp32 packed block, global) -> fp16 UMMA B operands + fp32 vectors in shared memory, and the two
// constant operand chunks (zero; fp16 one in K slot 0)
template <class C>
__device__ __forceinline__ void tc2_stage_weights(unsigned char* smem, float* vec, const float* __restrict__ m, int tid, int nt) {
  using ML = typename C::ML;
  constexpr int F = C::F, FP = C::FP;
  // [var | mean] rows of global_fc: k < F -> var_k, FP <= k < FP + F -> mean_(k - FP)
  stage_b2(smem + C::W_GS, 32, C::KS_S * 16, tid, nt, [&](int k, int n) {
    return k < F ? __ldg(m + ML::GLOB_W + (size_t)(F + k) * 32 + n)
                 : (k >= FP && k < FP + F ? __ldg(m + ML::GLOB_W + (size_t)(2 * F + k - FP) * 32 + n) : 0.f);
  });
  // [x_v | 1]: the constant-one slot carries global_fc's bias
  stage_b2(smem + C::W_GX, 32, C::KS_X * 16, tid, nt, [&](int k, int n) {
    return k < F ? __ldg(m + ML::GLOB_W + (size_t)k * 32 + n) : (k == F ? __ldg(m + ML::GLOB_B + n) : 0.f);
  });
  stage_b2(smem + C::W_FC, 16, 32, tid, nt, [&](int k, int n) { return __ldg(m + ML::FC_W + k * 16 + n); });
  // [vox (8) | img (16) | 1 | 0]
  stage_b2(smem + C::W_LR0, 64, 32, tid, nt, [&](int k, int n) {
    return k < 24 ? __ldg(m + ML::LR0_W + k * 64 + n) : (k == 24 ? __ldg(m + ML::LR0_B + n) : 0.f);
  });
  // [sigma | feat_head] as one N = 16 operand: n = 0 sigma, n = 1..8 geometry head
  stage_b2(smem + C::W_SH, 16, 64, tid, nt, [&](int k, int n) {
    return n == 0 ? __ldg(m + ML::SIG_W + k) : (n <= 8 ? __ldg(m + ML::FH_W + k * 8 + (n - 1)) : 0.f);
  });
  // [h (64) | vox (8) | img (16) | 1 | 0]
  stage_b2(smem + C::W_0S, 64, 96, tid, nt, [&](int k, int n) {
    return k < 88 ? __ldg(m + ML::W0_W + (size_t)k * 64 + n) : (k == 88 ? __ldg(m + ML::W0_B + n) : 0.f);
  });
  // [featrgb_v (F) | dir_v (4) | 0]
  stage_b2(smem + C::W_0V, 64, C::KS_FD * 16, tid, nt,
           [&](int k, int n) { return k < F + 4 ? __ldg(m + ML::W0_W + (size_t)(88 + k) * 64 + n) : 0.f; });
  for (int i = tid; i < 5 * FP; i += nt) vec[C::X_VIEW_W + i] = m[ML::VIEW_W + i];   // W [4][FP] + b [FP]
  for (int i = tid; i < 32; i += nt) vec[C::X_AGG_W + i] = m[ML::AGG_W + i];
  for (int i = tid; i < 16; i += nt) vec[C::X_FC_B + i] = m[ML::FC_B + i];
  for (int i = tid; i < 64; i += nt) vec[C::X_W2_W + i] = m[ML::W2_W + i];
  for (int i = tid; i < 8; i += nt) vec[C::X_FH_B + i] = m[ML::FH_B + i];
  if (tid == 0) { vec[C::X_SCAL + 0] = m[ML::AGG_B]; vec[C::X_SCAL + 1] = m[ML::SIG_B]; vec[C::X_SCAL + 2] = m[ML::W2_B]; }
  for (int i = tid; i < 128; i += nt) {
    *reinterpret_cast<uint4*>(smem + C::ZERO_OFF + i * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(smem + C::ONE_OFF + i * 16) = make_uint4(0x3C00u, 0, 0, 0);    // fp16 1.0 in K slot 0
  }
}

// one K step (16) of D (+)= A B: the two A chunks may live anywhere (a1 > a0), B K steps are contiguous
__device__ __forceinline__ void mma_step(uint32_t d_tmem, uint32_t a0, uint32_t a1, uint32_t b_addr, int N, uint32_t accumulate) {
  umma_f16(d_tmem, umma_desc(a0, a1 - a0, 128), umma_desc(b_addr, N * 16, 128), umma_idesc_f16(N), accumulate);
}
// `nch` consecutive chunks starting at `a` (odd counts pair the last chunk with the zero chunk)
__device__ __forceinline__ void mma_chunks(uint32_t d_tmem, uint32_t a, int nch, uint32_t zero_chunk, uint32_t b_addr, int N,
                                           uint32_t accumulate, int chb = 2048) {
  for (int ks = 0; 2 * ks < nch; ++ks) {
    const uint32_t a0 = a + ks * 2 * chb;
    const uint32_t a1 = (2 * ks + 1 < nch) ? a0 + chb : zero_chunk;
    mma_step(d_tmem, a0, a1, b_addr + ks * 2 * (N * 16), N, (ks > 0 || accumulate) ? 1u : 0u);
  }
}

// F.normalize(eps = 1e-12) with a reciprocal square root (2 ulp): the result feeds fp16 operands
__device__ __forceinline__ void unit3_fast(float& x, float& y, float& z) {
  const float inv = rsqrtf(fmaxf(x * x + y * y + z * z, 1e-24f));
  x *= inv; y *= inv; z *= inv;
}
// predicated 128-bit read-only load (zero when the predicate is off): no branch around it, so the compiler can keep a
// whole batch of gathers in flight behind one another instead of one dependent group per basic block
__device__ __forceinline__ float4 ldg4_if(const float4* ptr, bool pred) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}\n"
      : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w)
      : "l"(ptr), "r"((int)pred));
  return r;
}
// single-MUFU reciprocal / square root (1-2 ulp, no range-check branch and no slow-path call behind them)
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 h2_to_f2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

}  // namespace gdb
