// Shared helpers for the gdb_nerf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/gdb_nerf_b200.h"

namespace gdb {

// ------------------------------------------------------------------ errors --
inline char* err_buf() {
  static thread_local char buf[512] = "ok";
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline int cuda_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}
#define GDB_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return gdb::fail(code, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Per-device caches: one process may drive several GPUs (nn.DataParallel, render_sweep(device=...)), and both the SM count and
// the dynamic shared-memory opt-in of a kernel belong to ONE device.
constexpr int GDB_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & (GDB_MAX_DEVICES - 1);
}
inline int sm_count() {
  static int n[GDB_MAX_DEVICES] = {0};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}
// cudaFuncAttributeMaxDynamicSharedMemorySize, remembered per (kernel = owner of the SmemOptIn, device)
struct SmemOptIn {
  int bytes[GDB_MAX_DEVICES] = {0};
};
template <class K>
inline cudaError_t opt_in_smem(SmemOptIn& s, K kern, int bytes) {
  const int dev = current_device();
  if (bytes <= s.bytes[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) s.bytes[dev] = bytes;
  return e;
}

// ------------------------------------------------------------ camera block --
// Per batch: CAM_HEAD floats, then CAM_VIEW floats per source view.
constexpr int CAM_HEAD = 32;
constexpr int CAM_VIEW = 32;
// head
constexpr int CAM_M = 0;        // 3x3 ray matrix: dir = M * (x, y, 1)       (bundle_sampler.py:70)
constexpr int CAM_O = 9;        // ray origin / target camera centre          (:69)
constexpr int CAM_ZAXIS = 12;   // c2w[:3, 2]                                 (:68)
constexpr int CAM_PIXR = 15;    // 1/sqrt(fx*fy*pi)                           (:74)
constexpr int CAM_MINIV = 16;   // min sample interval                        (:227-229)
constexpr int CAM_NEAR = 17;
constexpr int CAM_FAR = 18;
constexpr int CAM_DISK = 19;    // b * pixel radius                           (:106)
// per view
constexpr int CV_E = 0;         // 3x4 world->camera
constexpr int CV_K = 12;        // 3x3 intrinsics
constexpr int CV_C = 21;        // camera centre in world                     (:305)
constexpr int CV_PIXR = 24;     // 1/sqrt(fx/b*fy/b*pi)                       (:313)

// ------------------------------------------------------ packed MLP params --
// Every linear layer is stored transposed, [K][N] row-major (N innermost), so
// that for a fixed input k the N weights are contiguous (broadcast LDS.128).
// N is padded to a multiple of 4 where noted.  F = feat_dim + 3.
template <int FEAT_DIM>
struct MlpLayout {
  static constexpr int F = FEAT_DIM + 3;
  static constexpr int FP = (F + 3) & ~3;             // padded
  static constexpr int VIEW_W = 0;                    // [4][FP]    view_fc.0   (nerf.py:20-23)
  static constexpr int VIEW_B = VIEW_W + 4 * FP;      // [FP]
  static constexpr int GLOB_W = VIEW_B + FP;          // [3F][32]   global_fc.0 (:25-28) rows: x | var | mean
  static constexpr int GLOB_B = GLOB_W + 3 * F * 32;  // [32]
  static constexpr int AGG_W = GLOB_B + 32;           // [32]       agg_w_fc.0  (:29-32)
  static constexpr int AGG_B = AGG_W + 32;            // [4] (1 used)
  static constexpr int FC_W = AGG_B + 4;              // [32][16]   fc.0        (:33-36)
  static constexpr int FC_B = FC_W + 32 * 16;         // [16]
  static constexpr int LR0_W = FC_B + 16;             // [24][64]   lr0.0       (:39-42) rows: vox(8) | img(16)
  static constexpr int LR0_B = LR0_W + 24 * 64;       // [64]
  static constexpr int SIG_W = LR0_B + 64;            // [64]       sigma.0     (:43-46)
  static constexpr int SIG_B = SIG_W + 64;            // [4] (1 used)
  static constexpr int W0_W = SIG_B + 4;              // [88+F+4][64] weight.0  (:47-52) rows: h(64)|vox(8)|img(16)|featrgb(F)|dir(4)
  static constexpr int W0_B = W0_W + (88 + F + 4) * 64;  // [64]
  static constexpr int W2_W = W0_B + 64;              // [64]       weight.2
  static constexpr int W2_B = W2_W + 64;              // [4] (1 used)
  static constexpr int FH_W = W2_B + 4;               // [64][8]    feat_head.0 (:53-56)
  static constexpr int FH_B = FH_W + 64 * 8;          // [8]
  static constexpr int TOTAL = FH_B + 8;
};

// ------------------------------------------------------------- device math --
// IEEE single operations that must not be contracted into FMAs: used where the
// result feeds a ceil()/floor() that defines an integer output of the reference.
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// torch.linspace(0, 1, n)[i] in float32 (symmetric evaluation, as ATen does)
__device__ __forceinline__ float linspace01(int i, int n) {
  if (n == 1) return 0.f;
  float step = fdiv(1.f, (float)(n - 1));
  return (i < n / 2) ? fmul(step, (float)i) : fsub(1.f, fmul(step, (float)(n - 1 - i)));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Packed fp32 FMA (FFMA2, sm_100): two IEEE fp32 FMAs per issue slot.  A three-register scalar FFMA issues every
// other cycle per scheduler on Blackwell, FFMA2 carries two lanes in the same slot, so FMA-bound inner loops
// (bilinear accumulation, the SIMT MLP) run them pairwise.  Results are bit-identical to two fmaf().
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// lerp(a, b, t) = a + t (b - a) on a pair
__device__ __forceinline__ unsigned long long lerp2(unsigned long long a, unsigned long long b, unsigned long long tt) {
  return fma2(tt, sub2(b, a), a);
}
// (a0, a1) += (v0, v1) * w
__device__ __forceinline__ void fma2_acc(float& a0, float& a1, float v0, float v1, unsigned long long ww) {
  unsigned long long r = fma2(pack2(v0, v1), ww, pack2(a0, a1));
  unpack2(r, a0, a1);
}

__device__ __forceinline__ float4 f4_scale_add(float4 acc, float4 v, float w) {
  const unsigned long long ww = pack2(w, w);
  fma2_acc(acc.x, acc.y, v.x, v.y, ww);
  fma2_acc(acc.z, acc.w, v.z, v.w, ww);
  return acc;
}

}  // namespace gdb
