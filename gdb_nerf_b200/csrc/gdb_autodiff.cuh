// Forward-mode dual numbers and warp-level gradient reductions shared by the backward kernels
// (gdb_render_bwd.cu, gdb_coarse.cu).
#pragma once
#include "gdb_common.cuh"

namespace gdb {

// ------------------------------------------------------------------ duals --
struct Dual {
  float v, d;
  __device__ __forceinline__ Dual() {}
  __device__ __forceinline__ Dual(float a) : v(a), d(0.f) {}
  __device__ __forceinline__ Dual(float a, float b) : v(a), d(b) {}
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  float q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ Dual dsqrt(Dual a) {
  float s = sqrtf(a.v);
  return Dual(s, a.d / (2.f * s));
}
__device__ __forceinline__ Dual dmax(Dual a, float lo) { return a.v < lo ? Dual(lo, 0.f) : a; }     // torch.clamp_min / max()
__device__ __forceinline__ Dual dclamp(Dual a, float lo, float hi) { return a.v < lo ? Dual(lo, 0.f) : (a.v > hi ? Dual(hi, 0.f) : a); }
__device__ __forceinline__ Dual dlog2(Dual a) { return Dual(log2f(a.v), a.d / (a.v * 0.6931471805599453f)); }
__device__ __forceinline__ void dunit3(Dual& x, Dual& y, Dual& z) {
  Dual n = dmax(dsqrt(x * x + y * y + z * z), 1e-12f);
  x = x / n; y = y / n; z = z / n;
}

// transpose-reduce: every lane passes 32 values; lane L gets sum over lanes of v[L]
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      float keep = up ? v[i + s] : v[i];
      float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// gsm[row0*Npad .. ] += sum over lanes x[k] * dy[n]  for k < K, n < N  ([K][Npad] block, N innermost)
__device__ __forceinline__ void accum_outer(float* gsm, int Npad, const float* x, int K, const float* dy, int N, int lane) {
  const int total = K * N;
  for (int base = 0; base < total; base += 32) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      int e = base + i;
      v[i] = e < total ? x[e / N] * dy[e % N] : 0.f;
    }
    float s = warp_transpose_reduce(v, lane);
    int e = base + lane;
    if (e < total) atomicAdd(gsm + (e / N) * Npad + (e % N), s);
  }
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
  return x;
}


}  // namespace gdb
