// Device helpers shared by the fused render kernels (fp32 SIMT and tcgen05 variants).
#pragma once
#include "gdb_sampling.cuh"

namespace gdb {

struct RenderParams {
  const float* rgba;      // (B*V, H, W, 4)
  const float* tex;       // mip chain, level k: (B*V, Hb>>k, Wb>>k, FP)
  const float* vol;       // (B, D, Hb, Wb, 8)
  const float* depth_range;
  const float* vol_range;
  const float* cam;
  const float* mlp;
  float* out_feat;        // (B, CT, Hb, Wb) planar, or (B, Hb, Wb, R) when out_cl
  float* out_dec;         // (B, Hb, Wb, F+8) when out_cl
  float* out_depth;
  float* out_opacity;
  // optional taps
  const int32_t* offsets;
  int64_t S_total;
  float* tap_rfd;
  float* tap_vox;
  float* tap_sigma;
  float* tap_feat;
  float* tap_w;
  int64_t tex_level[4];
  int cam_stride;
  int dec_stride;      // floats per bundle of out_dec (>= F + 8; the pad channels are written as zeros, the first as dec_pad0)
  float dec_pad0;      // 0, or 1: a constant-one channel that carries the bias of the decoder's first convolution
  int vol_stride;      // floats between consecutive voxels of vol (>= 8, multiple of 4)
  int64_t vol_sb, vol_sz, vol_sy, vol_sx;   // float strides of vol over (batch, depth, row, column): layout 0 (B,D,Hb,Wb,.) or 1 (B,Hb,Wb,D,.)
  int B, H, W, Hb, Wb, D, max_samples, L, inv_depth, adaptive, out_cl;
  int pix_lo, pix_hi;  // range of bundle indices (row-major over the Hb x Wb bundle map) rendered in every view: the image-tile split
  unsigned int* tile_counter;   // tensor-core kernels: tiles beyond a tile slot's first are handed out by this counter (zero at launch);
                                // nullptr = every tile slot strides through the tiles (static assignment)
};

// A zeroed tile counter for one launch on `st` (gdb_render.cu), or nullptr when none is available (the kernel then falls back
// to the static assignment).  The counters live in a __device__ array of the library (nothing is allocated): one per stream
// that ever launched (launches on a stream are ordered, so its counter is never shared by two kernels in flight), a fresh
// one for every launch recorded during a stream capture (a graph replay may overlap launches on the capturing stream).
unsigned int* acquire_tile_counter(cudaStream_t st);

template <int N>
__device__ __forceinline__ void axpy_row(float (&acc)[N], const float* __restrict__ wrow, float x) {
  const unsigned long long xx = pack2(x, x);
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    float4 w = *reinterpret_cast<const float4*>(wrow + n);
    fma2_acc(acc[n + 0], acc[n + 1], w.x, w.y, xx);        // FFMA2: two outputs per issue slot
    fma2_acc(acc[n + 2], acc[n + 3], w.z, w.w, xx);
  }
}
template <int N>
__device__ __forceinline__ void load_row(float (&acc)[N], const float* __restrict__ row) {
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    float4 w = *reinterpret_cast<const float4*>(row + n);
    acc[n + 0] = w.x; acc[n + 1] = w.y; acc[n + 2] = w.z; acc[n + 3] = w.w;
  }
}
template <int N>
__device__ __forceinline__ float dot_row(const float (&x)[N], const float* __restrict__ row) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int n = 0; n < N; n += 4) {
    float4 w = *reinterpret_cast<const float4*>(row + n);
    unsigned long long acc = fma2(pack2(w.x, w.y), pack2(x[n + 0], x[n + 1]), pack2(s0, s1));
    acc = fma2(pack2(w.z, w.w), pack2(x[n + 2], x[n + 3]), acc);
    unpack2(acc, s0, s1);
  }
  return s0 + s1;
}

__device__ __forceinline__ void unit3(float& x, float& y, float& z) {
  float n = fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);   // F.normalize(eps=1e-12)
  x /= n; y /= n; z /= n;
}

// border-clamped bilinear setup for grid_sample(align_corners=False, padding 'border')
struct Bilin {
  int o00, o10, o01, o11;   // texel offsets (in texels)
  float w00, w10, w01, w11;
};
__device__ __forceinline__ Bilin bilin_border(float gx, float gy, int Wd, int Hd) {
  float ix = fminf(fmaxf(((gx + 1.f) * (float)Wd - 1.f) * 0.5f, 0.f), (float)(Wd - 1));
  float iy = fminf(fmaxf(((gy + 1.f) * (float)Hd - 1.f) * 0.5f, 0.f), (float)(Hd - 1));
  float x0f = floorf(ix), y0f = floorf(iy);
  float tx = ix - x0f, ty = iy - y0f;
  int x0 = (int)x0f, y0 = (int)y0f;
  int x1 = min(x0 + 1, Wd - 1), y1 = min(y0 + 1, Hd - 1);   // weight is 0 whenever the clamp bites
  Bilin r;
  r.o00 = y0 * Wd + x0; r.o10 = y0 * Wd + x1; r.o01 = y1 * Wd + x0; r.o11 = y1 * Wd + x1;
  r.w00 = (1.f - tx) * (1.f - ty); r.w10 = tx * (1.f - ty); r.w01 = (1.f - tx) * ty; r.w11 = tx * ty;
  return r;
}

// nvdiffrast indexTextureLinear, boundary 'clamp'
struct TexTap {
  int o00, o10, o01, o11;
  float fu, fv;
};
__device__ __forceinline__ TexTap tex_tap(float u01, float v01, int w, int h) {
  float u = fminf(fmaxf(u01 * (float)w - 0.5f, 0.f), (float)(w - 1));
  float v = fminf(fmaxf(v01 * (float)h - 0.5f, 0.f), (float)(h - 1));
  bool cu = (u == 0.f) || (u == (float)(w - 1));
  bool cv = (v == 0.f) || (v == (float)(h - 1));
  int iu0 = (int)floorf(u), iv0 = (int)floorf(v);
  int iu1 = iu0 + (cu ? 0 : 1), iv1 = iv0 + (cv ? 0 : 1);
  TexTap t;
  t.fu = u - (float)iu0; t.fv = v - (float)iv0;
  t.o00 = iv0 * w + iu0; t.o10 = iv0 * w + iu1; t.o01 = iv1 * w + iu0; t.o11 = iv1 * w + iu1;
  return t;
}
__device__ __forceinline__ float lerpf(float a, float b, float t) { return fmaf(t, b - a, a); }
__device__ __forceinline__ float4 bilerp4(float4 a00, float4 a10, float4 a01, float4 a11, float fu, float fv) {
  // same operation order as the scalar lerpf form, two channels per instruction (FFMA2 / FADD2)
  const unsigned long long uu = pack2(fu, fu), vv = pack2(fv, fv);
  unsigned long long lo = lerp2(lerp2(pack2(a00.x, a00.y), pack2(a10.x, a10.y), uu), lerp2(pack2(a01.x, a01.y), pack2(a11.x, a11.y), uu), vv);
  unsigned long long hi = lerp2(lerp2(pack2(a00.z, a00.w), pack2(a10.z, a10.w), uu), lerp2(pack2(a01.z, a01.w), pack2(a11.z, a11.w), uu), vv);
  float4 r;
  unpack2(lo, r.x, r.y);
  unpack2(hi, r.z, r.w);
  return r;
}


// tensor-core (tcgen05) variant, gdb_render_tc.cu
int render_tc_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, int split, cudaStream_t st);
// second-generation tensor-core variant (fp16 operands), gdb_render_tc2.cu
int render_tc2_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, int gen, cudaStream_t st);
// fourth-generation tensor-core kernel (fp16 operands, the default of precision 1), gdb_render_tc3.cu
int render_tc3_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, cudaStream_t st);
// split-fp16 (fp32-class) arithmetic in the third-generation layout, gdb_render_tc4.cu: the production kernel of precision 2 where it
// covers the call (2x2 bundles, three source views, no parity taps); the first-generation split kernel handles the rest
bool render_tc4_covers(const RenderParams& p, int bundle_size, int feat_dim, int V);
int render_tc4_dispatch(const RenderParams& p, int bundle_size, int feat_dim, int V, cudaStream_t st);

}  // namespace gdb
