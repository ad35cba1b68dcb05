// Device functions shared by the sampling kernels and the fused render kernel:
// the per-bundle geometry and the depth-guided sample placement of
// networks/gdb_nerf/bundle_sampler.py:76-265.
#pragma once
#include "gdb_common.cuh"

namespace gdb {

// samples per bundle (bundle_sampler.py:152 fixed, :179 adaptive).  IEEE ops
// only: this integer is part of the bit-exact contract.
__device__ __forceinline__ int bundle_sample_count(float nr, float fr, float min_interval, int max_samples, int inv_depth,
                                                   int adaptive) {
  if (!adaptive) return max_samples;
  if (inv_depth) { nr = fdiv(1.f, nr); fr = fdiv(1.f, fr); }
  float c = ceilf(fdiv(fabsf(fsub(fr, nr)), min_interval));
  c = fminf(fmaxf(c, 1.f), (float)max_samples);   // clamp(1, max); NaN -> 1 like torch.clamp's min-then-max on CPU
  if (!(c >= 1.f)) c = 1.f;
  return (int)c;
}

// bin midpoint of slot s of n and its normalised volume coordinate (:183,246-251).
// nr/fr/vn/vf are already in the sampling domain (disparity if inv_depth).
__device__ __forceinline__ void sample_depth(float nr, float fr, float vn, float vf, int n, int s, int inv_depth, float& z,
                                             float& d) {
  float step = fdiv(fsub(fr, nr), (float)n);
  float t0 = fadd(nr, fmul(step, (float)s));
  float t1 = fadd(nr, fmul(step, (float)(s + 1)));
  z = fmul(0.5f, fadd(t0, t1));
  d = fsub(fdiv(fmul(2.f, fsub(z, vn)), fsub(vf, vn)), 1.f);
  if (inv_depth) z = fdiv(1.f, z);
}

template <int BS>
struct BundleGeom {
  float x0, y0;       // centre of the bundle's top-left pixel
  float u, v;         // bundle centre in [-1, 1]                       (:104)
  float unit_ball;    // ball radius per unit distance                 (:262)

  __device__ __forceinline__ void ray_dir(const float* __restrict__ head, int j, float& dx, float& dy, float& dz) const {
    float x = x0 + (float)(j % BS), y = y0 + (float)(j / BS);
    const float* M = head + CAM_M;
    dx = fmaf(x, M[0], fmaf(y, M[1], M[2]));
    dy = fmaf(x, M[3], fmaf(y, M[4], M[5]));
    dz = fmaf(x, M[6], fmaf(y, M[7], M[8]));
  }

  __device__ __forceinline__ void init(const float* __restrict__ head, int yb, int xb, int H, int W) {
    constexpr int BB = BS * BS;
    x0 = (float)(xb * BS) + 0.5f;
    y0 = (float)(yb * BS) + 0.5f;
    float su = 0.f, sv = 0.f, mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
    for (int j = 0; j < BB; ++j) {
      float x = x0 + (float)(j % BS), y = y0 + (float)(j / BS);
      su += fsub(fdiv(fmul(2.f, x), (float)W), 1.f);
      sv += fsub(fdiv(fmul(2.f, y), (float)H), 1.f);
      float dx, dy, dz;
      ray_dir(head, j, dx, dy, dz);
      mx += dx; my += dy; mz += dz;
    }
    const float inv = 1.f / (float)BB;
    u = su * inv;
    v = sv * inv;
    mx *= inv; my *= inv; mz *= inv;
    float nrm = sqrtf(mx * mx + my * my + mz * mz);
    float cs = (mx * head[CAM_ZAXIS + 0] + my * head[CAM_ZAXIS + 1] + mz * head[CAM_ZAXIS + 2]) / nrm;
    float disk = head[CAM_DISK];
    float tn = sqrtf(fmaxf(1.f / (cs * cs) - 1.f, 1e-12f)) - disk;
    unit_ball = disk * cs / sqrtf(tn * tn + 1.f);
  }
};

}  // namespace gdb
