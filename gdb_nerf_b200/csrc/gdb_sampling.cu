// Camera block + depth-guided bundle sampling (count / scan / emit).
// Reference: networks/gdb_nerf/bundle_sampler.py:30-265.
#include "gdb_sampling.cuh"

namespace gdb {

// ------------------------------------------------------------ camera block --
__device__ inline bool inv3x3(const double* a, double* o) {
  double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
  return det != 0.0;
}

__device__ inline void inv4x4_d(const float* a, double* out) {
  double m[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      m[i][j] = (double)a[i * 4 + j];
      m[i][j + 4] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    double best = fabs(m[c][c]);
    for (int r = c + 1; r < 4; ++r)
      if (fabs(m[r][c]) > best) { best = fabs(m[r][c]); piv = r; }
    if (piv != c)
      for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
    double inv = 1.0 / m[c][c];
    for (int j = 0; j < 8; ++j) m[c][j] *= inv;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        double f = m[r][c];
        for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
      }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][j + 4];
}

// one thread per (batch, slot): slot 0 = target head, slot 1+v = source view v
__global__ void camera_block_kernel(const float* __restrict__ tar_exts, const float* __restrict__ tar_ints,
                                    const float* __restrict__ src_exts, const float* __restrict__ src_ints,
                                    const float* __restrict__ near_far, int B, int V, int bsize, int gnd, int inv_depth,
                                    float* __restrict__ cam) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (V + 1)) return;
  int b = i / (V + 1), slot = i % (V + 1);
  const int stride = CAM_HEAD + CAM_VIEW * V;
  float* head = cam + (size_t)b * stride;
  const double PI = 3.14159265358979323846;
  if (slot == 0) {
    double c2w[16], K[9], Ki[9];
    inv4x4_d(tar_exts + b * 16, c2w);
    for (int k = 0; k < 9; ++k) K[k] = (double)tar_ints[b * 9 + k];
    inv3x3(K, Ki);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double a = 0.0;
        for (int k = 0; k < 3; ++k) a += c2w[r * 4 + k] * Ki[k * 3 + c];
        head[CAM_M + r * 3 + c] = (float)a;
      }
    for (int r = 0; r < 3; ++r) {
      head[CAM_O + r] = (float)c2w[r * 4 + 3];
      head[CAM_ZAXIS + r] = (float)c2w[r * 4 + 2];
    }
    float fxfy = fmul(fmul(tar_ints[b * 9 + 0], tar_ints[b * 9 + 4]), (float)PI);
    float pr = fdiv(1.f, sqrtf(fxfy));
    head[CAM_PIXR] = pr;
    float nr = near_far[b * 2 + 0], fr = near_far[b * 2 + 1];
    head[CAM_MINIV] = inv_depth ? fdiv(fsub(fdiv(1.f, nr), fdiv(1.f, fr)), (float)gnd) : fdiv(fsub(fr, nr), (float)gnd);
    head[CAM_NEAR] = nr;
    head[CAM_FAR] = fr;
    head[CAM_DISK] = fmul((float)bsize, pr);
    for (int k = CAM_DISK + 1; k < CAM_HEAD; ++k) head[k] = 0.f;
  } else {
    int v = slot - 1;
    float* cv = head + CAM_HEAD + CAM_VIEW * v;
    const float* E = src_exts + ((size_t)b * V + v) * 16;
    const float* K = src_ints + ((size_t)b * V + v) * 9;
    for (int k = 0; k < 12; ++k) cv[CV_E + k] = E[k];
    for (int k = 0; k < 9; ++k) cv[CV_K + k] = K[k];
    double c2w[16];
    inv4x4_d(E, c2w);
    for (int r = 0; r < 3; ++r) cv[CV_C + r] = (float)c2w[r * 4 + 3];
    float fb = (float)bsize;
    cv[CV_PIXR] = fdiv(1.f, sqrtf(fmul(fmul(fdiv(K[0], fb), fdiv(K[4], fb)), (float)PI)));
    for (int k = CV_PIXR + 1; k < CAM_VIEW; ++k) cv[k] = 0.f;
  }
}

// ------------------------------------------------------------------ count --
constexpr int SCAN_ITEMS = 4096;  // bundles per scan block (1024 threads x 4)

__global__ void __launch_bounds__(1024)
bundle_count_kernel(const float* __restrict__ depth_range, const float* __restrict__ cam, int cam_stride, int B, int HW,
                    int max_samples, int inv_depth, int adaptive, int32_t* __restrict__ counts,
                    int32_t* __restrict__ block_sums) {
  __shared__ int warp_sums[32];
  int NB = B * HW;
  int base = blockIdx.x * SCAN_ITEMS;
  int local = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int i = base + k * 1024 + threadIdx.x;
    if (i < NB) {
      int b = i / HW, p = i % HW;
      float nr = depth_range[(size_t)(b * 2 + 0) * HW + p], fr = depth_range[(size_t)(b * 2 + 1) * HW + p];
      int n = bundle_sample_count(nr, fr, cam[(size_t)b * cam_stride + CAM_MINIV], max_samples, inv_depth, adaptive);
      counts[i] = n;
      local += n;
    }
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_sums[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
  }
}

// ------------------------------------------------------------------- scan --
// pass 1 (one CTA): exclusive scan of the block sums in place, total -> offsets[NB]
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int32_t* __restrict__ block_sums, int nblocks, int NB,
                                                               int32_t* __restrict__ offsets) {
  __shared__ int carry_s;
  __shared__ int warp_tot[32];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nblocks ? block_sums[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x];
      int wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      warp_tot[threadIdx.x] = wi - w;  // exclusive
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_tot[threadIdx.x >> 5] + incl - v;
    if (i < nblocks) block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[NB] = carry_s;
}

// pass 2: per block exclusive scan of counts + block prefix
__global__ void __launch_bounds__(1024) scan_apply_kernel(const int32_t* __restrict__ counts,
                                                          const int32_t* __restrict__ block_sums, int NB,
                                                          int32_t* __restrict__ offsets) {
  __shared__ int warp_tot[32];
  int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
  int c[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k] = (base + k < NB) ? counts[base + k] : 0;
  int tsum = c[0] + c[1] + c[2] + c[3];
  int incl = tsum;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = warp_tot[threadIdx.x];
    int wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (threadIdx.x >= o) wi += t;
    }
    warp_tot[threadIdx.x] = wi - w;
  }
  __syncthreads();
  int run = block_sums[blockIdx.x] + warp_tot[threadIdx.x >> 5] + incl - tsum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < NB) offsets[base + k] = run;
    run += c[k];
  }
}

// ------------------------------------------------------------------- emit --
template <int BS>
__global__ void bundle_emit_kernel(const float* __restrict__ depth_range, const float* __restrict__ vol_range,
                                   const float* __restrict__ cam, int cam_stride, const int32_t* __restrict__ counts,
                                   const int32_t* __restrict__ offsets, int B, int Hb, int Wb, int inv_depth,
                                   int64_t* __restrict__ indices, float* __restrict__ z_vals, float* __restrict__ uvd,
                                   float* __restrict__ ball_radii, float* __restrict__ rays_xyz) {
  constexpr int BB = BS * BS;
  int HW = Hb * Wb;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  int b = i / HW, p = i % HW;
  const float* head = cam + (size_t)b * cam_stride;
  BundleGeom<BS> g;
  g.init(head, p / Wb, p % Wb, Hb * BS, Wb * BS);
  float nr = depth_range[(size_t)(b * 2 + 0) * HW + p], fr = depth_range[(size_t)(b * 2 + 1) * HW + p];
  float vn = vol_range[(size_t)(b * 2 + 0) * HW + p], vf = vol_range[(size_t)(b * 2 + 1) * HW + p];
  if (inv_depth) { nr = fdiv(1.f, nr); fr = fdiv(1.f, fr); vn = fdiv(1.f, vn); vf = fdiv(1.f, vf); }
  int n = counts[i];
  int off = offsets[i];
  for (int s = 0; s < n; ++s) {
    float z, d;
    sample_depth(nr, fr, vn, vf, n, s, inv_depth, z, d);
    size_t o = (size_t)off + s;
    if (indices) indices[o] = i;
    if (z_vals) z_vals[o] = z;
    if (uvd) { uvd[o * 3 + 0] = g.u; uvd[o * 3 + 1] = g.v; uvd[o * 3 + 2] = d; }
    if (ball_radii || rays_xyz) {
      float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
      for (int j = 0; j < BB; ++j) {
        float dx, dy, dz;
        g.ray_dir(head, j, dx, dy, dz);
        float x = fmaf(dx, z, head[CAM_O + 0]), y = fmaf(dy, z, head[CAM_O + 1]), zz = fmaf(dz, z, head[CAM_O + 2]);
        cx += x; cy += y; cz += zz;
        if (rays_xyz) {
          rays_xyz[(o * 3 + 0) * BB + j] = x;
          rays_xyz[(o * 3 + 1) * BB + j] = y;
          rays_xyz[(o * 3 + 2) * BB + j] = zz;
        }
      }
      if (ball_radii) {
        const float inv = 1.f / (float)BB;
        cx = cx * inv - head[CAM_O + 0]; cy = cy * inv - head[CAM_O + 1]; cz = cz * inv - head[CAM_O + 2];
        ball_radii[o] = sqrtf(cx * cx + cy * cy + cz * cz) * g.unit_ball;
      }
    }
  }
}

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_camera_block(const float* tar_exts, const float* tar_ints, const float* src_exts, const float* src_ints,
                                const float* near_far, int B, int V, int bundle_size, int global_num_depth, int inv_depth,
                                float* cam, void* stream) {
  GDB_REQUIRE(tar_exts && tar_ints && src_exts && src_ints && near_far && cam, GDB_E_BADARG, "gdb_camera_block: null pointer");
  GDB_REQUIRE(B > 0 && V > 0 && V <= GDB_MAX_VIEWS && bundle_size > 0 && global_num_depth > 0, GDB_E_BADARG,
              "gdb_camera_block: bad size (V must be 1..%d)", GDB_MAX_VIEWS);
  int n = B * (V + 1);
  camera_block_kernel<<<(n + 63) / 64, 64, 0, as_stream(stream)>>>(tar_exts, tar_ints, src_exts, src_ints, near_far, B, V,
                                                                  bundle_size, global_num_depth, inv_depth, cam);
  return cuda_check("gdb_camera_block");
}

extern "C" int gdb_bundle_count(const float* depth_range, const float* cam, int cam_stride, int B, int Hb, int Wb,
                                int max_samples, int inv_depth, int adaptive, int32_t* counts, int32_t* block_sums,
                                void* stream) {
  GDB_REQUIRE(depth_range && cam && counts && block_sums && B > 0 && Hb > 0 && Wb > 0, GDB_E_BADARG, "gdb_bundle_count: bad argument");
  GDB_REQUIRE(max_samples >= 1 && max_samples <= 32, GDB_E_BADARG, "gdb_bundle_count: max_samples must be 1..32");
  int NB = B * Hb * Wb;
  int blocks = (NB + SCAN_ITEMS - 1) / SCAN_ITEMS;
  bundle_count_kernel<<<blocks, 1024, 0, as_stream(stream)>>>(depth_range, cam, cam_stride, B, Hb * Wb, max_samples,
                                                             inv_depth, adaptive, counts, block_sums);
  return cuda_check("gdb_bundle_count");
}

extern "C" int gdb_bundle_scan(const int32_t* counts, int32_t* block_sums, int NB, int32_t* offsets, void* stream) {
  GDB_REQUIRE(counts && block_sums && offsets && NB > 0, GDB_E_BADARG, "gdb_bundle_scan: bad argument");
  int blocks = (NB + SCAN_ITEMS - 1) / SCAN_ITEMS;
  cudaStream_t st = as_stream(stream);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, blocks, NB, offsets);
  scan_apply_kernel<<<blocks, 1024, 0, st>>>(counts, block_sums, NB, offsets);
  return cuda_check("gdb_bundle_scan");
}

extern "C" int gdb_bundle_emit(const float* depth_range, const float* vol_range, const float* cam, int cam_stride,
                               const int32_t* counts, const int32_t* offsets, int B, int Hb, int Wb, int bundle_size,
                               int inv_depth, int64_t* indices, float* z_vals, float* uvd, float* ball_radii,
                               float* rays_xyz, void* stream) {
  GDB_REQUIRE(depth_range && vol_range && cam && counts && offsets && B > 0 && Hb > 0 && Wb > 0, GDB_E_BADARG,
              "gdb_bundle_emit: bad argument");
  int NB = B * Hb * Wb;
  cudaStream_t st = as_stream(stream);
  dim3 grid((NB + 127) / 128);
#define GDB_EMIT(BS)                                                                                                  \
  bundle_emit_kernel<BS><<<grid, 128, 0, st>>>(depth_range, vol_range, cam, cam_stride, counts, offsets, B, Hb, Wb, \
                                               inv_depth, indices, z_vals, uvd, ball_radii, rays_xyz)
  switch (bundle_size) {
    case 1: GDB_EMIT(1); break;
    case 2: GDB_EMIT(2); break;
    case 4: GDB_EMIT(4); break;
    default: return fail(GDB_E_UNSUPPORTED, "gdb_bundle_emit: bundle_size %d not in {1,2,4}", bundle_size);
  }
#undef GDB_EMIT
  return cuda_check("gdb_bundle_emit");
}
