// K2 fused with the probability head, TMA-fed (second generation of gdb_prob_head_depth_range_split_fwd).
//
// Reference: cost_reg_net.py:62-63 (prob = Conv3d(8 -> 1, 3x3x3, padding 1, no bias)) -> depth_net.py:172,479-514 (soft-max over
// depth, depth regression, confidence interval).
//
// The first-generation kernels (gdb_costvolume.cu) stage each haloed depth plane by hand (LDG -> STS, two barriers per plane) and
// read 54 data + 54 weight LDS.128 per output voxel: ncu showed them bound by shared-memory wavefronts and by the barrier chain.
// This kernel changes three things:
//  * the haloed planes arrive by TMA (`cp.async.bulk.tensor.5d`, one elected thread, a ring of PHT_STAGES planes with full /
//    empty mbarriers): the tensor map's out-of-bounds zero fill IS the convolution's padding in x, y and depth, nothing is staged
//    through registers and there is no CTA-wide barrier in the plane loop.  The 32-byte voxels are stored with the 32-byte
//    swizzle so that the float4 reads of neighbouring pixels fall on different banks;
//  * each input plane is read ONCE (18 LDS.128 per thread) and scattered into the three output planes it contributes to
//    (three rotating accumulator pairs) instead of being gathered three times: 3x fewer data wavefronts;
//  * the 216 weights sit in constant memory (a 864-byte device-to-device copy per call, a memcpy node under stream capture)
//    and enter the FMAs as constant operands: no weight loads at all.
// The depth axis stays split over CTAs with on-line soft-max partials merged by the last CTA of a tile, as in generation 1.
#include <cuda.h>

#include <mutex>

#include "gdb_common.cuh"

namespace gdb {

__device__ __forceinline__ float hypothesis_ph(float near_, float far_, int d, int D, int inv_depth) {
  if (inv_depth) {
    near_ = fdiv(1.f, near_);
    far_ = fdiv(1.f, far_);
  }
  return fadd(near_, fmul(fsub(far_, near_), linspace01(d, D)));
}

constexpr int PHT_TY = 8, PHT_TX = 16, PHT_THREADS = PHT_TY * PHT_TX, PHT_PW = PHT_TX + 2, PHT_PH = PHT_TY + 2;
constexpr int PHT_PLANE_BYTES = PHT_PH * PHT_PW * 32;                 // one haloed plane of 8-float voxels
constexpr int PHT_STAGE_BYTES = (PHT_PLANE_BYTES + 1023) / 1024 * 1024;
constexpr int PHT_STAGES = 4;

__constant__ float4 c_ph_w[54];                                       // [kd][ky][kx][2 halves of the 8 channels]

__device__ __forceinline__ uint32_t pht_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pht_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pht_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void pht_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pht_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one haloed plane: box (8 channels, PW columns, PH rows, 1 depth, 1 batch) at (0, x, y, d, b); out-of-range coordinates read zeros
__device__ __forceinline__ void pht_tma_plane(uint32_t dst, const CUtensorMap* map, int x, int y, int d, int b, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(map), "r"(0), "r"(x), "r"(y), "r"(d), "r"(b), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(PHT_THREADS) prob_head_tma_kernel(const __grid_constant__ CUtensorMap ymap, const float* __restrict__ range,
                                                                    int rh, int rw, int B, int D, int H, int W, int NCH, float ci_scale,
                                                                    int inv_depth, float4* __restrict__ scratch, int* __restrict__ counters,
                                                                    float* __restrict__ depth, float* __restrict__ ci,
                                                                    float* __restrict__ vol_range) {
  extern __shared__ __align__(1024) unsigned char ring[];            // PHT_STAGES planes
  __shared__ __align__(8) unsigned long long bars[2 * PHT_STAGES];   // full[s], empty[s]
  __shared__ int s_last;
  const int tid = threadIdx.x, ty = tid / PHT_TX, tx = tid % PHT_TX;
  const int b = blockIdx.z / NCH, chunk = blockIdx.z - b * NCH;
  const int y0 = blockIdx.y * PHT_TY, x0 = blockIdx.x * PHT_TX;
  const int HW = H * W;
  const int per = (D + NCH - 1) / NCH;
  const int dlo = chunk * per, dhi = min(dlo + per, D);
  const int nplanes = dhi - dlo + 2;                                   // input planes dlo - 1 .. dhi
  const uint32_t ring_s = pht_smem(ring);
  const uint32_t full0 = pht_smem(&bars[0]), empty0 = pht_smem(&bars[PHT_STAGES]);
  if (tid == 0) {
    for (int s = 0; s < PHT_STAGES; ++s) {
      pht_mbar_init(full0 + 8 * s, 1);
      pht_mbar_init(empty0 + 8 * s, PHT_THREADS / 32);                 // one arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&ymap) : "memory");
    for (int i = 0; i < PHT_STAGES && i < nplanes; ++i) {
      pht_expect_tx(full0 + 8 * i, PHT_PLANE_BYTES);
      pht_tma_plane(ring_s + i * PHT_STAGE_BYTES, &ymap, x0 - 1, y0 - 1, dlo - 1 + i, b, full0 + 8 * i);
    }
  }
  const int gy = y0 + ty, gx = x0 + tx;
  const bool live = gy < H && gx < W;
  const int ry = rh == 1 ? 0 : min(gy, H - 1), rx = rw == 1 ? 0 : min(gx, W - 1);
  const float near_ = range[((size_t)(b * 2 + 0) * rh + ry) * rw + rx];
  const float far_ = range[((size_t)(b * 2 + 1) * rh + ry) * rw + rx];
  const float first = hypothesis_ph(near_, far_, 0, D, inv_depth), last = hypothesis_ph(near_, far_, D - 1, D, inv_depth);
  const float cmid = 0.5f * (first + last);
  float m = -INFINITY, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  // byte offsets of my nine neighbours inside a plane (32-byte swizzle: the 16-byte half index is XORed with address bit 7)
  int off[9], swz[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int pl = (ty + ky) * PHT_PW + tx + kx;
      off[ky * 3 + kx] = pl * 32;
      swz[ky * 3 + kx] = ((pl >> 2) & 1) * 16;
    }
  // accumulator pairs of the output planes p - 1 (complete after this input plane), p and p + 1
  float a_prev0 = 0.f, a_prev1 = 0.f, a_cur0 = 0.f, a_cur1 = 0.f;
#pragma unroll 1
  for (int i = 0; i < nplanes; ++i) {
    const int s = i % PHT_STAGES;
    const int p = dlo - 1 + i;                                         // input plane
    pht_mbar_wait(full0 + 8 * s, (i / PHT_STAGES) & 1);
    const unsigned char* pl = ring + s * PHT_STAGE_BYTES;
    float4 xa[9], xc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      xa[t] = *reinterpret_cast<const float4*>(pl + off[t] + swz[t]);
      xc[t] = *reinterpret_cast<const float4*>(pl + off[t] + (swz[t] ^ 16));
    }
    float a_next0 = 0.f, a_next1 = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 a = xa[t], c = xc[t];
      {   // kd = 2: this plane is the far neighbour of output plane p - 1
        const float4 wa = c_ph_w[(18 + t) * 2], wc = c_ph_w[(18 + t) * 2 + 1];
        a_prev0 = fmaf(a.x, wa.x, a_prev0); a_prev1 = fmaf(a.y, wa.y, a_prev1);
        a_prev0 = fmaf(a.z, wa.z, a_prev0); a_prev1 = fmaf(a.w, wa.w, a_prev1);
        a_prev0 = fmaf(c.x, wc.x, a_prev0); a_prev1 = fmaf(c.y, wc.y, a_prev1);
        a_prev0 = fmaf(c.z, wc.z, a_prev0); a_prev1 = fmaf(c.w, wc.w, a_prev1);
      }
      {   // kd = 1
        const float4 wa = c_ph_w[(9 + t) * 2], wc = c_ph_w[(9 + t) * 2 + 1];
        a_cur0 = fmaf(a.x, wa.x, a_cur0); a_cur1 = fmaf(a.y, wa.y, a_cur1);
        a_cur0 = fmaf(a.z, wa.z, a_cur0); a_cur1 = fmaf(a.w, wa.w, a_cur1);
        a_cur0 = fmaf(c.x, wc.x, a_cur0); a_cur1 = fmaf(c.y, wc.y, a_cur1);
        a_cur0 = fmaf(c.z, wc.z, a_cur0); a_cur1 = fmaf(c.w, wc.w, a_cur1);
      }
      {   // kd = 0: the near neighbour of output plane p + 1
        const float4 wa = c_ph_w[t * 2], wc = c_ph_w[t * 2 + 1];
        a_next0 = fmaf(a.x, wa.x, a_next0); a_next1 = fmaf(a.y, wa.y, a_next1);
        a_next0 = fmaf(a.z, wa.z, a_next0); a_next1 = fmaf(a.w, wa.w, a_next1);
        a_next0 = fmaf(c.x, wc.x, a_next0); a_next1 = fmaf(c.y, wc.y, a_next1);
        a_next0 = fmaf(c.z, wc.z, a_next0); a_next1 = fmaf(c.w, wc.w, a_next1);
      }
    }
    // this warp is done with the stage (its values are in registers: the FMAs above consumed them)
    __syncwarp();
    if ((tid & 31) == 0) pht_mbar_arrive(empty0 + 8 * s);
    // refill the stage of the PREVIOUS plane (every warp has long passed it) with plane i - 1 + STAGES
    if (tid == 0 && i >= 1 && i - 1 + PHT_STAGES < nplanes) {
      const int sp = (i - 1) % PHT_STAGES;
      pht_mbar_wait(empty0 + 8 * sp, ((i - 1) / PHT_STAGES) & 1);
      pht_expect_tx(full0 + 8 * sp, PHT_PLANE_BYTES);
      pht_tma_plane(ring_s + sp * PHT_STAGE_BYTES, &ymap, x0 - 1, y0 - 1, dlo - 1 + (i - 1 + PHT_STAGES), b, full0 + 8 * sp);
    }
    // output plane p - 1 is complete once input plane p has been added
    const int o = p - 1;
    if (o >= dlo && o < dhi) {
      const float l = a_prev0 + a_prev1;
      const float x = hypothesis_ph(near_, far_, o, D, inv_depth) - cmid;
      if (l > m) {
        const float sc = expf(m - l);                                  // 0 on the first plane (m = -inf)
        s0 = fmaf(s0, sc, 1.f); s1 = fmaf(s1, sc, x); s2 = fmaf(s2, sc, x * x);
        m = l;
      } else {
        const float e = expf(l - m);
        s0 += e; s1 = fmaf(e, x, s1); s2 = fmaf(e, x * x, s2);
      }
    }
    a_prev0 = a_cur0; a_prev1 = a_cur1;
    a_cur0 = a_next0; a_cur1 = a_next1;
  }
  const int tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const size_t sbase = ((size_t)tile * NCH) * PHT_THREADS;
  scratch[sbase + (size_t)chunk * PHT_THREADS + tid] = make_float4(m, s0, s1, s2);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(counters + tile, 1);
    s_last = prev == NCH - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float M = -INFINITY;
  for (int c = 0; c < NCH; ++c) M = fmaxf(M, __ldcg(&scratch[sbase + (size_t)c * PHT_THREADS + tid]).x);
  float S0 = 0.f, S1 = 0.f, S2 = 0.f;
  for (int c = 0; c < NCH; ++c) {
    const float4 pt = __ldcg(&scratch[sbase + (size_t)c * PHT_THREADS + tid]);
    const float sc = expf(pt.x - M);
    S0 = fmaf(pt.y, sc, S0); S1 = fmaf(pt.z, sc, S1); S2 = fmaf(pt.w, sc, S2);
  }
  if (!live) return;
  const int pix = gy * W + gx;
  const float mx = S1 / S0;
  const float mean = cmid + mx;
  const float var = fmaxf(S2 / S0 - mx * mx, 0.f);
  const float half = fmul(ci_scale, sqrtf(fmaxf(var, 1e-12f)));
  float lo, hi, dep;
  if (inv_depth) {
    lo = fdiv(1.f, fminf(fadd(mean, half), first));
    hi = fdiv(1.f, fmaxf(fsub(mean, half), last));
    dep = fdiv(1.f, mean);
  } else {
    lo = fmaxf(fsub(mean, half), first);
    hi = fminf(fadd(mean, half), last);
    dep = mean;
  }
  depth[(size_t)b * HW + pix] = dep;
  ci[(size_t)(b * 2 + 0) * HW + pix] = lo;
  ci[(size_t)(b * 2 + 1) * HW + pix] = hi;
  if (vol_range) {
    vol_range[(size_t)(b * 2 + 0) * HW + pix] = first;
    vol_range[(size_t)(b * 2 + 1) * HW + pix] = last;
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace gdb

using namespace gdb;

extern "C" int64_t gdb_prob_head_tma_scratch_floats(int B, int h, int w, int nchunks) {
  const int64_t tiles = (int64_t)B * ((h + PHT_TY - 1) / PHT_TY) * ((w + PHT_TX - 1) / PHT_TX);
  return tiles * nchunks * PHT_THREADS * 4;
}
extern "C" int64_t gdb_prob_head_tma_counters(int B, int h, int w) {
  return (int64_t)B * ((h + PHT_TY - 1) / PHT_TY) * ((w + PHT_TX - 1) / PHT_TX);
}

extern "C" int gdb_prob_head_depth_range_tma_fwd(const float* y_cl, const float* weight, const float* depth_range, int rh, int rw, int B,
                                                 int C, int D, int h, int w, int nchunks, float ci_scale, int inv_depth, float* scratch,
                                                 int* counters, float* depth, float* ci, float* vol_range, void* stream) {
  GDB_REQUIRE(y_cl && weight && depth_range && depth && ci && scratch && counters && B > 0 && D > 0 && h > 0 && w > 0, GDB_E_BADARG,
              "gdb_prob_head_depth_range_tma_fwd: bad argument");
  GDB_REQUIRE(C == 8, GDB_E_UNSUPPORTED, "gdb_prob_head_depth_range_tma_fwd: C=%d not instantiated (8)", C);
  GDB_REQUIRE(nchunks >= 1 && nchunks <= D && (long)B * nchunks <= 65535, GDB_E_BADARG,
              "gdb_prob_head_depth_range_tma_fwd: nchunks %d outside [1, D] or B * nchunks > 65535", nchunks);
  GDB_REQUIRE((rh == 1 && rw == 1) || (rh == h && rw == w), GDB_E_BADARG,
              "gdb_prob_head_depth_range_tma_fwd: depth_range must be 1x1 or %dx%d, got %dx%d", h, w, rh, rw);
  GDB_REQUIRE(aligned16(y_cl) && aligned16(weight) && aligned16(scratch), GDB_E_ALIGN,
              "gdb_prob_head_depth_range_tma_fwd: y / weight / scratch must be 16-byte aligned");
  const int per = (D + nchunks - 1) / nchunks;
  GDB_REQUIRE((nchunks - 1) * per < D, GDB_E_BADARG, "gdb_prob_head_depth_range_tma_fwd: %d chunks of %d planes leave one empty (D = %d)",
              nchunks, per, D);
  EncodeTiledFn enc = encode_tiled();
  GDB_REQUIRE(enc != nullptr, GDB_E_UNSUPPORTED, "gdb_prob_head_depth_range_tma_fwd: the driver does not export cuTensorMapEncodeTiled");
  CUtensorMap map;
  {
    const cuuint64_t dims[5] = {8, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t strides[4] = {32, (cuuint64_t)32 * w, (cuuint64_t)32 * w * h, (cuuint64_t)32 * w * h * D};   // bytes, dims 1..4
    const cuuint32_t box[5] = {8, PHT_PW, PHT_PH, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(y_cl), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GDB_REQUIRE(r == CUDA_SUCCESS, GDB_E_UNSUPPORTED, "gdb_prob_head_depth_range_tma_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
  }
  cudaStream_t st = as_stream(stream);
  dim3 grid((w + PHT_TX - 1) / PHT_TX, (h + PHT_TY - 1) / PHT_TY, B * nchunks);
  {
    cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(int) * (size_t)grid.x * grid.y * B, st);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_ph_w, weight, 54 * sizeof(float4), 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail((int)e, "gdb_prob_head_depth_range_tma_fwd: memset / weight copy: %s", cudaGetErrorString(e));
  }
  constexpr int SMEM = PHT_STAGES * PHT_STAGE_BYTES;
  static SmemOptIn opt;
  {
    cudaError_t e = opt_in_smem(opt, prob_head_tma_kernel, SMEM);
    if (e != cudaSuccess) return fail((int)e, "gdb_prob_head_depth_range_tma_fwd: cudaFuncSetAttribute(%d B): %s", SMEM, cudaGetErrorString(e));
  }
  prob_head_tma_kernel<<<grid, PHT_THREADS, SMEM, st>>>(map, depth_range, rh, rw, B, D, h, w, nchunks, ci_scale, inv_depth,
                                                        reinterpret_cast<float4*>(scratch), counters, depth, ci, vol_range);
  return cuda_check("gdb_prob_head_depth_range_tma_fwd");
}
