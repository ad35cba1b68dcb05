// Gather-source preparation (feature+rgb texture pyramid, RGBA images) and the
// output assembly that follows the fused render kernel.
// Reference: networks/gdb_nerf/network.py:159-164,175-182 and the mip chain
// nvdiffrast builds inside texture() (bundle_sampler.py:355-359).
#include "gdb_common.cuh"

namespace gdb {

constexpr int TILE_H = 8, TILE_W = 32;

// One CTA = one 8x32 tile of level-0 texels of one (batch, view) image:
//   * transposes the planar FPN feature tile to channels-last,
//   * appends the bilinearly down-sampled source colours (no anti-aliasing:
//     the centre 2x2 of each bxb pixel block, network.py:163),
//   * writes the RGBA-interleaved full-resolution pixels of the tile,
//   * reduces the tile to mip levels 1..L with the 2x2 box
//     0.25*((a00+a10)+(a01+a11)) and writes them.
template <int FP>
__global__ void __launch_bounds__(256)
prepare_sources_kernel(const float* __restrict__ feat, const float* __restrict__ images, int Cf, int feat_cl, int Hb, int Wb,
                       int BS, int L, int64_t lvl1, int64_t lvl2, int64_t lvl3, float* __restrict__ tex, float* __restrict__ rgba) {
  __shared__ __align__(16) float s0[TILE_H * TILE_W * FP];
  __shared__ __align__(16) float s1[(TILE_H / 2) * (TILE_W / 2) * FP];
  __shared__ __align__(16) float s2[(TILE_H / 4) * (TILE_W / 4) * FP];
  __shared__ __align__(16) float s3[(TILE_H / 8) * (TILE_W / 8) * FP];

  const int bv = blockIdx.z;
  const int ty0 = blockIdx.y * TILE_H, tx0 = blockIdx.x * TILE_W;
  const int ncols = min(TILE_W, Wb - tx0), nrows = min(TILE_H, Hb - ty0);
  const int t = threadIdx.x;
  const int row = t / TILE_W, col = t % TILE_W;
  const bool in_tile = row < nrows && col < ncols;
  const int H = Hb * BS, W = Wb * BS;
  const int F = Cf + 3;

  // -- features: coalesced plane reads, channels-last in shared memory
  {
    float* sp = s0 + (row * TILE_W + col) * FP;
    if (feat_cl) {      // (BV, Hb, Wb, Cf): one texel is Cf contiguous floats
      const float* fp = feat + (((size_t)bv * Hb + ty0 + row) * Wb + tx0 + col) * Cf;
      for (int c = 0; c < Cf; c += 4) {
        float4 v = in_tile ? ldg4(fp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        sp[c] = v.x; sp[c + 1] = v.y; sp[c + 2] = v.z; sp[c + 3] = v.w;
      }
    } else {            // (BV, Cf, Hb, Wb): coalesced plane reads
      const float* fp = feat + (size_t)bv * Cf * Hb * Wb + (size_t)(ty0 + row) * Wb + tx0 + col;
      for (int c = 0; c < Cf; ++c) sp[c] = in_tile ? __ldg(fp + (size_t)c * Hb * Wb) : 0.f;
    }
    for (int c = F; c < FP; ++c) sp[c] = 0.f;
  }
  // -- full-resolution pixels of the tile -> RGBA
  {
    const float* ip = images + (size_t)bv * 3 * H * W;
    const int pw = ncols * BS, ph = nrows * BS;
    const int Y0 = ty0 * BS, X0 = tx0 * BS;
    for (int idx = t; idx < pw * ph; idx += 256) {
      int py = Y0 + idx / pw, px = X0 + idx % pw;
      size_t o = (size_t)py * W + px;
      float4 v = make_float4(__ldg(ip + o), __ldg(ip + (size_t)H * W + o), __ldg(ip + 2 * (size_t)H * W + o), 0.f);
      reinterpret_cast<float4*>(rgba)[(size_t)bv * H * W + o] = v;
    }
    // low-resolution colours: bilinear with half-pixel centres lands exactly between the two centre pixels
    float* sp = s0 + (row * TILE_W + col) * FP + Cf;
    if (in_tile) {
      int cy = (ty0 + row) * BS + (BS >> 1) - (BS > 1 ? 1 : 0), cx = (tx0 + col) * BS + (BS >> 1) - (BS > 1 ? 1 : 0);
      for (int c = 0; c < 3; ++c) {
        const float* pl = ip + (size_t)c * H * W + (size_t)cy * W + cx;
        if (BS > 1) {
          // four taps of weight 1/4 summed in tap order, the association ATen's CPU kernel uses (bit-exact with it)
          sp[c] = 0.25f * (((__ldg(pl) + __ldg(pl + 1)) + __ldg(pl + W)) + __ldg(pl + W + 1));
        } else {
          sp[c] = __ldg(pl);
        }
      }
    } else {
      sp[0] = sp[1] = sp[2] = 0.f;
    }
  }
  __syncthreads();

  constexpr int F4 = FP / 4;
  // -- level 0 out: each tile row is ncols*FP contiguous floats
  for (int idx = t; idx < nrows * ncols * F4; idx += 256) {
    int r = idx / (ncols * F4), k = idx % (ncols * F4);
    float4 v = reinterpret_cast<const float4*>(s0 + r * TILE_W * FP)[k];
    reinterpret_cast<float4*>(tex + (((size_t)bv * Hb + ty0 + r) * Wb + tx0) * FP)[k] = v;
  }
  // -- mip levels
  const float* src = s0;
  int sw = TILE_W;                       // row pitch (texels) of the source level in shared memory
  int lh = nrows, lw = ncols;            // valid extent of the source level inside this tile
  int gH = Hb, gW = Wb;                  // global extent of the source level
  for (int lvl = 1; lvl <= L; ++lvl) {
    float* dsts = lvl == 1 ? s1 : (lvl == 2 ? s2 : s3);
    int64_t base = lvl == 1 ? lvl1 : (lvl == 2 ? lvl2 : lvl3);
    lh >>= 1; lw >>= 1; gH >>= 1; gW >>= 1;
    int dw = sw >> 1;
    for (int idx = t; idx < lh * lw * FP; idx += 256) {
      int c = idx % FP, x = (idx / FP) % lw, y = idx / (FP * lw);
      const float* a = src + ((2 * y) * sw + 2 * x) * FP + c;
      float v = 0.25f * ((a[0] + a[FP]) + (a[sw * FP] + a[sw * FP + FP]));
      dsts[(y * dw + x) * FP + c] = v;
      int gy = (ty0 >> lvl) + y, gx = (tx0 >> lvl) + x;
      tex[base + (((size_t)bv * gH + gy) * gW + gx) * FP + c] = v;
    }
    __syncthreads();
    src = dsts;
    sw = dw;
  }
}

// ---------------------------------------------------------------------------
// output assembly (network.py:175-182 without the decoder CNN)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void up_axis(int o, int n_in, float inv_scale, int& i0, int& i1, float& l1) {
  float s = fmaxf(((float)o + 0.5f) * inv_scale - 0.5f, 0.f);
  i0 = min((int)s, n_in - 1);
  i1 = min(i0 + 1, n_in - 1);
  l1 = s - (float)i0;
}

__global__ void assemble_output_kernel(const float* __restrict__ feat, int Ctot, const float* __restrict__ dec,
                                       const float* __restrict__ bdepth, const float* __restrict__ bopac, int B, int Hb,
                                       int Wb, int BS, int reweighting, int layout, const float* __restrict__ dec_bias,
                                       float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ opacity) {
  const bool feat_cl = layout & 1, dec_cl = layout & 2, dec_ps = layout & 4;
  const int H = Hb * BS, W = Wb * BS;
  const size_t HW = (size_t)H * W, hw = (size_t)Hb * Wb;
  size_t n = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int b, y, x;
    if (n <= 0xffffffffull) {           // 32-bit index decomposition (64-bit divisions dominated this streaming kernel)
      const unsigned u = (unsigned)i, hw32 = (unsigned)HW;
      const unsigned bb = u / hw32, r = u - bb * hw32, yy = r / (unsigned)W;
      b = (int)bb; y = (int)yy; x = (int)(r - yy * (unsigned)W);
    } else {
      b = i / HW;
      y = (i % HW) / W; x = i % W;
    }
    int yb = y / BS, xb = x / BS, j = (y % BS) * BS + (x % BS);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float fine = feat_cl ? feat[((size_t)b * hw + (size_t)yb * Wb + xb) * Ctot + c * BS * BS + j]
                           : feat[((size_t)b * Ctot + c * BS * BS + j) * hw + (size_t)yb * Wb + xb];
      float dv;
      if (dec_ps) { // (B, H/2, W/2, 12): the decoder's last convolution before its pixel shuffle, channel = c*4 + (y%2)*2 + x%2
        const int ch = c * 4 + (y & 1) * 2 + (x & 1);
        dv = dec[((size_t)b * (HW / 4) + (size_t)(y >> 1) * (W >> 1) + (x >> 1)) * 12 + ch] + (dec_bias ? __ldg(dec_bias + ch) : 0.f);
      }
      else
        dv = dec_cl ? dec[((size_t)b * HW + (size_t)y * W + x) * 3 + c] : dec[((size_t)b * 3 + c) * HW + (size_t)y * W + x];
      float v = dv + fine;
      if (reweighting) v = 0.5f * (v + fine);
      rgb[((size_t)b * 3 + c) * HW + (size_t)y * W + x] = v;
    }
    int y0, y1, x0, x1;
    float ly, lx;
    up_axis(y, Hb, 1.f / (float)BS, y0, y1, ly);
    up_axis(x, Wb, 1.f / (float)BS, x0, x1, lx);
    const float* dp = bdepth + (size_t)b * hw;
    const float* op = bopac + (size_t)b * hw;
    float hy = 1.f - ly, hx = 1.f - lx;
    depth[i] = hy * (hx * dp[(size_t)y0 * Wb + x0] + lx * dp[(size_t)y0 * Wb + x1]) +
               ly * (hx * dp[(size_t)y1 * Wb + x0] + lx * dp[(size_t)y1 * Wb + x1]);
    opacity[i] = hy * (hx * op[(size_t)y0 * Wb + x0] + lx * op[(size_t)y0 * Wb + x1]) +
                 ly * (hx * op[(size_t)y1 * Wb + x0] + lx * op[(size_t)y1 * Wb + x1]);
  }
}

}  // namespace gdb

using namespace gdb;

static inline int padded_feat(int feat_dim) { return (feat_dim + 3 + 3) & ~3; }

extern "C" int64_t gdb_texture_floats(int BV, int Hb, int Wb, int feat_dim, int max_mip_level) {
  int64_t n = 0;
  for (int k = 0; k <= max_mip_level; ++k) n += (int64_t)BV * (Hb >> k) * (Wb >> k) * padded_feat(feat_dim);
  return n;
}

extern "C" int gdb_prepare_sources(const float* feat, int feat_channels_last, const float* images, int BV, int Cf, int Hb,
                                   int Wb, int bundle_size, int max_mip_level, float* tex, float* rgba, void* stream) {
  GDB_REQUIRE(feat && images && tex && rgba && BV > 0 && Cf > 0 && Hb > 0 && Wb > 0, GDB_E_BADARG, "gdb_prepare_sources: bad argument");
  GDB_REQUIRE(max_mip_level >= 0 && max_mip_level <= 3, GDB_E_UNSUPPORTED, "gdb_prepare_sources: max_mip_level must be 0..3");
  GDB_REQUIRE(bundle_size >= 1 && (bundle_size & (bundle_size - 1)) == 0, GDB_E_BADARG, "gdb_prepare_sources: bundle_size must be a power of two");
  const int m = 1 << max_mip_level;
  GDB_REQUIRE(Hb % m == 0 && Wb % m == 0, GDB_E_BADARG,
              "gdb_prepare_sources: bundle map %dx%d must be divisible by 2^max_mip_level=%d (as nvdiffrast requires)", Hb, Wb, m);
  GDB_REQUIRE(aligned16(tex) && aligned16(rgba), GDB_E_ALIGN, "gdb_prepare_sources: outputs not 16-byte aligned");
  GDB_REQUIRE(!feat_channels_last || (aligned16(feat) && Cf % 4 == 0), GDB_E_ALIGN, "gdb_prepare_sources: channels-last features need 16-byte alignment and Cf % 4 == 0");
  GDB_REQUIRE(BV <= 65535, GDB_E_UNSUPPORTED, "gdb_prepare_sources: B*V > 65535");
  const int FP = padded_feat(Cf);
  int64_t lvl[4] = {0, 0, 0, 0};
  for (int k = 1; k <= 3; ++k) lvl[k] = lvl[k - 1] + (int64_t)BV * (Hb >> (k - 1)) * (Wb >> (k - 1)) * FP;
  dim3 grid((Wb + TILE_W - 1) / TILE_W, (Hb + TILE_H - 1) / TILE_H, BV);
  cudaStream_t st = as_stream(stream);
  switch (FP) {
    case 12: prepare_sources_kernel<12><<<grid, 256, 0, st>>>(feat, images, Cf, feat_channels_last, Hb, Wb, bundle_size, max_mip_level, lvl[1], lvl[2], lvl[3], tex, rgba); break;
    case 20: prepare_sources_kernel<20><<<grid, 256, 0, st>>>(feat, images, Cf, feat_channels_last, Hb, Wb, bundle_size, max_mip_level, lvl[1], lvl[2], lvl[3], tex, rgba); break;
    case 36: prepare_sources_kernel<36><<<grid, 256, 0, st>>>(feat, images, Cf, feat_channels_last, Hb, Wb, bundle_size, max_mip_level, lvl[1], lvl[2], lvl[3], tex, rgba); break;
    default: return fail(GDB_E_UNSUPPORTED, "gdb_prepare_sources: feature width %d not in {8,16,32}", Cf);
  }
  return cuda_check("gdb_prepare_sources");
}

// ----------------------------------------------------------------------------------------------------------------
// 8-bit images -> float32 in [0, 1]: the reference's loaders do `img.astype(np.float32) / 255.` on the host
// (datasets/dataloader/dtu.py:84,135, llff.py:134, nerf.py:132) and ship 4 bytes per sample over PCIe; here the
// 8-bit samples cross the bus and the same IEEE division runs on the device (a 256-entry table of
// __fdiv_rn((float)u, 255.f): bit-identical to numpy's float32 division).
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uint4* __restrict__ in, float4* __restrict__ out, int64_t n16,
                                                         const unsigned char* __restrict__ in_tail, float* __restrict__ out_tail, int tail) {
  __shared__ float lut[256];
  lut[threadIdx.x] = fdiv((float)threadIdx.x, 255.f);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(in + i);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      __stcs(out + i * 4 + k, make_float4(lut[w[k] & 255u], lut[(w[k] >> 8) & 255u], lut[(w[k] >> 16) & 255u], lut[w[k] >> 24]));
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) out_tail[threadIdx.x] = lut[in_tail[threadIdx.x]];
}

extern "C" int gdb_u8_to_unit_f32(const unsigned char* src, float* dst, int64_t n, void* stream) {
  GDB_REQUIRE(src && dst && n > 0, GDB_E_BADARG, "gdb_u8_to_unit_f32: bad argument");
  GDB_REQUIRE(aligned16(src) && aligned16(dst), GDB_E_ALIGN, "gdb_u8_to_unit_f32: src / dst must be 16-byte aligned");
  const int64_t n16 = n / 16;
  const int tail = (int)(n - n16 * 16);
  int blocks = (int)std::min<int64_t>((n16 + 255) / 256 + 1, (int64_t)sm_count() * 8);
  u8_to_unit_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<float4*>(dst), n16,
                                                          src + n16 * 16, dst + n16 * 16, tail);
  return cuda_check("gdb_u8_to_unit_f32");
}

extern "C" int gdb_assemble_output(const float* feat, int Ctot, const float* dec, const float* bdepth, const float* bopacity,
                                   int B, int Hb, int Wb, int bundle_size, int reweighting, int layout, const float* dec_bias,
                                   float* rgb, float* depth, float* opacity, void* stream) {
  GDB_REQUIRE(feat && dec && bdepth && bopacity && rgb && depth && opacity, GDB_E_BADARG, "gdb_assemble_output: null pointer");
  GDB_REQUIRE(B > 0 && Hb > 0 && Wb > 0 && bundle_size > 0 && Ctot >= 3 * bundle_size * bundle_size, GDB_E_BADARG, "gdb_assemble_output: bad size");
  size_t n = (size_t)B * Hb * Wb * bundle_size * bundle_size;
  int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 16);
  assemble_output_kernel<<<blocks, 256, 0, as_stream(stream)>>>(feat, Ctot, dec, bdepth, bopacity, B, Hb, Wb, bundle_size,
                                                               reweighting, layout, (layout & 4) ? dec_bias : nullptr, rgb, depth, opacity);
  return cuda_check("gdb_assemble_output");
}
