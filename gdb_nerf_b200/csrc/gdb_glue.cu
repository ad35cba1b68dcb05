// Element-wise glue between the hand-written path and the cuDNN networks
// (SURVEY.md section 8f ranks 1-2): the epilogues PyTorch would otherwise run as
// three or four separate full-map passes each.  All tensors channels-last, fp32,
// C % 4 == 0, one float4 per thread.
//
//   gdb_bias_act_add : out = skip(+nearest x2) + act(x + bias)
//                      cost_reg_net.py:108-110 (y = s + relu(bn(deconv)))  and the
//                      FPN top-down step feature_net.py:52-58
//   gdb_gate_add     : out = x + y * gate[n, c]      decoder_rdn.py squeeze-excite residual
//   gdb_concat3      : channel concatenation of up to three maps (dense block inputs, decoder_rdn.py:36-41)
//   gdb_pixel_shuffle2: PixelShuffle(2) + the producing convolution's bias on channels-last maps (decoder_rdn.py:76-80)
//   gdb_channel_mean : per-image channel means (squeeze step of the squeeze-excite gate), deterministic two-stage sum
#include <algorithm>

#include "gdb_common.cuh"

namespace gdb {

// idx_t = unsigned when the element count fits 32 bits: the 64-bit divisions of the index decomposition made the
// up-sampling variant ALU-bound (0.136 ms for 567 MB; 4.2 TB/s) - with 32-bit ones it streams
template <typename idx_t>
__global__ void bias_act_add_kernel(const float4* __restrict__ x, const float* __restrict__ bias, const float4* __restrict__ skip,
                                    int C4_, int64_t n4_, int relu, int up2, int Hs_, int Ws_, float4* __restrict__ out) {
  // up2: skip is (N, Hs, Ws, C) and x/out are (N, 2Hs, 2Ws, C): nearest-neighbour x2 of skip
  const idx_t C4 = (idx_t)C4_, n4 = (idx_t)n4_, Hs = (idx_t)Hs_, Ws = (idx_t)Ws_;
  for (idx_t i = blockIdx.x * (idx_t)blockDim.x + threadIdx.x; i < n4; i += (idx_t)gridDim.x * blockDim.x) {
    const idx_t pix = i / C4;
    const idx_t c4 = i - pix * C4;
    float4 v = x[i];
    if (bias) {
      float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (skip) {
      idx_t j = i;
      if (up2) {
        const idx_t t = pix / (2 * Ws);
        const idx_t xo = pix - t * (2 * Ws);
        const idx_t n = t / (2 * Hs);
        const idx_t yo = t - n * (2 * Hs);
        j = ((n * Hs + (yo >> 1)) * Ws + (xo >> 1)) * C4 + c4;
      }
      float4 s = __ldg(skip + j);
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    out[i] = v;
  }
}

__global__ void gate_add_kernel(const float4* __restrict__ x, const float4* __restrict__ y, const float* __restrict__ gate,
                                const float4* __restrict__ extra, int C4, int64_t per_image4, int64_t n4, float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    int64_t n = i / per_image4;
    float4 g = __ldg(reinterpret_cast<const float4*>(gate) + n * C4 + c4);
    float4 a = x[i], b = y[i];
    float4 r = make_float4(fmaf(b.x, g.x, a.x), fmaf(b.y, g.y, a.y), fmaf(b.z, g.z, a.z), fmaf(b.w, g.w, a.w));
    if (extra) {                      // the decoder's global residual (decoder_rdn.py: y + blocks(y)) folded into the last block
      const float4 e = extra[i];
      r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
    }
    out[i] = r;
  }
}


__global__ void concat3_kernel(const float4* __restrict__ a, int a4, const float4* __restrict__ b, int b4, const float4* __restrict__ c,
                               int c4, int64_t n4, float4* __restrict__ out) {
  const int t4 = a4 + b4 + c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / t4;
    const int q = (int)(i - pix * t4);
    float4 v;
    if (q < a4) v = __ldcs(a + pix * a4 + q);
    else if (q < a4 + b4) v = __ldcs(b + pix * b4 + (q - a4));
    else v = __ldcs(c + pix * c4 + (q - a4 - b4));
    out[i] = v;
  }
}

// stage 1: CTA (chunk, n) sums its pixels per channel; 256 threads = (256 / C4) pixel lanes x C4 channel quads
__global__ void __launch_bounds__(256) channel_sum_kernel(const float4* __restrict__ x, int C4, int64_t S, int chunks, float4* __restrict__ partial) {
  __shared__ float4 red[256];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int q = threadIdx.x % C4, pl = threadIdx.x / C4, PL = 256 / C4;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = chunk * per, p1 = min(p0 + per, S);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pl < PL)
    for (int64_t p = p0 + pl; p < p1; p += PL) {
      const float4 v = __ldg(x + ((int64_t)n * S + p) * C4 + q);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (pl == 0) {
    for (int k = 1; k < PL; ++k) {
      const float4 v = red[k * C4 + q];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    partial[((int64_t)n * chunks + chunk) * C4 + q] = acc;
  }
}
// stage 2: fixed-order sum over the chunks
__global__ void channel_mean_finish_kernel(const float* __restrict__ partial, int C, int chunks, int64_t N, float inv_S, float* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int64_t n = i / C;
  const int c = (int)(i - n * C);
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[((int64_t)n * chunks + k) * C + c];
  out[i] = s * inv_S;
}

// PixelShuffle(2) of a channels-last map with the producing convolution's bias: in (N,H,W,4C) -> out (N,2H,2W,C),
// out[n, 2y+dy, 2x+dx, c] = in[n, y, x, 4c + 2dy + dx] + bias[4c + 2dy + dx]   (decoder_rdn.py:76-80)
// thread = (input pixel, four output channels): 64 contiguous bytes in, a 4x4 transpose, one float4 to each of the 4 output pixels
__global__ void pixel_shuffle2_kernel(const float4* __restrict__ in, const float* __restrict__ bias, int C4, int H, int W, int64_t n,
                                      float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / C4;
    const int q = (int)(i - pix * C4);                 // output channels 4q .. 4q+3
    const int x = (int)(pix % W);
    const int64_t t = pix / W;
    const int y = (int)(t % H);
    const int64_t img = t / H;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                      // v[k] = input channels 4(4q+k) .. +3 = (dy,dx) of output channel 4q+k
      v[k] = __ldcs(in + (pix * C4 + q) * 4 + k);
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + q * 4 + k);
        v[k].x += b.x; v[k].y += b.y; v[k].z += b.z; v[k].w += b.w;
      }
    }
    const int64_t o00 = ((img * 2 * H + 2 * y) * 2 * W + 2 * x) * C4 + q;
    out[o00] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
    out[o00 + C4] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
    out[o00 + (int64_t)2 * W * C4] = make_float4(v[0].z, v[1].z, v[2].z, v[3].z);
    out[o00 + (int64_t)2 * W * C4 + C4] = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
  }
}

}  // namespace gdb

using namespace gdb;

extern "C" int gdb_bias_act_add(const float* x, const float* bias, const float* skip, int64_t N, int64_t S, int C, int relu,
                                int skip_up2, int Hs, int Ws, float* out, void* stream) {
  GDB_REQUIRE(x && out && N > 0 && S > 0 && C > 0 && C % 4 == 0, GDB_E_BADARG, "gdb_bias_act_add: bad argument (C %% 4 must be 0)");
  GDB_REQUIRE(aligned16(x) && aligned16(out) && (!skip || aligned16(skip)) && (!bias || aligned16(bias)), GDB_E_ALIGN,
              "gdb_bias_act_add: pointers must be 16-byte aligned");
  GDB_REQUIRE(!skip_up2 || (skip && S == (int64_t)4 * Hs * Ws), GDB_E_BADARG, "gdb_bias_act_add: up2 needs skip and S == 4*Hs*Ws");
  int64_t n4 = N * S * (C / 4);
  int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 16);
  if (n4 + (int64_t)blocks * 256 < ((int64_t)1 << 32))        // the grid-stride index stays below 2^32
    bias_act_add_kernel<unsigned><<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), bias,
                                                                         reinterpret_cast<const float4*>(skip), C / 4, n4, relu, skip_up2,
                                                                         Hs, Ws, reinterpret_cast<float4*>(out));
  else
    bias_act_add_kernel<int64_t><<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), bias,
                                                                        reinterpret_cast<const float4*>(skip), C / 4, n4, relu, skip_up2,
                                                                        Hs, Ws, reinterpret_cast<float4*>(out));
  return cuda_check("gdb_bias_act_add");
}

extern "C" int gdb_gate_add(const float* x, const float* y, const float* gate, const float* extra, int64_t N, int64_t S, int C, float* out,
                            void* stream) {
  GDB_REQUIRE(x && y && gate && out && N > 0 && S > 0 && C > 0 && C % 4 == 0, GDB_E_BADARG, "gdb_gate_add: bad argument");
  GDB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gate) && aligned16(out) && (!extra || aligned16(extra)), GDB_E_ALIGN, "gdb_gate_add: pointers must be 16-byte aligned");
  int64_t n4 = N * S * (C / 4);
  int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 16);
  gate_add_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y), gate,
                                                         reinterpret_cast<const float4*>(extra), C / 4, S * (C / 4), n4,
                                                         reinterpret_cast<float4*>(out));
  return cuda_check("gdb_gate_add");
}

// Squeeze-excite gate computed in the prologue of the gated residual: every CTA of image n reduces the channel-sum partials
// (fixed order), runs the two tiny fully-connected layers (C -> R -> C, ReLU / sigmoid, modules.py:5-20) into shared memory and
// then streams out = x + y * gate (+ extra).  Replaces five launch-bound kernels (finish, 2 GEMV, clamp, sigmoid) per block.
__global__ void __launch_bounds__(256) se_gate_add_kernel(const float4* __restrict__ x, const float4* __restrict__ y,
                                                          const float* __restrict__ partial, int chunks, float inv_S,
                                                          const float* __restrict__ w1, const float* __restrict__ w2, int R,
                                                          const float4* __restrict__ extra, int C, int64_t per_image4,
                                                          float4* __restrict__ out, float4* __restrict__ out2, int out2_s4) {
  extern __shared__ float se_sm[];            // mean[C] | hid[R] | gate[C]
  float* mean = se_sm;
  float* hid = se_sm + C;
  float* gate = hid + ((R + 3) & ~3);
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += partial[((int64_t)n * chunks + k) * C + c];
    mean[c] = s * inv_S;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(w1[r * C + c], mean[c], s);
    hid[r] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s = fmaf(w2[c * R + r], hid[r], s);
    gate[c] = 1.f / (1.f + expf(-s));
  }
  __syncthreads();
  const int C4 = C / 4;
  const float4* g4 = reinterpret_cast<const float4*>(gate);
  const int64_t base = (int64_t)n * per_image4;
  const int64_t pix0 = (int64_t)n * (per_image4 / C4);        // first pixel of image n
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_image4; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned pl = (unsigned)i / (unsigned)C4, c4 = (unsigned)i - pl * (unsigned)C4;      // per_image4 < 2^32 (checked by the host)
    const float4 g = g4[c4];
    const float4 a = x[base + i], b = y[base + i];
    float4 r = make_float4(fmaf(b.x, g.x, a.x), fmaf(b.y, g.y, a.y), fmaf(b.z, g.z, a.z), fmaf(b.w, g.w, a.w));
    if (extra) {
      const float4 e = extra[base + i];
      r.x += e.x; r.y += e.y; r.z += e.z; r.w += e.w;
    }
    out[base + i] = r;
    if (out2) {        // second copy into the leading channels of the next block's concatenation buffer (row pitch out2_s4 float4)
      out2[(pix0 + pl) * out2_s4 + c4] = r;
    }
  }
}

// [a | b] written at channel offset `off4` of rows that are `stride4` float4 wide (the trailing slices of a concatenation
// buffer whose leading slice was written by the producer of that tensor)
__global__ void concat2_into_kernel(const float4* __restrict__ a, int a4, const float4* __restrict__ b, int b4, int64_t n4,
                                    float4* __restrict__ out, int stride4, int off4) {
  const int t4 = a4 + b4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / t4;
    const int q = (int)(i - pix * t4);
    const float4 v = q < a4 ? __ldcs(a + pix * a4 + q) : __ldcs(b + pix * b4 + (q - a4));
    out[pix * stride4 + off4 + q] = v;
  }
}

extern "C" int gdb_concat2_into(const float* a, int Ca, const float* b, int Cb, int64_t npix, float* out, int out_channels, int out_offset,
                                void* stream) {
  GDB_REQUIRE(a && b && out && npix > 0 && Ca > 0 && Cb > 0, GDB_E_BADARG, "gdb_concat2_into: bad argument");
  GDB_REQUIRE(Ca % 4 == 0 && Cb % 4 == 0 && out_channels % 4 == 0 && out_offset % 4 == 0 && out_offset >= 0 &&
                  out_offset + Ca + Cb <= out_channels,
              GDB_E_BADARG, "gdb_concat2_into: channel counts / offset must be multiples of 4 and fit the row");
  GDB_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), GDB_E_ALIGN, "gdb_concat2_into: pointers must be 16-byte aligned");
  const int64_t n4 = npix * ((Ca + Cb) / 4);
  int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 16);
  concat2_into_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a), Ca / 4, reinterpret_cast<const float4*>(b),
                                                             Cb / 4, n4, reinterpret_cast<float4*>(out), out_channels / 4, out_offset / 4);
  return cuda_check("gdb_concat2_into");
}

extern "C" int gdb_se_gate_add(const float* x, const float* y, const float* w1, const float* w2, int R, const float* extra, int64_t N,
                               int64_t S, int C, int chunks, float* partial, float* out, void* stream) {
  return gdb_se_gate_add_cat(x, y, w1, w2, R, extra, N, S, C, chunks, partial, out, nullptr, 0, stream);
}

extern "C" int gdb_se_gate_add_cat(const float* x, const float* y, const float* w1, const float* w2, int R, const float* extra, int64_t N,
                                   int64_t S, int C, int chunks, float* partial, float* out, float* out2, int out2_channels, void* stream) {
  GDB_REQUIRE(x && y && w1 && w2 && partial && out && N > 0 && S > 0 && chunks > 0 && R > 0 && R <= 256, GDB_E_BADARG, "gdb_se_gate_add: bad argument");
  GDB_REQUIRE(!out2 || (aligned16(out2) && out2_channels % 4 == 0 && out2_channels >= C), GDB_E_BADARG,
              "gdb_se_gate_add_cat: out2 must be 16-byte aligned with a row of out2_channels >= C floats (multiple of 4)");
  GDB_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024, GDB_E_BADARG, "gdb_se_gate_add: C must be a multiple of 4 in [4, 1024]");
  GDB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(out) && aligned16(partial) && (!extra || aligned16(extra)), GDB_E_ALIGN,
              "gdb_se_gate_add: pointers must be 16-byte aligned");
  dim3 grid1(chunks, (unsigned)N);
  channel_sum_kernel<<<grid1, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(y), C / 4, S, chunks, reinterpret_cast<float4*>(partial));
  const int64_t per4 = S * (C / 4);
  GDB_REQUIRE(per4 < ((int64_t)1 << 32), GDB_E_UNSUPPORTED, "gdb_se_gate_add: more than 2^32 float4 per image");
  int bx = (int)std::min<int64_t>((per4 + 255) / 256, std::max<int64_t>(1, (int64_t)sm_count() * 16 / N));
  dim3 grid2(bx, (unsigned)N);
  const int smem = (2 * C + ((R + 3) & ~3)) * (int)sizeof(float);
  // hid is padded to a multiple of 4 floats so that gate stays float4-aligned
  se_gate_add_kernel<<<grid2, 256, smem, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y), partial,
                                                               chunks, 1.f / (float)S, w1, w2, R, reinterpret_cast<const float4*>(extra), C,
                                                               per4, reinterpret_cast<float4*>(out), reinterpret_cast<float4*>(out2),
                                                               out2_channels / 4);
  return cuda_check("gdb_se_gate_add");
}

extern "C" int gdb_concat3(const float* a, int Ca, const float* b, int Cb, const float* c, int Cc, int64_t npix, float* out, void* stream) {
  GDB_REQUIRE(a && b && out && npix > 0 && Ca > 0 && Cb > 0 && Cc >= 0 && (Cc == 0 || c), GDB_E_BADARG, "gdb_concat3: bad argument");
  GDB_REQUIRE(Ca % 4 == 0 && Cb % 4 == 0 && Cc % 4 == 0, GDB_E_BADARG, "gdb_concat3: channel counts must be multiples of 4");
  GDB_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out) && (!c || aligned16(c)), GDB_E_ALIGN, "gdb_concat3: pointers must be 16-byte aligned");
  const int64_t n4 = npix * ((Ca + Cb + Cc) / 4);
  int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 16);
  concat3_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a), Ca / 4, reinterpret_cast<const float4*>(b), Cb / 4,
                                                        reinterpret_cast<const float4*>(c), Cc / 4, n4, reinterpret_cast<float4*>(out));
  return cuda_check("gdb_concat3");
}

extern "C" int gdb_channel_mean(const float* x, int64_t N, int64_t S, int C, int chunks, float* partial, float* out, void* stream) {
  GDB_REQUIRE(x && partial && out && N > 0 && S > 0 && chunks > 0, GDB_E_BADARG, "gdb_channel_mean: bad argument");
  GDB_REQUIRE(C % 4 == 0 && C >= 4 && C <= 1024, GDB_E_BADARG, "gdb_channel_mean: C must be a multiple of 4 in [4, 1024]");
  GDB_REQUIRE(aligned16(x) && aligned16(partial), GDB_E_ALIGN, "gdb_channel_mean: pointers must be 16-byte aligned");
  dim3 grid(chunks, (unsigned)N);
  channel_sum_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), C / 4, S, chunks, reinterpret_cast<float4*>(partial));
  const int64_t n = N * C;
  channel_mean_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(partial, C, chunks, N, 1.f / (float)S, out);
  return cuda_check("gdb_channel_mean");
}

extern "C" int gdb_pixel_shuffle2(const float* in, const float* bias, int64_t N, int H, int W, int C, float* out, void* stream) {
  GDB_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, GDB_E_BADARG, "gdb_pixel_shuffle2: bad argument (C %% 4 must be 0)");
  GDB_REQUIRE(aligned16(in) && aligned16(out) && (!bias || aligned16(bias)), GDB_E_ALIGN, "gdb_pixel_shuffle2: pointers must be 16-byte aligned");
  const int64_t n = N * H * W * (C / 4);
  int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
  pixel_shuffle2_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(in), bias, C / 4, H, W, n, reinterpret_cast<float4*>(out));
  return cuda_check("gdb_pixel_shuffle2");
}
