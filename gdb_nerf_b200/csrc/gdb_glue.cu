// Element-wise glue between the hand-written path and the cuDNN networks
// (SURVEY.md section 8f ranks 1-2): the epilogues PyTorch would otherwise run as
// three or four separate full-map passes each.  All tensors channels-last, fp32,
// C % 4 == 0, one float4 per thread.
//
//   gdb_bias_act_add : out = skip(+nearest x2) + act(x + bias)
//                      cost_reg_net.py:108-110 (y = s + relu(bn(deconv)))  and the
//                      FPN top-down step feature_net.py:52-58
//   gdb_gate_add     : out = x + y * gate[n, c]      decoder_rdn.py squeeze-excite residual
#include <algorithm>

#include "gdb_common.cuh"

namespace gdb {

__global__ void bias_act_add_kernel(const float4* __restrict__ x, const float* __restrict__ bias, const float4* __restrict__ skip,
                                    int C4, int64_t n4, int relu, int up2, int Hs, int Ws, float4* __restrict__ out) {
  // up2: skip is (N, Hs, Ws, C) and x/out are (N, 2Hs, 2Ws, C): nearest-neighbour x2 of skip
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    float4 v = x[i];
    if (bias) {
      float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (skip) {
      int64_t j = i;
      if (up2) {
        int64_t pix = i / C4;
        int xo = (int)(pix % (2 * Ws));
        int64_t t = pix / (2 * Ws);
        int yo = (int)(t % (2 * Hs));
        int64_t n = t / (2 * Hs);
        j = ((n * Hs + (yo >> 1)) * Ws + (xo >> 1)) * C4 + c4;
      }
      float4 s = __ldg(skip + j);
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    out[i] = v;
  }
}

__global__ void gate_add_kernel(const float4* __restrict__ x, const float4* __restrict__ y, const float* __restrict__ gate, int C4,
                                int64_t per_image4, int64_t n4, float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    int64_t n = i / per_image4;
    float4 g = __ldg(reinterpret_cast<const float4*>(gate) + n * C4 + c4);
    float4 a = x[i], b = y[i];
    out[i] = make_float4(fmaf(b.x, g.x, a.x), fmaf(b.y, g.y, a.y), fmaf(b.z, g.z, a.z), fmaf(b.w, g.w, a.w));
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
  bias_act_add_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), bias,
                                                             reinterpret_cast<const float4*>(skip), C / 4, n4, relu, skip_up2, Hs, Ws,
                                                             reinterpret_cast<float4*>(out));
  return cuda_check("gdb_bias_act_add");
}

extern "C" int gdb_gate_add(const float* x, const float* y, const float* gate, int64_t N, int64_t S, int C, float* out, void* stream) {
  GDB_REQUIRE(x && y && gate && out && N > 0 && S > 0 && C > 0 && C % 4 == 0, GDB_E_BADARG, "gdb_gate_add: bad argument");
  GDB_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gate) && aligned16(out), GDB_E_ALIGN, "gdb_gate_add: pointers must be 16-byte aligned");
  int64_t n4 = N * S * (C / 4);
  int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)sm_count() * 16);
  gate_add_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y), gate,
                                                         C / 4, S * (C / 4), n4, reinterpret_cast<float4*>(out));
  return cuda_check("gdb_gate_add");
}
