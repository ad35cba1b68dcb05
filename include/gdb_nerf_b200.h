/*
 * gdb_nerf_b200 - C ABI of the B200-native GDB-NeRF rendering hot path.
 *
 * The reference (KLMAV-CUC/GDB-NeRF) is pure Python/PyTorch and has no FFI of
 * its own; each entry point below replaces a Python function of the reference
 * (cited as file:line into the reference tree) and is what a ctypes binding in
 * networks/gdb_nerf/*.py would call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; tensors are dense,
 *     row-major, fp32 unless stated; indices are int64 where the reference's are.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work
 *     on it: no allocation, no synchronisation, no host read-back.
 *   - return value: 0 on success, a positive cudaError_t, or a negative
 *     GDB_E_* argument error.  gdb_last_error_string() describes the last
 *     failure of the calling thread.
 *   - "channels-last" (CL) means the channel index is innermost.
 */
#ifndef GDB_NERF_B200_H
#define GDB_NERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDB_ABI_VERSION 2

#define GDB_E_BADARG  (-1)   /* null pointer / non-positive size            */
#define GDB_E_UNSUPPORTED (-2) /* parameter combination not instantiated     */
#define GDB_E_ALIGN   (-3)   /* pointer not 16-byte aligned                  */

#define GDB_MAX_VIEWS 4

int gdb_abi_version(void);
const char* gdb_last_error_string(void);

/* number of floats in the packed MLP parameter block for a given 2-D feature
 * width (feat_dim = 16 or 32) - layout documented in DESIGN.md / mlp_pack.py */
int gdb_mlp_param_floats(int feat_dim);

/* ---------------------------------------------------------------- layout -- */
/* (N, C, S) planar -> (N, S, Cpad) channels-last, zero-filled pad channels.
 * Used for FPN feature maps (C=16/32) and the cost-regularised feature volume
 * (C=8).  C <= Cpad, Cpad % 4 == 0.                                          */
int gdb_planar_to_channels_last(const float* src, float* dst, int N, int C, int64_t S, int Cpad, void* stream);

/* 8-bit image samples -> float32 in [0, 1] on the device: dst[i] = (float)src[i] / 255.f in IEEE arithmetic, bit-identical
 * to the host-side `img.astype(np.float32) / 255.` of the reference's loaders (datasets/dataloader/dtu.py:84,135,
 * llff.py:134, nerf.py:132).  Lets a caller ship 8-bit source images over PCIe (4x fewer bytes).                        */
int gdb_u8_to_unit_f32(const unsigned char* src, float* dst, int64_t n, void* stream);

/* --------------------------------------------------------- cost volume ---- */
/* Homography matrices, replaces depth_net.py:453-457.
 * src_exts (B,V,4,4) src_ints (B,V,3,3) tar_exts (B,4,4) tar_ints (B,3,3);
 * intrinsic rows 0-1 are scaled by src_scale / tar_scale first
 * (depth_net.py:159-162).  out proj (B,V,3,4).                               */
int gdb_homography_mats(const float* src_exts, const float* src_ints, const float* tar_exts, const float* tar_ints,
                        float src_scale, float tar_scale, int B, int V, float* proj, void* stream);

/* Depth hypotheses, replaces get_depth_values depth_net.py:399-421.
 * depth_range (B,2,rh,rw) with (rh,rw) == (1,1) or (Ht,Wt) -> out (B,D,Ht,Wt) */
int gdb_depth_values(const float* depth_range, int rh, int rw, int B, int D, int Ht, int Wt, int inv_depth,
                     float* out, void* stream);

/* Homography warp + population variance over views, replaces
 * build_feature_volume depth_net.py:424-476 fused with get_depth_values.
 * feat_cl (B,V,Hs,Ws,C) channels-last, C in {8,16,32}; proj (B,V,3,4);
 * depth_range as above -> variance (B,C,D,Ht,Wt) (NCDHW, the reference's
 * layout); out_channels_last = 1: (B,D,Ht,Wt,C) (NDHWC: what cuDNN's
 * channels-last 3-D convolutions consume without a layout conversion);
 * out_channels_last = 2: (B,Ht,Wt,D,C), depth folded into the channels of a
 * channels-last 2-D map (the cost-regularisation net run on 2-D kernels).     */
int gdb_warp_variance_fwd(const float* feat_cl, const float* proj, const float* depth_range, int rh, int rw,
                          int B, int V, int C, int Hs, int Ws, int D, int Ht, int Wt, int inv_depth,
                          int out_channels_last, float* variance, void* stream);

/* Depth regression -> confidence interval, replaces depth_regression
 * depth_net.py:479-514 (+ vol_range = depth_values[:, [0,-1]], :179).
 * prob (B,D,h,w) -> depth (B,1,h,w), ci (B,2,h,w), vol_range (B,2,h,w).      */
int gdb_depth_range_fwd(const float* depth_range, int rh, int rw, const float* prob, int B, int D, int h, int w,
                        float ci_scale, int inv_depth, float* depth, float* ci, float* vol_range, void* stream);

/* K2 fused with the soft-max of the probability head (cost_reg_net.py:62-63,
 * 115-116 -> depth_net.py:172): logits element (b,d,pixel) lives at
 * logits[b*stride_b + d*stride_d + pixel*stride_pix] (so it may be one channel
 * of a channels-last multi-head convolution output).  D <= 64.  prob_out
 * (B,D,h,w) planar, optional (null: probabilities are not materialised).      */
int gdb_depth_range_from_logits_fwd(const float* depth_range, int rh, int rw, const float* logits, int64_t stride_b,
                                    int64_t stride_d, int64_t stride_pix, int B, int D, int h, int w, float ci_scale,
                                    int inv_depth, float* depth, float* ci, float* vol_range, float* prob_out, void* stream);

/* --------------------------------------------------------- sampling ------- */
/* Camera block for the sampling / render kernels: replaces
 * BundleSampler.build_rays bundle_sampler.py:30-74 and the per-view constants
 * of encode :303-313.  cam: (B, 32 + 32*V) floats (layout in DESIGN.md).     */
int gdb_camera_block(const float* tar_exts, const float* tar_ints, const float* src_exts, const float* src_ints,
                     const float* near_far, int B, int V, int bundle_size, int global_num_depth, int inv_depth,
                     float* cam, void* stream);

/* Samples per bundle, replaces bundle_sampler.py:152,179.
 * depth_range (B,2,Hb,Wb); counts int32 (NB); block_sums int32
 * (ceil(NB/4096)) partial sums for the scan.                                  */
int gdb_bundle_count(const float* depth_range, const float* cam, int cam_stride, int B, int Hb, int Wb,
                     int max_samples, int inv_depth, int adaptive, int32_t* counts, int32_t* block_sums, void* stream);

/* Exclusive scan of counts -> offsets int32 (NB+1); offsets[NB] = S.
 * The host may read offsets[NB] back once to size the packed outputs, or
 * use the NB*max_samples upper bound and stay sync-free.                      */
int gdb_bundle_scan(const int32_t* counts, int32_t* block_sums, int NB, int32_t* offsets, void* stream);

/* Packed sample list, replaces BundleSampler.sample bundle_sampler.py:193-265.
 * Any output pointer may be null.  indices int64 (S); z_vals (S); uvd (S,3);
 * ball_radii (S); rays_xyz (S,3,b*b).                                         */
int gdb_bundle_emit(const float* depth_range, const float* vol_range, const float* cam, int cam_stride,
                    const int32_t* counts, const int32_t* offsets, int B, int Hb, int Wb, int bundle_size,
                    int inv_depth, int64_t* indices, float* z_vals, float* uvd, float* ball_radii, float* rays_xyz,
                    void* stream);

/* --------------------------------------------------------- sources -------- */
/* Gather sources for the render kernel, replaces network.py:159-164 and the
 * mip construction nvdiffrast.torch.texture does internally
 * (bundle_sampler.py:355-359): feature+rgb texture with its mip chain
 * (channels-last, F=Cf+3 padded to FP=roundup4(F)) and RGBA-interleaved
 * full-resolution images.
 * feat (B,V,Cf,Hb,Wb) planar, or (B,V,Hb,Wb,Cf) when feat_channels_last != 0;
 * images (B,V,3,H,W) planar, H=Hb*b, W=Wb*b.
 * tex: levels 0..L concatenated, level k is (B*V, Hb>>k, Wb>>k, FP).
 * rgba (B*V,H,W,4).  Hb, Wb divisible by 2^L.                                 */
int64_t gdb_texture_floats(int BV, int Hb, int Wb, int feat_dim, int max_mip_level);
int gdb_prepare_sources(const float* feat, int feat_channels_last, const float* images, int BV, int Cf, int Hb, int Wb,
                        int bundle_size, int max_mip_level, float* tex, float* rgba, void* stream);

/* --------------------------------------------------------- render --------- */
/* Fused per-bundle render, replaces BundleSampler.sample/encode
 * (bundle_sampler.py:193-371), NeRF.forward (nerf.py:84-115),
 * render_weight_from_density / accumulate_value_along_rays
 * (utils.py:19-43,88-121) and Network.render_bundles (network.py:54-91).
 *
 * vol_cl (B,D,Hb,Wb,vol_stride) channels-last feature volume, the 8 feature
 * channels first (vol_stride >= 8, a multiple of 4: the volume may be the leading
 * channels of a wider multi-head convolution output); tex/rgba from
 * gdb_prepare_sources; depth_range, vol_range (B,2,Hb,Wb); mlp: packed
 * parameter block.
 * out_feat (B, 3b^2+F+8, Hb, Wb) planar (the reference's layout, out_dec
 * unused); or, with out_channels_last != 0, out_feat (B,Hb,Wb,3b^2) = the fine
 * colours and out_dec (B,Hb,Wb,F+8) = the decoder's input, both channels-last.
 * out_depth, out_opacity (B,Hb,Wb).
 *
 * Optional taps (any may be null; offsets required if any is non-null): the
 * reference's intermediates in its packed sample order -
 * rgbs_feat_dir (V,S,3b^2+F+4), vox_feat (S,8), sigma (S), feat (S,3b^2+F+8),
 * weights (S).  S_total is the row count of those tensors.
 *
 * Work distribution of the tensor-core kernels (precision 1 / 2, 2x2 bundles): persistent CTAs take their tiles from an atomic
 * counter that lives in a __device__ array of the library (nothing is allocated): one counter per stream that ever launched
 * (launches on a stream are ordered), a fresh one for every launch recorded during a stream capture; the call enqueues a
 * 4-byte cudaMemsetAsync in front of the kernel (a memset node inside a graph).  With no counter to hand out (more than 64
 * streams, 960 captured launches, cudaStreamPerThread, GDB_K3_STATIC=1) the CTAs stride through the tiles: same results, bit
 * for bit, either way.                                                        */
typedef struct gdb_render_taps {
  const int32_t* offsets;
  int64_t S_total;
  float* rgbs_feat_dir;
  float* vox_feat;
  float* sigma;
  float* feat;
  float* weights;
} gdb_render_taps;

int gdb_render_fused_fwd(const float* rgba, const float* tex, const float* vol_cl, const float* depth_range,
                         const float* vol_range, const float* cam, int cam_stride, const float* mlp,
                         int B, int V, int H, int W, int bundle_size, int feat_dim, int D, int vol_stride,
                         int vol_layout /* 0: vol_cl is (B,D,Hb,Wb,vol_stride); 1: (B,Hb,Wb,D,vol_stride) - the depth-folded
                                           output of the cost-regularisation net run on 2-D convolutions */,
                         int max_samples, int max_mip_level, int inv_depth, int adaptive,
                         int precision /* 0 = fp32 SIMT MLP (1e-4 class); 1 = fp16-operand tcgen05 MLP, fp32 accumulate (2e-3 class);
                                          2 = split-fp16 (hi+lo) tcgen05 MLP, three MMAs per K step, fp32 accumulate (1e-4 class);
                                          3 = the first-generation kernel of class 1 (kept for A/B measurements);
                                          4 = the round-1 kernel of class 1 (serial gathers; kept for A/B, bit-identical to 1);
                                          5 = 1; 6 = the fourth-generation kernel of class 1 (weighted-sum taps, unconditional
                                          gathers, padded operand chunks: a measured variant, within fp32 rounding of 1 in front of
                                          the fp16 operand rounding) */,
                         int out_channels_last /* 0: planar out_feat; 1: channels-last out_feat + out_dec; 2: as 1, and the first pad
                                                  channel of out_dec is written as 1.0 instead of 0: a constant-one input channel whose
                                                  centre-tap weights carry the bias of the decoder's first convolution (needs dec_stride > feat_dim + 11) */,
                         int dec_stride /* floats per bundle of out_dec: 0 = feat_dim + 11 (dense); up to +3 pad channels, written as zeros,
                                           so that the decoder's first convolution sees a float4-aligned pixel */,
                         float* out_feat, float* out_dec, float* out_depth, float* out_opacity,
                         int row_lo, int row_hi /* bundle-map rows [row_lo, row_hi) rendered by this call (the image-tile split of one
                                                   target view over several GPUs: every output buffer has the full size, only the
                                                   rows of the range are written); 0, Hb = everything; partial ranges need
                                                   precision 1, 4, 5 or 6 */,
                         const gdb_render_taps* taps, void* stream);

/* K2 fused with the probability head it follows (cost_reg_net.py:62-63 -> depth_net.py:172,479-514): logits =
 * Conv3d(C -> 1, 3x3x3, padding 1, no bias) of y_cl (B,D,h,w,C) channels-last, C = 8, weight (27, C) in [kd][ky][kx][c]
 * order; then the soft-max over depth and the depth regression of gdb_depth_range_fwd.  prob_out (B,D,h,w) may be null. */
int gdb_prob_head_depth_range_fwd(const float* y_cl, const float* weight, const float* depth_range, int rh, int rw, int B,
                                  int C, int D, int h, int w, float ci_scale, int inv_depth, float* depth, float* ci,
                                  float* vol_range, float* prob_out, void* stream);

/* The same operator with the depth axis split over `nchunks` CTAs per pixel tile (more parallelism for small maps): every CTA
 * keeps the soft-max statistics of its chunk on line (running maximum, sum e, sum e x, sum e x^2 about the interval's
 * midpoint), the last CTA of a tile merges them.  No probability output.  scratch: gdb_prob_head_split_scratch_floats
 * floats (16-byte aligned); counters: gdb_prob_head_split_counters ints, ZERO before the first launch (the kernel leaves
 * them zero).  Every chunk must hold a plane: (nchunks - 1) * ceil(D / nchunks) < D.                                     */
int64_t gdb_prob_head_split_scratch_floats(int B, int h, int w, int nchunks);
int64_t gdb_prob_head_split_counters(int B, int h, int w);
int gdb_prob_head_depth_range_split_fwd(const float* y_cl, const float* weight, const float* depth_range, int rh, int rw, int B,
                                        int C, int D, int h, int w, int nchunks, float ci_scale, int inv_depth, float* scratch,
                                        int* counters, float* depth, float* ci, float* vol_range, void* stream);

/* Second generation of the split operator (same arguments, same results to fp32 summation order): the haloed depth planes
 * arrive by TMA (cp.async.bulk.tensor.5d; the tensor map's out-of-bounds zero fill is the convolution's padding in x, y and
 * depth), every plane is read from shared memory once and scattered into the three output planes it contributes to, and the
 * 216 weights are constant operands (copied device-to-device into constant memory by every call: a memcpy node under stream
 * capture; concurrent calls on different streams of one device must use the same weights).  Pixel tiles are 8 x 16:
 * gdb_prob_head_tma_scratch_floats / _counters size the two work buffers.  Replaces cost_reg_net.py:62-63 + depth_net.py:479-514. */
int64_t gdb_prob_head_tma_scratch_floats(int B, int h, int w, int nchunks);
int64_t gdb_prob_head_tma_counters(int B, int h, int w);
int gdb_prob_head_depth_range_tma_fwd(const float* y_cl, const float* weight, const float* depth_range, int rh, int rw, int B,
                                      int C, int D, int h, int w, int nchunks, float ci_scale, int inv_depth, float* scratch,
                                      int* counters, float* depth, float* ci, float* vol_range, void* stream);

/* Output assembly, replaces network.py:175-182 minus the decoder CNN:
 * rgb = dec + pixel_shuffle(feat[:, :3b^2], b)  (reweighting: 0.5*(rgb + fine))
 * and the bilinear xb up-sampling of depth and opacity.
 * feat (B,Ctot,Hb,Wb), dec (B,3,H,W) -> rgb (B,3,H,W), depth/opacity (B,H,W).
 * layout bit 0: feat is channels-last (B,Hb,Wb,Ctot); bit 1: dec is (B,H,W,3);
 * bit 2: dec is (B,H/2,W/2,12) channels-last = the decoder's last convolution
 * composed with its 1x1 output convolution, pixel shuffle (decoder_rdn.py:78-81)
 * still pending: channel c*4 + (y%2)*2 + (x%2).                               */
int gdb_assemble_output(const float* feat, int Ctot, const float* dec, const float* bdepth, const float* bopacity,
                        int B, int Hb, int Wb, int bundle_size, int reweighting, int layout, const float* dec_bias /* bias of the composed last convolution (12), layout bit 2 only; may be null */,
                        float* rgb, float* depth, float* opacity, void* stream);

/* --------------------------------------------------------- backward ------- */
/* Training configuration (dtu_pretrain.yaml): adjoints of the kernels above,
 * what torch.autograd derives from the reference's Python (there is no
 * hand-written backward in the reference).  Buffers documented as
 * "accumulated" must be zeroed by the caller.                                 */

/* Backward of gdb_warp_variance_fwd (depth_net.py:459-474).  g_variance is
 * PLANAR (B,C,D,Ht,Wt).  d_feat_cl (B,V,Hs,Ws,C) accumulated; d_depth_range
 * (B,2,Ht,Wt) accumulated, may be null and is ignored for a 1x1 range.        */
int gdb_warp_variance_bwd(const float* feat_cl, const float* proj, const float* depth_range, int rh, int rw, int B,
                          int V, int C, int Hs, int Ws, int D, int Ht, int Wt, int inv_depth, const float* g_variance,
                          float* d_feat_cl, float* d_depth_range, void* stream);

/* Backward of gdb_depth_range_fwd (depth_net.py:479-514).  g_depth (B,1,h,w),
 * g_ci (B,2,h,w), g_vol_range (B,2,h,w): any may be null.  d_prob (B,D,h,w)
 * written; d_depth_range (B,2,h,w) written (null / ignored for a 1x1 range).  */
int gdb_depth_range_bwd(const float* depth_range, int rh, int rw, const float* prob, int B, int D, int h, int w,
                        float ci_scale, int inv_depth, const float* g_depth, const float* g_ci, const float* g_vol_range,
                        float* d_prob, float* d_depth_range, void* stream);

/* Backward of gdb_render_fused_fwd: recomputes the forward from the same
 * inputs (nothing is saved), see csrc/gdb_render_bwd.cu.  g_feat
 * (B,3b^2+F+8,Hb,Wb) planar; g_depth, g_opacity (B,Hb,Wb) or null.
 * d_mlp (packed block) / d_tex (mip chain) / d_vol (B,D,Hb,Wb,8) accumulated;
 * d_depth_range, d_vol_range (B,2,Hb,Wb) written.                             */
int gdb_render_fused_bwd(const float* rgba, const float* tex, const float* vol_cl, const float* depth_range,
                         const float* vol_range, const float* cam, int cam_stride, const float* mlp, int B, int V,
                         int H, int W, int bundle_size, int feat_dim, int D, int vol_stride, int max_samples,
                         int max_mip_level, int inv_depth, int adaptive, const float* g_feat, const float* g_depth,
                         const float* g_opacity, float* d_mlp, float* d_tex, float* d_vol, float* d_depth_range,
                         float* d_vol_range, void* stream);

/* Backward of gdb_prepare_sources with respect to the feature maps: pulls the
 * mip-level gradients down to level 0 (in place in d_tex) and writes the
 * feature channels as d_feat (B*V,Cf,Hb,Wb) planar.                           */
int gdb_prepare_sources_bwd(float* d_tex, int BV, int Cf, int Hb, int Wb, int max_mip_level, float* d_feat, void* stream);

/* --------------------------------------------------------- coarse render -- */
/* Training-only coarse render of every cascade stage but the last, replaces
 * DepthNet._render_rays depth_net.py:49-116 with build_rays :301-341,
 * get_img_feat_vectorized :344-396 and the coarse NeRF :248-298.
 * tex (B*V,Hs,Ws,FP): level 0 of gdb_prepare_sources for the stage's FPN level
 * (feature | bilinearly down-sampled rgb); vol_cl (B,D,Hi,Wi,8); ray_range,
 * vol_range (B,2,Hi,Wi); cam: gdb_camera_block built with the STAGE-scaled
 * intrinsics (bundle_size 1); mlp: packed like the fine MLP with the `color`
 * head in the weight.0 / weight.2 slots.  One ray per (Hi,Wi) pixel,
 * num_samples fixed samples.  -> rgb (B,3,Hi,Wi).                             */
int gdb_coarse_render_fwd(const float* tex, const float* vol_cl, const float* ray_range, const float* vol_range,
                          const float* cam, int cam_stride, const float* mlp, int B, int V, int Hi, int Wi, int Hs,
                          int Ws, int feat_dim, int D, int num_samples, int inv_depth, float* rgb, void* stream);
/* Its backward (recomputes the forward).  d_mlp, d_tex (B*V,Hs,Ws,FP), d_vol
 * accumulated; d_ray_range, d_vol_range (B,2,Hi,Wi) written.                  */
int gdb_coarse_render_bwd(const float* tex, const float* vol_cl, const float* ray_range, const float* vol_range,
                          const float* cam, int cam_stride, const float* mlp, int B, int V, int Hi, int Wi, int Hs,
                          int Ws, int feat_dim, int D, int num_samples, int inv_depth, const float* g_rgb, float* d_mlp,
                          float* d_tex, float* d_vol, float* d_ray_range, float* d_vol_range, void* stream);

/* --------------------------------------------------------- glue ----------- */
/* Element-wise epilogues between the kernels above and the cuDNN networks
 * (channels-last fp32, C % 4 == 0; x/out (N,S,C)):
 * out = skip + act(x + bias): bias (C) or null, relu != 0 applies ReLU, skip
 * null or (N,S,C); with skip_up2 != 0 skip is (N,Hs,Ws,C) and S == 4*Hs*Ws
 * (nearest-neighbour x2).  Replaces cost_reg_net.py:108-110 (y = s + relu(..))
 * and the FPN top-down step feature_net.py:52-58.                             */
int gdb_bias_act_add(const float* x, const float* bias, const float* skip, int64_t N, int64_t S, int C, int relu,
                     int skip_up2, int Hs, int Ws, float* out, void* stream);
/* out = x + y * gate[n,c] (+ extra): squeeze-excite residual of the decoder's dense
 * blocks (decoder_rdn.py:31-41).  gate (N,C).                                 */
int gdb_gate_add(const float* x, const float* y, const float* gate, const float* extra /* optional addend, same shape as x */,
                 int64_t N, int64_t S, int C, float* out, void* stream);
/* Squeeze-excite block with its residual in two launches (decoder_rdn.py:31-41, modules.py:5-20):
 * out = x + y * sigmoid(W2 relu(W1 mean_hw(y))) (+ extra).  x, y, extra, out channels-last (N,S,C); w1 (R,C), w2 (C,R)
 * row-major (the nn.Linear weights, no bias); partial (N,chunks,C) is scratch for the fixed-order channel sums of y.    */
int gdb_se_gate_add(const float* x, const float* y, const float* w1, const float* w2, int R, const float* extra, int64_t N,
                    int64_t S, int C, int chunks, float* partial, float* out, void* stream);
/* gdb_se_gate_add that also writes its result into the leading C channels of out2, a channels-last buffer with rows of
 * out2_channels floats (the next dense block's concatenation buffer: its h slice is then never copied); out2 may be null.
 * gdb_concat2_into fills the slices behind it: out[pix, out_offset : out_offset + Ca + Cb] = [a | b].                  */
int gdb_se_gate_add_cat(const float* x, const float* y, const float* w1, const float* w2, int R, const float* extra, int64_t N,
                        int64_t S, int C, int chunks, float* partial, float* out, float* out2, int out2_channels, void* stream);
int gdb_concat2_into(const float* a, int Ca, const float* b, int Cb, int64_t npix, float* out, int out_channels, int out_offset,
                     void* stream);
/* Channel concatenation of channels-last maps over npix pixels: out (npix, Ca+Cb+Cc) = [a | b | c] (c may be null with
 * Cc = 0): the inputs of the dense block's second and third convolutions (decoder_rdn.py:36-41).                          */
int gdb_concat3(const float* a, int Ca, const float* b, int Cb, const float* c, int Cc, int64_t npix, float* out, void* stream);
/* PixelShuffle(2) of a channels-last map fused with the producing convolution's bias (decoder_rdn.py:76-80):
 * in (N,H,W,4C), bias (4C) or null -> out (N,2H,2W,C), out[n,2y+dy,2x+dx,c] = in[n,y,x,4c+2dy+dx] + bias[4c+2dy+dx].          */
int gdb_pixel_shuffle2(const float* in, const float* bias, int64_t N, int H, int W, int C, float* out, void* stream);
/* Per-image channel means of a channels-last map x (N, S, C) -> out (N, C): the squeeze of the squeeze-excite gate
 * (modules.py, AdaptiveAvgPool2d(1)).  partial (N, chunks, C) is scratch; the summation order is fixed.                    */
int gdb_channel_mean(const float* x, int64_t N, int64_t S, int C, int chunks, float* partial, float* out, void* stream);

/* --------------------------------------------------------- optimiser ------ */
/* Tail of the training step on ONE flat fp32 buffer: replaces `torch.nn.utils.clip_grad_value_(parameters, 40)` +
 * `optimizer.step()` (train/trainers/trainer.py:63-65; torch.optim.Adam built by train/optimizer.py:13-29) and the division of
 * DistributedDataParallel's gradient average (trainer.py:16-22: the caller all-reduces `grad` with SUM, grad_scale = 1/world).
 * param, grad, exp_avg, exp_avg_sq: n floats each (n % 4 == 0, 16-byte aligned).  state[0] = completed steps (device-resident
 * so that the step can be captured into a CUDA graph); gdb_adam_advance adds one after the update.                          */
int gdb_adam_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, const float* state, int64_t n,
                       float lr, float beta1, float beta2, float eps, float weight_decay, float clip_value, float grad_scale,
                       void* stream);
int gdb_adam_advance(float* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDB_NERF_B200_H */
