/*
 * dis_b200.h -- C-ABI of libdis_b200.so: the B200 (sm_100a) implementation of
 * DepthInSpace's self-supervision hot path (LCN -> pattern / flow warp -> block-window
 * photometric loss -> edge-aware smoothness, forward and backward).
 *
 * This is the drop-in boundary.  The reference (idiap/DepthInSpace) binds this path
 * through the pybind module `ext_cuda` of the un-vendored Connecting-the-Dots torchext
 * (model/ext_functions.py:32-39) and through torch ops inside nn.Modules
 * (model/networks.py, model/multi_frame_networks.py).  Each entry point below names the
 * reference interface it replaces.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 NCHW data owned by the caller;
 *     nothing is retained across calls, re-entrant and thread-safe (the backward half runs on
 *     autograd's worker thread, model/ext_functions.py:130-140); the only process-wide state is a
 *     mutex-guarded record of which kernels have been opted in to > 48 KB shared memory;
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; work is enqueued, never
 *     synchronised;
 *   - return value: DIS_OK (0) or a negative dis_status; dis_status_string() names it.
 *     The Python shim raises (the reference raises Exception('invalid loss type'),
 *     model/ext_functions.py:153);
 *   - loss `type`: 0 mse, 1 sad, 2 census_mse, 3 census_sad (model/ext_functions.py:142-154);
 *   - block_size: odd, 1..15.
 */
#ifndef DIS_B200_H_
#define DIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DIS_API __attribute__((visibility("default")))
#else
#define DIS_API
#endif

typedef enum {
  DIS_OK = 0,
  DIS_ERR_INVALID_LOSS_TYPE = -1, /* model/ext_functions.py:153 */
  DIS_ERR_BAD_SHAPE = -2,
  DIS_ERR_UNSUPPORTED_BLOCK_SIZE = -3,
  DIS_ERR_NULL_POINTER = -4,
  DIS_ERR_CUDA_LAUNCH = -5,
  DIS_ERR_UNSUPPORTED_KSIZE = -6,
  DIS_ERR_WORKSPACE_TOO_SMALL = -7,
  DIS_ERR_UNSUPPORTED_COMBINATION = -8
} dis_status;

DIS_API int dis_abi_version(void);
DIS_API const char* dis_status_string(int status);
/* last CUDA error text seen by this thread inside the library (for DIS_ERR_CUDA_LAUNCH) */
DIS_API const char* dis_last_cuda_error(void);

/* ---- a1  LCN.tforward, model/networks.py:679-689 ------------------------------------
 * x [N,1,H,W] -> lcn, std [N,1,H,W].  radius 1..8.  Window sums are accumulated in fp64. */
DIS_API int dis_lcn_forward(const float* x, float* lcn, float* std_out, int N, int H, int W,
                            int radius, float eps, void* stream);

/* Worker pre-processing step, Worker.copy_data (model/worker.py:418-438), fused: x [bs,tl,1,H,W] as the
 * DataLoader delivers it -> im_cat [tl,bs,2,H,W] (channel 0 = LCN(x), channel 1 = x) and std [tl,bs,1,H,W];
 * replaces transpose + contiguous + LCN + torch.cat. */
DIS_API int dis_lcn_prepare_input(const float* x, float* im_cat, float* std_out, int bs, int tl, int H, int W,
                                  int radius, float eps, void* stream);

/* Backward of LCN for API completeness (the reference never needs it: LCN inputs are data,
 * model/worker.py:430-445).  x, lcn, std: forward input/outputs; g_lcn, g_std: upstream gradients (either may
 * be NULL); workspace: 2*N*H*W floats of scratch. */
DIS_API int dis_lcn_backward(const float* x, const float* lcn, const float* std_in, const float* g_lcn,
                             const float* g_std, float* grad_x, float* workspace, int N, int H, int W,
                             int radius, float eps, void* stream);

/* ---- a4  ext_cuda.photometric_loss_forward / _backward, model/ext_functions.py:124,137 --
 * es, ta [N,C,H,W] -> out [N,1,H,W];  grad_out [N,1,H,W] -> grad_es [N,C,H,W]
 * (gradient w.r.t. es only, model/ext_functions.py:140). */
DIS_API int dis_photometric_loss_forward(const float* es, const float* ta, float* out, int N, int C,
                                         int H, int W, int block_size, int type, float eps,
                                         void* stream);
DIS_API int dis_photometric_loss_backward(const float* es, const float* ta, const float* grad_out,
                                          float* grad_es, int N, int C, int H, int W,
                                          int block_size, int type, float eps, void* stream);

/* ---- a3  disparity-driven pattern warp, model/networks.py:356-367 --------------------
 * pattern [H,W] (batch-shared), disp [N,1,H,W] -> proj [N,1,H,W]; optional outputs:
 * dproj_ddisp [N,1,H,W] (autograd chain of :358-367), corner_x0 / corner_y0 int32 [N,1,H,W]
 * (floor()ed source indices, for bit-exact index checks). */
DIS_API int dis_pattern_warp_forward(const float* disp, const float* pattern, float* proj,
                                     float* dproj_ddisp, int32_t* corner_x0, int32_t* corner_y0,
                                     int N, int H, int W, void* stream);

/* ---- a3+a4 fused  RectifiedPatternSimilarityLoss.tforward, model/networks.py:354-377 --
 * One pass: warp -> k x k window loss -> sigma-weighted partial sums, and (when grad_num != NULL)
 * the un-normalised gradient  d(sum(mask*diff)) / d disp  [N,1,H,W].
 *   std         may be NULL (mask = ones, :368-370)
 *   proj, diff  optional [N,1,H,W] outputs (pattern_proj :367, per-pixel loss map :376)
 *   partials    float[2 * dis_pattern_loss_num_partials(N,H,W)]: per-CTA (num, den) pairs
 * Follow with dis_reduce_pairs() to obtain (num, den, num/den) on the device. */
DIS_API int dis_pattern_loss_num_partials(int N, int H, int W);
DIS_API int dis_pattern_loss_forward(const float* disp, const float* im, const float* std_in,
                                     const float* pattern, float* proj, float* diff, float* grad_num,
                                     float* partials, int N, int H, int W, int block_size, int type,
                                     float eps, void* stream);
/* Same with grad_num already multiplied by the DEVICE scalar *grad_scale (weight / sum(sigma)): final gradient. */
DIS_API int dis_pattern_loss_forward_scaled(const float* disp, const float* im, const float* std_in,
                                            const float* pattern, float* proj, float* diff, float* grad_num,
                                            const float* grad_scale, float* partials, int N, int H, int W,
                                            int block_size, int type, float eps, void* stream);

/* Deterministic fixed-order reduction of n (a,b) pairs: out = {sum a, sum b, sum a / sum b}. */
DIS_API int dis_reduce_pairs(const float* partials, int n, float* out3, void* stream);
/* `count` independent segments of n pairs each (segment i at partials + 2*n*i) -> out3[3*i .. 3*i+2]. */
DIS_API int dis_reduce_pairs_batched(const float* partials, int n, int count, float* out3, void* stream);

/* ---- a8  photometric loop of the loss assembly, model/single_frame_worker.py:108-115 -------------
 * S (2 or 4) disparity maps of the SAME frames against one (im, std): one launch evaluates the
 * scale-independent half of the soft census (target term, sigma, tile loads) once and the estimate side
 * with packed fp32x2 arithmetic, two scales per instruction.  census_mse / census_sad only
 * (DIS_ERR_UNSUPPORTED_COMBINATION otherwise: use dis_pattern_loss_forward per scale).
 *   disps, grad_nums   HOST arrays of S device pointers ([N,1,H,W] each); grad_nums may be NULL
 *   partials           float[2 * S * dis_pattern_loss_multi_num_partials(N,H,W)], scale-major;
 *                      reduce with dis_reduce_pairs_batched(partials, num_partials, S, out3). */
DIS_API int dis_pattern_loss_multi_num_partials(int N, int H, int W);
DIS_API int dis_pattern_loss_multi_forward(const float* const* disps, int S, const float* im,
                                           const float* std_in, const float* pattern,
                                           float* const* grad_nums, float* partials, int N, int H, int W,
                                           int block_size, int type, float eps, void* stream);
/* Same, with the stored gradients already final: grad_nums[s] = grad_scale[s] * d(num_s)/d disp_s, grad_scale a DEVICE
 * array of S floats (typically weight_s / sum(sigma), sum(sigma) from dis_l1_forward(std, NULL)): the loss assembly
 * (model/single_frame_worker.py:108-115) knows its weights, so no separate scaling pass over the gradients is needed. */
DIS_API int dis_pattern_loss_multi_forward_scaled(const float* const* disps, int S, const float* im,
                                                  const float* std_in, const float* pattern,
                                                  float* const* grad_nums, const float* grad_scale,
                                                  float* partials, int N, int H, int W, int block_size,
                                                  int type, float eps, void* stream);

/* mse / sad pattern loss of S = 1..4 disparity maps of the same frames WITHOUT the per-pixel loss map, as a point-wise
 * kernel behind one box filter of the weights (replaces the same reference path as dis_pattern_loss_forward,
 * model/networks.py:354-377 + model/ext_functions.py:156-168, for types DIS_LOSS_MSE / DIS_LOSS_SAD):
 *   sum_p w(p) out(p) = sum_q s(q) M(q),  M(q) = (1/k^2) sum_{(p,o): clamp(p+o)=q} w(p),  d num / d e(q) = s'(q) M(q).
 *   disps, projs, grad_nums  HOST arrays of S device pointers ([N,1,H,W] each); projs / grad_nums and any of their
 *                      entries may be NULL
 *   grad_scale         optional DEVICE array of S floats multiplied into grad_nums[s]
 *   workspace          float[2 * N*H*W]: scratch in the first half, M in the second half; reuse_wbox != 0 says the
 *                      second half already holds M of these frames (same std_in, block_size): the box passes are skipped
 *   partials           float[2 * S * dis_pattern_loss_point_num_partials(N,H,W)], scale-major (num_s, den) pairs;
 *                      reduce with dis_reduce_pairs_batched(partials, num_partials, S, out3). */
DIS_API int dis_pattern_loss_point_num_partials(int N, int H, int W);
DIS_API int dis_pattern_loss_point_forward(const float* const* disps, int S, const float* im, const float* std_in,
                                           const float* pattern, float* const* projs, float* const* grad_nums,
                                           const float* grad_scale, float* workspace, int reuse_wbox,
                                           float* partials, int N, int H, int W, int block_size, int type,
                                           void* stream);

/* out[i] = in[i] * (*numer) / (*denom) (denom may be NULL = 1).  Scalars live on the device
 * so no host synchronisation is needed between forward and backward. */
DIS_API int dis_scale_by_device_scalar(const float* in, float* out, size_t n, const float* numer,
                                       const float* denom, void* stream);
/* Auxiliary L1 terms of the loss assembly, mean|a - b| (single_frame_worker.py:152-155, multi_frame_worker.py:
 * 160-165) in one pass: partials = float[2 * dis_l1_num_partials(n)] of (sum|a-b|, count) pairs (reduce with
 * dis_reduce_pairs -> mean in out3[2]); sign_out (optional) = sign(a - b), the gradient of the sum w.r.t. a. */
DIS_API int dis_l1_num_partials(size_t n);
DIS_API int dis_l1_forward(const float* a, const float* b, float* sign_out, float* partials, size_t n, void* stream);
/* b may be NULL (treated as 0): sum |a|, e.g. the sigma normaliser of the photometric terms. */
/* SGM warm-up term of the single-frame worker (model/single_frame_worker.py:158-163):
 *   valid = (b > threshold);  partials = (sum |a - b + noise| * valid, sum valid);  sign_out = sign(a - b + noise) * valid
 * noise may be NULL; partials sized like dis_l1_forward's. */
DIS_API int dis_masked_l1_forward(const float* a, const float* b, const float* noise, float threshold,
                                  float* sign_out, float* partials, size_t n, void* stream);
/* out[i] = a[i] * b[i] */
DIS_API int dis_mul(const float* a, const float* b, float* out, size_t n, void* stream);

/* ---- a5  SobelFilter / DisparitySmoothLoss, model/networks.py:697-730, 419-431 ---------
 * dis_sobel_forward: x [N,1,H,W] -> out [N,2,H,W] (gx, gy), replicate padding, ksize 3 or 5.
 * dis_sobel_backward: adjoint, grad_out [N,2,H,W] -> grad_x [N,1,H,W].
 * dis_smooth_loss_forward: one pass producing per-CTA partial sums of
 *   | sobel(disp) * exp(-|255 sobel(im)|) |  (pairs (sum, element count = 2 per pixel)) and, when grad_sum != NULL,
 *   d(sum)/d disp [N,1,H,W]  (divide by 2*N*H*W for the mean, :431). */
DIS_API int dis_sobel_forward(const float* x, float* out, int N, int H, int W, int ksize, void* stream);
DIS_API int dis_sobel_backward(const float* grad_out, float* grad_x, int N, int H, int W, int ksize,
                               void* stream);
DIS_API int dis_smooth_loss_num_partials(int N, int H, int W);
DIS_API int dis_smooth_loss_forward(const float* disp, const float* im, float* grad_sum,
                                    float* partials, int N, int H, int W, void* stream);
/* grad_sum multiplied by grad_scale on the way out (weight / (2*N*H*W) makes it the final gradient of the term);
 * accumulate != 0: added to what grad_sum already holds (the photometric gradient of scale 0) instead of overwriting. */
DIS_API int dis_smooth_loss_forward_scaled(const float* disp, const float* im, float* grad_sum,
                                           float* partials, int N, int H, int W, float grad_scale,
                                           int accumulate, void* stream);

/* ---- a6  warp(x, flow), model/multi_frame_networks.py:83-99 ---------------------------
 * x [N,C,H,W], flow [N,2,H,W] -> out [N,C,H,W]; bilinear, zeros padding, align_corners.
 * backward: grad_x (zero-filled here, then accumulated) and/or grad_flow may be NULL.
 * fb_mask_out (optional, forward only, requires C==2): the forward-backward consistency
 * mask of model/multi_frame_networks.py:205-207 computed from (flow, out). */
DIS_API int dis_flow_warp_forward(const float* x, const float* flow, float* out, float* fb_mask_out,
                                  int32_t* corner_x0, int32_t* corner_y0, int N, int C, int H, int W,
                                  void* stream);
DIS_API int dis_flow_warp_backward(const float* x, const float* flow, const float* grad_out,
                                   float* grad_x, float* grad_flow, int N, int C, int H, int W,
                                   void* stream);

/* ---- a7  gather step of FuseNet: gather_warped_xyz / Block2D3D.gather_warped_feat,
 *          model/multi_frame_networks.py:187-214, 347-360 -------------------------------------------
 * x, out, grad_out, grad_x [tl,bs,C,H,W] (tl <= 8).  One launch builds the stacked tensor the reference
 * assembles with tl-1 warp() calls + torch.stack:
 *   out[0] = x[tidx];  out[k] = warp(x[j_k], flows[k-1]),  j_k = the k-th frame index != tidx in
 *   increasing order, flows[k-1] = flow_{tidx, j_k} [bs,2,H,W] (HOST array of tl-1 device pointers).
 * backward: grad_x[tidx] = grad_out[0]; grad_x[j_k] = warp^T(grad_out[k]) (zero-filled here, RED.ADD). */
DIS_API int dis_flow_warp_gather_forward(const float* x, const float* const* flows, float* out, int tl,
                                         int tidx, int bs, int C, int H, int W, void* stream);
DIS_API int dis_flow_warp_gather_backward(const float* const* flows, const float* grad_out,
                                          float* grad_x, int tl, int tidx, int bs, int C, int H, int W,
                                          void* stream);

/* The same for ALL target frames at once (the tidx loop of Block2D3D.fwd_3d_1 / fwd_3d_2, :376-404, whose results the
 * reference stacks into [tl,tl,bs,C,H,W]): out[tidx*tl + k] as above for every tidx; flows = HOST array of tl*tl
 * device pointers, flows[i*tl + j] = flow_{ij} [bs,2,H,W] (diagonal ignored).  backward: grad_out [tl*tl,bs,C,H,W] ->
 * grad_x [tl,bs,C,H,W] = every frame's own-slot gradient plus the tl-1 scattered ones, no zero-fill pass. */
DIS_API int dis_flow_warp_gather_all_forward(const float* x, const float* const* flows, float* out, int tl,
                                             int bs, int C, int H, int W, void* stream);
DIS_API int dis_flow_warp_gather_all_backward(const float* const* flows, const float* grad_out,
                                              float* grad_x, int tl, int bs, int C, int H, int W, void* stream);

/* ---- a9  flow-consistency (geometric) loss, ONE direction (frame 0 -> frame 1) -------------------------
 * Single_Frame_Flow_Consistency_Loss.fwd / Multi_Frame_Flow_Consistency_Loss.fwd, model/networks.py:619-655,
 * 564-601 (+ ProjectionBaseLoss :455-488).  depth*, amb* [bs,C,H,W] (depth C=1), flow* [bs,2,H,W],
 * R* [bs,3,3], t* [bs,3], K [3,3], ray [H*W,3] (= uv1 @ Ki^T, :446-449), all device fp32.
 *   primary_depth1  NULL for the single-frame variant; else adds the < 1 px reprojection mask (:591-595)
 *   clamp           > 0: diff clamped to [0, clamp] (:637-638);  fb_scale 0.02 (:645)
 *   loss_mask, orig_mask  optional [bs,1,H,W] float 0/1 outputs (:646-651, :640)
 *   grad_depth0     optional: d sum(diff*mask) / d depth0 (per pixel)
 *   grad_depth1     optional: d sum(diff*mask) / d depth1 (zero-filled here, then RED.ADD scatter)
 *   partials        float[2 * dis_flow_consistency_num_partials(bs,H,W)] (sum(diff*mask), sum(mask)) pairs
 * loss = sum(diff*mask) / (sum(mask) + 1e-8): reduce with dis_reduce_pairs, scale gradients with dis_combine2. */
DIS_API int dis_flow_consistency_num_partials(int bs, int H, int W);
DIS_API int dis_flow_consistency_forward(const float* depth0, const float* depth1, const float* R0,
                                         const float* t0, const float* R1, const float* t1,
                                         const float* flow0, const float* flow1, const float* amb0,
                                         const float* amb1, int amb_channels, const float* primary_depth1,
                                         const float* K, const float* ray, float clamp, float fb_scale,
                                         float* loss_mask, float* orig_mask, float* grad_depth0,
                                         float* grad_depth1, float* partials, int bs, int H, int W,
                                         void* stream);
/* out = (*numer) * (a / (*den_a + eps) + b / (*den_b + eps)); b, den_b may be NULL. */
DIS_API int dis_combine2(const float* a, const float* b, float* out, size_t n, const float* numer,
                         const float* den_a, const float* den_b, float eps, void* stream);

/* Gradient of ALL geometric terms of the loss assembly (the pair loops of single_frame_worker.py:127-149 /
 * multi_frame_worker.py:128-157) w.r.t. the disparity maps, in one pass, DispToDepth (model/networks.py:311-319)
 * included:   grad_disp[f] = d depth/d disp(f) * sum_{k: frame_of[k] == f} scale[k] * planes[k]
 *   planes    HOST array of n_terms device pointers: the grad_depth0 / grad_depth1 outputs of
 *             dis_flow_consistency_forward ([bs,1,H,W] each); frame_of: HOST int[n_terms], the frame each belongs to
 *   scale     DEVICE float[n_terms]: upstream * weight / (den + 1e-8) of the term the plane came from
 *   disp, grad_disp  [tl,bs,1,H,W] (tl <= 8, at most 16 planes per frame); depth = baseline_focal / (max(disp,0) + 1e-12) */
DIS_API int dis_geometric_grad_combine(const float* const* planes, const int* frame_of, int n_terms,
                                       const float* scale, const float* disp, float baseline_focal,
                                       float* grad_disp, int tl, int bs, int H, int W, void* stream);

/* ---- (next) the ext ops the reference wraps but never calls, model/ext_functions.py:41-110 ----------------------
 * Replace ext_cuda.nn_cuda (:46), crosscheck_cuda (:64), proj_nn_cuda (:81), xcorrvol_cuda (:100).  Their definitions
 * live in the un-vendored Connecting-the-Dots torchext: PARITY UNPINNED (semantics in csrc/ext_misc.cu).
 *   nn:         in0 [n0,dim], in1 [n1,dim] (dim <= 8) -> out int64 [n0], index of the nearest row of in1 (-1 if n1 == 0)
 *   crosscheck: in0 int64 [n0] (indices into in1), in1 int64 [n1] -> out uint8 [n0] = (in1[in0[i]] == i)
 *   proj_nn:    xyz0, xyz1 [bs,H,W,3], K [3,3] row-major -> out int64 [bs,H,W], flat index into xyz1's points or -1
 *   xcorrvol:   in0, in1 [C,H,W] -> out [n_disps,H,W] */
DIS_API int dis_ext_nn(const float* in0, const float* in1, int64_t* out, int64_t n0, int64_t n1, int dim, void* stream);
DIS_API int dis_ext_crosscheck(const int64_t* in0, const int64_t* in1, uint8_t* out, int64_t n0, int64_t n1, void* stream);
DIS_API int dis_ext_proj_nn(const float* xyz0, const float* xyz1, const float* K, int64_t* out, int bs, int H, int W,
                            int patch_size, void* stream);
DIS_API int dis_ext_xcorrvol(const float* in0, const float* in1, float* out, int C, int H, int W, int n_disps,
                             int block_size, void* stream);

/* ---- (next) resize_like / resize_flow_like / resize_flow_masks_like, model/multi_frame_networks.py:42-81 ----------
 * Bilinear resize with align_corners=True of `count` tensors [N,C,H,W] -> [N,C,oh,ow] in ONE launch (ins / outs: HOST
 * arrays of `count` device pointers: the entries of the reference's flow / mask dicts), fused with what follows it:
 *   mode 0  plain                                             (resize_like, :42-52)
 *   mode 1  flow: channel 0 *= ow/W, channel 1 *= oh/H; C == 2   (resize_flow_like, :54-68)
 *   mode 2  mask: out = (value > 0.5) ? 1 : 0                 (resize_flow_masks_like, :70-81; (resize_like(m) > 0.5) :394)
 * Same arithmetic as ATen's upsample_bilinear2d CUDA kernel.  backward: adjoint of mode 0 w.r.t. the input. */
DIS_API int dis_resize_bilinear_forward(const float* const* ins, float* const* outs, int count, int N, int C, int H,
                                        int W, int oh, int ow, int mode, void* stream);
DIS_API int dis_resize_bilinear_backward(const float* grad_out, float* grad_in, int N, int C, int H, int W, int oh,
                                         int ow, void* stream);

/* ---- (next) neighbour selection + gather of FuseNet's Conv3D, model/multi_frame_networks.py:469-501 -------
 * xyz [tl,bs,3,h,w], feat [tl,bs,C,h,w], mask [tl,bs,1,h,w] -> for each of the M = bs*oh*ow output pixels
 * (ksize x ksize window, zero padding (ksize-1)/2, given stride) the `neighbors` candidates (of ksize^2*tl <= 64)
 * closest to the centre ray in the normalised image plane:
 *   xyz_nb [M,neighbors,3] = xyz_local, feat_nb [M,neighbors,C], idx [M,neighbors] uint8 candidate index
 *   ((ky*ksize + kx)*tl + t, ascending key, ties -> lowest index);
 *   scratch: dis_conv3d_scratch_elems(tl, bs, h, w) floats on the device (global maximum + the normalised image-plane
 *   coordinates xyz / (z + 1e-12), computed once per source element).   [ABI 2: was 1 float in ABI 1]
 * backward: deterministic gather; g_xyz / g_feat may be NULL. */
DIS_API int dis_conv3d_out_size(int n, int ksize, int stride);
DIS_API size_t dis_conv3d_scratch_elems(int tl, int bs, int h, int w);
DIS_API int dis_conv3d_gather_forward(const float* xyz, const float* feat, const float* mask, float* xyz_nb,
                                      float* feat_nb, uint8_t* idx, float* scratch, int tl, int bs, int C, int h,
                                      int w, int ksize, int stride, int neighbors, void* stream);
/* The two halves of dis_conv3d_gather_forward.  The selection (xyz_nb, idx) depends on xyz and mask only: the Conv3D layers
 * of one FuseNet level (multi_frame_networks.py:520-560 call the layer twice per level on the same xyz / mask) and the
 * checkpoint recompute can rank once and gather features per call. */
DIS_API int dis_conv3d_rank(const float* xyz, const float* mask, float* xyz_nb, uint8_t* idx, float* scratch, int tl,
                            int bs, int h, int w, int ksize, int stride, int neighbors, void* stream);
DIS_API int dis_conv3d_gather_features(const float* feat, const uint8_t* idx, float* feat_nb, int tl, int bs, int C,
                                       int h, int w, int ksize, int stride, int neighbors, void* stream);
DIS_API int dis_conv3d_gather_backward(const float* g_xyz_nb, const float* g_feat_nb, const uint8_t* idx,
                                       float* g_xyz, float* g_feat, int tl, int bs, int C, int h, int w,
                                       int ksize, int stride, int neighbors, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIS_B200_H_ */
