/* rangedet_b200.h -- C-ABI of librangedet_b200.so (sm_100a CUDA kernels for the RangeDet hot path).
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch / MXNet types; every tensor pointer is a DEVICE pointer
 *     owned by the caller (the reference's engine owns all buffers too:
 *     operator_cxx/contrib/decode_3d_bbox-inl.h:279-283), dense row-major fp32 unless stated
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); kernels are
 *     enqueued on it and the call returns without synchronising unless stated
 *   - return 0 on success, non-zero on error; rd_last_error() gives a thread-local message
 *     (the reference's CHECK_* / LOG(FATAL) -> dmlc::Error, decode_3d_bbox.cc:30-60)
 *   - no allocation inside: scratch space is a caller-provided workspace whose size the
 *     matching *_workspace_bytes() query returns
 *   - there is NO CPU fallback: without an sm_100 device every compute entry point fails
 */
#ifndef RANGEDET_B200_H_
#define RANGEDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rd_stream_t;

/* ---- library ---------------------------------------------------------------------------- */
int rd_version(void);                 /* ABI version, currently 1 */
const char* rd_last_error(void);      /* thread-local, never NULL */
/* 0 if the current CUDA device is compute capability 10.x, else non-zero (+ rd_last_error). */
int rd_check_device(void);
/* Number of kernel launches issued by this library on the calling thread since load
 * (bench.py's `gpu_launches`). */
uint64_t rd_launch_count(void);

/* ---- Meta-Kernel -------------------------------------------------------------------------
 * Replaces MetaKernel.meta_baseline_bias, rangedet/symbol/backbone/meta_kernel.py:166-240
 * (two mx.sym.im2col + two 1x1 Convolution + broadcast_minus + elemwise mul, all MXNet):
 *   out[b, c*9+k, h, w] = data[b,c,h+dy,w+dx] * (W1 . relu(W0 . rel + b0) + b1)[c]
 *   rel = coord[b,:,h+dy,w+dx] (0 outside the image) - coord[b,:,h,w],  k = ky*3+kx
 * data (B,C,H,W)  coord (B,3,H,W)  w0 (32,3)  b0 (32)  w1 (C,32)  b1 (C)  out (B,9C,H,W)
 * C must be a multiple of 8, <= 64.  impl: 0 = default, 1 = CUDA-core fp32, 2 = tcgen05 (C == 64),
 * 3 = TMA-fed warp-specialised tcgen05 (C == 64, W % 4 == 0).
 */
int rd_meta_kernel_fwd(const float* data, const float* coord, const float* w0, const float* b0,
                       const float* w1, const float* b1, float* out,
                       int B, int C, int H, int W, int impl, rd_stream_t stream);

/* Backward of the same op w.r.t. data and the four MLP parameters (coord is a graph input with
 * grad_req null, rangedet/symbol/head/builder.py:20-37).  grad_* are overwritten (kWriteTo). */
size_t rd_meta_kernel_bwd_workspace_bytes(int B, int C, int H, int W);
int rd_meta_kernel_bwd(const float* grad_out, const float* data, const float* coord,
                       const float* w0, const float* b0, const float* w1, const float* b1,
                       float* grad_data, float* grad_w0, float* grad_b0, float* grad_w1,
                       float* grad_b1, void* workspace, size_t workspace_bytes,
                       int B, int C, int H, int W, int impl, rd_stream_t stream);

/* Forward fused with the BatchNorm(9C)+ReLU that follows it in meta_kernel_conv
 * (rangedet/symbol/backbone/dla_backbone.py:92-94): y = relu?(out * scale + shift), written as zero-haloed
 * NHWC bf16 y_pad [B][H+2][W+2][9C] with TAP-MAJOR channels: y[..., k*C + c] = f(out[b, c*9+k, h, w]);
 * scale/shift are indexed the same way (k*C + c).  C == 64, W % 4 == 0.  The halo is not written.
 */
int rd_meta_kernel_fwd_nhwc_bf16(const float* data, const float* coord, const float* w0, const float* b0,
                                 const float* w1, const float* b1, const float* scale, const float* shift,
                                 int relu, void* y_pad, int B, int C, int H, int W, rd_stream_t stream);

/* Backward of the Meta-Kernel inside the training graph: grad_out arrives as the zero-haloed NHWC tensor
 * [B][H+2][W+2][9C] with TAP-MAJOR channels (k*C + c) -- the layout rd_meta_kernel_fwd_nhwc_* writes and the BatchNorm
 * backward (rd_bn_act_bwd_*) returns -- instead of the (B, 9C, H, W) fp32 tensor of the reference op boundary
 * (rangedet/symbol/backbone/meta_kernel.py:232-239): half the gradient bytes and no layout pass in between.  Same
 * arithmetic as rd_meta_kernel_bwd on the widened values (bit-identical results).  C == 64, W % 4 == 0.  grad_data may be
 * NULL (input needs no gradient); the four parameter gradients may be NULL together. */
int rd_meta_kernel_bwd_nhwc_bf16(const void* grad_out_pad, const float* data, const float* coord, const float* w0,
                                 const float* b0, const float* w1, const float* b1, float* grad_data, float* grad_w0,
                                 float* grad_b0, float* grad_w1, float* grad_b1, void* workspace, size_t workspace_bytes,
                                 int B, int C, int H, int W, rd_stream_t stream);
int rd_meta_kernel_bwd_nhwc_f16(const void* grad_out_pad, const float* data, const float* coord, const float* w0,
                                const float* b0, const float* w1, const float* b1, float* grad_data, float* grad_w0,
                                float* grad_b0, float* grad_w1, float* grad_b1, void* workspace, size_t workspace_bytes,
                                int B, int C, int H, int W, rd_stream_t stream);

/* The two halves of rd_meta_kernel_bwd, separately callable (grad_req 'null' on either side). */
int rd_meta_kernel_bwd_data(const float* grad_out, const float* coord, const float* w0,
                            const float* b0, const float* w1, const float* b1, float* grad_data,
                            int B, int C, int H, int W, int impl, rd_stream_t stream);
int rd_meta_kernel_bwd_params(const float* grad_out, const float* data, const float* coord,
                              const float* w0, const float* b0, const float* w1, const float* b1,
                              float* grad_w0, float* grad_b0, float* grad_w1, float* grad_b1,
                              void* workspace, size_t workspace_bytes,
                              int B, int C, int H, int W, int impl, rd_stream_t stream);

/* ---- Decode3DBbox ------------------------------------------------------------------------
 * Replaces _contrib_Decode3DBbox: Decode3DBboxForward, operator_cxx/contrib/
 * decode_3d_bbox-inl.h:279-305 (functors :169-277 and, is_bin, :64-167).
 * delta (n_total, is_bin?7:8)  pc (n_total,3)  out (n_total,10); n_total = B*N.
 */
int rd_decode_3d_bbox(const float* delta, const float* pc, float* out, int64_t n_total,
                      int is_bin, rd_stream_t stream);

/* ---- RotatedIOU --------------------------------------------------------------------------
 * Replaces _contrib_RotatedIOU: RotatedIOUForward, operator_cxx/contrib/rotated_iou-inl.h:525-547.
 * boxes1 (n1,T) boxes2 (n2,T) -> ious (n1,n2);  T = box_type in {5,7,8}.
 */
int rd_rotated_iou(const float* boxes1, const float* boxes2, float* ious, int64_t n1, int64_t n2,
                   int box_type, rd_stream_t stream);

/* Replaces the Python CustomOp 'batch_rotated_iou' (operator_py/batch_rotated_iou.py:11-49):
 * per image RotatedIOU(proposal[:, :8], gt) -> NaN/inf/>1/<0 -> 0 -> max over GT, fused.
 * proposal (B,N,10)  gt (B,G,8) [iou_type 0 = 'bev'] or (B,G,7) [1 = '3d']  out (B,N).
 */
int rd_batch_rotated_iou_max(const float* proposal, const float* gt, float* out, int B, int64_t N,
                             int G, int iou_type, rd_stream_t stream);

/* ---- RPN loss head (training) ----------------------------------------------------------------
 * Replaces, per pyramid level, RangeRpnHead.get_iou_target + get_vfl_loss + get_normalize_reg_loss
 * (rangedet/symbol/head/builder.py:155-197, :350-422; rangedet/symbol/head/loss.py:4-30): ~45 MXNet ops
 * incl. Decode3DBbox and the Python CustomOp 'batch_rotated_iou' -> one reduction + one fused kernel.
 *   cls_logit (B,1,H,Wl) = (B,N)   reg_delta (B,8,H,Wl) = (B,8,N) planar, as the head produces them
 *   pc (B,N,3)   gt (B,G,8) corners [or (B,G,7) for iou_type 1]   mask (B,N)
 *   reg_target / reg_weight / reg_norm_weight (B,8,N)
 * Outputs (any may be NULL to skip):
 *   iou_target (B,N)  = max_g sanitise(IoU(decode(reg_delta, pc), gt_g))      (stop_gradient in the graph)
 *   cls_loss (B,N)    = VFL(cls_logit, iou_target; alpha, gamma) * mask / (sum(mask)+1)
 *   reg_loss (B,8,N)  = smooth_l1(reg_delta - reg_target; scalar) * weight * norm_weight /
 *                       (sum(norm_weight)+1) * reg_loss_weight
 *   d_cls (B,N), d_reg (B,8,N) = gradients of cls_grad_scale*sum(cls_loss) + reg_grad_scale*sum(reg_loss)
 *                       w.r.t. cls_logit / reg_delta, i.e. what MakeLoss(grad_scale=...) back-propagates
 *                       (cls: scale_loss_shift*cls_loss_weight, reg: scale_loss_shift; builder.py:374-378,417-421).
 * iou_type: 0 = 'bev', 1 = '3d'.  workspace: rd_rpn_loss_workspace_bytes(), 8-byte aligned. */
size_t rd_rpn_loss_workspace_bytes(void);
int rd_rpn_loss(const float* cls_logit, const float* reg_delta, const float* pc, const float* gt,
                const float* mask, const float* reg_target, const float* reg_weight,
                const float* reg_norm_weight, int B, int64_t N, int G, int iou_type, float alpha, float gamma,
                float smooth_l1_scalar, float cls_grad_scale, float reg_loss_weight, float reg_grad_scale,
                float* iou_target, float* cls_loss, float* reg_loss, float* d_cls, float* d_reg,
                void* workspace, size_t workspace_bytes, rd_stream_t stream);
/* The same, with the head outputs and their gradients in the layout the head convolutions read and write (training graph):
 * cls_pad / reg_pad = zero-haloed NHWC [B][H+2][W+2][cpad] of bf16 / fp16, logit = channel 0 of cls_pad, deltas = channels
 * 0..7 of reg_pad (the 1x1 head convs of builder.py:247-262 with their channels padded to cpad); the gradients are
 * written, rounded to the storage type, to the same channels of dcls_pad / dreg_pad and nothing else is touched (their
 * other channels and halos must be zero).  Replaces rd_nhwc_*_to_nchw_f32 -> rd_rpn_loss -> rd_nchw_f32_to_nhwc_* per
 * head output: same values (the fp32 planar tensors only ever held widened 16-bit numbers). */
int rd_rpn_loss_nhwc_bf16(const void* cls_pad, const void* reg_pad, int H, int W, int cpad, const float* pc,
                          const float* gt, const float* mask, const float* reg_target, const float* reg_weight,
                          const float* reg_norm_weight, int B, int G, int iou_type, float alpha, float gamma,
                          float smooth_l1_scalar, float cls_grad_scale, float reg_loss_weight, float reg_grad_scale,
                          float* iou_target, float* cls_loss, float* reg_loss, void* dcls_pad, void* dreg_pad,
                          void* workspace, size_t workspace_bytes, rd_stream_t stream);
int rd_rpn_loss_nhwc_f16(const void* cls_pad, const void* reg_pad, int H, int W, int cpad, const float* pc,
                         const float* gt, const float* mask, const float* reg_target, const float* reg_weight,
                         const float* reg_norm_weight, int B, int G, int iou_type, float alpha, float gamma,
                         float smooth_l1_scalar, float cls_grad_scale, float reg_loss_weight, float reg_grad_scale,
                         float* iou_target, float* cls_loss, float* reg_loss, void* dcls_pad, void* dreg_pad,
                         void* workspace, size_t workspace_bytes, rd_stream_t stream);

/* ---- Training-target assignment (data-loader side of the training graph) -------------------------
 * Replace processing_cxx.assign3D_v2 / get_point_num (operator_cxx/src_cxx/assigner.h:11-87, :89-109; called from
 * rangedet/core/input.py:311-320, :433) and GenerateTarget.get_rpn_reg_target (+ normalisation / dimension
 * weights, input.py:345-372, 430-506, num_classes == 1).  One frame per call, like the reference.
 *   pc (N,3) vehicle-frame points   bbox (M,24) = 8 corners xyz   bbox_center (M,3)   bbox_radius (M)
 *   mask (N), is_in_nlz (N): point skipped if mask < 0.5 or nlz > 0; the six extent floats and max_dist as in
 *   the reference (distances are SQUARED distances there, and so here).
 *   result (N) int32: index of the first containing box, -1 if none.  */
int rd_assign3d_v2(const float* pc, const float* bbox, const float* bbox_center, const float* bbox_radius,
                   const float* mask, const float* is_in_nlz, float max_x, float min_x, float max_y, float min_y,
                   float max_z, float min_z, float max_dist, int64_t n_points, int n_boxes, int* result,
                   rd_stream_t stream);
/* bbox_inds_each_pt (N) FLOAT indices as the reference passes them (input.py:433-435) -> out (N) float: number
 * of points sharing the point's box, -1 where the index is negative (MAX_BOX_NUM = 500, assigner.h:94; indices
 * >= 500 are out of bounds in the reference and yield -1 here).  The workspace holds the 500 int32 counts
 * afterwards (input of rd_rpn_reg_target). */
size_t rd_get_point_num_workspace_bytes(void);
int rd_get_point_num(const float* bbox_inds_each_pt, int64_t n_points, float* out, void* workspace,
                     size_t workspace_bytes, rd_stream_t stream);
/* gt_box7 (M,7) [x,y,z,l,w,h,yaw] ("gt_bbox_csa"), bbox_ind (N) from rd_assign3d_v2, point_hist = the counts
 * left in rd_get_point_num's workspace, reg_dim_weight (8) -> reg_target (N,8) =
 * [dx,dy (signed sqrt, azimuth frame), log w, log l, cos, sin (yaw - azimuth), bottom z, log h],
 * reg_normalize_weight (N,8) = 1 / points-in-box, reg_weight (N,8) = reg_dim_weight inside boxes; 0 elsewhere. */
int rd_rpn_reg_target(const float* pc, const float* gt_box7, const int* bbox_ind, const int* point_hist,
                      const float* reg_dim_weight, int64_t n_points, int n_boxes, float* reg_target,
                      float* reg_normalize_weight, float* reg_weight, rd_stream_t stream);

/* ---- weighted NMS ------------------------------------------------------------------------
 * Replaces processing_cxx.wnms_4c: point4_wnms_4c / trtplus::wnms_4c,
 * operator_cxx/src_cxx/nms.h:781-794, :452-577.
 * dets (n,12) [8 BEV corner coords, yaw, z0, h, score] -> out_dets (K,12), keep_inds (K) indices
 * into the INPUT array, in descending-score order.  *out_count (HOST int) = K.  Synchronises the
 * stream before returning (K is data dependent).  out_dets / keep_inds must hold n rows.
 */
size_t rd_wnms_4c_workspace_bytes(int n);
int rd_wnms_4c(const float* dets, int n, float thresh, float thresh_vote, int is_3d,
               int hash_scale, float* out_dets, int32_t* keep_inds, int* out_count,
               void* workspace, size_t workspace_bytes, rd_stream_t stream);

/* ---- NMS3D (hard NMS) ---------------------------------------------------------------------
 * Replaces _contrib_NMS3D: NMS3DForward<gpu>, operator_cxx/contrib/nms_3d.cu:470-534.
 * boxes (B,N,10) [4 BEV corners, z0, z1] sorted by score -> keep_idx (B,max_keep) int32 indices in
 * keep order, filled with -1; boxes_out (B,max_keep,10) the kept boxes, filled with 0.
 * A box j > i is removed when iou(i, j) > iou_thres, iou = volumetric rotated IoU (nms_3d.cu:342-368)
 * or, normal_iou != 0, axis-aligned IoU of boxes[:, :4] (:370-378).
 */
size_t rd_nms3d_workspace_bytes(int B, int N);
int rd_nms3d(const float* boxes, int B, int N, float iou_thres, int max_keep, int normal_iou,
             int32_t* keep_idx, float* boxes_out, void* workspace, size_t workspace_bytes,
             rd_stream_t stream);

/* ---- get_sorted_foreground (test-time proposal selection) --------------------------------------
 * Replaces the Python CustomOp GetSortedFGOperator.forward, operator_py/get_sorted_foreground.py:11-40
 * (registered as 'get_sorted_foreground' :47-84; call site rangedet/symbol/head/builder.py:512-521).
 * cls_score (B,N), bbox_delta (B,N,8), pc (B,N,3), mask (B,N) ->
 * out_score (B,K) = the K = num_fgs largest cls_score*mask per row in descending order,
 * out_delta (B,K,8), out_pc (B,K,3) = the rows of those points.  Equal scores keep ascending point
 * index (stable; MXNet's tie order is unspecified).  num_fgs > N is an error (:66).  No gradient (:42-44).
 */
size_t rd_get_sorted_foreground_workspace_bytes(int B, int N);
int rd_get_sorted_foreground(const float* cls_score, const float* bbox_delta, const float* pc,
                             const float* mask, int B, int N, int num_fgs, float* out_score,
                             float* out_delta, float* out_pc, void* workspace, size_t workspace_bytes,
                             rd_stream_t stream);

/* ---- Convolution family (DLA backbone / RPN head) -------------------------------------------
 * Replaces mx.sym.Convolution / mx.sym.Deconvolution (+ inference-form BatchNorm, ReLU, residual
 * add) as emitted by mxnext/simple.py:123-158,545-580 for rangedet/symbol/backbone/
 * dla_backbone.py:17-56,95,116-127 and rangedet/symbol/head/builder.py:198-266.
 *   x_pad        bf16 [N][H+2][W+2][Cin]    NHWC with a one-pixel ZERO halo
 *   w_packed     bf16 [taps][Cout][Cin]     conv: tap = ky*ksize + kx (cross-correlation);
 *                                            deconv: tap = ky*kw + kx of the (Cin,Cout,3,kw) weight
 *   scale, shift fp32 [Cout] or NULL (-> 1, 0)
 *   residual_pad bf16 haloed NHWC at OUTPUT resolution, or NULL
 *   y_pad        bf16 haloed NHWC; interior written, halo left untouched (keep it zero)
 * rd_conv2d:   y = relu?( conv(x) * scale + shift + residual ); ksize 3 (pad 1) or 1; H-stride 1,
 *              W-stride stride_w in {1,2} (W even); output width W / stride_w.
 * rd_deconv2d: y = relu?( deconv(x) * scale + shift ) + residual; kernel (3,kw), stride (1,kw/2),
 *              pad (1,kw/4) for kw in {8,4} (the two shapes of agg_stage); output width W * kw/2.
 *              kw = 3: kernel (3,3), stride (1,2), pad (1,1), output width 2*W -- the data gradient of
 *              the 3x3 W-stride-2 convolutions (weights [ky*3+kx][Cin_of_conv][Cout_of_conv], no flip).
 * Cin: multiple of 64 (<= 1024); Cout: 64 or 128.  Pad narrower layers with zero channels.
 */
int rd_conv2d_nhwc_bf16(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                        const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout,
                        int ksize, int stride_w, int relu, rd_stream_t stream);
int rd_deconv2d_nhwc_bf16(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                          const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout,
                          int kw, int relu, rd_stream_t stream);
/* rd_conv2d without residual, writing channels [y_coff, y_coff+Cout) of a haloed NHWC tensor that has
 * y_ctotal channels (wide outputs, e.g. the 576-channel data gradient of the Meta-Kernel unit's 1x1
 * aggregation conv, are produced as several 64/128-channel slices). */
int rd_conv2d_nhwc_bf16_slice(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                              void* y_pad, int N, int H, int W, int Cin, int Cout, int ksize, int stride_w,
                              int relu, int y_ctotal, int y_coff, rd_stream_t stream);

/* ---- Training path of the convolution family ---------------------------------------------------
 * What the reference gets from MXNet autograd + cuDNN for every conv / BatchNorm / ReLU / add of the
 * DLA backbone and RPN head in training (mxnext/simple.py:123-158,545-580, mxnext/complicate.py:14,
 * 32-43; graph: rangedet/symbol/backbone/dla_backbone.py:17-127, rangedet/symbol/head/builder.py:
 * 198-266).  The data gradient of a convolution is itself a member of the forward family
 * (rd_conv2d / rd_deconv2d with transposed weights); the pieces below are the rest.
 * All activation tensors: zero-haloed NHWC bf16 [N][H+2][W+2][C], interior written only.
 *
 * rd_conv2d_wgrad_nhwc_bf16:  G[tap][a][b] = sum_{n,h,w} A[n,h,w][a] * B[n, h+ky-pad, w*stride_w+kx-pad][b]
 *   (tap = ky*ksize+kx, pad = ksize/2), fp32 [ksize*ksize][CA][CB] -- the packed layout of the forward
 *   weights when A = gradient of the conv output (CA = Cout) and B = the conv input (CB = Cin).
 *   A is at resolution (H, W), B at (H, W*stride_w).  CA in {64,128}; CB multiple of 64, <= 1024;
 *   ksize in {1,3}; stride_w in {1,2}.  tcgen05 (both operands MN-major), deterministic.
 */
size_t rd_conv2d_wgrad_workspace_bytes(int N, int H, int W, int CA, int CB, int ksize, int stride_w);
int rd_conv2d_wgrad_nhwc_bf16(const void* a_pad, const void* b_pad, float* g, int N, int H, int W, int CA,
                              int CB, int ksize, int stride_w, void* workspace, size_t workspace_bytes,
                              rd_stream_t stream);

/* Training-mode BatchNorm (batch statistics per GPU, biased variance, eps added to the variance,
 * moving = moving*momentum + batch*(1-momentum)) fused with ReLU and the residual adds.
 * coef: fp32 [6][C] = a (gamma*invstd) | b (beta - mean*a) | mean | invstd | var | sum -- produced by
 * rd_bn_train_stats, consumed by rd_bn_act_fwd / rd_bn_act_bwd.  gamma/beta NULL -> 1/0; moving_* NULL ->
 * not updated.  Workspace: rd_bn_workspace_bytes(C) (+ 8*C floats for the backward / channel sums).
 *   rd_bn_act_fwd : y = relu?(z*a + b + res_before) + res_after         (either residual may be NULL)
 *   rd_bn_act_bwd : g = dy * mask; dz = a*(g - mean(g) - xhat*mean(g*xhat)); dgamma, dbeta;
 *                   mask_mode 0: none, 1: (y_mask > 0), 2: (z*a + b > 0) recomputed (res_after layers);
 *                   dz is written with a W halo of dz_halo_w pixels (S for the phase-grouped view a
 *                   transposed convolution's backward reads, else 1); g_out (optional) receives g,
 *                   the gradient that flows into res_before.
 *   rd_channel_sums: sums[c] = sum over pixels of x (bias gradient of the un-normalised head convs)
 *   rd_add_nhwc_bf16: y = x0 + x1 (gradient accumulation at fan-out points)
 */
size_t rd_bn_workspace_bytes(int C);
int rd_bn_train_stats_nhwc_bf16(const void* z_pad, int N, int H, int W, int C, const float* gamma,
                                const float* beta, float eps, float momentum, float* moving_mean,
                                float* moving_var, float* coef, void* workspace, size_t workspace_bytes,
                                rd_stream_t stream);
int rd_bn_act_fwd_nhwc_bf16(const void* z_pad, const float* coef, const void* res_before,
                            const void* res_after, void* y_pad, int N, int H, int W, int C, int relu,
                            rd_stream_t stream);
int rd_bn_act_bwd_nhwc_bf16(const void* dy_pad, const void* y_mask_pad, const void* z_pad, const float* coef,
                            int mask_mode, void* dz_pad, int dz_halo_w, void* g_out_pad, float* dgamma,
                            float* dbeta, int N, int H, int W, int C, void* workspace,
                            size_t workspace_bytes, rd_stream_t stream);
int rd_channel_sums_nhwc_bf16(const void* x_pad, int N, int H, int W, int C, float* sums, void* workspace,
                              size_t workspace_bytes, rd_stream_t stream);
int rd_add_nhwc_bf16(const void* x0_pad, const void* x1_pad, void* y_pad, int N, int H, int W, int C,
                     rd_stream_t stream);

/* Training forward with the batch statistics fused into the convolution epilogue: y = conv(x) (raw, no scale / shift /
 * ReLU / residual) as rd_conv2d_nhwc_*, and the per-channel sums of the STORED values (sum z, sum z^2) accumulated by
 * the epilogue warps straight from the staged output tile -- the separate full read of z by rd_bn_train_stats goes away
 * (mxnext/complicate.py:32-43: mx.sym.BatchNorm with batch statistics, following every conv of dla_backbone.py:17-56
 * and builder.py:198-246).  stats_partial: rd_bn_workspace_bytes(Cout) bytes; *stats_slots receives the number of
 * partial slots written.  rd_bn_train_finalize turns them into `coef` and updates the moving statistics exactly like
 * the second half of rd_bn_train_stats. */
int rd_conv2d_nhwc_bf16_stats(const void* x_pad, const void* w_packed, void* y_pad, int N, int H, int W, int Cin,
                              int Cout, int ksize, int stride_w, float* stats_partial, size_t stats_bytes,
                              int* stats_slots, rd_stream_t stream);
int rd_bn_train_finalize(const float* partial, int nslots, int N, int H, int W, int C, const float* gamma,
                         const float* beta, float eps, float momentum, float* moving_mean, float* moving_var,
                         float* coef, rd_stream_t stream);

/* The same fusion for the backward pass.  In a conv -> BatchNorm -> ReLU -> conv chain (dla_backbone.py:23-41: conv1 /
 * bn1 / relu / conv2 of every basic block; builder.py:198-246: the head towers) the gradient dy of the BatchNorm output is
 * the data gradient of the NEXT convolution and has no other contributor.  rd_conv2d_nhwc_*_bwdstats is that data-gradient
 * convolution (3x3, stride 1, 128 output channels = the channels of the BatchNorm below; x_pad = the next layer's dz,
 * w_packed = its flipped / transposed weight) with the two sums the BatchNorm backward needs accumulated by the epilogue
 * from the gradient it stores: S1 = sum g, S2 = sum g (z - mean), g = dy (bn_mask_mode 0) or dy where z*a + b > 0 (2);
 * bn_z_pad is the BatchNorm's input z, bn_coef its coefficient block.  rd_bn_act_bwd_apply_* then finishes
 * rd_bn_act_bwd from those sums (finalize + apply: dz, dgamma, dbeta) -- the separate reduction pass over dy and z
 * goes away.  sums_partial: rd_bn_workspace_bytes(Cout) bytes; workspace of the apply call: 2*C floats. */
int rd_conv2d_nhwc_bf16_bwdstats(const void* x_pad, const void* w_packed, void* y_pad, const void* bn_z_pad,
                                 const float* bn_coef, int bn_mask_mode, int N, int H, int W, int Cin, int Cout,
                                 float* sums_partial, size_t sums_bytes, int* sums_slots, rd_stream_t stream);
int rd_bn_act_bwd_apply_nhwc_bf16(const void* dy_pad, const void* y_mask_pad, const void* z_pad, const float* coef,
                                  int mask_mode, const float* sums_partial, int sums_slots, void* dz_pad,
                                  int dz_halo_w, void* g_out_pad, float* dgamma, float* dbeta, int N, int H, int W,
                                  int C, void* workspace, size_t workspace_bytes, rd_stream_t stream);

/* Layout conversions across the Meta-Kernel op boundary (the reference is NCHW throughout; channel
 * index of the (B,9C,H,W) Meta-Kernel tensors is c*9+k, meta_kernel.py:232-239):
 *   src_pad / dst_pad : zero-haloed NHWC bf16 [N][H+2][W+2][C_src or C_dst], interior touched only
 *   dst / src         : NCHW fp32 [N][C][H][W]
 *   chmap 0: same channel order; chmap 1: NHWC channel k*(C/9)+c  <->  NCHW channel c*9+k (tap-major). */
int rd_nhwc_bf16_to_nchw_f32(const void* src_pad, float* dst, int N, int H, int W, int C_src, int C, int chmap,
                             rd_stream_t stream);
int rd_nchw_f32_to_nhwc_bf16(const float* src, void* dst_pad, int N, int H, int W, int C, int C_dst, int chmap,
                             rd_stream_t stream);

/* ---- Parameter plumbing of the training step --------------------------------------------------
 * One launch over flat buffers instead of one per parameter tensor (optim.cu).
 * rd_gather_f32_to_bf16: dst[i] = bf16(idx[i] >= 0 ? src[idx[i]] : 0) -- re-packs every bf16 conv operand
 *   (forward / data-gradient / deconv layouts) from the flat fp32 master parameters.
 * rd_gather_f32:         dst[i] = idx[i] >= 0 ? src[idx[i]] : 0      -- collects every parameter gradient from
 *   the kernels' native output layouts into the flat buffer that is all-reduced (tools/train.py:364-368).
 * rd_sgd_mom_update: MXNet sgd_mom_update with fp32 masters (tools/train.py:306-319, 359-361):
 *   g = clip(rescale*grad, +-clip) + wd[i]*w;  mom = momentum*mom - lr*g;  w += mom
 *   hyper = DEVICE pointer to {lr, momentum, rescale_grad, clip_gradient (<= 0: off)}; wd[i] is per element
 *   (0 for *_bias / *_beta like Optimizer.set_wd_mult). */
int rd_gather_f32_to_bf16(const float* src, const int* idx, void* dst, int64_t n, rd_stream_t stream);
int rd_gather_f32(const float* src, const int* idx, float* dst, int64_t n, rd_stream_t stream);
/* Channel-slice copy between two pixel-major tensors of 2-byte elements (either storage type): for each of npix pixels
 * dst[p][dst_off + c] = src[p][src_off + c], c < nchan; counts and offsets multiples of 8.  This is the
 * mx.sym.concat(data, agg3) feeding the level-0 head towers (rangedet/symbol/head/builder.py:198-266 via
 * dla_backbone.py:150-161), written straight into the 128-channel operand buffer. */
int rd_copy_channels_16b(const void* src, int src_ctotal, int src_off, void* dst, int dst_ctotal, int dst_off, int nchan,
                         int64_t npix, rd_stream_t stream);
int rd_sgd_mom_update(float* weight, const float* grad, float* mom, const float* wd, const float* hyper,
                      int64_t n, rd_stream_t stream);

/* ---- fp16-storage twins ------------------------------------------------------------------------
 * The reference trains this graph with fp16 storage and loss scale 128 (config/rangedet/
 * rangedet_veh_wo_aug_4_18e.py:35-36; casts at rangedet/symbol/backbone/dla_backbone.py:136-137 and
 * meta_kernel.py:193-196, un-scaled by rescale_grad at tools/train.py:359-361).  Every entry point above whose
 * name carries `bf16` has a twin with identical arguments and semantics in which the stored activations /
 * operands (x_pad, w_packed, residual_pad, y_pad, z_pad, dy_pad, ... ) are IEEE fp16 instead of bf16: same
 * kernels compiled from the same source with the other storage type (csrc/act_type.cuh), same tcgen05
 * kind::f16 MMA with fp32 accumulation, fp32 statistics / parameter gradients. */
int rd_meta_kernel_fwd_nhwc_f16(const float* data, const float* coord, const float* w0, const float* b0,
                                const float* w1, const float* b1, const float* scale, const float* shift,
                                int relu, void* y_pad, int B, int C, int H, int W, rd_stream_t stream);
int rd_conv2d_nhwc_f16(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                       const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout,
                       int ksize, int stride_w, int relu, rd_stream_t stream);
int rd_deconv2d_nhwc_f16(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                         const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout,
                         int kw, int relu, rd_stream_t stream);
int rd_conv2d_nhwc_f16_slice(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                             void* y_pad, int N, int H, int W, int Cin, int Cout, int ksize, int stride_w,
                             int relu, int y_ctotal, int y_coff, rd_stream_t stream);
int rd_conv2d_nhwc_f16_stats(const void* x_pad, const void* w_packed, void* y_pad, int N, int H, int W, int Cin,
                             int Cout, int ksize, int stride_w, float* stats_partial, size_t stats_bytes,
                             int* stats_slots, rd_stream_t stream);
int rd_conv2d_wgrad_nhwc_f16(const void* a_pad, const void* b_pad, float* g, int N, int H, int W, int CA,
                             int CB, int ksize, int stride_w, void* workspace, size_t workspace_bytes,
                             rd_stream_t stream);
int rd_bn_train_stats_nhwc_f16(const void* z_pad, int N, int H, int W, int C, const float* gamma,
                               const float* beta, float eps, float momentum, float* moving_mean,
                               float* moving_var, float* coef, void* workspace, size_t workspace_bytes,
                               rd_stream_t stream);
int rd_bn_act_fwd_nhwc_f16(const void* z_pad, const float* coef, const void* res_before,
                           const void* res_after, void* y_pad, int N, int H, int W, int C, int relu,
                           rd_stream_t stream);
int rd_bn_act_bwd_nhwc_f16(const void* dy_pad, const void* y_mask_pad, const void* z_pad, const float* coef,
                           int mask_mode, void* dz_pad, int dz_halo_w, void* g_out_pad, float* dgamma,
                           float* dbeta, int N, int H, int W, int C, void* workspace,
                           size_t workspace_bytes, rd_stream_t stream);
int rd_conv2d_nhwc_f16_bwdstats(const void* x_pad, const void* w_packed, void* y_pad, const void* bn_z_pad,
                                const float* bn_coef, int bn_mask_mode, int N, int H, int W, int Cin, int Cout,
                                float* sums_partial, size_t sums_bytes, int* sums_slots, rd_stream_t stream);
int rd_bn_act_bwd_apply_nhwc_f16(const void* dy_pad, const void* y_mask_pad, const void* z_pad, const float* coef,
                                 int mask_mode, const float* sums_partial, int sums_slots, void* dz_pad,
                                 int dz_halo_w, void* g_out_pad, float* dgamma, float* dbeta, int N, int H, int W,
                                 int C, void* workspace, size_t workspace_bytes, rd_stream_t stream);
int rd_channel_sums_nhwc_f16(const void* x_pad, int N, int H, int W, int C, float* sums, void* workspace,
                             size_t workspace_bytes, rd_stream_t stream);
int rd_add_nhwc_f16(const void* x0_pad, const void* x1_pad, void* y_pad, int N, int H, int W, int C,
                    rd_stream_t stream);
int rd_nhwc_f16_to_nchw_f32(const void* src_pad, float* dst, int N, int H, int W, int C_src, int C, int chmap,
                            rd_stream_t stream);
int rd_nchw_f32_to_nhwc_f16(const float* src, void* dst_pad, int N, int H, int W, int C, int C_dst, int chmap,
                            rd_stream_t stream);
int rd_gather_f32_to_f16(const float* src, const int* idx, void* dst, int64_t n, rd_stream_t stream);

/* ---- Launch mode ------------------------------------------------------------------------------------
 * The conv / wgrad / BatchNorm kernels are launched with programmatic dependent launch (each kernel's prologue and
 * launch latency overlap the tail of its predecessor; csrc/rd_common.cuh).  rd_set_pdl(0) switches to plain stream
 * order (diagnostics, A/B timing); returns the previous setting.  Default: on, unless the environment has RD_PDL=0. */
int rd_set_pdl(int on);
/* 3x3 / stride-1 convolutions with 128 output channels run in the transposed GEMM orientation of csrc/conv_t.cu
 * (M = Cout, N = 256 flattened pixels) by default; rd_set_conv_t(0) (or RD_CONV_T=0) keeps them on the M = pixels kernel of
 * csrc/conv_tc.cu.  Same results (same K order); returns the previous setting.  on = 160 / 192 / 224 / 256 additionally fixes
 * the pixel-tile width (default: chosen per tensor shape so that the last round of tiles fills the SMs). */
int rd_set_conv_t(int on);

/* ---- tcgen05 self-test -------------------------------------------------------------------
 * D(128 x n) = A(128 x k) . B(n x k)^T with bf16 operands staged in shared memory in the
 * canonical no-swizzle K-major core-matrix layout, tcgen05.mma into TMEM, tcgen05.ld back.
 * Used by tests to validate the descriptor encodings the fused kernels rely on.
 * mn_major = 1 stages both operands in the MN-major canonical layout instead (same math).
 * a (128,k) b (n,k) fp32 (rounded to bf16 inside), d (128,n) fp32; k % 16 == 0, k <= 128,
 * n % 16 == 0, 16 <= n <= 256.
 */
int rd_tc_probe_gemm(const float* a, const float* b, float* d, int n, int k, int mn_major,
                     rd_stream_t stream);

/* ---- TMA self-test -----------------------------------------------------------------------
 * Loads the box (box_w, 1, 64) of a (W, H, C) fp32 tensor at signed coordinates (c0, c1, c2) through
 * cp.async.bulk.tensor into shared memory (zero fill outside the tensor), copies it to dst
 * (64 x box_w) and TMA-stores it to the same coordinates of dst2 (same shape as src; out-of-bound
 * parts are dropped).  W % 4 == 0, box_w % 4 == 0, 4 <= box_w <= 256, C >= 64.
 */
int rd_tma_probe(const float* src, float* dst, float* dst2, int W, int H, int C, int box_w, int c0,
                 int c1, int c2, rd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RANGEDET_B200_H_ */
