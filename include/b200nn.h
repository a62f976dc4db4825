/*
 * b200nn.h -- C ABI of libb200nn.so: hand-written sm_100a kernels for the 3-D-convolutional
 * hot path of kondratevakate/mri-epilepsy-diagnosis.
 *
 * The reference has no FFI layer of its own: its operator API *is* torch.nn, bound at import
 * by `import torch.nn as nn` (segmentation/models/unet3d.py:1, classification/models/AE_model.py:2,
 * classification/models/cnn_model.py:3, segmentation/models/modified_3dunet.py:1,
 * detection/model_utils.py:3).  Each entry point below replaces the ATen/cuDNN call that one of
 * those torch.nn modules issues; the comment on each names the reference call sites.  The Python
 * drop-in modules (mri_epilepsy_diagnosis_b200.nn) bind these symbols with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every activation tensor is channels-last: physical N,D,H,W,C contiguous (what
 *     torch.channels_last_3d gives a logical N,C,D,H,W tensor; 2-D ops are the D=1 case);
 *   - dtypes are B200_F32 or B200_BF16; accumulation and all statistics are fp32;
 *   - plain pointers and sizes only; the callee never allocates, frees or retains device memory:
 *     outputs, workspaces, packed weights, saved statistics and argmax codes are caller-owned;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*), never synchronises;
 *   - return 0 on success, non-zero on error with a thread-local message in b200_last_error();
 *     an unsupported descriptor is an error -- there is no CPU fallback and no cuDNN dispatch.
 */
#ifndef B200NN_H
#define B200NN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NN_VERSION 100

enum { B200_F32 = 0, B200_BF16 = 1 };
enum { B200_ACT_NONE = 0, B200_ACT_RELU = 1, B200_ACT_LEAKY = 2 };
enum { B200_NORM_BATCH = 0, B200_NORM_INSTANCE = 1, B200_NORM_GROUP = 2 };
enum { B200_UP_NEAREST = 0, B200_UP_TRILINEAR = 1, B200_UP_TRILINEAR_ALIGNED = 2 };
enum { B200_PASS_FWD = 0, B200_PASS_DGRAD = 1, B200_PASS_WGRAD = 2 };
enum { B200_ALGO_SIMT = 0, B200_ALGO_UMMA = 1, B200_ALGO_ROW = 2 };

int b200_version(void);
const char* b200_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t b200_launch_count(void);

/* ------------------------------------------------------------------ convolution
 * nn.Conv3d / nn.Conv2d (cross-correlation, zero padding, groups=1): unet3d.py:30-36,57-63,99-101;
 * AE_model.py:9-26,74-91,160-164,218-234,268-284; cnn_model.py:14,49-81,212-240; model_utils.py:47;
 * nn.ConvTranspose3d: AE_model.py:62-68,159 (transposed=1: x is the small tensor, y the large one,
 * and `Ci`/`Co` still name the channels of x / y).
 */
typedef struct {
    int32_t x_dtype, y_dtype;            /* activation dtypes of the conv's input and output      */
    int32_t N;
    int32_t Ci, Di, Hi, Wi;              /* input  (x)                                            */
    int32_t Co, Do, Ho, Wo;              /* output (y)                                            */
    int32_t kd, kh, kw;
    int32_t sd, sh, sw;
    int32_t pd, ph, pw;
    int32_t dd, dh, dw;
    int32_t transposed;                  /* 0: Conv3d, 1: ConvTranspose3d                         */
    int32_t allow_umma;                  /* 0 forces the SIMT fp32-FMA kernels (tf32-off mode)    */
} b200_conv_desc;

/* Which kernel family the library will use for (desc, pass): B200_ALGO_SIMT or B200_ALGO_UMMA. */
int b200_conv_algo(const b200_conv_desc* d, int pass);
/* Packed-weight size for (desc, pass in {FWD, DGRAD}); layout is private to the library. */
size_t b200_conv_packed_bytes(const b200_conv_desc* d, int pass);
/* Pack the fp32 PyTorch-layout parameter (Conv: Co,Ci,kd,kh,kw; ConvTranspose: Ci,Co,kd,kh,kw). */
int b200_conv_pack_weights(const b200_conv_desc* d, int pass, const float* w, void* packed, void* stream);

/* Batched re-pack of many (layer, pass) weight copies in ONE launch.  Each entry gathers `count` packed elements through a
 * precomputed permutation: dst[i] = idx[i] < 0 ? 0 : src[idx[i]], src = src0 (n0 elements) followed by src1 (may be null), dst bf16
 * or fp32.  The host derives idx once per (descriptor, pass) by running b200_conv_pack_weights on index-coded weights
 * (mri_epilepsy_diagnosis_b200/functional.py:PackPlan).  Entries live in DEVICE memory; entry k owns blocks
 * [first_block, first_block + ceil(count / 2048)).  Replaces the per-module re-pack the reference gets for free from cuDNN reading
 * nn.Conv3d.weight in place (segmentation/models/unet3d.py:20-79: every conv of the step). */
typedef struct {
    const float* src0;
    const float* src1;
    const int32_t* idx;
    void* dst;
    int64_t count;
    int64_t first_block;
    int32_t n0;
    int32_t dst_bf16;
} b200_pack_entry;
int b200_pack_batched(const b200_pack_entry* entries_dev, int n_entries, int64_t total_blocks, void* stream);
size_t b200_conv_workspace_bytes(const b200_conv_desc* d, int pass);
/* y = conv(x, w) + bias.  `bias` is fp32 [Co] or NULL. */
int b200_conv_fwd(const b200_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y,
                  void* workspace, size_t ws_bytes, void* stream);
/* Convolution + the statistics of the BatchNorm that follows (unet3d.py:42-46 conv->bn pairs), fused into the conv epilogue:
 * `stat_partial` receives fp32 [chunks][2][Co] per-block (sum, sum of squares) of the fp32 outputs, chunks =
 * b200_conv_stats_chunks(d) (0 = this descriptor has no fused-statistics kernel; use b200_norm_stats instead).
 * Feed the partials to b200_norm_stats_from_partial. */
int b200_conv_stats_chunks(const b200_conv_desc* d);
int b200_conv_fwd_stats(const b200_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y,
                        float* stat_partial, void* workspace, size_t ws_bytes, void* stream);
/* Two convolutions of the SAME input evaluated as one with concatenated output channels (d->Co = Co_a + Co_b) when only the
 * statistics of the first Co_a = first_stored_channel outputs are needed (unet3d.py:43-46: conv2 feeds a branch whose value is
 * discarded, conv3 the live one): statistics cover all Co channels, `y_tail` receives channels [first_stored_channel, Co) as a
 * dense (N, Co - first_stored_channel, ...) tensor.  first_stored_channel must be a multiple of 16, Co <= 64. */
int b200_conv_fwd_stats_tail(const b200_conv_desc* d, const void* x, const void* w_packed, const float* bias, void* y_tail,
                             int first_stored_channel, float* stat_partial, void* workspace, size_t ws_bytes, void* stream);
/* dx = d(loss)/dx given dy (dtypes: dy has y_dtype, dx has x_dtype). */
int b200_conv_dgrad(const b200_conv_desc* d, const void* dy, const void* w_packed_dgrad, void* dx,
                    void* workspace, size_t ws_bytes, void* stream);
/* dw (fp32, PyTorch parameter layout) and optionally dbias (fp32 [Co], may be NULL); overwritten, not accumulated. */
int b200_conv_wgrad(const b200_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                    void* workspace, size_t ws_bytes, void* stream);

/* --------------------------------------------------- normalisation (+ fused activation / residual)
 * nn.BatchNorm3d/2d (unet3d.py:10; AE_model.py:30,94; cnn_model.py; model_utils.py:48), nn.InstanceNorm3d
 * (unet3d.py:14; modified_3dunet.py:20-43), nn.GroupNorm(4,C) (unet3d.py:12).  eps/momentum are the
 * module's; statistics are fp32 with a shifted two-pass-free formulation.
 *
 *   stats : per-group mean and biased variance from x; when `running_mean` is given also the
 *           BatchNorm running update (momentum, UNBIASED variance) -- torch.nn semantics.
 *   apply : y = act( (x-mean)*rstd*gamma + beta [+ residual] )
 *   bwd   : dx, dgamma, dbeta (act' folded in through the saved OUTPUT y when act != NONE)
 * groups: BATCH -> C groups; INSTANCE -> N*C; GROUP -> N*G.  `mean`/`rstd` are fp32 [groups].
 */
typedef struct {
    int32_t dtype;
    int32_t N, C;
    int64_t S;                 /* D*H*W */
    int32_t kind;              /* B200_NORM_*                       */
    int32_t G;                 /* GROUP only                        */
    float eps, momentum;
    int32_t act;               /* B200_ACT_* fused into apply / bwd */
    float slope;               /* LEAKY                              */
} b200_norm_desc;

size_t b200_norm_workspace_bytes(const b200_norm_desc* d);
int b200_norm_stats(const b200_norm_desc* d, const void* x, float* mean, float* rstd,
                    float* running_mean, float* running_var, void* workspace, size_t ws_bytes, void* stream);
/* BatchNorm batch statistics (and running update) from the partial sums written by b200_conv_fwd_stats */
int b200_norm_stats_from_partial(const b200_norm_desc* d, const float* partial, int chunks, float* mean, float* rstd,
                                 float* running_mean, float* running_var, void* stream);
/* eval-mode BatchNorm: mean/rstd derived from the running statistics */
int b200_norm_stats_from_running(const b200_norm_desc* d, const float* running_mean, const float* running_var,
                                 float* mean, float* rstd, void* stream);
/* SyncBN (data parallel): packed[2C] = (mean, var + mean^2) of this rank's statistics -> the caller all-reduces
 * `packed` with AVG over the ranks -> finalize writes the global (mean, rstd) and updates the running statistics with the unbiased
 * variance for `count_global` = world x N x S elements. */
int b200_syncbn_pack(int C, float eps, const float* mean, const float* rstd, float* packed, void* stream);
int b200_syncbn_finalize(int C, float eps, float momentum, double count_global, const float* packed, float* mean, float* rstd,
                         float* running_mean, float* running_var, void* stream);
int b200_norm_apply(const b200_norm_desc* d, const void* x, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, const void* residual, void* y, void* stream);
/* training=1: batch statistics take part in the gradient; 0: eval-mode BN (mean/rstd are constants).
 * `y` (the saved OUTPUT) supplies the activation gate; it may be NULL when act != NONE and the forward had NO residual: the
 * gate is then recomputed from x through the same affine form as the forward pass (needs `beta`), one tensor read less. */
int b200_norm_bwd(const b200_norm_desc* d, int training, const void* x, const void* y, const void* dy,
                  const float* mean, const float* rstd, const float* gamma, const float* beta,
                  void* dx, void* dresidual, float* dgamma, float* dbeta,
                  void* workspace, size_t ws_bytes, void* stream);

/* The same backward split in two so that a data-parallel caller can all-reduce `sums` (fp32 [NB*C*2]: sum dy',
 * sum dy'*xhat per (reduce-batch, channel); NB = 1 for BATCH, N otherwise) between the steps -- SyncBN.
 * `world` multiplies the BATCH element count (number of ranks whose sums were added). */
int b200_norm_bwd_reduce(const b200_norm_desc* d, const void* x, const void* y, const void* dy,
                         const float* mean, const float* rstd, const float* gamma, const float* beta, float* sums,
                         void* workspace, size_t ws_bytes, void* stream);
int b200_norm_bwd_apply(const b200_norm_desc* d, int training, int world, const void* x, const void* y, const void* dy,
                        const float* mean, const float* rstd, const float* gamma, const float* beta, const float* sums,
                        void* dx, void* dresidual, float* dgamma, float* dbeta,
                        void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ activations
 * nn.ReLU(inplace) unet3d.py:28,66; nn.LeakyReLU() AE_model.py:32; nn.PReLU(1) unet.UNet (routine.py:355).
 * y may alias x (inplace).  bwd for RELU/LEAKY reads the OUTPUT y; PReLU reads the input x. */
int b200_act_fwd(int dtype, int act, float slope, int64_t n, const void* x, void* y, void* stream);
int b200_act_bwd(int dtype, int act, float slope, int64_t n, const void* y, const void* dy, void* dx, void* stream);
int b200_prelu_fwd(int dtype, int64_t n, const void* x, const float* a, void* y, void* stream);
size_t b200_prelu_workspace_bytes(int64_t n);
int b200_prelu_bwd(int dtype, int64_t n, const void* x, const float* a, const void* dy, void* dx, float* da,
                   void* workspace, size_t ws_bytes, void* stream);
/* y = act(a + b) -- residual joins unet3d.py:47, cnn_model.py:37-38 */
int b200_add_act_fwd(int dtype, int act, float slope, int64_t n, const void* a, const void* b, void* y, void* stream);

/* ------------------------------------------------------------------ fused softmax + soft-Dice loss
 * The loss of the segmentation loop (segmentation/routine.py:272-274 + get_dice_score/get_dice_loss :239-253):
 *   p = softmax(logits, dim=1); per (n,c): tp = sum p*t, fp = sum p*(1-t), fn = sum (1-p)*t with t = targets (N,1,...) broadcast
 *   over the channels (the reference scores both channels against the same mask); loss = mean(1 - 2tp/(2tp+fp+fn+eps)).
 * logits: channels-last (N,S,C) fp32 or bf16, C <= 8; targets fp32 (N,S).  `sums` (fp32 [N][2C+1]: sum p_c, sum p_c*t, sum t)
 * is written by fwd and read by bwd; `dloss` is the upstream gradient (a DEVICE scalar); dlogits has the logits' dtype. */
typedef struct {
    int32_t dtype;
    int32_t N, C;
    int64_t S;
    float eps;
} b200_dice_desc;
size_t b200_softmax_dice_workspace_bytes(const b200_dice_desc* d);
int b200_softmax_dice_fwd(const b200_dice_desc* d, const void* logits, const float* targets, float* sums, float* loss,
                          void* workspace, size_t ws_bytes, void* stream);
int b200_softmax_dice_bwd(const b200_dice_desc* d, const void* logits, const float* targets, const float* sums,
                          const float* dloss, void* dlogits, void* stream);

/* ------------------------------------------------------------------ max pooling
 * nn.MaxPool3d(k, s) no padding, floor mode: unet3d.py:25; AE_model.py:27; cnn_model.py:115-148,221,232;
 * nn.MaxPool2d(2) model_utils.py:29.  Ties -> first in (d,h,w) raster order; NaN wins.
 * `code` (uint8, one per output element) is the window-local argmax used by the backward pass;
 * `indices` (int64, optional) are torch's flat offsets inside the input D*H*W plane. */
typedef struct {
    int32_t dtype;
    int32_t N, C, Di, Hi, Wi, Do, Ho, Wo;
    int32_t kd, kh, kw, sd, sh, sw;
} b200_pool_desc;
int b200_maxpool_fwd(const b200_pool_desc* d, const void* x, void* y, uint8_t* code, int64_t* indices, void* stream);
int b200_maxpool_bwd(const b200_pool_desc* d, const void* dy, const uint8_t* code, void* dx, void* stream);
/* dx = maxpool_bwd(dy) + add, for a pooled tensor that has a second consumer (U-Net skip connection, unet3d.py:113-121 with
 * torch.cat at :76): `add` is that consumer's gradient, channels-last with `add_ctot` channels per voxel (>= C; the pointer
 * already addresses the first channel of the window).  Replaces autograd's accumulation kernel.  kernel = stride = 2 only. */
int b200_maxpool_bwd_add(const b200_pool_desc* d, const void* dy, const uint8_t* code, const void* add, int add_ctot, void* dx, void* stream);

/* ------------------------------------------------------------------ upsample (+ skip concat)
 * F.upsample/nn.Upsample trilinear align_corners=False (unet3d.py:73,85; unet.UNet), =True
 * (3d_bayes_layers.py:65), nearest (AE_model.py:70-73,119; modified_3dunet.py:13).  The result is written
 * into channels [c_off, c_off+C) of an output whose channel count is Ctot, so `torch.cat([skip, up], 1)`
 * (unet3d.py:76) costs no extra pass: the skip is copied by b200_copy_channels. */
typedef struct {
    int32_t dtype;
    int32_t mode;              /* B200_UP_*               */
    int32_t N, C, Di, Hi, Wi, Do, Ho, Wo;
    int32_t Ctot, c_off;       /* output channel stride and offset */
} b200_up_desc;
int b200_upsample_fwd(const b200_up_desc* d, const void* x, void* y, void* stream);
int b200_upsample_bwd(const b200_up_desc* d, const void* dy, void* dx, void* stream);
/* dst[v, dst_off:dst_off+C] = src[v, src_off:src_off+C] for V voxels (concat / split along channels) */
int b200_copy_channels(int dtype, int64_t V, int32_t C, const void* src, int32_t src_ctot, int32_t src_off,
                       void* dst, int32_t dst_ctot, int32_t dst_off, void* stream);

/* ------------------------------------------------------------------ layout / dtype
 * NCDHW <-> NDHWC transposes with dtype conversion (first layer reads the loader's fp32 NCDHW batch,
 * segmentation/routine.py:190). */
int b200_to_channels_last(int src_dtype, int dst_dtype, int32_t N, int32_t C, int64_t S, const void* src, void* dst, void* stream);
int b200_from_channels_last(int src_dtype, int dst_dtype, int32_t N, int32_t C, int64_t S, const void* src, void* dst, void* stream);

/* min-max normalisation of get_image_patches (detection/patch_utils.py:196): out = (x - min(x)) / (max(x) - min(x)), float64,
 * the same two correctly rounded operations per element as numpy (bit-exact).  workspace: b200_minmax_workspace_bytes() bytes. */
size_t b200_minmax_workspace_bytes(void);
int b200_minmax_normalize(const double* x, int64_t n, double* out, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ FCD mask post-processing (row f-4)
 * FCDMaskGenerator (detection/model_utils.py:118-228) around the batched patch classifier: `plan` is the (rows, 5) int32 output of
 * b200_patch_plan for the template; patch_map is the reference's (4, Y/h, Z) int64 array, mask its (X, Y, Z) int64 result.
 * b200_fcd_scatter_labels  :160-178   patch_map[slot][row0/h][slice] = labels[row] (patch_map must be zeroed by the caller)
 * b200_fcd_vote            :182-193   out-of-place 4-neighbour vote; fixed = 0 reproduces the reference's int-array-as-index
 *                                     behaviour (slabs 0 / 1 overwritten), fixed = 1 the boolean-mask vote; flags4: 4 zeroed int32
 * b200_fcd_paint           :195-216   paints the patch labels back (mask must be zeroed by the caller) */
int b200_fcd_scatter_labels(const int32_t* plan, int64_t rows, const int64_t* labels, int X, int Y, int Z, int h, int w, int64_t* patch_map, void* stream);
int b200_fcd_vote(const int64_t* patch_map, int ny, int Z, int fixed, int64_t* out, int32_t* flags4, void* stream);
int b200_fcd_paint(const int32_t* plan, int64_t rows, const int64_t* patch_map, int X, int Y, int Z, int h, int w, int64_t* mask, void* stream);

/* ------------------------------------------------------------------ surface distances (row f-2, the surface half)
 * compute_surface_distances, segmentation/metrics.py:25-178 (neighbour-code correlate :123-130, borders :133-135, exact Euclidean
 * distance transform :139-149, surfel lists :157-160), called from segmentation/routine.py:206-214.  All arrays live on the corner
 * grid (D+1, H+1, W+1) of a (D, H, W) mask volume; voxels != 0 are inside.
 * b200_surface_codes:   code u8 per corner + the number of border corners (code != 0 && != 255) added to *nborder (zero it first).
 * b200_surface_edt:     exact squared distance of every corner to the nearest border corner of `code`; spacing == NULL: int32
 *                       result in voxel units (exact), else fp64 with the three spacings; `scratch` has the result's size.
 * b200_surface_collect: for every border corner of `code`: its code and dist2_other[corner], appended at out[atomic counter++]
 *                       (unordered; *counter must be zeroed first; outputs sized by the b200_surface_codes count). */
int b200_surface_codes(const uint8_t* mask, int D, int H, int W, uint8_t* code, uint32_t* nborder, void* stream);
int b200_surface_edt(const uint8_t* code, int D1, int H1, int W1, const double* spacing, void* dist2, void* scratch, void* stream);
int b200_surface_collect(const uint8_t* code, const void* dist2_other, int is_f64, int64_t corners, void* out_dist2, uint8_t* out_code,
                         uint32_t* counter, void* stream);

/* ------------------------------------------------------------------ sliding-grid patch inference / random-patch sampling (row f-3)
 * torchio.inference.GridSampler / GridAggregator and torchio.Queue(ImageSampler) as called at
 * segmentation/pretraining_3d_unet.ipynb [cell 26, 35] and segmentation/routine.py:150-178 (third-party torchio, version
 * unpinned: the host side restates its published <= 0.16 algorithm, see mri_epilepsy_diagnosis_b200/grid.py).
 * b200_grid_gather:    out[l][c][:,:,:] = vol[c][z0:z0+pd, y0:y0+ph, x0:x0+pw] for the L windows loc[l] = (z0,y0,x0,z1,y1,x1).
 * b200_grid_aggregate: vol (D,H,W) receives, voxel by voxel, the value of the LAST window (in `loc` order) whose box cropped by
 *                      `border` on every side contains it (== the reference's sequential overwrites); other voxels are untouched.
 * elem_bytes: 1, 2, 4 or 8 (plain copies: bit-exact for every dtype). */
int b200_grid_gather(int elem_bytes, const void* vol, const int32_t* loc, int64_t L, int C, int D, int H, int W, int pd, int ph, int pw,
                     void* out, void* stream);
int b200_grid_aggregate(int elem_bytes, const void* labels, const int32_t* loc, int L, int D, int H, int W, int pd, int ph, int pw,
                        int bd, int bh, int bw, void* vol, void* stream);

/* ------------------------------------------------------------------ sliding-window patch gather
 * detection/patch_utils.py: get_only_patches :142-191, get_all_patches_and_labels :17-140.
 * Volumes are C-order (X,Y,Z) float64 exactly as the reference holds them.  Two steps:
 *   plan  : per (slice, strip) decisions + order-preserving compaction -> int32 plan rows
 *           {slice, row0, c0, c1, label}; *count receives the number of rows; error flags in *status
 *           (bit0: start_idx==0 assertion, bit1: ragged strip would be emitted).
 *   gather: copy the planned (2,h,w) windows (channel 1 mirrored) -> out (P,2,h,w) float64 or float32.
 */
typedef struct {
    int32_t X, Y, Z, h, w;
    int32_t with_mask;         /* labels from a uint8 mask volume            */
    int32_t upsample_passes;   /* 1: also the k=1..h-1 positive-only passes  */
} b200_patch_desc;
size_t b200_patch_workspace_bytes(const b200_patch_desc* d);
int64_t b200_patch_max_rows(const b200_patch_desc* d);
int b200_patch_plan(const b200_patch_desc* d, const double* gmpm, const uint8_t* mask, int32_t* plan,
                    int32_t* count, int32_t* status, void* workspace, size_t ws_bytes, void* stream);
int b200_patch_gather(const b200_patch_desc* d, const double* target, const int32_t* plan, int64_t rows,
                      int out_dtype_is_f32, void* out, void* stream);

/* ---- validation overlap metrics (the counting part) ----------------------------------------------
 * One pass over the predicted label volume and the ground truth (uint8, as validate_dsc_asd passes them,
 * segmentation/routine.py:216-237) -> counts5 (device, 5 x uint64):
 *   [0] gt.sum()  [1] pred.sum()  [2] (gt & pred).sum()      -- compute_dice_coefficient, segmentation/metrics.py:312-329
 *   [3] #(pred > 0 and gt > 0)    [4] #(pred > 0 or gt > 0)  -- get_iou_score, segmentation/routine.py:198-204 */
int b200_overlap_counts(const uint8_t* pred, const uint8_t* gt, int64_t n, uint64_t* counts5, void* stream);

/* ---- intensity preprocessing -------------------------------------------------------------------
 * Histogram standardisation of one volume = `normalize(tensor, landmarks, mask, cutoff, epsilon)` of
 * classification/train_ENC_CLF.ipynb [cell 9] (the collate function of the classification loaders, cell 9 `default_collate`):
 *   percentile_values = np.percentile(data[mask], percentiles)          exact order statistics, numpy 'linear' interpolation
 *   piecewise-linear map of the percentiles listed in range_idx onto `landmarks`; np.digitize + slope * x + intercept in float64
 * q: the percentiles as fractions in [0, 1] (ascending); landmarks: the trained mapping, one value per percentile;
 * range_idx: which of them take part in the map (the notebook uses 11 of its 13).  x: n float32 values on the device, mask:
 * optional n bytes (non-zero = selected for the percentiles; every element is mapped), out: n float32, percentiles_out: optional
 * nq doubles on the device.  Finite inputs; n < 2^32. */
typedef struct {
    double q[16];
    double landmarks[16];
    int32_t range_idx[16];
    int32_t nq, nrange;
    double eps;
} b200_histstd_desc;
size_t b200_histstd_workspace_bytes(void);
int b200_histstd_normalize(const b200_histstd_desc* d, const float* x, const uint8_t* mask, int64_t n, float* out,
                           double* percentiles_out, void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NN_H */
