/* yolov5m_b200 -- C ABI of the B200-native hot path of AlessandroMondin/YOLOV5m.
 *
 * The reference is pure Python/PyTorch: it has no FFI of its own.  The boundary it
 * offers is the set of Python call signatures consumed by train.py / detect.py
 * (SURVEY.md 8b).  This header is the C ABI that the drop-in Python classes in
 * yolov5m_b200/ bind with ctypes; every entry point names the reference
 * interface (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions: extern "C"; plain pointers and sizes; all pointers are DEVICE
 * pointers owned by the caller unless stated; `stream` is a cudaStream_t passed
 * as void*; return 0 on success, negative on error (yb_last_error() describes it);
 * no hidden device allocation; activations are NHWC bf16 ("pitch" = elements
 * between consecutive pixels, so a tensor may be a channel slice of a wider
 * concat buffer).
 */
#ifndef YOLOV5M_B200_H
#define YOLOV5M_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* yb_last_error(void);
int yb_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t yb_launch_count(void);
/* number of per-CTA partial rows a conv launch may write to `stats` (= max grid = #SMs) */
int yb_conv_max_partials(void);

/* ---- Conv2d forward (model.py:16 CBL conv; model.py:162 head conv) ----------------
 * y[n,ho,wo,co] = epilogue( sum_{kh,kw,ci} x[n, ho*s+kh-p, wo*s+kw-p, ci] * w[co,(kh*ks+kw)*Cin+ci] )
 * ks in {1,3}, p = ks/2, s in {1,2}; ks = 31 means a 3x1 kernel (three vertical taps, stride 1, w_packed [Cout][3*Cin]).
 * ks = 31 with x_pitch < Cin (the stem: Cin = 48, x_pitch = 16): x is the row-padded 16-channel staging (N,H,W+2,16) of
 * yb_prep_input and the 48 "channels" of pixel w are the 48 contiguous values that start at padded column w (an
 * overlapping-window tensor-map view; the same convention holds for yb_conv_wgrad_plan).
 * w_packed: bf16 [Cout][ks*ks*Cin].
 * epilogue: v = acc; if scale: v = v*scale[co]+shift[co]; elif shift: v += shift[co];
 *           if act: v = SiLU(v); if addend: v += addend[n,ho,wo,co]; store.
 * out_kind 0: bf16 NHWC (y_pitch); 1: fp32 head layout (B,na,H,W,no) of model.py:173
 *          (Cout = na*no, need not be a multiple of 16); 2: fp32 NHWC; 3 / 4: like 2 / 1 but out += value
 *          (fp32 accumulate: the parity mode sums several bf16-split passes of one convolution, see yb_p32_* below).
 * stats (optional): fp32 [yb_conv_max_partials()][2][Cout]; row r receives CTA r's sum and
 * sum of squares of the raw accumulator per out channel (training-mode BatchNorm2d batch
 * statistics, model.py:17); *stats_rows = rows written. */
int yb_conv2d_fwd(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* w_packed, int Cout,
                  int ks, int stride, void* y, int64_t y_pitch, int out_kind, const float* scale,
                  const float* shift, int act, const void* addend, int64_t addend_pitch, float* stats,
                  int* stats_rows, int head_na, int head_no, void* stream);

/* ---- Conv2d data gradient (autograd of model.py:16 / :162) ---------------------------
 * dx[n,h,w,ci] = sum_{kh,kw,co} dy[n,(h+p-kh)/s,(w+p-kw)/s,co] * wt[ci,(kh*ks+kw)*Cout+co]
 * (terms with non-integral or out-of-range dy coordinates vanish).  wt_packed: bf16
 * [Cin][ks*ks*Cout].  H, W are the INPUT (dx) spatial dims.  out_kind 0 (bf16), 2 (fp32) or 3 (fp32 accumulate). */
int yb_conv2d_dgrad(const void* dy, int N, int H, int W, int Cout, int64_t dy_pitch, const void* wt_packed, int Cin,
                    int ks, int stride, void* dx, int64_t dx_pitch, const void* addend, int64_t addend_pitch,
                    int out_kind, void* stream);


/* ---- cached launch plans (TMA descriptors are encoded once per tensor geometry) ------------------------------
 * yb_conv_fwd_plan / yb_conv_dgrad_plan take the arguments of yb_conv2d_fwd / yb_conv2d_dgrad (minus the stream),
 * return an opaque handle (NULL on error -> yb_last_error()).  yb_plan_run launches it.  The pointers baked into a
 * plan must stay valid for its lifetime. */
void* yb_conv_fwd_plan(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* w_packed, int Cout,
                       int ks, int stride, void* y, int64_t y_pitch, int out_kind, const float* scale,
                       const float* shift, int act, const void* addend, int64_t addend_pitch, float* stats,
                       int* stats_rows, int head_na, int head_no);
void* yb_conv_dgrad_plan(const void* dy, int N, int H, int W, int Cout, int64_t dy_pitch, const void* wt_packed, int Cin,
                         int ks, int stride, void* dx, int64_t dx_pitch, const void* addend, int64_t addend_pitch,
                         int out_kind);
int yb_plan_run(void* plan, void* stream);
/* Kernel selection for 3x3 convolutions planned AFTER the call (tuning / test knob; default 0, or $YB_CONV_PATCH):
 * 0 = heuristic, 1 = halo-patch kernel (conv_patch.cu) wherever legal, 2 = same plus multi-tile super-tiles on small
 * problems, -1 = generic kernel (conv_igemm.cu) only.  Results are identical up to fp32 summation order. */
void yb_set_conv_patch_mode(int mode);
/* same for the weight-gradient kernels planned after the call (conv_wgrad_patch.cu vs conv_wgrad.cu; $YB_WGRAD_PATCH) */
void yb_set_wgrad_patch_mode(int mode);
void yb_plan_destroy(void* plan);

/* ---- Conv2d weight gradient (autograd of model.py:16 / :162) ---------------------------------------------------
 * dw[co][(kh*ks+kw)*Cin+ci] (+)= sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*s+kh-p, wo*s+kw-p, ci]      (fp32 output)
 * x: conv input (N,H,W,Cin) bf16 NHWC; dy: (N,H/s,W/s,Cout) bf16 NHWC (Cout a multiple of 16; the head pads 255->256
 * and passes out_rows = 255).  workspace: fp32 scratch for the split-K partial tiles, at least
 * yb_conv_wgrad_workspace_floats(Cin, Cout, ks) floats (one partial tile, channels padded to the 64-wide TMA boxes);
 * more lets the planner split the pixel reduction over more CTAs.  index_map (optional, int32 [Cout*ks*ks*Cin]): element
 * i of the packed gradient is written to dw[index_map[i]] (skipped when negative) -- used to fold the space-to-depth
 * stem back to its 6x6 layout.  Deterministic: partials are reduced in a fixed order. */
int64_t yb_conv_wgrad_workspace_floats(int Cin, int Cout, int ks);
void* yb_conv_wgrad_plan(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* dy, int Cout,
                         int64_t dy_pitch, int ks, int stride, float* workspace, int64_t workspace_floats,
                         int max_splits);
int yb_wgrad_plan_run(void* plan, float* dw, int out_rows, const int* index_map, int accumulate, void* stream);
/* the same in two halves, so that a caller can order them across streams: phase 1 = the split-K tensor-core kernel (writes the
 * plan's workspace), phase 2 = the fixed-order reduction of the partials into dw (reads the workspace), phase 0 = both.
 * The workspace may be overwritten by another plan only after phase 2 has completed. */
int yb_wgrad_plan_run_phase(void* plan, float* dw, int out_rows, const int* index_map, int accumulate, int phase, void* stream);
int yb_conv2d_wgrad(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* dy, int Cout,
                    int64_t dy_pitch, int ks, int stride, float* workspace, int64_t workspace_floats, int max_splits,
                    float* dw, int out_rows, const int* index_map, int accumulate, void* stream);

/* ---- BatchNorm2d(eps, momentum) + SiLU around the convs (model.py:17,23; residual add model.py:50;
 *      nearest 2x upsample model.py:225 fused as a second store) -----------------------------------------------
 * yb_bn_finalize: training=1: batch mean / biased variance from the conv's stats partials ([rows][2][C]), updates
 *   running_mean/var (unbiased variance, momentum) and num_batches_tracked when given; training=0: uses the running
 *   statistics.  Writes scale = gamma*invstd, shift = beta - mean*scale (and mean, invstd when non-NULL).
 * yb_bn_act_fwd:   out = SiLU(y*scale+shift) (+ res);  out_up (optional) receives the 2x nearest-upsampled copy.
 * backward (two passes): reduce -> per-channel sums of dz = da*SiLU'(z) and dz*xhat as [rows][2][C] partials;
 *   finalize -> dgamma, dbeta (fp32, optional accumulate) and coef[2][C] = sums/count; apply -> dy (bf16). */
int yb_bn_finalize(const float* stats, int rows, int C, double count, const float* gamma, const float* beta, float eps,
                   float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, float* scale,
                   float* shift, float* mean, float* invstd, int training, void* stream);
/* same, for a layer whose statistics are a channel slice of a wider fused convolution (C3's c1 || c_skipped run as one GEMM,
 * model.py:90-92): `stats` points at the slice's first channel, a partial row holds 2 * stats_ld floats */
int yb_bn_finalize_ld(const float* stats, int rows, int C, int stats_ld, double count, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* mean, float* invstd, int training, void* stream);
/* inference (model.eval(), model.py:17 with running statistics): the training=0 form of yb_bn_finalize for ALL the
 * BatchNorms of the network in one launch.  items (device): n rows of 8 x int64 = {gamma, beta, running_mean, running_var,
 * scale, shift (device pointers), C, eps (float bits in the low 32)} */
int yb_bn_fold_batch(const void* items, int n, void* stream);
int yb_bn_act_fwd(const void* y, int64_t y_pitch, int N, int H, int W, int C, const float* scale, const float* shift,
                  const void* res, int64_t res_pitch, void* out, int64_t out_pitch, void* out_up, int64_t up_pitch,
                  void* stream);
int yb_bwd_reduce_max_rows(void);
int yb_bn_act_bwd_reduce(const void* da, int64_t da_pitch, const void* y, int64_t y_pitch, int64_t npix, int C,
                         const float* scale, const float* shift, const float* mean, const float* invstd, float* partial,
                         int* rows, void* stream);
int yb_bn_bwd_finalize(const float* partial, int rows, int C, double count, float* dgamma, float* dbeta, float* coef,
                       int accumulate, void* stream);
int yb_bn_act_bwd_apply(const void* da, int64_t da_pitch, const void* y, int64_t y_pitch, int64_t npix, int C,
                        const float* scale, const float* shift, const float* mean, const float* invstd,
                        const float* coef, void* dy, int64_t dy_pitch, void* stream);
/* per-channel column sums of a bf16 NHWC tensor as [rows][2][C] partials (row 0 of each pair); head bias gradient */
int yb_colsum(const void* x, int64_t x_pitch, int64_t npix, int C, float* partial, int* rows, void* stream);
int yb_reduce_rows(const float* partial, int rows, int64_t stride, int n, float* out, int accumulate, void* stream);

/* ---- glue ops of YOLOV5m.forward (model.py:210-239) and their backward ----------------------------------------- */
int yb_upsample2x_fwd(const void* src, int64_t src_pitch, int N, int H, int W, int C, void* dst, int64_t dst_pitch,
                      void* stream);                                   /* Resize(NEAREST), model.py:225 */
int yb_upsample2x_bwd(const void* dup, int64_t dup_pitch, int N, int H, int W, int C, void* dsrc, int64_t dsrc_pitch,
                      int accumulate, void* stream);
int yb_add_into(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t npix, int C, int accumulate,
                void* stream);                                         /* gradient fan-in (residual, model.py:50) */
int yb_maxpool5_fwd(const void* x, int64_t x_pitch, int N, int H, int W, int C, void* y, int64_t y_pitch,
                    uint8_t* argmax, void* stream);                    /* nn.MaxPool2d(5,1,2), model.py:103 */
int yb_maxpool5_bwd(const void* dy, int64_t dy_pitch, const uint8_t* argmax, int N, int H, int W, int C, void* dx,
                    int64_t dx_pitch, int accumulate, void* stream);
/* SPPF (model.py:96-112) pools its input three times in a chain: y1 = p(x), y2 = p(y1), y3 = p(y2), p = MaxPool2d(5,1,2).
 * One launch: the H x W x 16-channel tile lives in shared memory for the three pools (separable row / column maximum with the
 * reference's "first maximum wins" order).  y1..y3 share y_pitch (the three channel slices of the concat buffer); am1..am3
 * (optional, all or none) receive the arg-max window slots for the backward pass.  Returns 1 (nothing launched) when the map
 * does not fit one CTA's shared memory: chain yb_maxpool5_fwd then.  yb_sppf_pool3_bwd is the whole backward chain:
 * g0 (+)= scatter(g1 + scatter(g2 + scatter(g3, am3), am2), am1), g0..g3 = the gradient slices [x | y1 | y2 | y3]. */
int yb_sppf_pool3_fwd(const void* x, int64_t x_pitch, int N, int H, int W, int C, void* y1, void* y2, void* y3, int64_t y_pitch,
                      uint8_t* am1, uint8_t* am2, uint8_t* am3, void* stream);
int yb_sppf_pool3_bwd(const void* g1, const void* g2, const void* g3, int64_t g_pitch, const uint8_t* am1, const uint8_t* am2,
                      const uint8_t* am3, int N, int H, int W, int C, void* g0, int64_t g0_pitch, int accumulate, void* stream);
/* x (N,3,H,W) NCHW, dtype 0 = float32 in [0,1], 1 = uint8 (divided by 255, training_utils.py:98)
 * -> out (N,H/2,W/2+2,16) bf16: space-to-depth (12 -> 16 channels: (r*2+s)*3+c = x[c][2h+r][2w+s]) at column w+1 of rows
 * that carry one zero pixel on either side.  The three horizontal taps of the stem are then the 48 contiguous values
 * starting at padded column w (channel kw*16+j = s2d pixel w+kw-1, zero outside), so the 6x6/s2 stem (model.py:184) is a
 * 3x1 convolution (ks code 31 of yb_conv_fwd_plan / yb_conv_wgrad_plan, Cin = 48, x_pitch = 16) with weights
 * [Cout][3][48] = yb_repack_stem, and the 48-channel tap-gathered tensor is never written. */
int yb_prep_input(const void* x, int dtype, int N, int H, int W, void* out, void* stream);
/* same staging, but the (N,3,Hs,Ws) image is first resampled to (H,W) like the reference's multi_scale()
 * (utils/training_utils.py:11-28: F.interpolate(img, size=(H,W), mode="bilinear", align_corners=False) of the float image);
 * the resized float image is never materialised */
int yb_prep_input_resized(const void* x, int dtype, int N, int Hs, int Ws, int H, int W, void* out, void* stream);
/* dense gradient of a head output (B,na,H,W,no) fp32 -> bf16 NHWC (B,H,W,Cpad), channel a*no+o (model.py:173 backward);
 * accumulate = 1 adds to what dy already holds (a second gradient contribution to the same head output) */
int yb_head_grad_pack(const float* g, int B, int na, int H, int W, int no, void* dy, int Cpad, int accumulate, void* stream);

/* ---- fp32 parity mode (BASELINE.json configs[1]: "fp32 vs reference", 1e-3) --------------------------
 * Activations and gradients stay fp32; a convolution is six passes of the SAME tcgen05 kernels above over the 3-way bf16
 * split of both operands (x = x0+x1+x2, 24 mantissa bits), accumulating into one fp32 tensor (out_kind 2 then 3; wgrad
 * accumulate = 1).  "planes" = three bf16 tensors of the fp32 tensor's geometry, plane_stride elements apart.  The
 * element-wise operators of model.py:17,23,50,103,225 and their backward run as plain fp32 kernels (exact SiLU, double
 * accumulation of the BatchNorm sums).  partial / rows feed yb_bn_finalize, yb_bn_bwd_finalize and yb_reduce_rows. */
int yb_p32_split_flat(const void* src, int dtype, int64_t n, float* o0, float* o1, float* o2, void* stream);
int yb_p32_planes(const float* src, int64_t src_pitch, int64_t npix, int C, void* planes, int64_t pl_pitch,
                  int64_t plane_stride, void* stream);
int yb_p32_bn_stats(const float* y, int64_t y_pitch, int64_t npix, int C, float* partial, int max_rows, int* rows,
                    void* stream);
int yb_p32_bn_act_fwd(const float* y, int64_t y_pitch, int N, int H, int W, int C, const float* scale, const float* shift,
                      const float* res, int64_t res_pitch, float* out, void* out_planes, int64_t out_pitch,
                      int64_t out_plane_stride, float* up, void* up_planes, int64_t up_pitch, int64_t up_plane_stride,
                      void* stream);
int yb_p32_bn_act_bwd_reduce(const float* da, int64_t da_pitch, const float* y, int64_t y_pitch, int64_t npix, int C,
                             const float* scale, const float* shift, const float* mean, const float* invstd,
                             float* partial, int max_rows, int* rows, void* stream);
int yb_p32_bn_act_bwd_apply(const float* da, int64_t da_pitch, const float* y, int64_t y_pitch, int64_t npix, int C,
                            const float* scale, const float* shift, const float* mean, const float* invstd,
                            const float* coef, float* dy, void* dy_planes, int64_t dy_pitch, int64_t dy_plane_stride,
                            void* stream);
int yb_p32_maxpool5_fwd(const float* x, int64_t x_pitch, int N, int H, int W, int C, float* y, void* y_planes,
                        int64_t y_pitch, int64_t y_plane_stride, uint8_t* argmax, void* stream);
int yb_p32_maxpool5_bwd(const float* dy, int64_t dy_pitch, const uint8_t* argmax, int N, int H, int W, int C, float* dx,
                        int64_t dx_pitch, int accumulate, void* stream);
int yb_p32_upsample2x_bwd(const float* dup, int64_t dup_pitch, int N, int H, int W, int C, float* dsrc, int64_t dsrc_pitch,
                          int accumulate, void* stream);
int yb_p32_add_into(const float* src, int64_t src_pitch, float* dst, int64_t dst_pitch, int64_t npix, int C, int accumulate,
                    void* stream);
int yb_p32_head_grad_pack(const float* g, int B, int na, int H, int W, int no, float* dy, void* dy_planes, int Cpad,
                          int64_t plane_stride, int accumulate, void* stream);

/* ---- parameter packing + optimiser tail on the flat fp32 master / gradient buffers ------------------------------
 * The master weights stay in the reference's state_dict tensors (fp32, conv weights channels-last = [Cout][kh*kw][Cin]),
 * laid out back to back in one flat buffer; the tcgen05 operands are bf16 copies:
 *   yb_cast_bf16     forward operands: element-wise cast of the whole flat buffer (same offsets)
 *   yb_repack_dgrad  dgrad operands [Cin][tap][Cout_pad]; table = int64 [nlayers][8]
 *                    {src_off, dst_off, Cout, taps, Cin, Cout_pad, cum_begin, cum_end} (offsets in elements)
 *   yb_repack_stem   6x6/s2 stem [Cout][6][6][3] -> space-to-depth 3x3 [Cout][9][16]
 * yb_grad_norm / yb_adam_step replace scaler.unscale_ + clip_grad_norm_(max_norm) + Adam.step of
 * utils/training_utils.py:114-122 (train.py:61: Adam(lr, weight_decay) = L2 added to the gradient, all parameters):
 *   norm_out[0] = grad_scale * ||g||_2 ;  clip = min(1, max_norm/(norm+1e-6)) (max_norm <= 0: no clipping)
 *   step counts from 1 (host value), or is read from step_dev when non-NULL (CUDA-graph replay; yb_counter_inc). */
int yb_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
int yb_repack_dgrad(const float* src, void* dst, const int64_t* table, int nlayers, int64_t total, void* stream);
int yb_repack_stem(const float* w6, void* w3, int Cout, void* stream);
int yb_grad_norm(const float* g, int64_t n, float grad_scale, float* partial, int partial_len, float* norm_out,
                 void* stream);
/* counter += 1 unless norm (optional, device) is inf / NaN: a step with non-finite gradients is skipped like
 * GradScaler.step (training_utils.py:119); yb_adam_step given the same norm leaves p / m / v / w_bf16 untouched then */
int yb_counter_inc(int64_t* counter, const float* norm, void* stream);
/* dst += src over a flat fp32 bucket: gradient accumulation over micro-batches (training_utils.py:88-90,:116) */
int yb_accumulate_f32(float* dst, const float* src, int64_t n, void* stream);
int yb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int64_t step, const int64_t* step_dev, float grad_scale, float max_norm,
                 const float* norm, void* w_bf16, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (BASELINE.json configs[3]; the reference is single-GPU) ------
 * peer_ptrs_host: HOST array of `world` device pointers, entry r = base of rank r's fp32 gradient bucket (symmetric memory,
 * peer-mapped, 16-byte aligned, n floats each, n % 4 == 0).  Rank `rank` loads its 1/world slice from every peer, sums in
 * rank order (bit-identical on every rank) and stores the sum into that slice of every peer's bucket: reduce-scatter and
 * all-gather in one kernel.  The caller orders it across ranks with a barrier before (all buckets complete) and after
 * (all stores landed).  ctas <= 0: default grid. */
int yb_allreduce_p2p(const uint64_t* peer_ptrs_host, int rank, int world, int64_t n, int ctas, void* stream);

/* ---- ComputeLoss (ultralytics_loss.py:17-311) ---------------------------------------------------------------------
 * One yb_loss_level per detection level; all pointers are device pointers owned by the caller, `cap` = row capacity of
 * the per-level arrays (>= 5*na*nt, the exact upper bound of build_targets). */
typedef struct yb_loss_level {
  const float* p;       /* (B,na,H,W,no) fp32 raw logits of this level (model.py:173 layout) */
  int32_t H, W;
  float balance;        /* ultralytics_loss.py:37 */
  int32_t pad_;
  int64_t* idx;         /* [4][cap]: image b, anchor a, grid row gj, grid column gi   (ultralytics_loss.py:285) */
  float* tbox;          /* [cap][4]: gx-gi, gy-gj, gw, gh in grid units               (:296) */
  float* anch;          /* [cap][2]: matched anchor                                   (:301) */
  int64_t* tcls;        /* [cap]:    class index                                      (:306) */
  float* row_val;       /* [3][cap]: GIoU.clamp(0), 1-GIoU, sum_c BCE(cls)            (scratch) */
  float* row_grad;      /* [cap][no]: unit-scale d(row loss)/d(logits)                (scratch) */
  int32_t* row_prev;    /* [cap]: previous row that hit the same cell (1-based, 0 = none) */
  int32_t* cell_head;   /* [B*na*H*W]: last-linked row of each cell (1-based, 0 = none) */
  float* obj_partial;   /* [yb_loss_obj_rows()] per-CTA objectness BCE partial sums */
  float* grad_f32;      /* optional out: dL/dp, (B,na,H,W,no) fp32 */
  void* grad_bf16;      /* optional out: dL/dp as the head conv's gradient operand, (B,H,W,cpad) bf16, channel a*no+o */
} yb_loss_level;

int yb_loss_obj_rows(void);
/* intersection_over_union (utils/bboxes_utils.py:33-87): (n,4) x (n,4) fp32 -> (n,) IoU or GIoU; midpoint = xywh boxes */
int yb_box_iou(const float* boxes_preds, const float* boxes_labels, int64_t n, int midpoint, int giou, float eps, float* out,
               void* stream);
/* build_targets (ultralytics_loss.py:122-311): targets (nt,6) fp32 [img,cls,x,y,w,h] normalised; anchors [nl][na][2]
 * (stride-divided, model.py:156-157).  Writes idx/tbox/anch/tcls of every level in the reference's row order
 * (offset-major, then anchor, then target -- the order boolean-mask indexing produces) and counts[nl]. */
int yb_build_targets(const float* targets, int nt, const float* anchors, const yb_loss_level* levels, int nl, int na,
                     float anchor_t, int64_t cap, int* counts, void* stream);
/* ComputeLoss.__call__ (ultralytics_loss.py:60-120): out4 = [(lbox+lobj+lcls)*B, lbox, lobj, lcls] (weighted parts).
 * counts[nl] = row slots to scan per level.  Rows with image index < 0 are unused slots; rows with tcls < 0 are "ignore"
 * rows (objectness target -1, no box / class term: YOLO_LOSS, loss.py:190).  nobj (optional, device [nl]) = number of
 * object rows = denominator of the box / class means (NULL: counts).  nan_on_empty = 1 reproduces the NaN the reference's
 * YOLO_LOSS returns for a level without objects (mean of an empty tensor, loss.py:211). */
int yb_loss_fwd(const yb_loss_level* levels, int nl, int B, int na, int no, int64_t cap, const int* counts,
                const int* nobj, int nan_on_empty, float lam_box, float lam_obj, float lam_cls, float* out4, void* stream);
/* its backward: gout = device scalar dLoss (NULL = 1).  Needs the scratch yb_loss_fwd left in the levels. */
int yb_loss_bwd(const yb_loss_level* levels, int nl, int B, int na, int no, int64_t cap, const int* counts,
                const int* nobj, float lam_box, float lam_obj, float lam_cls, const float* gout, int cpad, void* stream);
/* YOLO_LOSS.build_targets (loss.py:101-192), the matcher train.py uses without --ultralytics_loss: sequential per image.
 * labels: device float64 [nt][5] = class, x, y, w, h (normalised; np.loadtxt rows); offsets: device int32 [B+1].
 * anchor_table: device fp32 [T][9][2], row k = the reference's anchor tensor after k in-place divisions by 640
 * (utils/bboxes_utils.py:18), times the level stride; the label row r (0-based in this batch) uses table row
 * min(decay_base + (r+1)*decay_stride, T-1), or row 1 when decay_stride = 0 (anchors normalised once = the bug fixed).
 * Writes the row lists of the three levels (3 slots per label row; unused slots get image index -1), counts[3] = 3*nt,
 * nobj[3]; state: int8 scratch of sum_l B*3*H*W bytes; dense (optional): the reference's dense target tensors, level after
 * level, each (B,3,H,W,6) fp32 = [x_cell, y_cell, w_cell, h_cell, objectness (1 / -1 / 0), class]. */
int yb_yolo_build_targets(const double* labels, const int* offsets, int B, int nt, const float* anchor_table, int T,
                          int64_t decay_base, int decay_stride, const float* head_anchors, const yb_loss_level* levels,
                          int nl, int na, float ignore_thr, int64_t cap, int8_t* state, float* dense, int* counts,
                          int* nobj, void* stream);

/* ---- cells_to_bboxes(is_pred=True) (utils/plot_utils.py:10-40) and non_max_suppression (utils/bboxes_utils.py:175-209)
 * yb_decode_level: p (B,na,H,W,no) logits of one level -> rows [cls, sigmoid(obj), cx, cy, w, h] (pixels) written at
 *   out[(b*rows_per_image + level_off + (a*H+y)*W+x)*6]; anchors_px = anchors[level]*stride, [na][2].  is_pred = 0:
 *   the target-tensor branch (plot_utils.py:29-34; no = 6, no sigmoid, class id in channel 5).
 * yb_nms_batched: boxes (B,N,6) rows as above.  Per image: keep score > threshold, xywh->xyxy, +cls offset, stable
 *   descending sort, greedy suppression (IoU > iou_threshold), first max_det survivors.  out [B][max_det][6] rows
 *   [cls, score, x1, y1, x2, y2]; out_count [B]; out_index (optional) [B][max_det] = row index into N; cand_count
 *   (optional) [B] = candidates above threshold.  scratch: yb_nms_scratch_bytes(B,N) bytes. */
int yb_decode_level(const float* p, int B, int na, int H, int W, int no, float stride, const float* anchors_px, int is_pred,
                    float* out, int64_t rows_per_image, int64_t level_off, void* stream);
int64_t yb_nms_scratch_bytes(int B, int64_t N);
/* YOLO_EVAL.check_class_accuracy for one level (utils/validation_utils.py:58-69): p (cells, no) logits, y (cells, ny) label
 * rows [x, y, w, h, objectness, class]; over the cells with y[4] == 1: counters[0] += 1, counters[1] += (argmax(p[5:]) ==
 * y[5]), counters[2] += (sigmoid(p[0]) > conf) -- channel 0, as the reference has it.  counters: device uint64 [3]. */
int yb_class_accuracy(const float* p, const float* y, int64_t cells, int no, int ny, float conf, uint64_t* counters,
                      void* stream);
int yb_nms_batched(const float* boxes, int B, int64_t N, float iou_threshold, float threshold, int max_det, void* scratch,
                   float* out, int* out_count, int* out_index, int* cand_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif
