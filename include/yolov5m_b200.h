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
/* number of per-CTA partial rows a conv launch may write to `stats` (= max grid = #SMs) */
int yb_conv_max_partials(void);

/* ---- Conv2d forward (model.py:16 CBL conv; model.py:162 head conv) ----------------
 * y[n,ho,wo,co] = epilogue( sum_{kh,kw,ci} x[n, ho*s+kh-p, wo*s+kw-p, ci] * w[co,(kh*ks+kw)*Cin+ci] )
 * ks in {1,3}, p = ks/2, s in {1,2}.  w_packed: bf16 [Cout][ks*ks*Cin].
 * epilogue: v = acc; if scale: v = v*scale[co]+shift[co]; elif shift: v += shift[co];
 *           if act: v = SiLU(v); if addend: v += addend[n,ho,wo,co]; store.
 * out_kind 0: bf16 NHWC (y_pitch); 1: fp32 head layout (B,na,H,W,no) of model.py:173
 *          (Cout = na*no, need not be a multiple of 16); 2: fp32 NHWC.
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
 * [Cin][ks*ks*Cout].  H, W are the INPUT (dx) spatial dims.  Same epilogue options. */
int yb_conv2d_dgrad(const void* dy, int N, int H, int W, int Cout, int64_t dy_pitch, const void* wt_packed, int Cin,
                    int ks, int stride, void* dx, int64_t dx_pitch, const void* addend, int64_t addend_pitch,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif
