// fp32 parity mode of the network (BASELINE.json configs[1]: "fp32 vs reference", 1e-3).
//
// The reference computes in fp32 (model.py:210-239 under plain PyTorch); the production path here stores activations in
// bf16.  Parity mode keeps every activation / gradient in fp32 and runs each convolution as SIX passes of the SAME tcgen05
// kernels (conv_igemm.cu / conv_patch.cu / conv_wgrad*.cu) over a 3-way bf16 split of both operands,
//     x = x0 + x1 + x2,   x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)          (24 mantissa bits in total)
//     x * w ~= x0 w0 + (x0 w1 + x1 w0) + (x1 w1 + x0 w2 + x2 w0)                           (dropped terms <= 2^-24 |x w|)
// each pass accumulating in fp32 (TMEM) and into the fp32 output (OUT_F32_ACC).  bf16 x bf16 products are exact in fp32, so
// the result carries fp32-level error: the tensor-core kernels themselves are what the parity tests exercise.
// The element-wise operators around the convolutions are the plain fp32 kernels below (exact expf-based SiLU instead of
// tanh.approx, double accumulation of the BatchNorm statistics).  Nothing here is tuned: parity mode is a numerics
// instrument (YB_PARITY=1 / model.parity = True), not the throughput path.
#include "../../include/yolov5m_b200.h"
#include "common.cuh"

#include <algorithm>

namespace yb {

static inline int p32_blocks(long n) { return (int)std::max<long>(1, std::min<long>((n + 255) / 256, 148L * 16)); }

__device__ __forceinline__ void split3(float x, float& a, float& b, float& c) {
  a = __bfloat162float(__float2bfloat16_rn(x));
  const float r1 = x - a;  // exact (Sterbenz-like: a is x rounded to 8 bits)
  b = __bfloat162float(__float2bfloat16_rn(r1));
  c = __bfloat162float(__float2bfloat16_rn(r1 - b));
}

// flat split into three fp32 arrays whose values are bf16-representable (weights; the NCHW input image, dtype 1 = uint8/255)
__global__ void p32_split_flat_kernel(const void* __restrict__ src, int dtype, long n, float* __restrict__ o0,
                                      float* __restrict__ o1, float* __restrict__ o2) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float x = dtype == 0 ? reinterpret_cast<const float*>(src)[i]
                               : (float)reinterpret_cast<const uint8_t*>(src)[i] / 255.f;  // training_utils.py:98
    float a, b, c;
    split3(x, a, b, c);
    o0[i] = a; o1[i] = b; o2[i] = c;
  }
}

__device__ __forceinline__ void store_planes(bf16* pl, long off, long ps, float x) {
  float a, b, c;
  split3(x, a, b, c);
  pl[off] = __float2bfloat16_rn(a);
  pl[off + ps] = __float2bfloat16_rn(b);
  pl[off + 2 * ps] = __float2bfloat16_rn(c);
}

// NHWC fp32 view -> three bf16 planes of the same geometry
__global__ void p32_planes_kernel(const float* __restrict__ src, long src_pitch, long npix, int C, bf16* __restrict__ pl,
                                  long pl_pitch, long ps) {
  const long total = npix * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / C;
    const int c = (int)(i - pix * C);
    store_planes(pl, pix * pl_pitch + c, ps, src[pix * src_pitch + c]);
  }
}

// per-channel sum / sum of squares (mode 0) of an fp32 NHWC view; block = 32 channels x 8 pixel lanes; grid = (C/32, R).
// partial[R][2][C] (fp32 rows, accumulated in double) feeds yb_bn_finalize / yb_reduce_rows.
// mode 1: BatchNorm+SiLU backward reduction: sums of dz = da * SiLU'(y*scale+shift) and dz * xhat, xhat = (y-mean)*invstd
__global__ void __launch_bounds__(256) p32_colstats_kernel(const float* __restrict__ a, long a_pitch, const float* __restrict__ y,
                                                           long y_pitch, long npix, int C, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, float* __restrict__ partial,
                                                           int mode) {
  __shared__ double sh[2][8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0, q = 0.0;
  if (c < C) {
    float sc = 0.f, sh_ = 0.f, mu = 0.f, is = 0.f;
    if (mode == 1) { sc = scale[c]; sh_ = shift[c]; mu = mean[c]; is = invstd[c]; }
    for (long p = (long)blockIdx.y * 8 + ry; p < npix; p += (long)gridDim.y * 8) {
      const float v = a[p * a_pitch + c];
      if (mode == 0) {
        s += (double)v;
        q += (double)v * (double)v;
      } else {
        const float yv = y[p * y_pitch + c];
        const float z = fmaf(yv, sc, sh_);
        const float sg = 1.f / (1.f + expf(-z));
        const float dz = v * (sg * (1.f + z * (1.f - sg)));
        s += (double)dz;
        q += (double)dz * (double)((yv - mu) * is);
      }
    }
  }
  sh[0][ry][cx] = s;
  sh[1][ry][cx] = q;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int r = 1; r < 8; ++r) { s += sh[0][r][cx]; q += sh[1][r][cx]; }
    partial[(size_t)blockIdx.y * 2 * C + c] = (float)s;
    partial[(size_t)blockIdx.y * 2 * C + C + c] = (float)q;
  }
}

// out = SiLU(y*scale+shift) (+ res), written as fp32 and as bf16 planes; optional nearest 2x upsampled copy (model.py:225)
__global__ void p32_bn_act_fwd_kernel(const float* __restrict__ y, long y_pitch, int H, int W, int C, long npix,
                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                      const float* __restrict__ res, long res_pitch, float* __restrict__ out,
                                      bf16* __restrict__ out_pl, long out_pitch, long out_ps, float* __restrict__ up,
                                      bf16* __restrict__ up_pl, long up_pitch, long up_ps) {
  const long total = npix * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / C;
    const int c = (int)(i - pix * C);
    const float z = fmaf(y[pix * y_pitch + c], scale[c], shift[c]);
    float v = z / (1.f + expf(-z));
    if (res != nullptr) v += res[pix * res_pitch + c];
    out[pix * out_pitch + c] = v;
    if (out_pl != nullptr) store_planes(out_pl, pix * out_pitch + c, out_ps, v);
    if (up != nullptr) {
      const long n = pix / ((long)H * W);
      const long r = pix - n * (long)H * W;
      const int h = (int)(r / W), w = (int)(r - (long)h * W);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const long upix = (n * 2 * H + 2 * h + (d >> 1)) * 2 * W + 2 * w + (d & 1);
        up[upix * up_pitch + c] = v;
        if (up_pl != nullptr) store_planes(up_pl, upix * up_pitch + c, up_ps, v);
      }
    }
  }
}

// dy = scale * (dz - coef0 - xhat * coef1) written as bf16 planes (the dgrad / wgrad operands) and optionally as fp32
__global__ void p32_bn_act_bwd_apply_kernel(const float* __restrict__ da, long da_pitch, const float* __restrict__ y,
                                            long y_pitch, long npix, int C, const float* __restrict__ scale,
                                            const float* __restrict__ shift, const float* __restrict__ mean,
                                            const float* __restrict__ invstd, const float* __restrict__ coef,
                                            float* __restrict__ dy, bf16* __restrict__ dy_pl, long dy_pitch, long dy_ps) {
  const long total = npix * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / C;
    const int c = (int)(i - pix * C);
    const float yv = y[pix * y_pitch + c];
    const float z = fmaf(yv, scale[c], shift[c]);
    const float sg = 1.f / (1.f + expf(-z));
    const float dz = da[pix * da_pitch + c] * (sg * (1.f + z * (1.f - sg)));
    const float xhat = (yv - mean[c]) * invstd[c];
    const float v = scale[c] * (dz - coef[c] - xhat * coef[C + c]);
    if (dy != nullptr) dy[pix * dy_pitch + c] = v;
    store_planes(dy_pl, pix * dy_pitch + c, dy_ps, v);
  }
}

// nn.MaxPool2d(5, 1, 2) (model.py:103): first maximum in row-major window order wins (ATen max_pool2d), argmax = window slot
__global__ void p32_maxpool5_fwd_kernel(const float* __restrict__ x, long x_pitch, int H, int W, int C, long npix,
                                        float* __restrict__ yo, bf16* __restrict__ y_pl, long y_pitch, long y_ps,
                                        uint8_t* __restrict__ argmax) {
  const long total = npix * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / C;
    const int c = (int)(i - pix * C);
    const long n = pix / ((long)H * W);
    const long r = pix - n * (long)H * W;
    const int h = (int)(r / W), w = (int)(r - (long)h * W);
    float best = -INFINITY;
    int slot = 12;
    bool first = true;
    for (int dh = -2; dh <= 2; ++dh) {
      const int hh = h + dh;
      if (hh < 0 || hh >= H) continue;
      for (int dw = -2; dw <= 2; ++dw) {
        const int ww = w + dw;
        if (ww < 0 || ww >= W) continue;
        const float v = x[((n * H + hh) * W + ww) * x_pitch + c];
        if (first || v > best || v != v) {
          best = v;
          slot = (dh + 2) * 5 + dw + 2;
          first = false;
        }
      }
    }
    yo[pix * y_pitch + c] = best;
    if (y_pl != nullptr) store_planes(y_pl, pix * y_pitch + c, y_ps, best);
    if (argmax != nullptr) argmax[pix * C + c] = (uint8_t)slot;
  }
}

// gather form of the scatter: dx[h,w] = sum over the <= 25 outputs whose window holds (h,w) and whose argmax points at it
__global__ void p32_maxpool5_bwd_kernel(const float* __restrict__ dy, long dy_pitch, const uint8_t* __restrict__ argmax, int H,
                                        int W, int C, long npix, float* __restrict__ dx, long dx_pitch, int accumulate) {
  const long total = npix * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / C;
    const int c = (int)(i - pix * C);
    const long n = pix / ((long)H * W);
    const long r = pix - n * (long)H * W;
    const int h = (int)(r / W), w = (int)(r - (long)h * W);
    float s = 0.f;
    for (int dh = -2; dh <= 2; ++dh) {
      const int ho = h - dh;  // output row whose window slot (dh, dw) is this input pixel
      if (ho < 0 || ho >= H) continue;
      for (int dw = -2; dw <= 2; ++dw) {
        const int wo = w - dw;
        if (wo < 0 || wo >= W) continue;
        const long opix = (n * H + ho) * W + wo;
        if (argmax[opix * C + c] == (uint8_t)((dh + 2) * 5 + dw + 2)) s += dy[opix * dy_pitch + c];
      }
    }
    float* d = dx + pix * dx_pitch + c;
    *d = accumulate ? *d + s : s;
  }
}

// backward of the nearest 2x upsample: dsrc[h,w] (+)= sum of the four dup pixels
__global__ void p32_upsample2x_bwd_kernel(const float* __restrict__ dup, long dup_pitch, int H, int W, int C, long npix,
                                          float* __restrict__ dsrc, long dsrc_pitch, int accumulate) {
  const long total = npix * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / C;
    const int c = (int)(i - pix * C);
    const long n = pix / ((long)H * W);
    const long r = pix - n * (long)H * W;
    const int h = (int)(r / W), w = (int)(r - (long)h * W);
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) s += dup[((n * 2 * H + 2 * h + (d >> 1)) * 2 * W + 2 * w + (d & 1)) * dup_pitch + c];
    float* o = dsrc + pix * dsrc_pitch + c;
    *o = accumulate ? *o + s : s;
  }
}

__global__ void p32_add_into_kernel(const float* __restrict__ src, long src_pitch, float* __restrict__ dst, long dst_pitch,
                                    long npix, int C, int accumulate) {
  const long total = npix * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / C;
    const int c = (int)(i - pix * C);
    const float v = src[pix * src_pitch + c];
    float* d = dst + pix * dst_pitch + c;
    *d = accumulate ? *d + v : v;
  }
}

// g (B,na,H,W,no) fp32 -> dy (B,H,W,Cpad) fp32 + bf16 planes, channel a*no+o; channels >= na*no are zero
__global__ void p32_head_grad_pack_kernel(const float* __restrict__ g, int na, long hw, int no, float* __restrict__ dy,
                                          bf16* __restrict__ dy_pl, int Cpad, long ps, long total, int accumulate) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long t = i / Cpad;
    const long pix = t % hw, b = t / hw;
    float v = 0.f;
    if (c < na * no) {
      const int a = c / no, o = c - a * no;
      v = g[((b * na + a) * hw + pix) * no + o];
    }
    if (accumulate) v += dy[i];
    dy[i] = v;
    store_planes(dy_pl, i, ps, v);
  }
}

}  // namespace yb

using namespace yb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define B16(p) reinterpret_cast<bf16*>(p)
#define F32(p) reinterpret_cast<float*>(p)
#define CF32(p) reinterpret_cast<const float*>(p)

extern "C" {

int yb_p32_split_flat(const void* src, int dtype, int64_t n, float* o0, float* o1, float* o2, void* stream) {
  YB_REQUIRE(dtype == 0 || dtype == 1, "p32_split_flat: dtype %d (0 = float32, 1 = uint8)", dtype);
  p32_split_flat_kernel<<<p32_blocks(n), 256, 0, ST(stream)>>>(src, dtype, n, o0, o1, o2);
  YB_LAUNCHED();
  return 0;
}

int yb_p32_planes(const float* src, int64_t src_pitch, int64_t npix, int C, void* planes, int64_t pl_pitch,
                  int64_t plane_stride, void* stream) {
  p32_planes_kernel<<<p32_blocks(npix * C), 256, 0, ST(stream)>>>(src, src_pitch, npix, C, B16(planes), pl_pitch, plane_stride);
  YB_LAUNCHED();
  return 0;
}

static int colstats_rows(long npix, int C, int max_rows) {
  const int cb = (C + 31) / 32;
  long r = std::max<long>(1, (148L * 4) / cb);
  r = std::min<long>(r, (npix + 7) / 8);
  return (int)std::max<long>(1, std::min<long>(r, max_rows));
}

int yb_p32_bn_stats(const float* y, int64_t y_pitch, int64_t npix, int C, float* partial, int max_rows, int* rows,
                    void* stream) {
  const int R = colstats_rows(npix, C, max_rows);
  p32_colstats_kernel<<<dim3((C + 31) / 32, R), 256, 0, ST(stream)>>>(y, y_pitch, nullptr, 0, npix, C, nullptr, nullptr,
                                                                      nullptr, nullptr, partial, 0);
  YB_LAUNCHED();
  if (rows) *rows = R;
  return 0;
}

int yb_p32_bn_act_fwd(const float* y, int64_t y_pitch, int N, int H, int W, int C, const float* scale, const float* shift,
                      const float* res, int64_t res_pitch, float* out, void* out_planes, int64_t out_pitch,
                      int64_t out_plane_stride, float* up, void* up_planes, int64_t up_pitch, int64_t up_plane_stride,
                      void* stream) {
  const long npix = (long)N * H * W;
  p32_bn_act_fwd_kernel<<<p32_blocks(npix * C), 256, 0, ST(stream)>>>(y, y_pitch, H, W, C, npix, scale, shift, res, res_pitch,
                                                                      out, B16(out_planes), out_pitch, out_plane_stride, up,
                                                                      B16(up_planes), up_pitch, up_plane_stride);
  YB_LAUNCHED();
  return 0;
}

int yb_p32_bn_act_bwd_reduce(const float* da, int64_t da_pitch, const float* y, int64_t y_pitch, int64_t npix, int C,
                             const float* scale, const float* shift, const float* mean, const float* invstd,
                             float* partial, int max_rows, int* rows, void* stream) {
  const int R = colstats_rows(npix, C, max_rows);
  p32_colstats_kernel<<<dim3((C + 31) / 32, R), 256, 0, ST(stream)>>>(da, da_pitch, y, y_pitch, npix, C, scale, shift, mean,
                                                                      invstd, partial, 1);
  YB_LAUNCHED();
  if (rows) *rows = R;
  return 0;
}

int yb_p32_bn_act_bwd_apply(const float* da, int64_t da_pitch, const float* y, int64_t y_pitch, int64_t npix, int C,
                            const float* scale, const float* shift, const float* mean, const float* invstd,
                            const float* coef, float* dy, void* dy_planes, int64_t dy_pitch, int64_t dy_plane_stride,
                            void* stream) {
  p32_bn_act_bwd_apply_kernel<<<p32_blocks(npix * C), 256, 0, ST(stream)>>>(da, da_pitch, y, y_pitch, npix, C, scale, shift,
                                                                            mean, invstd, coef, dy, B16(dy_planes), dy_pitch,
                                                                            dy_plane_stride);
  YB_LAUNCHED();
  return 0;
}

int yb_p32_maxpool5_fwd(const float* x, int64_t x_pitch, int N, int H, int W, int C, float* y, void* y_planes,
                        int64_t y_pitch, int64_t y_plane_stride, uint8_t* argmax, void* stream) {
  const long npix = (long)N * H * W;
  p32_maxpool5_fwd_kernel<<<p32_blocks(npix * C), 256, 0, ST(stream)>>>(x, x_pitch, H, W, C, npix, y, B16(y_planes), y_pitch,
                                                                        y_plane_stride, argmax);
  YB_LAUNCHED();
  return 0;
}

int yb_p32_maxpool5_bwd(const float* dy, int64_t dy_pitch, const uint8_t* argmax, int N, int H, int W, int C, float* dx,
                        int64_t dx_pitch, int accumulate, void* stream) {
  const long npix = (long)N * H * W;
  p32_maxpool5_bwd_kernel<<<p32_blocks(npix * C), 256, 0, ST(stream)>>>(dy, dy_pitch, argmax, H, W, C, npix, dx, dx_pitch,
                                                                        accumulate);
  YB_LAUNCHED();
  return 0;
}

int yb_p32_upsample2x_bwd(const float* dup, int64_t dup_pitch, int N, int H, int W, int C, float* dsrc, int64_t dsrc_pitch,
                          int accumulate, void* stream) {
  const long npix = (long)N * H * W;
  p32_upsample2x_bwd_kernel<<<p32_blocks(npix * C), 256, 0, ST(stream)>>>(dup, dup_pitch, H, W, C, npix, dsrc, dsrc_pitch,
                                                                          accumulate);
  YB_LAUNCHED();
  return 0;
}

int yb_p32_add_into(const float* src, int64_t src_pitch, float* dst, int64_t dst_pitch, int64_t npix, int C, int accumulate,
                    void* stream) {
  p32_add_into_kernel<<<p32_blocks(npix * C), 256, 0, ST(stream)>>>(src, src_pitch, dst, dst_pitch, npix, C, accumulate);
  YB_LAUNCHED();
  return 0;
}

int yb_p32_head_grad_pack(const float* g, int B, int na, int H, int W, int no, float* dy, void* dy_planes, int Cpad,
                          int64_t plane_stride, int accumulate, void* stream) {
  YB_REQUIRE(Cpad >= na * no, "p32_head_grad_pack: Cpad");
  const long total = (long)B * H * W * Cpad;
  p32_head_grad_pack_kernel<<<p32_blocks(total), 256, 0, ST(stream)>>>(g, na, (long)H * W, no, dy, B16(dy_planes), Cpad,
                                                                       plane_stride, total, accumulate);
  YB_LAUNCHED();
  return 0;
}

}  // extern "C"
