// HBM-bound NHWC bf16 passes around the tcgen05 convolutions (sm_100a):
//   BatchNorm2d batch-stat finalize / apply + SiLU (+ residual, + fused nearest 2x upsample)   model.py:17,23,50,225
//   their backward (two-pass BN backward with the SiLU derivative folded in)                     autograd of the above
//   SPPF 5x5/1/2 max pooling fwd/bwd (argmax kept as a 1-byte window offset)                     model.py:103,108-110
//   input staging (NCHW float/uint8 -> space-to-depth NHWC bf16 for the 6x6/s2 stem)             model.py:184, training_utils.py:98
//   dense head-gradient repack                                                                   model.py:173 (permute backward)
// Every thread moves 16-byte vectors (8 bf16 channels); channel counts are multiples of 8 by construction.
#include "../../include/yolov5m_b200.h"
#include "common.cuh"

#include <algorithm>

namespace yb {

static inline int ew_blocks(long n, int threads = 256, int per_sm = 8) {
  int sms = 148;
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        sms > 0)
      cached = sms;
    else
      cached = 148;
  }
  sms = cached;
  long b = (n + threads - 1) / threads;
  return (int)std::max<long>(1, std::min<long>(b, (long)sms * per_sm));
}

// resident-CTA-exact grid for a grid-stride kernel: #SMs x occupancy (one full wave, no tail), capped by the work
template <typename K>
static int ew_wave_grid(K kernel, int threads, long work_items) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0) != cudaSuccess || occ <= 0) occ = 2;
  return ew_blocks(work_items, threads, occ);
}

struct V8 {
  float v[8];
};
__device__ __forceinline__ V8 ld8(const bf16* p) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  V8 o;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    o.v[2 * j] = f.x;
    o.v[2 * j + 1] = f.y;
  }
  return o;
}
__device__ __forceinline__ void st8(bf16* p, const V8& a) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(a.v[2 * j], a.v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
// 16 channels (32 bytes, 32-byte aligned) in ONE store: a full L2 sector
__device__ __forceinline__ void st16(bf16* p, const V8& a, const V8& b) {
  uint4 r0, r1;
  __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&r0);
  __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&r1);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h0[j] = __floats2bfloat162_rn(a.v[2 * j], a.v[2 * j + 1]);
    h1[j] = __floats2bfloat162_rn(b.v[2 * j], b.v[2 * j + 1]);
  }
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r0.x), "r"(r0.y), "r"(r0.z),
               "r"(r0.w), "r"(r1.x), "r"(r1.y), "r"(r1.z), "r"(r1.w)
               : "memory");
}
__device__ __forceinline__ V8 cvt8(const uint4& r) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  V8 o;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    o.v[2 * j] = f.x;
    o.v[2 * j + 1] = f.y;
  }
  return o;
}
__device__ __forceinline__ uint4 ldraw(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
// streaming read of a tensor the kernel never writes: read-only path, no L1 allocation, 256-byte L2 prefetch granule
__device__ __forceinline__ uint4 ldstream(const bf16* p) {
#ifdef YB_EW_PLAIN_LOADS
  return *reinterpret_cast<const uint4*>(p);
#else
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
#endif
}
__device__ __forceinline__ uint2 ldg_stream8(const uint8_t* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
// SiLU / its derivative through ONE special-function op per element: sigmoid(z) = 0.5 * tanh(z / 2) + 0.5 with
// tanh.approx.f32 (max relative error 2^-11, below the bf16 rounding of every value these passes store).  exp + reciprocal
// would be two MUFU ops per element, and at 16 MUFU results / clock / SM that -- not HBM -- bounded these passes.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float z) { return fmaf(0.5f, tanh_fast(0.5f * z), 0.5f); }
__device__ __forceinline__ float silu_fast(float z) {
  const float hz = 0.5f * z;
  return fmaf(hz, tanh_fast(hz), hz);
}
__device__ __forceinline__ V8 half8(V8 a) {
#pragma unroll
  for (int j = 0; j < 8; ++j) a.v[j] *= 0.5f;
  return a;
}
// SiLU'(z) = (1 + w) / 2  with  w = t + hz * (1 - t^2),  t = tanh(hz),  hz = z / 2:  the backward passes carry 2 * dz =
// g * (1 + w) = fma(g, w, g) and fold the 1/2 into a per-channel constant (three FMAs after the MUFU instead of six ops)
__device__ __forceinline__ float dsilu2_w(float hz) {
  const float t = tanh_fast(hz);
  return fmaf(hz, fmaf(-t, t, 1.f), t);
}
__device__ __forceinline__ V8 ldf8(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  V8 o;
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
  return o;
}

// ------------------------------------------------------------------------------------------------ BN finalize
// block = 32 channels x 32 row-lanes: the [rows][2][C] partials are summed in double (fixed order: row-lane strided, then a
// shared-memory tree) so the result is deterministic; 128-byte coalesced reads.
// ld = channels per partial row (>= C: the partials may belong to a wider fused convolution, `part` then points at the
// first channel of this layer's slice)
__device__ __forceinline__ void colsum2_block(const float* __restrict__ part, int rows, int ld, int c, bool cvalid,
                                              double& s_out, double& q_out) {
  __shared__ double sh_s[32][33], sh_q[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  double s = 0.0, q = 0.0;
  if (cvalid) {
    for (int r = ry; r < rows; r += 32) {
      s += (double)part[(size_t)r * 2 * ld + c];
      q += (double)part[(size_t)r * 2 * ld + ld + c];
    }
  }
  sh_s[ry][cx] = s;
  sh_q[ry][cx] = q;
  __syncthreads();
  for (int off = 16; off > 0; off >>= 1) {
    if (ry < off) {
      sh_s[ry][cx] += sh_s[ry + off][cx];
      sh_q[ry][cx] += sh_q[ry + off][cx];
    }
    __syncthreads();
  }
  s_out = sh_s[0][cx];
  q_out = sh_q[0][cx];
}

// training: mean/var from the per-CTA partial sums written by the conv epilogue.
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ stats, int rows, int C, int ld, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, long long* nbt,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int training) {
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const bool cvalid = c < C;
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && nbt != nullptr) *nbt += 1;
  double s = 0.0, q = 0.0;
  if (training) colsum2_block(stats, rows, ld, c, cvalid, s, q);
  if (!cvalid || threadIdx.x >= 32) return;
  float mean, var;
  if (training) {
    const double m = s / count;
    double v = q / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean != nullptr) {
      const double unb = count > 1.0 ? v * count / (count - 1.0) : v;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float inv = rsqrtf(var + eps);
  const float sc = gamma[c] * inv;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (invstd_out) invstd_out[c] = inv;
}

// inference: fold the running statistics of EVERY BatchNorm of the network in one launch (one CTA per layer) instead of one
// bn_finalize launch in front of every conv: 79 launches and as many kernel boundaries less per forward.
// items: n rows of 8 x int64 = {gamma, beta, running_mean, running_var, scale, shift (device pointers), C, eps (float bits)}
__global__ void __launch_bounds__(256) bn_fold_batch_kernel(const long long* __restrict__ items) {
  const long long* it = items + (size_t)blockIdx.x * 8;
  const float* gamma = reinterpret_cast<const float*>(it[0]);
  const float* beta = reinterpret_cast<const float*>(it[1]);
  const float* rm = reinterpret_cast<const float*>(it[2]);
  const float* rv = reinterpret_cast<const float*>(it[3]);
  float* scale = reinterpret_cast<float*>(it[4]);
  float* shift = reinterpret_cast<float*>(it[5]);
  const int C = (int)it[6];
  const float eps = __int_as_float((int)it[7]);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {  // same arithmetic as bn_finalize_kernel's inference branch
    const float inv = rsqrtf(rv[c] + eps);
    const float sc = gamma[c] * inv;
    scale[c] = sc;
    shift[c] = beta[c] - rm[c] * sc;
  }
}

// ------------------------------------------------------------------------------------------------ BN apply + SiLU
// kEwThreads = 192 = 2^6 * 3: for every channel count of the network (C/8 in {2,6,12,24,48,96,192}) the grid-stride
// (gridDim * 192) is a multiple of C/8, so a thread keeps ONE channel vector for its whole life: the per-channel
// parameters are loaded once, the pixel index advances by a constant, and kEwUnroll independent 16-byte loads are in
// flight per thread and operand.  FIXED = false is the generic (any C % 8 == 0) path.
static constexpr int kEwThreads = 192;
static constexpr int kEwUnroll = 4;

template <bool FIXED>
__global__ void __launch_bounds__(kEwThreads, 4) bn_act_fwd_kernel(const bf16* __restrict__ y, long y_pitch, int H, int W, int C, long npix,
                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                  const bf16* __restrict__ res, long res_pitch, bf16* __restrict__ out, long out_pitch,
                                  bf16* __restrict__ out_up, long up_pitch) {
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed
  const unsigned cv = C >> 3;
  auto finish = [&](long pix, unsigned c, V8 v, const V8& sc, const V8& sh, const uint4* rraw) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // sc / sh are HALF the BN scale / shift: hz = z / 2, SiLU(z) = hz * tanh(hz) + hz
      const float hz = fmaf(v.v[j], sc.v[j], sh.v[j]);
      v.v[j] = fmaf(hz, tanh_fast(hz), hz);
    }
    if (rraw != nullptr) {
      const V8 r = cvt8(*rraw);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] += r.v[j];
    }
    st8(out + pix * out_pitch + c, v);
    if (out_up != nullptr) {
      const long w = pix % W, t = pix / W;
      const long h = t % H, n = t / H;
      bf16* u = out_up + (((n * 2 * H + 2 * h) * 2 * W) + 2 * w) * up_pitch + c;
      st8(u, v);
      st8(u + up_pitch, v);
      st8(u + 2 * W * up_pitch, v);
      st8(u + (2 * W + 1) * up_pitch, v);
    }
  };
  if (FIXED) {
    const unsigned T = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned c = (tid % cv) << 3;
    const long pstep = T / cv;
    const V8 sc = half8(ldf8(scale + c)), sh = half8(ldf8(shift + c));
    long pix = tid / cv;
    for (; pix + (kEwUnroll - 1) * pstep < npix; pix += kEwUnroll * pstep) {
      uint4 yr[kEwUnroll], rr[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) yr[u] = ldstream(y + (pix + u * pstep) * y_pitch + c);
      if (res != nullptr) {
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) rr[u] = ldstream(res + (pix + u * pstep) * res_pitch + c);
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) finish(pix + u * pstep, c, cvt8(yr[u]), sc, sh, res != nullptr ? &rr[u] : nullptr);
    }
    for (; pix < npix; pix += pstep) {
      const uint4 yr = ldstream(y + pix * y_pitch + c);
      uint4 rr = make_uint4(0, 0, 0, 0);
      if (res != nullptr) rr = ldstream(res + pix * res_pitch + c);
      finish(pix, c, cvt8(yr), sc, sh, res != nullptr ? &rr : nullptr);
    }
  } else {
    const long total = npix * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const long pix = i / cv;
      const unsigned c = (unsigned)(i - pix * cv) << 3;
      const uint4 yr = ldstream(y + pix * y_pitch + c);
      uint4 rr = make_uint4(0, 0, 0, 0);
      if (res != nullptr) rr = ldstream(res + pix * res_pitch + c);
      finish(pix, c, cvt8(yr), half8(ldf8(scale + c)), half8(ldf8(shift + c)), res != nullptr ? &rr : nullptr);
    }
  }
}

// plain nearest 2x upsample (eval path: the producer conv already applied BN+SiLU in its epilogue)
__global__ void upsample2x_fwd_kernel(const bf16* __restrict__ src, long src_pitch, int H, int W, int C, long npix,
                                      bf16* __restrict__ dst, long dst_pitch) {
  const int cv = C >> 3;
  const long total = npix * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / cv;
    const int c = (int)(i - pix * cv) << 3;
    const uint4 v = *reinterpret_cast<const uint4*>(src + pix * src_pitch + c);
    const long w = pix % W, t = pix / W;
    const long h = t % H, n = t / H;
    bf16* u = dst + (((n * 2 * H + 2 * h) * 2 * W) + 2 * w) * dst_pitch + c;
    *reinterpret_cast<uint4*>(u) = v;
    *reinterpret_cast<uint4*>(u + dst_pitch) = v;
    *reinterpret_cast<uint4*>(u + 2 * W * dst_pitch) = v;
    *reinterpret_cast<uint4*>(u + (2 * W + 1) * dst_pitch) = v;
  }
}

// d(src)[n,h,w] (+)= sum of the 2x2 block of d(up)
__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dup, long dup_pitch, int H, int W, int C, long npix,
                                      bf16* __restrict__ dsrc, long dsrc_pitch, int accumulate) {
  const int cv = C >> 3;
  const long total = npix * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / cv;
    const int c = (int)(i - pix * cv) << 3;
    const long w = pix % W, t = pix / W;
    const long h = t % H, n = t / H;
    const bf16* u = dup + (((n * 2 * H + 2 * h) * 2 * W) + 2 * w) * dup_pitch + c;
    const V8 a = ld8(u), b = ld8(u + dup_pitch), d = ld8(u + 2 * W * dup_pitch), e = ld8(u + (2 * W + 1) * dup_pitch);
    V8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = (a.v[j] + b.v[j]) + (d.v[j] + e.v[j]);
    bf16* dp = dsrc + pix * dsrc_pitch + c;
    if (accumulate) {
      const V8 p = ld8(dp);
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] += p.v[j];
    }
    st8(dp, o);
  }
}

__global__ void add_into_kernel(const bf16* __restrict__ src, long src_pitch, bf16* __restrict__ dst, long dst_pitch,
                                long npix, int C, int accumulate) {
  const int cv = C >> 3;
  const long total = npix * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / cv;
    const int c = (int)(i - pix * cv) << 3;
    V8 v = ld8(src + pix * src_pitch + c);
    bf16* dp = dst + pix * dst_pitch + c;
    if (accumulate) {
      const V8 p = ld8(dp);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] += p.v[j];
    }
    st8(dp, v);
  }
}

// ------------------------------------------------------------------------------------------------ BN + SiLU backward

#ifndef YB_REDUCE_OCC
#define YB_REDUCE_OCC 2
#endif
static constexpr int kReduceOcc = YB_REDUCE_OCC;  // resident CTAs per SM of the reduction pass (register cap 65536/256/occ)

// pass 1: per-channel sums of dz and dz*xhat, dz = da * silu'(y*scale+shift), xhat = (y-mean)*invstd.
// block = (rows x cv) threads; every block owns a contiguous pixel range; partial[block][2][C].
// mode 1: plain column sums of `da` (head bias gradient): partial[block][0][C] only.
__global__ void __launch_bounds__(256, kReduceOcc) bn_act_bwd_reduce_kernel(const bf16* __restrict__ da, long da_pitch, const bf16* __restrict__ y,
                                         long y_pitch, long npix, int C, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, float* __restrict__ partial, int rows_pb,
                                         int mode) {
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed
  extern __shared__ float sred[];  // [rows_pb][2][C]
  const int cv = C >> 3;
  const int tid = threadIdx.x;
  const int row = tid / cv, c = (tid - row * cv) << 3;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (row < rows_pb) {
    const long per = (npix + gridDim.x - 1) / gridDim.x;
    const long p0 = (long)blockIdx.x * per, p1 = min(npix, p0 + per);
    V8 sc, sh;  // hz = y*sc + sh (half scale / shift)
    if (mode == 0) {
      sc = half8(ldf8(scale + c)); sh = half8(ldf8(shift + c));
    }
    // the loop accumulates s1 = sum 2*dz and s2 = sum 2*dz*y (RAW y): 9 instructions per element; the centring
    // (xhat = y*xa + xb) is applied once per thread after the loop
    auto accum = [&](const uint4& graw, const uint4& yraw) {
      const V8 g = cvt8(graw);
      if (mode == 0) {
        const V8 yv = cvt8(yraw);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gw = fmaf(g.v[j], dsilu2_w(fmaf(yv.v[j], sc.v[j], sh.v[j])), g.v[j]);
          s1[j] += gw;
          s2[j] = fmaf(gw, yv.v[j], s2[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] += g.v[j];
      }
    };
    long p = p0 + row;
    for (; p + (kEwUnroll - 1) * (long)rows_pb < p1; p += kEwUnroll * (long)rows_pb) {
      uint4 gr[kEwUnroll], yr[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) gr[u] = ldstream(da + (p + u * (long)rows_pb) * da_pitch + c);
      if (mode == 0) {
#pragma unroll
        for (int u = 0; u < kEwUnroll; ++u) yr[u] = ldstream(y + (p + u * (long)rows_pb) * y_pitch + c);
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) accum(gr[u], yr[u]);
    }
    for (; p < p1; p += rows_pb) {
      const uint4 gr = ldstream(da + p * da_pitch + c);
      uint4 yr = make_uint4(0, 0, 0, 0);
      if (mode == 0) yr = ldstream(y + p * y_pitch + c);
      accum(gr, yr);
    }
    if (mode == 0) {  // sum dz = s1 / 2;  sum dz * xhat = xa * (s2 / 2) + xb * (s1 / 2),  xa = invstd, xb = -mean * invstd
      const V8 xa = ldf8(invstd + c), mu = ldf8(mean + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] *= 0.5f;
        s2[j] = xa.v[j] * fmaf(-mu.v[j], s1[j], 0.5f * s2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sred[(size_t)row * 2 * C + c + j] = s1[j];
      sred[(size_t)row * 2 * C + C + c + j] = s2[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < rows_pb; ++r) a += sred[(size_t)r * 2 * C + i];
    partial[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

// dgamma = sum dz*xhat, dbeta = sum dz; coef[0][c] = dbeta/m, coef[1][c] = dgamma/m
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                       int accumulate) {
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const bool cvalid = c < C;
  double s, q;
  colsum2_block(partial, rows, C, c, cvalid, s, q);
  if (!cvalid || threadIdx.x >= 32) return;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)q : (float)q;
  if (coef) {
    coef[c] = (float)(s / count);
    coef[C + c] = (float)(q / count);
  }
}

// pass 2: dy = scale * (dz - coef0 - xhat*coef1)
template <bool FIXED>
__global__ void __launch_bounds__(kEwThreads, 5) bn_act_bwd_apply_kernel(const bf16* __restrict__ da, long da_pitch, const bf16* __restrict__ y,
                                        long y_pitch, long npix, int C, const float* __restrict__ scale,
                                        const float* __restrict__ shift, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, const float* __restrict__ coef,
                                        bf16* __restrict__ dy, long dy_pitch) {
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed
  const unsigned cv = C >> 3;
  // dy = scale * (dz - coef0 - xhat * coef1),  xhat = (y - mean) * invstd   ==>   dy = scale * dz + (y * m1 + m0)
  //  scale * dz = (scale / 2) * (2 dz) = hsc * fma(g, w, g): the half scale that builds hz is also the output factor
  struct Par {
    V8 hsc, hsh, m1, m0;
  };
  auto load_par = [&](unsigned c) {
    Par q;
    const V8 sc = ldf8(scale + c);
    q.hsc = half8(sc);
    q.hsh = half8(ldf8(shift + c));
    const V8 mu = ldf8(mean + c), is = ldf8(invstd + c), c0 = ldf8(coef + c), c1 = ldf8(coef + C + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float k1 = is.v[j] * c1.v[j];
      q.m1.v[j] = -sc.v[j] * k1;
      q.m0.v[j] = -sc.v[j] * (c0.v[j] - mu.v[j] * k1);
    }
    return q;
  };
  auto finish = [&](long pix, unsigned c, const uint4& graw, const uint4& yraw, const Par& q) {
    const V8 g = cvt8(graw), yv = cvt8(yraw);
    V8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gw = fmaf(g.v[j], dsilu2_w(fmaf(yv.v[j], q.hsc.v[j], q.hsh.v[j])), g.v[j]);
      o.v[j] = fmaf(q.hsc.v[j], gw, fmaf(yv.v[j], q.m1.v[j], q.m0.v[j]));
    }
    st8(dy + pix * dy_pitch + c, o);
  };
  if (FIXED) {
    const unsigned T = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned c = (tid % cv) << 3;
    const long pstep = T / cv;
    const Par q = load_par(c);
    long pix = tid / cv;
    for (; pix + (kEwUnroll - 1) * pstep < npix; pix += kEwUnroll * pstep) {
      uint4 gr[kEwUnroll], yr[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) {
        gr[u] = ldstream(da + (pix + u * pstep) * da_pitch + c);
        yr[u] = ldstream(y + (pix + u * pstep) * y_pitch + c);
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; ++u) finish(pix + u * pstep, c, gr[u], yr[u], q);
    }
    for (; pix < npix; pix += pstep) finish(pix, c, ldstream(da + pix * da_pitch + c), ldstream(y + pix * y_pitch + c), q);
  } else {
    const long total = npix * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const long pix = i / cv;
      const unsigned c = (unsigned)(i - pix * cv) << 3;
      finish(pix, c, ldstream(da + pix * da_pitch + c), ldstream(y + pix * y_pitch + c), load_par(c));
    }
  }
}

// ------------------------------------------------------------------------------------------------ SPPF max pool 5/1/2
__global__ void maxpool5_fwd_kernel(const bf16* __restrict__ x, long x_pitch, int H, int W, int C, long npix,
                                    bf16* __restrict__ out, long out_pitch, uint8_t* __restrict__ argmax) {
  const int cv = C >> 3;
  const long total = npix * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / cv;
    const int c = (int)(i - pix * cv) << 3;
    const int w = (int)(pix % W);
    const long t = pix / W;
    const int h = (int)(t % H);
    const long n = t / H;
    // fully unrolled, branch-free window: every tap loads from a clamped (always valid) address, so the 25 loads are
    // independent and in flight together; a tap outside the image is masked out of the comparison.  `arg` starts at the
    // first in-image tap, which is what "first max wins" yields when every value is -inf.
    float best[8];
    int arg[8];
    const int arg0 = max(0, 2 - h) * 5 + max(0, 2 - w);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      arg[j] = arg0;
    }
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {
      const int ih = h + kh - 2;
      const bool okh = ih >= 0 && ih < H;
      const int ihc = min(max(ih, 0), H - 1);
#pragma unroll
      for (int kw = 0; kw < 5; ++kw) {
        const int iw = w + kw - 2;
        const bool ok = okh && iw >= 0 && iw < W;
        const int iwc = min(max(iw, 0), W - 1);
        const V8 v = ld8(x + ((n * H + ihc) * W + iwc) * x_pitch + c);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (ok && (v.v[j] > best[j] || v.v[j] != v.v[j])) {  // first max wins; NaN propagates (ATen max_pool2d)
            best[j] = v.v[j];
            arg[j] = kh * 5 + kw;
          }
      }
    }
    V8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = best[j];
    st8(out + pix * out_pitch + c, o);
    if (argmax != nullptr) {
      uint2 pk;
      pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      pk.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      *reinterpret_cast<uint2*>(argmax + pix * C + c) = pk;
    }
  }
}

// gather form (deterministic): dx[i] (+)= sum over the <=25 outputs whose recorded argmax is i
__global__ void maxpool5_bwd_kernel(const bf16* __restrict__ dy, long dy_pitch, const uint8_t* __restrict__ argmax, int H,
                                    int W, int C, long npix, bf16* __restrict__ dx, long dx_pitch, int accumulate) {
  const int cv = C >> 3;
  const long total = npix * cv;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / cv;
    const int c = (int)(i - pix * cv) << 3;
    const int w = (int)(pix % W);
    const long t = pix / W;
    const int h = (int)(t % H);
    const long n = t / H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {  // unrolled and branch-free like the forward pass (clamped addresses, masked taps)
      const int oh = h - kh + 2;
      const bool okh = oh >= 0 && oh < H;
      const int ohc = min(max(oh, 0), H - 1);
#pragma unroll
      for (int kw = 0; kw < 5; ++kw) {
        const int ow = w - kw + 2;
        const bool ok = okh && ow >= 0 && ow < W;
        const int owc = min(max(ow, 0), W - 1);
        const long op = (n * H + ohc) * W + owc;
        const uint2 pk = *reinterpret_cast<const uint2*>(argmax + op * C + c);
        const V8 g = ld8(dy + op * dy_pitch + c);
        const unsigned k = kh * 5 + kw;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned a = ((j < 4 ? pk.x : pk.y) >> (8 * (j & 3))) & 0xffu;
          if (ok && a == k) acc[j] += g.v[j];
        }
      }
    }
    bf16* dp = dx + pix * dx_pitch + c;
    V8 o;
    if (accumulate) {
      o = ld8(dp);
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] += acc[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = acc[j];
    }
    st8(dp, o);
  }
}


// ------------------------------------------------------------------------------------------------ fused SPPF pooling
// SPPF (model.py:96-112) chains three 5/1/2 max pools: cat[x, p(x), p(p(x)), p(p(p(x)))].  One CTA owns one image x 16
// channels: the H x W tile is loaded into shared memory once and the three pools run back to back there, each as a
// separable pass (row maximum over kw, then column maximum over kh of the row results: 10 shared-memory reads per output
// instead of 25 global loads; "first maximum in row-major window order wins, NaN propagates" is preserved by the
// separation: the row pass keeps the first maximum of each row, the column pass the first row that holds the window
// maximum).  Outputs are the three channel slices of the concat buffer plus the 1-byte arg-max window slots for backward.
// Replaces three maxpool5_fwd launches that each re-read their input 25 times through L1/L2 (0.31 ms -> one launch).
struct Px16 {
  uint4 lo, hi;  // 16 bf16 channels
};
__device__ __forceinline__ void unpack16(const Px16& p, float (&v)[16]) {
  const V8 a = cvt8(p.lo), b = cvt8(p.hi);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = a.v[j];
    v[8 + j] = b.v[j];
  }
}
__device__ __forceinline__ Px16 pack16(const float (&v)[16]) {
  Px16 p;
  __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&p.lo);
  __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&p.hi);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    h1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
  }
  return p;
}

__global__ void __launch_bounds__(256) sppf_pool3_fwd_kernel(const bf16* __restrict__ x, long x_pitch, int H, int W, int C,
                                                             bf16* __restrict__ y1, bf16* __restrict__ y2, bf16* __restrict__ y3,
                                                             long y_pitch, uint8_t* __restrict__ am1, uint8_t* __restrict__ am2,
                                                             uint8_t* __restrict__ am3) {
  extern __shared__ __align__(16) uint8_t sp_smem[];
  const int HW = H * W;
  Px16* cur = reinterpret_cast<Px16*>(sp_smem);            // [HW] input of the current pool
  Px16* rmv = cur + HW;                                     // [HW] row maxima
  uint4* rma = reinterpret_cast<uint4*>(rmv + HW);          // [HW] kw slot of each row maximum, 16 x uint8
  const int cg = C >> 4;
  const long n = blockIdx.x / cg;
  const int c0 = (int)(blockIdx.x - n * cg) << 4;
  const long pix0 = n * HW;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const bf16* src = x + (pix0 + p) * x_pitch + c0;
    cur[p].lo = ldraw(src);
    cur[p].hi = ldraw(src + 8);
  }
  __syncthreads();
  for (int stage = 0; stage < 3; ++stage) {
    bf16* yo = stage == 0 ? y1 : (stage == 1 ? y2 : y3);
    uint8_t* am = stage == 0 ? am1 : (stage == 1 ? am2 : am3);
    // row pass: first maximum over kw = 0..4 (w - 2 .. w + 2) inside the image
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      float best[16];
      unsigned arg[16];
      const unsigned a0 = (unsigned)max(0, 2 - w);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        best[j] = -INFINITY;
        arg[j] = a0;
      }
#pragma unroll
      for (int kw = 0; kw < 5; ++kw) {
        const int iw = w + kw - 2;
        if (iw < 0 || iw >= W) continue;
        float v[16];
        unpack16(cur[h * W + iw], v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (v[j] > best[j] || v[j] != v[j]) {
            best[j] = v[j];
            arg[j] = kw;
          }
      }
      rmv[p] = pack16(best);  // values are bf16 already: exact
      uint4 pk;
      pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      pk.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      pk.z = arg[8] | (arg[9] << 8) | (arg[10] << 16) | (arg[11] << 24);
      pk.w = arg[12] | (arg[13] << 8) | (arg[14] << 16) | (arg[15] << 24);
      rma[p] = pk;
    }
    __syncthreads();
    // column pass: first row (kh = 0..4) whose row maximum is the window maximum; slot = kh * 5 + kw
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      float best[16];
      unsigned arg[16];
      const unsigned a0 = (unsigned)(max(0, 2 - h) * 5 + max(0, 2 - w));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        best[j] = -INFINITY;
        arg[j] = a0;
      }
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const int ih = h + kh - 2;
        if (ih < 0 || ih >= H) continue;
        float v[16];
        unpack16(rmv[ih * W + w], v);
        const uint4 ka = rma[ih * W + w];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (v[j] > best[j] || v[j] != v[j]) {
            const unsigned word = j < 4 ? ka.x : (j < 8 ? ka.y : (j < 12 ? ka.z : ka.w));
            best[j] = v[j];
            arg[j] = (unsigned)kh * 5u + ((word >> (8 * (j & 3))) & 0xffu);
          }
      }
      const Px16 o = pack16(best);
      bf16* dst = yo + (pix0 + p) * y_pitch + c0;
      *reinterpret_cast<uint4*>(dst) = o.lo;
      *reinterpret_cast<uint4*>(dst + 8) = o.hi;
      if (am != nullptr) {
        uint4 pk;
        pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
        pk.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
        pk.z = arg[8] | (arg[9] << 8) | (arg[10] << 16) | (arg[11] << 24);
        pk.w = arg[12] | (arg[13] << 8) | (arg[14] << 16) | (arg[15] << 24);
        *reinterpret_cast<uint4*>(am + (pix0 + p) * C + c0) = pk;
      }
      // the output of this pool is the input of the next one; `cur` is no longer read in this stage (the row pass is done)
      cur[p] = o;
    }
    __syncthreads();
  }
}

// inference form (no arg-max planes): packed bf16x2 maxima, NaN-propagating like ATen's `val > max || isnan(val)`; a maximum
// is a selection, so the result is the chain's bit for bit (up to the sign of a zero that ties with its negative).
__device__ __forceinline__ uint4 max8(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) pr[j] = __hmax2_nan(pa[j], pb[j]);
  return r;
}
__global__ void __launch_bounds__(512) sppf_pool3_eval_kernel(const bf16* __restrict__ x, long x_pitch, int H, int W, int C,
                                                              bf16* __restrict__ y1, bf16* __restrict__ y2, bf16* __restrict__ y3,
                                                              long y_pitch) {
  extern __shared__ __align__(16) uint8_t sp_smem[];
  const int HW = H * W;
  Px16* cur = reinterpret_cast<Px16*>(sp_smem);  // [HW] input of the current pool
  Px16* rmv = cur + HW;                           // [HW] row maxima
  const int cg = C >> 4;
  const long n = blockIdx.x / cg;
  const int c0 = (int)(blockIdx.x - n * cg) << 4;
  const long pix0 = n * HW;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const bf16* src = x + (pix0 + p) * x_pitch + c0;
    cur[p].lo = ldraw(src);
    cur[p].hi = ldraw(src + 8);
  }
  __syncthreads();
  for (int stage = 0; stage < 3; ++stage) {
    bf16* yo = stage == 0 ? y1 : (stage == 1 ? y2 : y3);
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      const int lo = max(w - 2, 0), hi = min(w + 2, W - 1);
      Px16 m = cur[h * W + lo];
      for (int iw = lo + 1; iw <= hi; ++iw) {
        const Px16 v = cur[h * W + iw];
        m.lo = max8(m.lo, v.lo);
        m.hi = max8(m.hi, v.hi);
      }
      rmv[p] = m;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      const int lo = max(h - 2, 0), hi = min(h + 2, H - 1);
      Px16 m = rmv[lo * W + w];
      for (int ih = lo + 1; ih <= hi; ++ih) {
        const Px16 v = rmv[ih * W + w];
        m.lo = max8(m.lo, v.lo);
        m.hi = max8(m.hi, v.hi);
      }
      bf16* dst = yo + (pix0 + p) * y_pitch + c0;
      *reinterpret_cast<uint4*>(dst) = m.lo;
      *reinterpret_cast<uint4*>(dst + 8) = m.hi;
      cur[p] = m;  // input of the next pool (the row pass of this stage is complete)
    }
    __syncthreads();
  }
}

// backward of the chain in one launch: g2 += scatter(g3, am3); g1 += scatter(g2, am2); g0 += scatter(g1, am1), the
// intermediate sums kept in fp32 in shared memory (gather form: every input pixel sums the <= 25 outputs that point at
// it).  g0..g3 = gradient slices of the concat buffer ([x | p1 | p2 | p3]); only g0 is written (g1 / g2 have no other reader).
__global__ void __launch_bounds__(256) sppf_pool3_bwd_kernel(const bf16* __restrict__ g1, const bf16* __restrict__ g2,
                                                             const bf16* __restrict__ g3, long g_pitch, const uint8_t* __restrict__ am1,
                                                             const uint8_t* __restrict__ am2, const uint8_t* __restrict__ am3, int H, int W,
                                                             int C, bf16* __restrict__ g0, long g0_pitch, int accumulate) {
  extern __shared__ __align__(16) uint8_t sp_smem[];
  const int HW = H * W;
  float* up = reinterpret_cast<float*>(sp_smem);            // [HW][16] gradient w.r.t. the output of the current pool
  float* dn = up + (size_t)HW * 16;                         // [HW][16] gradient w.r.t. its input
  uint4* am = reinterpret_cast<uint4*>(dn + (size_t)HW * 16);  // [HW] arg-max slots of the current pool
  const int cg = C >> 4;
  const long n = blockIdx.x / cg;
  const int c0 = (int)(blockIdx.x - n * cg) << 4;
  const long pix0 = n * HW;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    Px16 v;
    const bf16* src = g3 + (pix0 + p) * g_pitch + c0;
    v.lo = ldraw(src);
    v.hi = ldraw(src + 8);
    float f[16];
    unpack16(v, f);
#pragma unroll
    for (int j = 0; j < 16; ++j) up[p * 16 + j] = f[j];
  }
  for (int stage = 2; stage >= 0; --stage) {
    const uint8_t* ams = stage == 2 ? am3 : (stage == 1 ? am2 : am1);
    for (int p = threadIdx.x; p < HW; p += blockDim.x) am[p] = *reinterpret_cast<const uint4*>(ams + (pix0 + p) * C + c0);
    __syncthreads();
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {
      const int h = p / W, w = p - h * W;
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const int oh = h - kh + 2;
        if (oh < 0 || oh >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const int ow = w - kw + 2;
          if (ow < 0 || ow >= W) continue;
          const int op = oh * W + ow;
          const uint4 ka = am[op];
          const unsigned k = kh * 5 + kw;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const unsigned word = j < 4 ? ka.x : (j < 8 ? ka.y : (j < 12 ? ka.z : ka.w));
            if (((word >> (8 * (j & 3))) & 0xffu) == k) acc[j] += up[op * 16 + j];
          }
        }
      }
      // add the gradient the concat slice of this pool's INPUT received directly from c_out (stage 0: x's slice, g0)
      if (stage > 0) {
        Px16 v;
        const bf16* src = (stage == 2 ? g2 : g1) + (pix0 + p) * g_pitch + c0;
        v.lo = ldraw(src);
        v.hi = ldraw(src + 8);
        float f[16];
        unpack16(v, f);
#pragma unroll
        for (int j = 0; j < 16; ++j) dn[p * 16 + j] = acc[j] + f[j];
      } else {
        bf16* dst = g0 + (pix0 + p) * g0_pitch + c0;
        if (accumulate) {
          Px16 v;
          v.lo = ldraw(dst);
          v.hi = ldraw(dst + 8);
          float f[16];
          unpack16(v, f);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += f[j];
        }
        const Px16 o = pack16(acc);
        *reinterpret_cast<uint4*>(dst) = o.lo;
        *reinterpret_cast<uint4*>(dst + 8) = o.hi;
      }
    }
    __syncthreads();
    float* t = up;
    up = dn;
    dn = t;
  }
}

// ------------------------------------------------------------------------------------------------ input staging
// x (N,3,H,W) float in [0,1] (or uint8, divided by 255) -> out (N,H/2,W/2+2,16) bf16:
//   space-to-depth: s2d[ho][wo][(r*2+s)*3+c] = x[c][2ho+r][2wo+s] (12 channels, padded to 16), stored at column wo + 1 of a
//   row that carries one zero pixel on either side.
// The 6x6/s2 stem (model.py:184) = 3x3/s1 over s2d = THREE vertical taps with K = 48 per tap over the VIEW
//   g[ho][wo][kw*16 + j] = s2d[ho][wo+kw-1][j] = 48 contiguous values of `out` starting at padded column wo
// (tensor-map dims (48, W/2, H/2, N) with a 32-byte pixel stride: overlapping windows, verified by
// tools/tma_overlap_probe.cu): a third of the TMA boxes / MMA steps of the nine-tap form and full-width pipeline stages,
// without materialising the 48-channel tensor (3x the bytes: 5.0 GB at bs=128, 1280x1280).
// RESIZE: the image is first resampled from (Hs, Ws) to (H, W) exactly like the reference's multi_scale()
// (utils/training_utils.py:11-28: nn.functional.interpolate(img, size, mode="bilinear", align_corners=False) on the
// float image): src = max(scale * (dst + 0.5) - 0.5, 0), scale = in / out, neighbours (i, min(i + 1, in - 1)).
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_coord(int dst, float scale, int in) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  Lerp r;
  r.i0 = min((int)src, in - 1);
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}
template <typename T>
__device__ __forceinline__ float px_value(const T* p) {
  if constexpr (sizeof(T) == 1) return (float)(*p) / 255.f;
  else return *p;
}

template <typename T, bool RESIZE>
__global__ void prep_input_kernel(const T* __restrict__ x, int N, int H, int W, bf16* __restrict__ out, int Hs, int Ws,
                                  float scale_h, float scale_w) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long total = (long)N * Ho * Wo;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const long t = i / Wo;
    const int ho = (int)(t % Ho);
    const long n = t / Ho;
    float v[16];
#pragma unroll
    for (int j = 12; j < 16; ++j) v[j] = 0.f;
    if constexpr (RESIZE) {
      Lerp lx[2], ly[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        ly[r] = lerp_coord(2 * ho + r, scale_h, Hs);
        lx[r] = lerp_coord(2 * wo + r, scale_w, Ws);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const T* plane = x + (n * 3 + c) * (long)Hs * Ws;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const T* r0 = plane + (long)ly[r].i0 * Ws;
          const T* r1 = plane + (long)ly[r].i1 * Ws;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float top = lx[q].l0 * px_value(r0 + lx[q].i0) + lx[q].l1 * px_value(r0 + lx[q].i1);
            const float bot = lx[q].l0 * px_value(r1 + lx[q].i0) + lx[q].l1 * px_value(r1 + lx[q].i1);
            v[(r * 2 + q) * 3 + c] = ly[r].l0 * top + ly[r].l1 * bot;
          }
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const T* p = x + ((n * 3 + c) * H + 2 * ho + r) * (long)W + 2 * wo;
          float a, b;
          if constexpr (sizeof(T) == 1) {
            a = (float)p[0] / 255.f;
            b = (float)p[1] / 255.f;
          } else {
            const float2 f = *reinterpret_cast<const float2*>(p);
            a = f.x;
            b = f.y;
          }
          v[(r * 2 + 0) * 3 + c] = a;
          v[(r * 2 + 1) * 3 + c] = b;
        }
    }
    V8 lo, hi, z;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      lo.v[j] = v[j];
      hi.v[j] = v[8 + j];
      z.v[j] = 0.f;
    }
    // one pixel = 16 channels = 32 bytes = one L2 sector (st.global.v8); rows carry one zero pixel on either side, so the
    // stem's three horizontal taps are the 48 contiguous values starting one pixel to the left (an overlapping-window
    // tensor-map view, see yb_prep_input in the header): the 48-channel tap-gathered tensor never exists in HBM
    bf16* me = out + ((n * Ho + ho) * (long)(Wo + 2) + wo + 1) * 16;
    st16(me, lo, hi);
    if (wo == 0) st16(me - 16, z, z);
    if (wo + 1 == Wo) st16(me + 16, z, z);
  }
}

// uint8 images at W % 8 == 0 (the detect / train batches): FOUR output pixels per thread -- six 8-byte loads in flight per
// thread instead of six 2-byte ones (the 1-pixel form ran at 1.65 TB/s: too few bytes in flight), and bf16(v / 255) comes
// from a 256-entry shared-memory table (same IEEE division, done once per value and CTA) instead of 12 divisions a pixel.
__global__ void __launch_bounds__(256) prep_input_u8x4_kernel(const uint8_t* __restrict__ x, int N, int H, int W,
                                                              bf16* __restrict__ out) {
  __shared__ unsigned short lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = __bfloat16_as_ushort(__float2bfloat16((float)i / 255.f));
  __syncthreads();
  const int Ho = H >> 1, Wo = W >> 1, Wq = Wo >> 2;
  const long total = (long)N * Ho * Wq;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int wq = (int)(i % Wq);
    const long t = i / Wq;
    const int ho = (int)(t % Ho);
    const long n = t / Ho;
    uint2 raw[3][2];  // [c][r]: input pixels 8 wq .. 8 wq + 7 of row 2 ho + r, channel c
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        raw[c][r] = ldg_stream8(x + ((n * 3 + c) * H + 2 * ho + r) * (long)W + 8 * wq);
    bf16* me = out + ((n * Ho + ho) * (long)(Wo + 2) + 4 * wq + 1) * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // output pixel 4 wq + q: input columns 2 q, 2 q + 1 of the 8 loaded ones
      unsigned short ch[16];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const unsigned word = q < 2 ? raw[c][r].x : raw[c][r].y;
          const unsigned two = (word >> (16 * (q & 1))) & 0xffffu;
          ch[(r * 2 + 0) * 3 + c] = lut[two & 0xffu];
          ch[(r * 2 + 1) * 3 + c] = lut[two >> 8];
        }
#pragma unroll
      for (int j = 12; j < 16; ++j) ch[j] = 0;
      uint4 lo, hi;
      lo.x = ch[0] | ((unsigned)ch[1] << 16); lo.y = ch[2] | ((unsigned)ch[3] << 16);
      lo.z = ch[4] | ((unsigned)ch[5] << 16); lo.w = ch[6] | ((unsigned)ch[7] << 16);
      hi.x = ch[8] | ((unsigned)ch[9] << 16); hi.y = ch[10] | ((unsigned)ch[11] << 16);
      hi.z = 0u; hi.w = 0u;
      *reinterpret_cast<uint4*>(me + q * 16) = lo;
      *reinterpret_cast<uint4*>(me + q * 16 + 8) = hi;
    }
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    if (wq == 0) {
      *reinterpret_cast<uint4*>(me - 16) = z4;
      *reinterpret_cast<uint4*>(me - 8) = z4;
    }
    if (wq + 1 == Wq) {
      *reinterpret_cast<uint4*>(me + 64) = z4;
      *reinterpret_cast<uint4*>(me + 72) = z4;
    }
  }
}

// g (B,na,H,W,no) fp32 -> dy (B,H,W,Cpad) bf16, channel a*no+o; channels >= na*no are zero
__global__ void head_grad_pack_kernel(const float* __restrict__ g, int na, long hw, int no, bf16* __restrict__ dy, int Cpad,
                                      long total, int accumulate) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long t = i / Cpad;
    const long pix = t % hw, b = t / hw;
    float v = 0.f;
    if (c < na * no) {
      const int a = c / no, o = c - a * no;
      v = g[((b * na + a) * hw + pix) * no + o];
    }
    if (accumulate) v += __bfloat162float(dy[i]);
    dy[i] = __float2bfloat16(v);
  }
}

// out[i] (+)= sum_r partial[r*stride + i]
__global__ void reduce_rows_kernel(const float* __restrict__ partial, int rows, long stride, int n, float* __restrict__ out,
                                   int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += (double)partial[(size_t)r * stride + i];
  out[i] = accumulate ? out[i] + (float)s : (float)s;
}

}  // namespace yb

using namespace yb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define B16(p) reinterpret_cast<bf16*>(p)
#define CB16(p) reinterpret_cast<const bf16*>(p)
#define LAUNCH_OK() YB_LAUNCHED()

extern "C" {

int yb_bn_fold_batch(const void* items, int n, void* stream) {
  if (n <= 0) return 0;
  YB_REQUIRE(items != nullptr, "bn_fold_batch: items");
  bn_fold_batch_kernel<<<n, 256, 0, ST(stream)>>>(reinterpret_cast<const long long*>(items));
  LAUNCH_OK();
  return 0;
}

int yb_bn_finalize(const float* stats, int rows, int C, double count, const float* gamma, const float* beta, float eps,
                   float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, float* scale,
                   float* shift, float* mean, float* invstd, int training, void* stream) {
  return yb_bn_finalize_ld(stats, rows, C, C, count, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked,
                           scale, shift, mean, invstd, training, stream);
}

int yb_bn_finalize_ld(const float* stats, int rows, int C, int stats_ld, double count, const float* gamma, const float* beta,
                      float eps, float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* mean, float* invstd, int training, void* stream) {
  YB_REQUIRE(training ? (stats != nullptr && rows > 0) : (running_mean && running_var), "bn_finalize: missing inputs");
  YB_REQUIRE(stats_ld >= C, "bn_finalize: stats_ld=%d < C=%d", stats_ld, C);
  YB_CHECK_CUDA(launch_pdl(bn_finalize_kernel, dim3((C + 31) / 32), dim3(1024), 0, ST(stream), stats, rows, C, stats_ld, count, gamma, beta, eps, momentum,
                                                               running_mean, running_var,
                                                               reinterpret_cast<long long*>(num_batches_tracked), scale,
                                                               shift, mean, invstd, training));
  LAUNCH_OK();
  return 0;
}

int yb_bn_act_fwd(const void* y, int64_t y_pitch, int N, int H, int W, int C, const float* scale, const float* shift,
                  const void* res, int64_t res_pitch, void* out, int64_t out_pitch, void* out_up, int64_t up_pitch,
                  void* stream) {
  YB_REQUIRE(C % 8 == 0 && y_pitch % 8 == 0 && out_pitch % 8 == 0 && res_pitch % 8 == 0 && up_pitch % 8 == 0,
             "bn_act_fwd: C and pitches must be multiples of 8");
  const long npix = (long)N * H * W;
  static const int occ_grid = ew_wave_grid(bn_act_fwd_kernel<true>, kEwThreads, 1L << 40);
  const int grid = std::min(occ_grid, ew_blocks(npix * (C / 8), kEwThreads, 64));
  if (kEwThreads % (C / 8) == 0)
    YB_CHECK_CUDA(launch_pdl(bn_act_fwd_kernel<true>, dim3(grid), dim3(kEwThreads), 0, ST(stream), CB16(y), y_pitch, H, W, C, npix, scale, shift, CB16(res),
                                                                 res_pitch, B16(out), out_pitch, B16(out_up), up_pitch));
  else
    YB_CHECK_CUDA(launch_pdl(bn_act_fwd_kernel<false>, dim3(grid), dim3(kEwThreads), 0, ST(stream), CB16(y), y_pitch, H, W, C, npix, scale, shift, CB16(res),
                                                                  res_pitch, B16(out), out_pitch, B16(out_up), up_pitch));
  LAUNCH_OK();
  return 0;
}

int yb_upsample2x_fwd(const void* src, int64_t src_pitch, int N, int H, int W, int C, void* dst, int64_t dst_pitch,
                      void* stream) {
  YB_REQUIRE(C % 8 == 0 && src_pitch % 8 == 0 && dst_pitch % 8 == 0, "upsample2x_fwd: alignment");
  const long npix = (long)N * H * W;
  upsample2x_fwd_kernel<<<ew_blocks(npix * (C / 8)), 256, 0, ST(stream)>>>(CB16(src), src_pitch, H, W, C, npix, B16(dst),
                                                                            dst_pitch);
  LAUNCH_OK();
  return 0;
}

int yb_upsample2x_bwd(const void* dup, int64_t dup_pitch, int N, int H, int W, int C, void* dsrc, int64_t dsrc_pitch,
                      int accumulate, void* stream) {
  YB_REQUIRE(C % 8 == 0 && dup_pitch % 8 == 0 && dsrc_pitch % 8 == 0, "upsample2x_bwd: alignment");
  const long npix = (long)N * H * W;
  upsample2x_bwd_kernel<<<ew_blocks(npix * (C / 8)), 256, 0, ST(stream)>>>(CB16(dup), dup_pitch, H, W, C, npix, B16(dsrc),
                                                                            dsrc_pitch, accumulate);
  LAUNCH_OK();
  return 0;
}

int yb_add_into(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t npix, int C, int accumulate,
                void* stream) {
  YB_REQUIRE(C % 8 == 0 && src_pitch % 8 == 0 && dst_pitch % 8 == 0, "add_into: alignment");
  add_into_kernel<<<ew_blocks(npix * (C / 8)), 256, 0, ST(stream)>>>(CB16(src), src_pitch, B16(dst), dst_pitch, npix, C,
                                                                      accumulate);
  LAUNCH_OK();
  return 0;
}

static constexpr int kReduceRows = 148 * kReduceOcc;
static int reduce_geometry(int C, long npix, int& threads, int& rows_pb, int& grid, size_t& smem) {
  const int cv = C / 8;
  YB_REQUIRE(C % 8 == 0 && cv <= 256, "bwd_reduce: C=%d unsupported", C);
  rows_pb = std::max(1, 256 / cv);
  while (rows_pb > 1 && (size_t)rows_pb * 2 * C * sizeof(float) > 96 * 1024) --rows_pb;
  threads = ((rows_pb * cv + 31) / 32) * 32;
  smem = (size_t)rows_pb * 2 * C * sizeof(float);
  const long want = (npix + (long)rows_pb * 16 - 1) / ((long)rows_pb * 16);  // >= 16 pixels per thread row
  grid = (int)std::max<long>(1, std::min<long>(want, kReduceRows));  // one resident wave: 148 SMs x kReduceOcc CTAs
  return 0;
}

int yb_bwd_reduce_max_rows(void) { return kReduceRows; }

int yb_bn_act_bwd_reduce(const void* da, int64_t da_pitch, const void* y, int64_t y_pitch, int64_t npix, int C,
                         const float* scale, const float* shift, const float* mean, const float* invstd, float* partial,
                         int* rows, void* stream) {
  int threads, rows_pb, grid;
  size_t smem;
  if (reduce_geometry(C, npix, threads, rows_pb, grid, smem)) return -1;
  YB_REQUIRE(da_pitch % 8 == 0 && y_pitch % 8 == 0, "bn_act_bwd_reduce: alignment");
  static bool attr = false;
  if (!attr) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(bn_act_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  YB_CHECK_CUDA(launch_pdl(bn_act_bwd_reduce_kernel, dim3(grid), dim3(threads), smem, ST(stream), CB16(da), da_pitch, CB16(y), y_pitch, npix, C, scale,
                                                                shift, mean, invstd, partial, rows_pb, 0));
  LAUNCH_OK();
  if (rows) *rows = grid;
  return 0;
}

int yb_colsum(const void* x, int64_t x_pitch, int64_t npix, int C, float* partial, int* rows, void* stream) {
  int threads, rows_pb, grid;
  size_t smem;
  if (reduce_geometry(C, npix, threads, rows_pb, grid, smem)) return -1;
  static bool attr = false;
  if (!attr) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(bn_act_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  YB_CHECK_CUDA(launch_pdl(bn_act_bwd_reduce_kernel, dim3(grid), dim3(threads), smem, ST(stream), CB16(x), x_pitch, nullptr, 0, npix, C, nullptr, nullptr,
                                                                nullptr, nullptr, partial, rows_pb, 1));
  LAUNCH_OK();
  if (rows) *rows = grid;
  return 0;
}

int yb_bn_bwd_finalize(const float* partial, int rows, int C, double count, float* dgamma, float* dbeta, float* coef,
                       int accumulate, void* stream) {
  YB_CHECK_CUDA(launch_pdl(bn_bwd_finalize_kernel, dim3((C + 31) / 32), dim3(1024), 0, ST(stream), partial, rows, C, count, dgamma, dbeta, coef,
                                                                   accumulate));
  LAUNCH_OK();
  return 0;
}

int yb_bn_act_bwd_apply(const void* da, int64_t da_pitch, const void* y, int64_t y_pitch, int64_t npix, int C,
                        const float* scale, const float* shift, const float* mean, const float* invstd,
                        const float* coef, void* dy, int64_t dy_pitch, void* stream) {
  YB_REQUIRE(C % 8 == 0 && da_pitch % 8 == 0 && y_pitch % 8 == 0 && dy_pitch % 8 == 0, "bn_act_bwd_apply: alignment");
  static const int occ_grid = ew_wave_grid(bn_act_bwd_apply_kernel<true>, kEwThreads, 1L << 40);
  const int grid = std::min(occ_grid, ew_blocks(npix * (C / 8), kEwThreads, 64));
  if (kEwThreads % (C / 8) == 0)
    YB_CHECK_CUDA(launch_pdl(bn_act_bwd_apply_kernel<true>, dim3(grid), dim3(kEwThreads), 0, ST(stream), CB16(da), da_pitch, CB16(y), y_pitch, npix, C, scale,
                                                                       shift, mean, invstd, coef, B16(dy), dy_pitch));
  else
    YB_CHECK_CUDA(launch_pdl(bn_act_bwd_apply_kernel<false>, dim3(grid), dim3(kEwThreads), 0, ST(stream), CB16(da), da_pitch, CB16(y), y_pitch, npix, C, scale,
                                                                        shift, mean, invstd, coef, B16(dy), dy_pitch));
  LAUNCH_OK();
  return 0;
}

int yb_maxpool5_fwd(const void* x, int64_t x_pitch, int N, int H, int W, int C, void* y, int64_t y_pitch,
                    uint8_t* argmax, void* stream) {
  YB_REQUIRE(C % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0, "maxpool5_fwd: alignment");
  const long npix = (long)N * H * W;
  maxpool5_fwd_kernel<<<ew_blocks(npix * (C / 8)), 256, 0, ST(stream)>>>(CB16(x), x_pitch, H, W, C, npix, B16(y), y_pitch,
                                                                          argmax);
  LAUNCH_OK();
  return 0;
}

int yb_maxpool5_bwd(const void* dy, int64_t dy_pitch, const uint8_t* argmax, int N, int H, int W, int C, void* dx,
                    int64_t dx_pitch, int accumulate, void* stream) {
  YB_REQUIRE(C % 8 == 0 && dy_pitch % 8 == 0 && dx_pitch % 8 == 0, "maxpool5_bwd: alignment");
  const long npix = (long)N * H * W;
  maxpool5_bwd_kernel<<<ew_blocks(npix * (C / 8)), 256, 0, ST(stream)>>>(CB16(dy), dy_pitch, argmax, H, W, C, npix,
                                                                          B16(dx), dx_pitch, accumulate);
  LAUNCH_OK();
  return 0;
}

int yb_sppf_pool3_fwd(const void* x, int64_t x_pitch, int N, int H, int W, int C, void* y1, void* y2, void* y3, int64_t y_pitch,
                      uint8_t* am1, uint8_t* am2, uint8_t* am3, void* stream) {
  YB_REQUIRE(C % 16 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0, "sppf_pool3_fwd: alignment (C %% 16, pitches %% 8)");
  if (am1 == nullptr && am2 == nullptr && am3 == nullptr) {  // inference: no arg-max planes
    const size_t smem_e = (size_t)H * W * 64;
    if (smem_e > 200 * 1024) return 1;
    static size_t attr_e = 0;
    if (smem_e > attr_e) {
      YB_CHECK_CUDA(cudaFuncSetAttribute(sppf_pool3_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
      attr_e = smem_e;
    }
    sppf_pool3_eval_kernel<<<N * (C / 16), H * W > 512 ? 512 : 256, smem_e, ST(stream)>>>(CB16(x), x_pitch, H, W, C, B16(y1),
                                                                                         B16(y2), B16(y3), y_pitch);
    LAUNCH_OK();
    return 0;
  }
  const size_t smem = (size_t)H * W * (32 + 32 + 16);
  if (smem > 200 * 1024) return 1;  // map too large for one CTA's shared memory: the caller chains yb_maxpool5_fwd instead
  static size_t attr = 0;
  if (smem > attr) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(sppf_pool3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  sppf_pool3_fwd_kernel<<<N * (C / 16), 256, smem, ST(stream)>>>(CB16(x), x_pitch, H, W, C, B16(y1), B16(y2), B16(y3), y_pitch,
                                                                am1, am2, am3);
  LAUNCH_OK();
  return 0;
}

int yb_sppf_pool3_bwd(const void* g1, const void* g2, const void* g3, int64_t g_pitch, const uint8_t* am1, const uint8_t* am2,
                      const uint8_t* am3, int N, int H, int W, int C, void* g0, int64_t g0_pitch, int accumulate, void* stream) {
  YB_REQUIRE(C % 16 == 0 && g_pitch % 8 == 0 && g0_pitch % 8 == 0, "sppf_pool3_bwd: alignment (C %% 16, pitches %% 8)");
  const size_t smem = (size_t)H * W * (64 + 64 + 16);
  if (smem > 200 * 1024) return 1;
  static size_t attr = 0;
  if (smem > attr) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(sppf_pool3_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  sppf_pool3_bwd_kernel<<<N * (C / 16), 256, smem, ST(stream)>>>(CB16(g1), CB16(g2), CB16(g3), g_pitch, am1, am2, am3, H, W, C,
                                                                B16(g0), g0_pitch, accumulate);
  LAUNCH_OK();
  return 0;
}

int yb_prep_input(const void* x, int dtype, int N, int H, int W, void* out, void* stream) {
  YB_REQUIRE(H % 2 == 0 && W % 2 == 0, "prep_input: odd image size");
  const long total = (long)N * (H / 2) * (W / 2);
  if (dtype == 0)
    prep_input_kernel<float, false><<<ew_blocks(total), 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(x), N, H, W,
                                                                               B16(out), H, W, 1.f, 1.f);
  else if (dtype == 1 && W % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0)
    prep_input_u8x4_kernel<<<ew_blocks(total / 4), 256, 0, ST(stream)>>>(reinterpret_cast<const uint8_t*>(x), N, H, W, B16(out));
  else if (dtype == 1)
    prep_input_kernel<uint8_t, false><<<ew_blocks(total), 256, 0, ST(stream)>>>(reinterpret_cast<const uint8_t*>(x), N, H,
                                                                                 W, B16(out), H, W, 1.f, 1.f);
  else
    YB_REQUIRE(false, "prep_input: dtype %d (0 = float32, 1 = uint8)", dtype);
  LAUNCH_OK();
  return 0;
}

int yb_prep_input_resized(const void* x, int dtype, int N, int Hs, int Ws, int H, int W, void* out, void* stream) {
  YB_REQUIRE(H % 2 == 0 && W % 2 == 0 && Hs > 0 && Ws > 0, "prep_input_resized: sizes");
  const long total = (long)N * (H / 2) * (W / 2);
  const float sh = (float)Hs / (float)H, sw = (float)Ws / (float)W;  // ATen area_pixel_compute_scale (align_corners=False)
  if (dtype == 0)
    prep_input_kernel<float, true><<<ew_blocks(total), 256, 0, ST(stream)>>>(reinterpret_cast<const float*>(x), N, H, W,
                                                                              B16(out), Hs, Ws, sh, sw);
  else if (dtype == 1)
    prep_input_kernel<uint8_t, true><<<ew_blocks(total), 256, 0, ST(stream)>>>(reinterpret_cast<const uint8_t*>(x), N, H, W,
                                                                                B16(out), Hs, Ws, sh, sw);
  else
    YB_REQUIRE(false, "prep_input_resized: dtype %d (0 = float32, 1 = uint8)", dtype);
  LAUNCH_OK();
  return 0;
}

int yb_head_grad_pack(const float* g, int B, int na, int H, int W, int no, void* dy, int Cpad, int accumulate, void* stream) {
  YB_REQUIRE(Cpad >= na * no, "head_grad_pack: Cpad");
  const long total = (long)B * H * W * Cpad;
  head_grad_pack_kernel<<<ew_blocks(total), 256, 0, ST(stream)>>>(g, na, (long)H * W, no, B16(dy), Cpad, total, accumulate);
  LAUNCH_OK();
  return 0;
}

int yb_reduce_rows(const float* partial, int rows, int64_t stride, int n, float* out, int accumulate, void* stream) {
  reduce_rows_kernel<<<(n + 127) / 128, 128, 0, ST(stream)>>>(partial, rows, stride, n, out, accumulate);
  LAUNCH_OK();
  return 0;
}

}  // extern "C"
