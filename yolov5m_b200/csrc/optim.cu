// Parameter-side kernels on the flat fp32 master buffer (sm_100a, HBM-bound):
//   * bf16 operand packing for the tcgen05 convs (forward [Cout][tap][Cin], dgrad [Cin][tap][Cout], stem space-to-depth)
//   * the optimiser tail of the reference training step (utils/training_utils.py:114-122, train.py:61):
//     GradScaler.unscale_ + clip_grad_norm_(max_norm) + Adam(lr, weight_decay as L2-in-gradient) fused over the flat
//     (all-reduced) gradient bucket, writing the bf16 forward operands in the same pass.
#include "../../include/yolov5m_b200.h"
#include "common.cuh"

#include <algorithm>

namespace yb {

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
  const long n4 = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16(src[i]);
}

// table[l] = {src_off, dst_off, Cout, taps, Cin, Cout_pad, dst_cum_begin, dst_cum_end}
// dst[dst_off + (ci*taps + tap)*Cout_pad + co] = co < Cout ? src[src_off + (co*taps + tap)*Cin + ci] : 0
// A per-(layer, tap) [Cout x Cin] -> [Cin x Cout_pad] transpose in 32 x 32 tiles through shared memory: 128-byte coalesced
// fp32 reads along ci, 64-byte coalesced bf16 writes along co.  (The first version decoded every element on its own -- a
// binary search over the layer table, two 64-bit divisions and a 4-byte gather whose neighbours sit taps*Cin*4 bytes apart:
// 0.29 ms per step for 21 M weights.)  Persistent blocks walk the tile list; the layer of a tile is found by advancing a
// cursor, since tile ids are ordered by layer.
__device__ __forceinline__ long repack_layer_tiles(const long long* t) {
  const long cout_pad = t[5], taps = t[3], cin = t[4];
  return taps * ((cout_pad + 31) / 32) * ((cin + 31) / 32);
}
__global__ void __launch_bounds__(256) repack_dgrad_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                           const long long* __restrict__ table, int nlayers) {
  __shared__ float tile[32][33];
  __shared__ long s_info[2];  // layer of the current tile, tile index inside that layer
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps: warp w handles tile rows w, w + 8, w + 16, w + 24
  int layer = 0;        // cursor (thread 0 only)
  long layer_begin = 0;  // first tile id of `layer`
  for (long t = blockIdx.x;; t += gridDim.x) {
    if (threadIdx.x == 0) {
      while (layer < nlayers) {
        const long n = repack_layer_tiles(table + layer * 8);
        if (t < layer_begin + n) break;
        layer_begin += n;
        ++layer;
      }
      s_info[0] = layer;
      s_info[1] = t - layer_begin;
    }
    __syncthreads();
    const int l = (int)s_info[0];
    long rem = s_info[1];
    if (l >= nlayers) break;  // uniform: past the last tile
    const long long* e = table + l * 8;
    const int cout = (int)e[2], taps = (int)e[3], cin = (int)e[4], cop = (int)e[5];
    const int tiles_ci = (cin + 31) / 32, tiles_co = (cop + 31) / 32;
    const int tci = (int)(rem % tiles_ci);
    rem /= tiles_ci;
    const int tco = (int)(rem % tiles_co);
    const int tap = (int)(rem / tiles_co);
    const int co0 = tco * 32, ci0 = tci * 32;
    const float* sp = src + e[0];
    bf16* dp = dst + e[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int co = co0 + ty + 8 * k, ci = ci0 + tx;
      tile[ty + 8 * k][tx] = (co < cout && ci < cin) ? sp[((long)co * taps + tap) * cin + ci] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ci = ci0 + ty + 8 * k, co = co0 + tx;
      if (ci < cin && co < cop) dp[((long)ci * taps + tap) * cop + co] = __float2bfloat16(tile[tx][ty + 8 * k]);
    }
    // the next iteration's first __syncthreads orders these tile reads before the next tile's writes
  }
}

// stem: w6 fp32 [Cout][6][6][3] (channels-last OIHW) -> w3 bf16 [Cout][3*3][16], channel (r*2+s)*3+c, 12..15 zero
__global__ void repack_stem_kernel(const float* __restrict__ w6, bf16* __restrict__ w3, int Cout) {
  const int total = Cout * 9 * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ch = i & 15;
    const int tap = (i >> 4) % 9;
    const int co = (i >> 4) / 9;
    float v = 0.f;
    if (ch < 12) {
      const int c = ch % 3, rs = ch / 3, r = rs >> 1, s = rs & 1;
      const int a = tap / 3, b = tap % 3;
      v = w6[((co * 6 + 2 * a + r) * 6 + 2 * b + s) * 3 + c];
    }
    w3[i] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------------------------- clip + Adam
__global__ void sqnorm_partial_kernel(const float* __restrict__ g, long n, float* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  const long n4 = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
  }
}

// norm_out[0] = grad_scale * sqrt(sum partial)  (the norm of the UNSCALED, averaged gradient)
// the optimiser's step counter advances only when the gradient norm is finite: a step with inf / NaN gradients is skipped
// entirely, like GradScaler.step (utils/training_utils.py:119)
__global__ void counter_inc_kernel(long long* c, const float* norm) {
  if (norm == nullptr || isfinite(norm[0])) c[0] += 1;
}

// dst += src over the flat fp32 gradient bucket (gradient accumulation over micro-batches, training_utils.py:88-90,116)
__global__ void accumulate_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, long n) {
  const long n4 = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
  for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

__global__ void sqnorm_final_kernel(const float* __restrict__ partial, int n, float grad_scale, float* __restrict__ norm_out) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += (double)partial[i];
  norm_out[0] = grad_scale * (float)sqrt(s);
}

// g' = g*grad_scale*clip,  clip = min(1, max_norm/(norm+1e-6))   (torch.nn.utils.clip_grad_norm_)
// Adam (torch.optim.Adam, weight_decay = L2 added to the gradient):
//   g' += wd*p; m = b1*m+(1-b1)*g'; v = b2*v+(1-b2)*g'^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float wd,
                                 float bc1, float bc2_sqrt, const long long* __restrict__ step_dev, float grad_scale,
                                 float max_norm, const float* __restrict__ norm, bf16* __restrict__ w_bf16) {
  if (step_dev != nullptr) {  // device-resident step counter (CUDA-graph replay): bias corrections computed here
    const double t = (double)step_dev[0];
    bc1 = 1.f - (float)pow((double)b1, t);
    bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  }
  // inf / NaN gradients (norm is computed from sum g^2, so any non-finite element makes it non-finite): skip the whole
  // update -- parameters, moments and bf16 operands stay as they are (GradScaler.step semantics, training_utils.py:119)
  if (norm != nullptr && !isfinite(norm[0])) return;
  float clip = 1.f;
  if (max_norm > 0.f && norm != nullptr) clip = fminf(1.f, max_norm / (norm[0] + 1e-6f));
  const float gs = grad_scale * clip;
  const float step = lr / bc1;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * gs);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
    if (w_bf16 != nullptr) w_bf16[i] = __float2bfloat16(pi);
  }
}

static int blocks_for(long n, int per = 8) {
  int sms = 148, dev = 0;
  static int cached = 0;
  if (!cached) {
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      cached = sms;
    else
      cached = 148;
  }
  return (int)std::max<long>(1, std::min<long>((n + 255) / 256, (long)cached * per));
}

}  // namespace yb

using namespace yb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int yb_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  cast_bf16_kernel<<<blocks_for(n / 4 + 1), 256, 0, ST(stream)>>>(src, reinterpret_cast<bf16*>(dst), n);
  YB_LAUNCHED();
  return 0;
}

int yb_repack_dgrad(const float* src, void* dst, const int64_t* table, int nlayers, int64_t total, void* stream) {
  (void)total;
  repack_dgrad_kernel<<<148 * 8, 256, 0, ST(stream)>>>(src, reinterpret_cast<bf16*>(dst),
                                                       reinterpret_cast<const long long*>(table), nlayers);
  YB_LAUNCHED();
  return 0;
}

int yb_repack_stem(const float* w6, void* w3, int Cout, void* stream) {
  repack_stem_kernel<<<(Cout * 144 + 255) / 256, 256, 0, ST(stream)>>>(w6, reinterpret_cast<bf16*>(w3), Cout);
  YB_LAUNCHED();
  return 0;
}

int yb_grad_norm(const float* g, int64_t n, float grad_scale, float* partial, int partial_len, float* norm_out,
                 void* stream) {
  const int blocks = std::min(blocks_for(n / 4 + 1, 4), partial_len);
  YB_REQUIRE(blocks >= 1, "grad_norm: partial_len");
  sqnorm_partial_kernel<<<blocks, 256, 0, ST(stream)>>>(g, n, partial);
  YB_LAUNCHED();
  sqnorm_final_kernel<<<1, 1, 0, ST(stream)>>>(partial, blocks, grad_scale, norm_out);
  YB_LAUNCHED();
  return 0;
}

int yb_counter_inc(int64_t* counter, const float* norm, void* stream) {
  counter_inc_kernel<<<1, 1, 0, ST(stream)>>>(reinterpret_cast<long long*>(counter), norm);
  YB_LAUNCHED();
  return 0;
}

int yb_accumulate_f32(float* dst, const float* src, int64_t n, void* stream) {
  accumulate_f32_kernel<<<blocks_for(n / 4 + 1), 256, 0, ST(stream)>>>(dst, src, n);
  YB_LAUNCHED();
  return 0;
}

int yb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int64_t step, const int64_t* step_dev, float grad_scale, float max_norm,
                 const float* norm, void* w_bf16, void* stream) {
  YB_REQUIRE(step >= 1 || step_dev != nullptr, "adam_step: step must start at 1");
  if (step < 1) step = 1;
  const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
  const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adam_step_kernel<<<blocks_for(n), 256, 0, ST(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s,
                                                           reinterpret_cast<const long long*>(step_dev), grad_scale, max_norm, norm, reinterpret_cast<bf16*>(w_bf16));
  YB_LAUNCHED();
  return 0;
}

}  // extern "C"
