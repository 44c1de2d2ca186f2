// Weight-gradient (wgrad) convolution on tcgen05 + TMA (sm_100a).
//
// Replaces the cuDNN wgrad that autograd runs for every nn.Conv2d of the reference
// (model.py:16 CBL conv, model.py:162 head conv; backward triggered at utils/training_utils.py:114).
//
//   dW[co][tap][ci] = sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*s + kh - p, wo*s + kw - p, ci]
//
// GEMM view per tap: D[M = co][N = ci] += A[M][K = pixel] * B[N][K = pixel].  Both operands sit in
// HBM as NHWC, i.e. the GEMM-K dimension (pixels) is the *slow* one: they are MN-major UMMA operands.
// TMA drops [KP pixels x KC channels] boxes (KC*2 bytes = the swizzle span) into shared memory; the
// canonical MN-major layout then is   ((KC,n),(8,k)) : ((1,LBO),(KC,SBO))   with
//   SBO = 8 * KC * 2 bytes  (next group of 8 pixels inside a box),
//   LBO = KP * KC * 2 bytes (next KC-channel box).
// The reduction over pixels is split across CTAs (split-K); each work item writes an fp32 partial tile
// and a second kernel reduces the partials in a fixed order (deterministic, no atomics).
//
// CTA = 256 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..7 epilogue.
#include "conv_wgrad.cuh"

#include <algorithm>
#include <cstring>

namespace yb {

static constexpr int kThreads = 256;
static constexpr int kMaxStages = 8;
static constexpr int kBarRegion = 1024;

struct WItem {
  int mt, nt, tap, sp;
};

__device__ __forceinline__ WItem decode_item(const WgradKParams& p, int t) {
  WItem c;
  c.sp = t % p.splits;
  t /= p.splits;
  c.tap = t % p.ntaps;
  t /= p.ntaps;
  c.nt = t % p.n_tiles;
  c.mt = t / p.n_tiles;
  return c;
}

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* a_smem = smem + kBarRegion;  // A stages first: over-reads of unused M rows stay inside the CTA's smem
  uint8_t* b_smem = a_smem + (size_t)p.stages * p.a_stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.m_tiles * p.n_tiles * p.ntaps * p.splits;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmB[i]);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const WItem it = decode_item(p, t);
        const ConvTap tap = p.taps[it.tap];
        const int pt0 = (int)(((long)it.sp * p.ptiles) / p.splits);
        const int pt1 = (int)(((long)(it.sp + 1) * p.ptiles) / p.splits);
        const int a_boxes = min(p.a_boxes, (p.Cout - it.mt * 128 + p.KCA - 1) / p.KCA);
        const uint32_t tx = (uint32_t)(a_boxes * p.KCA + p.b_boxes * p.KCB) * 2u * p.KP;
        for (int pt = pt0; pt < pt1; ++pt) {
          int m = pt;
          const int w0 = (m % p.tiles_w) * p.PW;
          m /= p.tiles_w;
          const int h0 = (m % p.tiles_h) * p.PH;
          const int n0 = (m / p.tiles_h) * p.PN;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], tx);
          uint8_t* as = a_smem + (size_t)s * p.a_stage_bytes;
          uint8_t* bs = b_smem + (size_t)s * p.b_stage_bytes;
          for (int j = 0; j < a_boxes; ++j)
            tma_load_4d(&p.tmA, &full_bar[s], as + (size_t)j * p.KP * 2 * p.KCA, it.mt * 128 + j * p.KCA, w0, h0, n0);
          for (int j = 0; j < p.b_boxes; ++j)
            tma_load_4d(&p.tmB[tap.map], &full_bar[s], bs + (size_t)j * p.KP * 2 * p.KCB,
                        it.nt * p.BLOCK_N + j * p.KCB, w0 + tap.dw, h0 + tap.dh, n0);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.BLOCK_N, 1, 1);  // both operands MN-major
      const uint32_t rbA = 2u * p.KCA, rbB = 2u * p.KCB;
      const uint32_t ltA = swizzle_layout_type(rbA), ltB = swizzle_layout_type(rbB);
      const uint32_t lboA = (uint32_t)p.KP * rbA, lboB = (uint32_t)p.KP * rbB;
      const int kinner = p.KP / 16;
      int s = 0;
      uint32_t ph = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
        const WItem it = decode_item(p, t);
        const int pt0 = (int)(((long)it.sp * p.ptiles) / p.splits);
        const int pt1 = (int)(((long)(it.sp + 1) * p.ptiles) / p.splits);
        const int ab = iter & 1;
        const uint32_t aph = (iter >> 1) & 1;
        mbar_wait(&tempty_bar[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * 256;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_smem + (size_t)s * p.a_stage_bytes);
          const uint32_t b_addr = smem_u32(b_smem + (size_t)s * p.b_stage_bytes);
          for (int k = 0; k < kinner; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * 16 * rbA, lboA, 8 * rbA, ltA);
            const uint64_t db = make_smem_desc(b_addr + k * 16 * rbB, lboB, 8 * rbB, ltB);
            umma_bf16(d_tmem, da, db, idesc, (pt != pt0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull_bar[ab]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> fp32 partial tile
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int iter = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
      const WItem it = decode_item(p, t);
      const int co = it.mt * 128 + r;
      const int ab = iter & 1;
      const uint32_t aph = (iter >> 1) & 1;
      float* dst = p.partial + ((size_t)it.sp * p.Cout + co) * p.ldo + p.taps[it.tap].kbase + it.nt * p.BLOCK_N;
      const int ncols = min(p.BLOCK_N, p.Cin - it.nt * p.BLOCK_N);
      mbar_wait(&tfull_bar[ab], aph);  // the plan guarantees pt1 > pt0 (splits <= pixel tiles)
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256;
      for (int cc = 0; cc * 16 < ncols; ++cc) {
        uint32_t vr[16];
        tmem_ld16(t_addr + cc * 16, vr);
        tmem_ld_wait();
        if (co < p.Cout) {
          float4* o = reinterpret_cast<float4*>(dst + cc * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_float4(__uint_as_float(vr[4 * j]), __uint_as_float(vr[4 * j + 1]), __uint_as_float(vr[4 * j + 2]),
                               __uint_as_float(vr[4 * j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[ab]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[map ? map[i] : i] (+)= sum_s partial[s][i]   (fixed summation order => deterministic)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, long n, long nfull,
                                    float* __restrict__ out,
                                    const int* __restrict__ map, int accumulate) {
  const long stride = (long)gridDim.x * blockDim.x;
  if (map == nullptr) {
    const long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 acc = reinterpret_cast<const float4*>(partial)[i];
      for (int s = 1; s < splits; ++s) {
        const float4 v = reinterpret_cast<const float4*>(partial + (size_t)s * nfull)[i];
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
      float4* o = reinterpret_cast<float4*>(out) + i;
      if (accumulate) {
        const float4 c = *o;
        acc.x += c.x;
        acc.y += c.y;
        acc.z += c.z;
        acc.w += c.w;
      }
      *o = acc;
    }
  } else {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int d = map[i];
      if (d < 0) continue;
      float acc = partial[i];
      for (int s = 1; s < splits; ++s) acc += partial[(size_t)s * nfull + i];
      out[d] = accumulate ? out[d] + acc : acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
static int pick_kc(int C) { return (C % 64 == 0) ? 64 : (C % 32 == 0 ? 32 : 16); }

// pixel patch with PW*PH*PN a multiple of 16 and <= maxp, minimising the padded pixel count
static void choose_kpatch(int W, int H, int NB, int maxp, int& PW, int& PH, int& PN) {
  double best = -1;
  for (int pw = 1; pw <= std::min(W, 256); ++pw)
    for (int ph = 1; ph <= std::min(H, 256) && pw * ph <= maxp; ++ph)
      for (int pn = 1; pn <= std::min(NB, 256) && pw * ph * pn <= maxp; ++pn) {
        if (pn > 1 && (ph < H || pw < W)) continue;  // span images only with whole-image patches
        if (ph > 1 && pw < W) continue;              // span rows only with whole-row patches
        const int kp = pw * ph * pn;
        if (kp % 16) continue;
        const double padded = (double)((W + pw - 1) / pw) * pw * ((H + ph - 1) / ph) * ph * ((NB + pn - 1) / pn) * pn;
        const double tiles = padded / kp;
        const double score = padded + tiles * 24;  // fewer, fuller tiles
        if (best < 0 || score < best) {
          best = score;
          PW = pw;
          PH = ph;
          PN = pn;
        }
      }
  if (best < 0) {  // fall back: partial rows (any pw with pw % 16 == 0 works through OOB fill)
    PW = 16;
    PH = 1;
    PN = 1;
  }
}

static int make_map(CUtensorMap* m, const TView& v, int KC, int PW, int PH, int PN, int py, int px, int sy, int sx) {
  const bf16* base = reinterpret_cast<const bf16*>(v.ptr) + ((long)py * v.W + px) * v.pitch;
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)(v.W / sx), (uint64_t)(v.H / sy), (uint64_t)v.N};
  uint64_t strides[3] = {(uint64_t)v.pitch * sx * 2, (uint64_t)v.pitch * v.W * sy * 2,
                         (uint64_t)v.pitch * v.W * v.H * 2};
  uint32_t box[4] = {(uint32_t)KC, (uint32_t)PW, (uint32_t)PH, (uint32_t)PN};
  return encode_tmap(m, base, 4, dims, strides, box, 2 * KC, 2);
}

int wgrad_max_grid();

size_t wgrad_workspace_floats(int Cout, int Cin, int ks, int splits) { return (size_t)splits * Cout * ks * ks * Cin; }

int wgrad_plan(WgradPlan& pl, const TView& x, const TView& dy, int ks, int stride, float* partial, size_t partial_floats,
               int max_splits) {
  memset(&pl, 0, sizeof(pl));
  WgradKParams& kp = pl.kp;
  YB_REQUIRE((ks == 1 || ks == 3) && (stride == 1 || stride == 2), "wgrad: ks=%d stride=%d unsupported", ks, stride);
  YB_REQUIRE(x.C % 16 == 0 && dy.C % 16 == 0 && x.pitch % 8 == 0 && dy.pitch % 8 == 0, "wgrad: channel alignment");
  YB_REQUIRE(dy.H == x.H / stride && dy.W == x.W / stride && dy.N == x.N, "wgrad: geometry");
  kp.Cout = dy.C;
  kp.Cin = x.C;
  kp.KCA = pick_kc(dy.C);
  kp.KCB = pick_kc(x.C);
  kp.a_boxes = 128 / kp.KCA;
  kp.m_tiles = (dy.C + 127) / 128;
  {
    const int parts = (x.C + 255) / 256;
    int bn = ((x.C + parts - 1) / parts + kp.KCB - 1) / kp.KCB * kp.KCB;
    kp.BLOCK_N = bn;
    kp.n_tiles = (x.C + bn - 1) / bn;
    kp.b_boxes = bn / kp.KCB;
  }
  YB_REQUIRE(kp.BLOCK_N % 16 == 0 && kp.BLOCK_N <= 256, "wgrad: BLOCK_N=%d", kp.BLOCK_N);
  YB_REQUIRE(x.C % kp.BLOCK_N == 0, "wgrad: Cin=%d not a multiple of the N tile %d", x.C, kp.BLOCK_N);
  // stage = (128 + BLOCK_N) channels x KP pixels x 2 bytes: shrink the pixel patch until >= 3 stages fit
  for (int maxp = 128; maxp >= 16; maxp -= 16) {
    choose_kpatch(dy.W, dy.H, dy.N, maxp, kp.PW, kp.PH, kp.PN);
    kp.KP = kp.PW * kp.PH * kp.PN;
    if ((size_t)(128 + kp.BLOCK_N) * 2 * kp.KP * 3 <= 220 * 1024) break;
  }
  YB_REQUIRE(kp.KP % 16 == 0 && kp.KP <= 256, "wgrad: pixel patch %dx%dx%d", kp.PW, kp.PH, kp.PN);
  kp.tiles_w = (dy.W + kp.PW - 1) / kp.PW;
  kp.tiles_h = (dy.H + kp.PH - 1) / kp.PH;
  kp.tiles_n = (dy.N + kp.PN - 1) / kp.PN;
  kp.ptiles = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  if (make_map(&kp.tmA, dy, kp.KCA, kp.PW, kp.PH, kp.PN, 0, 0, 1, 1)) return -1;
  const int pad = ks / 2;
  int nt = 0;
  if (stride == 1) {
    if (make_map(&kp.tmB[0], x, kp.KCB, kp.PW, kp.PH, kp.PN, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmB[i] = kp.tmB[0];
    for (int kh = 0; kh < ks; ++kh)
      for (int kw = 0; kw < ks; ++kw) {
        kp.taps[nt] = ConvTap{0, (int8_t)(kw - pad), (int8_t)(kh - pad), 0, (int32_t)((kh * ks + kw) * x.C)};
        ++nt;
      }
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        if (make_map(&kp.tmB[py * 2 + px], x, kp.KCB, kp.PW, kp.PH, kp.PN, py, px, 2, 2)) return -1;
    for (int kh = 0; kh < ks; ++kh)
      for (int kw = 0; kw < ks; ++kw) {
        const int oy = kh - pad, ox = kw - pad;
        const int py = oy & 1, px = ox & 1;
        const int dh = (oy - py) / 2, dw = (ox - px) / 2;
        kp.taps[nt] = ConvTap{(int8_t)(py * 2 + px), (int8_t)dw, (int8_t)dh, 0, (int32_t)((kh * ks + kw) * x.C)};
        ++nt;
      }
  }
  kp.ntaps = nt;
  kp.ldo = nt * x.C;
  kp.a_stage_bytes = 128u * 2u * kp.KP;
  kp.b_stage_bytes = (uint32_t)kp.BLOCK_N * 2u * kp.KP;
  kp.a_stage_bytes = (kp.a_stage_bytes + 1023u) & ~1023u;
  kp.b_stage_bytes = (kp.b_stage_bytes + 1023u) & ~1023u;
  const size_t budget = 227 * 1024 - 1024 - kBarRegion;
  int stages = (int)(budget / (kp.a_stage_bytes + kp.b_stage_bytes));
  stages = std::min(stages, kMaxStages);
  YB_REQUIRE(stages >= 2, "wgrad: tile does not fit in shared memory");
  kp.stages = stages;
  pl.smem = (int)(1024 + kBarRegion + (size_t)stages * (kp.a_stage_bytes + kp.b_stage_bytes));
  pl.smem = std::max(pl.smem, 120 * 1024);
  // split-K: fill the machine, at least ~4 pixel tiles per item
  const int base_items = kp.m_tiles * kp.n_tiles * kp.ntaps;
  const int sms = wgrad_max_grid();
  int splits = std::max(1, (2 * sms + base_items - 1) / base_items);
  splits = std::min(splits, std::max(1, kp.ptiles / 4));  // >= 4 pixel tiles per item (never an empty split)
  if (max_splits > 0) splits = std::min(splits, max_splits);
  const size_t per_split = (size_t)kp.Cout * kp.ldo;
  splits = (int)std::min<size_t>(splits, partial_floats / per_split);
  YB_REQUIRE(splits >= 1, "wgrad: workspace too small (%zu floats, need >= %zu)", partial_floats, per_split);
  kp.splits = splits;
  kp.partial = partial;
  pl.grid = std::min(base_items * splits, sms);
  return 0;
}

static int g_sms = 0;
int wgrad_max_grid() {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms <= 0)
      g_sms = 148;
  }
  return g_sms;
}

int wgrad_run(const WgradPlan& pl, float* out, int out_rows, const int* map, int accumulate, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  conv_wgrad_kernel<<<pl.grid, kThreads, pl.smem, st>>>(pl.kp);
  YB_LAUNCHED();
  YB_REQUIRE(out_rows > 0 && out_rows <= pl.kp.Cout, "wgrad: out_rows=%d", out_rows);
  const long n = (long)out_rows * pl.kp.ldo;
  const long nfull = (long)pl.kp.Cout * pl.kp.ldo;
  const int blocks = (int)std::min<long>((n / 4 + 255) / 256 + 1, 4L * wgrad_max_grid());
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(pl.kp.partial, pl.kp.splits, n, nfull, out, map, accumulate);
  YB_LAUNCHED();
  return 0;
}

}  // namespace yb
