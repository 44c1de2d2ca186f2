// Weight-gradient (wgrad) convolution on tcgen05 + TMA (sm_100a).
//
// Replaces the cuDNN wgrad that autograd runs for every nn.Conv2d of the reference
// (model.py:16 CBL conv, model.py:162 head conv; backward triggered at utils/training_utils.py:114).
//
//   dW[co][tap][ci] = sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*s + kh - p, wo*s + kw - p, ci]
//
// GEMM view:  D[M = (tap, ci)][N = co] += A[M][K = pixel] * B[N][K = pixel]
//   A = the conv input x, shifted per tap;  B = dy.  Both sit in HBM as NHWC, i.e. the GEMM-K dimension (pixels) is the
//   slow one: both are MN-major UMMA operands.  TMA drops [KP pixels x 64 channels] boxes (128 bytes per pixel row = the
//   swizzle span) into shared memory; channel counts that are not multiples of 64 are zero-padded by the TMA
//   out-of-bounds fill (no bytes fetched for them).  Canonical MN-major SWIZZLE_128B layout of a stack of boxes:
//       ((64, n), (8, k)) : ((1, LBO), (64, SBO)),  SBO = 8 * 128 B (next 8 pixels inside a box),  LBO = KP * 128 B (next box)
//   An M tile (128 rows) = two boxes = two (tap, 64-channel chunk) pairs, so small-Cin layers waste no tensor-core rows
//   on padding between taps, and the dy box(es) of a pipeline stage are shared by up to MT M tiles whose accumulators
//   live side by side in TMEM (MT * BLOCK_N <= 512 columns): dy is fetched once per stage instead of once per tap.
// The reduction over pixels is split across CTAs (split-K).  Work items that cover the same pixel range run
// concurrently (item index = split * (m_groups * n_tiles) + ...), so x / dy tiles are served from L2 after the first
// touch.  Each item writes an fp32 partial tile [Mpad][Npad]; a second kernel reduces the partials in a fixed order
// (deterministic, no atomics) and transposes them into dW's [co][tap][ci] layout.
//
// CTA = 512 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..15 epilogue.
#include "conv_wgrad.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace yb {

static constexpr int kThreads = 512;
static constexpr int kEpiGroups = 3;  // epilogue warps per TMEM lane quarter (interleaved over 16-column chunks)
static constexpr int kMaxStages = 6;
static constexpr int kBarRegion = 1024;
static constexpr int kBoxC = 64;  // channels per TMA box (128-byte swizzle span)

struct WItem {
  int sp, mg, nt;
};

__device__ __forceinline__ WItem decode_item(const WgradKParams& p, int t) {
  WItem c;
  int m;
  fdivmod(t, p.fd_nt, m, c.nt);
  fdivmod(m, p.fd_mg, c.sp, c.mg);
  return c;
}

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);
  uint8_t* stage_smem = smem + kBarRegion;  // per stage: [2*MT x boxes][nb dy boxes]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.splits * p.m_groups * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4 * kEpiGroups);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmX[i]);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, elected lane issues)
    {
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const WItem it = decode_item(p, t);
        const int pt0 = (int)(((long)it.sp * p.ptiles) / p.splits);
        const int pt1 = (int)(((long)(it.sp + 1) * p.ptiles) / p.splits);
        const int box0 = it.mg * p.MT * 2;
        const int box1 = min(p.boxes_total, box0 + p.MT * 2);
        const uint32_t tx = (uint32_t)(box1 - box0 + p.nb) * p.box_bytes;
        for (int pt = pt0; pt < pt1; ++pt) {
          int m, wi, hi, ni;
          fdivmod(pt, p.fd_tw, m, wi);
          fdivmod(m, p.fd_th, ni, hi);
          const int w0 = wi * p.PW, h0 = hi * p.PH, n0 = ni * p.PN;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) {
            mbar_expect_tx(&full_bar[s], tx);
            uint8_t* as = stage_smem + (size_t)s * p.stage_bytes;
            uint8_t* bs = as + (size_t)2 * p.MT * p.box_bytes;
            for (int j = 0; j < p.nb; ++j)
              tma_load_4d(&p.tmDY, &full_bar[s], bs + (size_t)j * p.box_bytes, it.nt * p.BLOCK_N + j * kBoxC, w0, h0, n0);
            int tapi, cb;
            fdivmod(box0, p.fd_cb, tapi, cb);
            for (int b = box0; b < box1; ++b) {
              tma_load_4d(&p.tmX[p.taps[tapi].map], &full_bar[s], as + (size_t)(b - box0) * p.box_bytes, cb * kBoxC,
                          w0 + p.taps[tapi].dw, h0 + p.taps[tapi].dh, n0);
              if (++cb == p.cboxes) {
                cb = 0;
                ++tapi;
              }
            }
          }
          __syncwarp();
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // whole warp, warp-uniform values, one elected lane issues (see conv_igemm.cu); descriptors advance by integer adds
    {
      const bool leader = elect_one();
      const uint32_t lt = swizzle_layout_type(128);
      const int kinner = p.KP / 16;
      const uint64_t desc0 = make_smem_desc(smem_u32(stage_smem), p.box_bytes, 1024, lt);
      const uint32_t stage_step = p.stage_bytes >> 4, tile_step = (2u * p.box_bytes) >> 4;
      const uint32_t b_off = (2u * p.MT * p.box_bytes) >> 4;
      int s = 0;
      uint32_t ph = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
        const WItem it = decode_item(p, t);
        const int pt0 = (int)(((long)it.sp * p.ptiles) / p.splits);
        const int pt1 = (int)(((long)(it.sp + 1) * p.ptiles) / p.splits);
        const int mts = min(p.MT, p.m_tiles - it.mg * p.MT);
        // MMA N = the real output channels of this N tile (rounded up to 16), not the 64-channel box multiple: the tail
        // of the last dy box is OOB zero fill and never read (Cout = 96: N = 96 instead of 128, Cout = 48: 48 instead of 64)
        const int n_mma = p.exact_n ? min(p.BLOCK_N, ((p.Cout + 15) & ~15) - it.nt * p.BLOCK_N) : p.BLOCK_N;
        const uint32_t idesc = make_idesc_bf16(128, n_mma, 1, 1);  // both operands MN-major
        mbar_wait(tempty_bar, (iter & 1) ^ 1);
        tc_fence_after();
        uint32_t acc = 0;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t a_desc = desc0 + (uint64_t)(s * stage_step);
          const uint64_t b_desc = a_desc + b_off;
          if (leader) {
            for (int mt = 0; mt < mts; ++mt) {
              const uint32_t d_tmem = tmem_base + mt * p.BLOCK_N;
              uint64_t da = a_desc + (uint64_t)(mt * tile_step), db = b_desc;
              umma_bf16(d_tmem, da, db, idesc, acc);
              for (int k = 1; k < kinner; ++k) {
                da += 128;  // 16 pixels x 128 B = 2048 B
                db += 128;
                umma_bf16(d_tmem, da, db, idesc, 1u);
              }
            }
            umma_commit(&empty_bar[s]);
          }
          __syncwarp();
          acc = 1u;
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        if (leader) umma_commit(tfull_bar);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> fp32 partial tile
    const int q = warp & 3;
    const int eg = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    int iter = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
      const WItem it = decode_item(p, t);
      const int mts = min(p.MT, p.m_tiles - it.mg * p.MT);
      const int box0 = it.mg * p.MT * 2;
      const int ncols = min(p.BLOCK_N, ((p.Cout + 15) & ~15) - it.nt * p.BLOCK_N);
      mbar_wait(tfull_bar, iter & 1);  // the plan guarantees pt1 > pt0 (splits <= pixel tiles)
      tc_fence_after();
      for (int mt = 0; mt < mts; ++mt) {
        const int box = box0 + mt * 2 + (r >> 6);
        int tapi, cbi;
        fdivmod(box, p.fd_cb, tapi, cbi);
        const int ci = cbi * kBoxC + (r & 63);
        const bool valid = box < p.boxes_total && ci < p.Cin;
        float* dst = p.partial + ((size_t)it.sp * p.Mpad + (size_t)tapi * p.Cin_pad + ci) * p.Npad + it.nt * p.BLOCK_N;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + mt * p.BLOCK_N;
        for (int cc = eg; cc * 16 < ncols; cc += kEpiGroups) {
          uint32_t vr[16];
          tmem_ld16(t_addr + cc * 16, vr);
          tmem_ld_wait();
          if (valid) {
            // two 32-byte stores (st.global.v8): full L2 sectors; rows of the partial tile are Npad * 4 bytes apart
            float* o = dst + cc * 16;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "r"(vr[0]), "r"(vr[1]),
                         "r"(vr[2]), "r"(vr[3]), "r"(vr[4]), "r"(vr[5]), "r"(vr[6]), "r"(vr[7])
                         : "memory");
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o + 8), "r"(vr[8]), "r"(vr[9]),
                         "r"(vr[10]), "r"(vr[11]), "r"(vr[12]), "r"(vr[13]), "r"(vr[14]), "r"(vr[15])
                         : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[map ? map[i] : i] (+)= sum_s partial[s][tap*Cin_pad + ci][co],  i = co*ldo + tap*Cin + ci
// Block = 32 (co) x 32 threads; the 32 rows of threads are split into KR k-rows x SL split-lanes (KR * SL = 32): lane sl
// sums splits sl, sl+SL, ... in order, then a fixed shared-memory tree combines the lanes -> deterministic for a given
// plan.  Small outputs (many splits) use many split-lanes so that the reduction still fills the machine; reads are
// coalesced along co, the transposed store goes through shared memory.
__global__ void __launch_bounds__(1024) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Mpad, int Npad, int Cin,
                                    int Cin_pad, int ldo, int out_rows, int SL, float* __restrict__ out,
                                    const int* __restrict__ map, int accumulate) {
  __shared__ float red[32][33];
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int KR = 32 / SL;
  const int kk = ty / SL, sl = ty - kk * SL;
  const int k = blockIdx.x * KR + kk, co = blockIdx.y * 32 + tx;
  const size_t sstride = (size_t)Mpad * Npad;
  float acc = 0.f;
  if (k < ldo && co < out_rows) {
    const int tapi = k / Cin;
    const float* src = partial + ((size_t)tapi * Cin_pad + (k - tapi * Cin)) * Npad + co;
    int s = sl;
    for (; s + 3 * SL < splits; s += 4 * SL) {
      const float v0 = src[(size_t)s * sstride], v1 = src[(size_t)(s + SL) * sstride];
      const float v2 = src[(size_t)(s + 2 * SL) * sstride], v3 = src[(size_t)(s + 3 * SL) * sstride];
      acc = (((acc + v0) + v1) + v2) + v3;
    }
    for (; s < splits; s += SL) acc += src[(size_t)s * sstride];
  }
  red[ty][tx] = acc;
  __syncthreads();
  for (int off = SL >> 1; off > 0; off >>= 1) {
    if (sl < off) red[ty][tx] += red[ty + off][tx];
    __syncthreads();
  }
  // transposed store: thread (a = ty, b = tx) -> co = co0 + a, k-row b (b < KR); totals sit in red[b * SL][a]
  const int co2 = blockIdx.y * 32 + ty, k2 = blockIdx.x * KR + tx;
  if (tx < KR && k2 < ldo && co2 < out_rows) {
    const float v = red[tx * SL][ty];
    long i = (long)co2 * ldo + k2;
    if (map != nullptr) i = map[i];
    if (i >= 0) out[i] = accumulate ? out[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ host
// pixel patch with PW*PH*PN a multiple of 16 and <= maxp, minimising the padded pixel count
static bool choose_kpatch(int W, int H, int NB, int maxp, int& PW, int& PH, int& PN) {
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
    return maxp >= 16;
  }
  return true;
}

static int make_map(CUtensorMap* m, const TView& v, int PW, int PH, int PN, int py, int px, int sy, int sx) {
  const bf16* base = reinterpret_cast<const bf16*>(v.ptr) + (long)py * v.rowp() + (long)px * v.pitch;
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)(v.W / sx), (uint64_t)(v.H / sy), (uint64_t)v.N};
  uint64_t strides[3] = {(uint64_t)v.pitch * sx * 2, (uint64_t)v.rowp() * sy * 2,
                         (uint64_t)v.rowp() * v.H * 2};
  uint32_t box[4] = {(uint32_t)kBoxC, (uint32_t)PW, (uint32_t)PH, (uint32_t)PN};
  return encode_tmap(m, base, 4, dims, strides, box, 128, 2);
}

static int pick_block_n(int Cout, int& n_tiles) {
  const int c64 = (Cout + kBoxC - 1) / kBoxC;  // dy boxes in total
  n_tiles = (c64 + 3) / 4;                     // <= 4 boxes (256 columns) per N tile
  return (c64 + n_tiles - 1) / n_tiles * kBoxC;
}

size_t wgrad_min_workspace_floats(int Cin, int Cout, int ks) {
  int n_tiles;
  const int bn = pick_block_n(Cout, n_tiles);
  const size_t cin_pad = (size_t)(Cin + kBoxC - 1) / kBoxC * kBoxC;
  return (size_t)(ks == 31 ? 3 : ks * ks) * cin_pad * n_tiles * bn;
}

int wgrad_plan(WgradPlan& pl, const TView& x, const TView& dy, int ks, int stride, float* partial, size_t partial_floats,
               int max_splits) {
  memset(&pl, 0, sizeof(pl));
  {
    const int rc = wgrad_patch_plan(pl, x, dy, ks, stride, partial, partial_floats, max_splits);
    if (rc <= 0) return rc;
    pl.kind = 0;
  }
  WgradKParams& kp = pl.kp;
  YB_REQUIRE(((ks == 1 || ks == 3) && (stride == 1 || stride == 2)) || (ks == 31 && stride == 1),
             "wgrad: ks=%d stride=%d unsupported", ks, stride);
  const int ksh = ks == 31 ? 3 : ks, ksw = ks == 31 ? 1 : ks;
  YB_REQUIRE(x.C % 8 == 0 && dy.C % 8 == 0 && x.pitch % 8 == 0 && dy.pitch % 8 == 0, "wgrad: channel alignment");
  YB_REQUIRE(dy.H == x.H / stride && dy.W == x.W / stride && dy.N == x.N, "wgrad: geometry");
  YB_REQUIRE((reinterpret_cast<uintptr_t>(partial) & 31) == 0, "wgrad: workspace must be 32-byte aligned");
  YB_REQUIRE(x.H % stride == 0 && x.W % stride == 0, "wgrad: odd input for stride 2");
  kp.Cout = dy.C;
  kp.Cin = x.C;
  kp.ntaps = ksh * ksw;
  kp.cboxes = (x.C + kBoxC - 1) / kBoxC;
  kp.Cin_pad = kp.cboxes * kBoxC;
  kp.boxes_total = kp.ntaps * kp.cboxes;
  kp.m_tiles = (kp.boxes_total + 1) / 2;
  kp.BLOCK_N = pick_block_n(dy.C, kp.n_tiles);
  kp.exact_n = wgrad_exact_n();
  kp.nb = kp.BLOCK_N / kBoxC;
  kp.Mpad = kp.ntaps * kp.Cin_pad;
  kp.Npad = kp.n_tiles * kp.BLOCK_N;
  kp.ldo = kp.ntaps * x.C;
  // pixel patch (GEMM-K chunk) and the number of M tiles that share one dy stage: minimise the L2 -> shared-memory
  // bytes per MMA (nb + 2*MT boxes feed MT tiles) under >= 2 (preferably >= 3) pipeline stages of shared memory
  const size_t budget = wgrad_smem_budget();
  const int mt_max = std::max(1, std::min(512 / kp.BLOCK_N, kp.m_tiles));
  double best = -1;
  const double real_px = (double)dy.W * dy.H * dy.N;
  for (int maxp = 128; maxp >= 16; maxp -= 16) {
    int PW, PH, PN;
    if (!choose_kpatch(dy.W, dy.H, dy.N, maxp, PW, PH, PN)) continue;
    const int KP = PW * PH * PN;
    if (KP % 16 || KP > 256) continue;
    const double padded = (double)((dy.W + PW - 1) / PW) * PW * ((dy.H + PH - 1) / PH) * PH * ((dy.N + PN - 1) / PN) * PN;
    for (int mt = mt_max; mt >= 1; --mt) {
      const size_t stage = (size_t)(kp.nb + 2 * mt) * KP * 128;
      const int stages = (int)std::min<size_t>(kMaxStages, budget / stage);
      if (stages < 2) continue;
      const int groups = (kp.m_tiles + mt - 1) / mt;
      const double boxes_per_tile = (double)(groups * kp.nb + 2 * kp.m_tiles) / kp.m_tiles;
      const double score = boxes_per_tile * (padded / real_px) * (1.0 + 12.0 / KP) * (stages >= 3 ? 1.0 : 1.3);
      if (best < 0 || score < best) {
        best = score;
        kp.PW = PW;
        kp.PH = PH;
        kp.PN = PN;
        kp.KP = KP;
        kp.MT = mt;
        kp.stages = stages;
      }
    }
  }
  YB_REQUIRE(best >= 0, "wgrad: no pixel patch fits in shared memory (W=%d H=%d N=%d)", dy.W, dy.H, dy.N);
  kp.m_groups = (kp.m_tiles + kp.MT - 1) / kp.MT;
  kp.MT = (kp.m_tiles + kp.m_groups - 1) / kp.m_groups;  // balance the groups
  kp.box_bytes = (uint32_t)kp.KP * 128u;
  kp.stage_bytes = (uint32_t)(kp.nb + 2 * kp.MT) * kp.box_bytes;
  kp.stages = (int)std::min<size_t>(kMaxStages, budget / kp.stage_bytes);
  kp.tiles_w = (dy.W + kp.PW - 1) / kp.PW;
  kp.tiles_h = (dy.H + kp.PH - 1) / kp.PH;
  kp.tiles_n = (dy.N + kp.PN - 1) / kp.PN;
  kp.ptiles = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  if (make_map(&kp.tmDY, dy, kp.PW, kp.PH, kp.PN, 0, 0, 1, 1)) return -1;
  const int pad = ks / 2;
  int nt = 0;
  if (stride == 1) {
    if (make_map(&kp.tmX[0], x, kp.PW, kp.PH, kp.PN, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmX[i] = kp.tmX[0];
    for (int kh = 0; kh < ksh; ++kh)
      for (int kw = 0; kw < ksw; ++kw) {
        kp.taps[nt] = ConvTap{0, (int8_t)(kw - ksw / 2), (int8_t)(kh - ksh / 2), 0, (int32_t)((kh * ksw + kw) * x.C)};
        ++nt;
      }
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        if (make_map(&kp.tmX[py * 2 + px], x, kp.PW, kp.PH, kp.PN, py, px, 2, 2)) return -1;
    for (int kh = 0; kh < ks; ++kh)
      for (int kw = 0; kw < ks; ++kw) {
        const int oy = kh - pad, ox = kw - pad;
        const int py = oy & 1, px = ox & 1;
        const int dh = (oy - py) / 2, dw = (ox - px) / 2;
        kp.taps[nt] = ConvTap{(int8_t)(py * 2 + px), (int8_t)dw, (int8_t)dh, 0, (int32_t)((kh * ks + kw) * x.C)};
        ++nt;
      }
  }
  pl.smem = (int)(1024 + kBarRegion + (size_t)kp.stages * kp.stage_bytes);
  pl.smem = std::max(pl.smem, 120 * 1024);  // one CTA per SM: each CTA allocates all 512 TMEM columns
  // split-K: ONE wave of work items (one item per SM), >= 8 pixel tiles per item (never an empty split).  All TMEM columns
  // belong to the MT accumulators of the current item, so the epilogue of an item is not overlapped with the MMAs of the
  // next one: a second wave only adds an exposed epilogue and doubles the partial tiles the reduction has to read
  // (5.68 -> 5.20 ms over all layers, tools/bench_conv.py)
  const int base_items = kp.m_groups * kp.n_tiles;
  const int sms = wgrad_max_grid();
  int splits = std::max(1, (wgrad_waves() * sms) / base_items);
  splits = std::min(splits, std::max(1, kp.ptiles / 8));
  if (max_splits > 0) splits = std::min(splits, max_splits);
  const size_t per_split = (size_t)kp.Mpad * kp.Npad;
  splits = (int)std::min<size_t>(splits, partial_floats / per_split);
  {  // avoid a mostly empty last wave: items = base_items * splits should fill whole multiples of the SM count
    const int items = base_items * splits;
    const int waves = (items + sms - 1) / sms;
    if (waves > 1 && items < waves * sms * 0.85) splits = std::max(1, (waves - 1) * sms / base_items);
  }
  YB_REQUIRE(splits >= 1, "wgrad: workspace too small (%zu floats, need >= %zu)", partial_floats, per_split);
  kp.splits = splits;
  kp.partial = partial;
  kp.fd_nt = make_fdiv(kp.n_tiles);
  kp.fd_mg = make_fdiv(kp.m_groups);
  kp.fd_tw = make_fdiv(kp.tiles_w);
  kp.fd_th = make_fdiv(kp.tiles_h);
  kp.fd_cb = make_fdiv(kp.cboxes);
  pl.grid = std::min(base_items * splits, sms);
  return 0;
}

static int g_sms = 0;
int wgrad_waves() {
  static const int v = [] {
    const char* e = getenv("YB_WGRAD_WAVES");
    return std::max(1, e != nullptr ? atoi(e) : 1);
  }();
  return v;
}

int wgrad_exact_n() {
  static const int v = [] {
    const char* e = getenv("YB_WGRAD_EXACT_N");
    return e != nullptr ? atoi(e) : 1;
  }();
  return v;
}

size_t wgrad_smem_budget() {
  static const size_t budget = [] {
    const char* e = getenv("YB_WGRAD_SMEM_KB");
    int kb = e != nullptr ? atoi(e) : 227;
    kb = std::min(227, std::max(128, kb));
    return (size_t)kb * 1024 - 1024 - kBarRegion;
  }();
  return budget;
}

int wgrad_max_grid() {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms <= 0)
      g_sms = 148;
  }
  return g_sms;
}

// phase: 0 = both kernels on `st`; 1 = the tcgen05 split-K kernel only; 2 = the partial reduction only (the caller orders the
// two across streams: the reduction is a small L2-resident kernel that can run beside the next layer's backward passes)
int wgrad_run(const WgradPlan& pl, float* out, int out_rows, const int* map, int accumulate, cudaStream_t st, int phase) {
  static bool attr_set = false;
  if (!attr_set) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  YB_REQUIRE(out_rows > 0 && out_rows <= pl.kp.Cout, "wgrad: out_rows=%d", out_rows);
  if (phase != 2) {
    if (pl.kind == 1) {
      if (wgrad_patch_launch(pl, st)) return -2;
    } else {
      YB_CHECK_CUDA(launch_pdl(conv_wgrad_kernel, dim3(pl.grid), dim3(kThreads), pl.smem, st, pl.kp));
      YB_LAUNCHED();
    }
  }
  if (phase == 1) return 0;
  const WgradKParams& kp = pl.kp;  // the patch planner fills the reduce-relevant fields of kp as well
  // split-lanes: enough threads to fill the machine even when the output is tiny and the split count large
  int SL = 1;
  while (SL < 32 && SL * 2 <= kp.splits && (long)kp.ldo * out_rows * SL < 148L * 2048) SL *= 2;
  const int KR = 32 / SL;
  dim3 grid((kp.ldo + KR - 1) / KR, (out_rows + 31) / 32);
  YB_CHECK_CUDA(launch_pdl(wgrad_reduce_kernel, dim3(grid), dim3(1024), 0, st, kp.partial, kp.splits, kp.Mpad, kp.Npad, kp.Cin, kp.Cin_pad, kp.ldo, out_rows,
                                              SL, out, map, accumulate));
  YB_LAUNCHED();
  return 0;
}

}  // namespace yb
