// Weight gradient of 3x3 / stride-1 convolutions with tap reuse from ONE halo patch (tcgen05 + TMA, sm_100a).
//
// Same math, split-K scheme, partial-tile layout and reduce kernel as conv_wgrad.cu (reference model.py:16 autograd
// wgrad); different operand traffic.  conv_wgrad.cu fetches nine shifted [KP px x 64 ch] boxes of x per pixel tile.
// Here a pixel tile is PH x PW pixels with PW a multiple of 8, ONE TMA box brings its halo patch [(PH+2) x (PW+2) px x
// 64 ch] into shared memory (128 bytes per pixel row, SWIZZLE_128B), and the MN-major A operand of tap (kh, kw) is the
// same patch read from row (kh * pitch + kw) on: a K = 16 MMA step covers two groups of 8 consecutive pixels of an image
// row, i.e. 8 consecutive patch rows each, so its descriptor is   start = patch + (tap_off + step offset),  SBO = byte
// distance between the two groups (1024 when both lie in the same image row, pitch * 128 when PW = 8).  tcgen05 applies
// the swizzle on absolute shared-memory address bits, so any 128-byte row is a legal start (tools/umma_shift_probe.cu).
// An M tile = two taps of the same 64-channel chunk (LBO = distance between their start rows); a work item owns up to
// MT tiles (MT * BLOCK_N <= 512 TMEM columns) of one chunk, so x costs one patch per stage instead of nine boxes and dy is
// shared by up to nine taps.
//
// CTA = 512 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..15 epilogue.
#include <algorithm>
#include <cstring>

#include "conv_wgrad.cuh"

namespace yb {

static constexpr int kThreads = 512;
static constexpr int kEpiGroups = 3;
static constexpr int kMaxStages = 6;
static constexpr int kBarRegion = 1024;
static constexpr int kTilesPerChunk = 5;  // taps (0,1) (2,3) (4,5) (6,7) (8,-)

struct WPItem {
  int sp, cb, mg, nt;
};
__device__ __forceinline__ WPItem decode_witem(const WPatchKParams& p, int t) {
  WPItem c;
  int m;
  fdivmod(t, p.fd_nt, m, c.nt);
  fdivmod(m, p.fd_mg, m, c.mg);
  fdivmod(m, p.fd_cb, c.sp, c.cb);
  return c;
}

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_patch_kernel(const __grid_constant__ WPatchKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);
  uint8_t* stage_smem = smem + kBarRegion;  // per stage: [x patch][nb dy boxes]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.splits * p.cboxes * p.m_groups * p.n_tiles;

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
    tma_prefetch_desc(&p.tmX);
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
      const uint32_t tx = p.patch_tx + (uint32_t)p.nb * p.box_bytes;  // patch_bytes is the 1024-aligned slot size
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const WPItem it = decode_witem(p, t);
        const int pt0 = (int)(((long)it.sp * p.ptiles) / p.splits);
        const int pt1 = (int)(((long)(it.sp + 1) * p.ptiles) / p.splits);
        for (int pt = pt0; pt < pt1; ++pt) {
          int m, wi, hi, ni;
          fdivmod(pt, p.fd_tw, m, wi);
          fdivmod(m, p.fd_th, ni, hi);
          const int w0 = wi * p.PW, h0 = hi * p.PH;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) {
            mbar_expect_tx(&full_bar[s], tx);
            uint8_t* as = stage_smem + (size_t)s * p.stage_bytes;
            uint8_t* bs = as + p.patch_bytes;
            tma_load_4d(&p.tmX, &full_bar[s], as, it.cb * 64, w0 - 1, h0 - 1, ni);
            for (int j = 0; j < p.nb; ++j)
              tma_load_4d(&p.tmDY, &full_bar[s], bs + (size_t)j * p.box_bytes, it.nt * p.BLOCK_N + j * 64, w0, h0, ni);
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
    // whole warp, warp-uniform values, one elected lane issues (see conv_igemm.cu); integer-add descriptors
    {
      const bool leader = elect_one();
      const int ksteps = p.KP / 16;
      const uint32_t stage_step = p.stage_bytes >> 4;
      const uint32_t base16 = smem_u32(stage_smem) >> 4;
      const uint64_t b_hi = make_smem_desc(0, p.box_bytes, 1024, 2);
      const uint32_t b_off16 = p.patch_bytes >> 4;
      int s = 0;
      uint32_t ph = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
        const WPItem it = decode_witem(p, t);
        const int pt0 = (int)(((long)it.sp * p.ptiles) / p.splits);
        const int pt1 = (int)(((long)(it.sp + 1) * p.ptiles) / p.splits);
        const int tile0 = it.mg * p.MT;
        const int mts = min(p.MT, kTilesPerChunk - tile0);
        // MMA N = real output channels of this N tile rounded up to 16 (see conv_wgrad.cu)
        const int n_mma = p.exact_n ? min(p.BLOCK_N, ((p.Cout + 15) & ~15) - it.nt * p.BLOCK_N) : p.BLOCK_N;
        const uint32_t idesc = make_idesc_bf16(128, n_mma, 1, 1);  // both operands MN-major
        // per tile: descriptor high part (LBO = distance between its two taps) + start of the first tap
        uint64_t a_tile[kTilesPerChunk];
#pragma unroll
        for (int mt = 0; mt < kTilesPerChunk; ++mt) {
          const int t0 = min(2 * (tile0 + mt), 8), t1 = 2 * (tile0 + mt) + 1;
          const uint32_t lbo = t1 < 9 ? (uint32_t)(p.tap_off16[t1 < 9 ? t1 : 8] - p.tap_off16[t0]) * 16u : 1024u;
          a_tile[mt] = make_smem_desc(0, lbo, (uint32_t)p.a_sbo, 2) + (uint64_t)p.tap_off16[t0];
        }
        mbar_wait(tempty_bar, (iter & 1) ^ 1);
        tc_fence_after();
        uint32_t acc = 0;
        for (int pt = pt0; pt < pt1; ++pt) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st16 = base16 + s * stage_step;
          const uint64_t b_desc = b_hi + (uint64_t)(st16 + b_off16);
          if (leader) {
#pragma unroll
            for (int mt = 0; mt < kTilesPerChunk; ++mt) {
              if (mt >= mts) break;
              const uint32_t d_tmem = tmem_base + mt * p.BLOCK_N;
              uint64_t da = a_tile[mt] + st16, db = b_desc;
              umma_bf16(d_tmem, da, db, idesc, acc);
              for (int k = 1; k < ksteps; ++k) {
                da += p.a_step16;
                db += 128;  // 16 pixels x 128 B
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
      const WPItem it = decode_witem(p, t);
      const int tile0 = it.mg * p.MT;
      const int mts = min(p.MT, kTilesPerChunk - tile0);
      const int ncols = min(p.BLOCK_N, ((p.Cout + 15) & ~15) - it.nt * p.BLOCK_N);
      const int ci = it.cb * 64 + (r & 63);
      mbar_wait(tfull_bar, iter & 1);
      tc_fence_after();
      for (int mt = 0; mt < mts; ++mt) {
        const int tapi = 2 * (tile0 + mt) + (r >> 6);
        const bool valid = tapi < 9 && ci < p.Cin;
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

// ------------------------------------------------------------------------------------------------ host
static int g_wpatch_mode = -2;  // $YB_WGRAD_PATCH: 0 auto, 1 wherever legal, -1 never
static int wpatch_mode() {
  if (g_wpatch_mode == -2) {
    const char* e = getenv("YB_WGRAD_PATCH");
    g_wpatch_mode = e ? atoi(e) : 0;
  }
  return g_wpatch_mode;
}
void set_wgrad_patch_mode(int m) { g_wpatch_mode = m; }

static int make_box_map(CUtensorMap* m, const TView& v, int bw, int bh) {
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.N};
  uint64_t strides[3] = {(uint64_t)v.pitch * 2, (uint64_t)v.rowp() * 2, (uint64_t)v.rowp() * v.H * 2};
  uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, 1u};
  return encode_tmap(m, v.ptr, 4, dims, strides, box, 128, 2);
}

int wgrad_patch_plan(WgradPlan& pl, const TView& x, const TView& dy, int ks, int stride, float* partial,
                     size_t partial_floats, int max_splits) {
  if (wpatch_mode() < 0 || ks != 3 || stride != 1) return 1;
  if (x.C % 8 || dy.C % 8 || x.pitch % 8 || dy.pitch % 8) return 1;
  if (dy.H != x.H || dy.W != x.W || dy.N != x.N) return 1;
  // pixel tile: PW = 16 (two 8-pixel groups of a K = 16 step in one image row) or 8 (in two consecutive rows)
  const int W = dy.W, H = dy.H;
  int PW = 0, PH = 0;
  double best = -1;
  const int cand[4][2] = {{16, 8}, {8, 16}, {8, 8}, {16, 4}};
  for (int i = 0; i < 4; ++i) {
    const int pw = cand[i][0], ph = cand[i][1];
    const double eff = (double)W * H / ((double)((W + pw - 1) / pw * pw) * ((H + ph - 1) / ph * ph));
    const double score = eff * (pw * ph >= 128 ? 1.0 : 0.9);  // prefer full 128-pixel stages
    if (score > best + 1e-9) {
      best = score;
      PW = pw;
      PH = ph;
    }
  }
  const double eff = (double)W * H / ((double)((W + PW - 1) / PW * PW) * ((H + PH - 1) / PH * PH));
  if (eff < 0.8 && wpatch_mode() == 0) return 1;
  WPatchKParams& kp = pl.pp;
  memset(&kp, 0, sizeof(kp));
  kp.Cout = dy.C;
  kp.Cin = x.C;
  kp.cboxes = (x.C + 63) / 64;
  kp.Cin_pad = kp.cboxes * 64;
  {
    const int c64 = (dy.C + 63) / 64;
    kp.n_tiles = (c64 + 3) / 4;
    kp.BLOCK_N = (c64 + kp.n_tiles - 1) / kp.n_tiles * 64;
    kp.exact_n = wgrad_exact_n();
  }
  kp.nb = kp.BLOCK_N / 64;
  kp.Mpad = 9 * kp.Cin_pad;
  kp.Npad = kp.n_tiles * kp.BLOCK_N;
  kp.PW = PW;
  kp.PH = PH;
  kp.KP = PW * PH;
  kp.pitch = PW + 2;
  kp.patch_tx = (uint32_t)(PW + 2) * (PH + 2) * 128u;
  kp.patch_bytes = (kp.patch_tx + 1023u) & ~1023u;
  kp.box_bytes = (uint32_t)kp.KP * 128u;
  kp.stage_bytes = kp.patch_bytes + (uint32_t)kp.nb * kp.box_bytes;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) kp.tap_off16[kh * 3 + kw] = (kh * kp.pitch + kw) * 8;
  if (PW % 16 == 0) {  // step j = (row j / (PW/16), 16-pixel segment): consecutive segments of a row, then the next row
    YB_REQUIRE(PW == 16, "wgrad patch: PW=%d", PW);
    kp.a_step16 = kp.pitch * 8;  // next image row
    kp.a_sbo = 1024;
  } else {                       // PW = 8: step j = image rows 2j, 2j+1
    kp.a_step16 = 2 * kp.pitch * 8;
    kp.a_sbo = kp.pitch * 128;
  }
  const size_t budget = wgrad_smem_budget();
  kp.stages = (int)std::min<size_t>(kMaxStages, budget / kp.stage_bytes);
  if (kp.stages < 2) return 1;
  kp.MT = std::max(1, std::min(512 / kp.BLOCK_N, kTilesPerChunk));
  kp.m_groups = (kTilesPerChunk + kp.MT - 1) / kp.MT;
  kp.MT = (kTilesPerChunk + kp.m_groups - 1) / kp.m_groups;
  // measured (profiles/conv_layers_r1m.json): the patch variant wins where all nine taps of a chunk share one dy stage
  // (N tile <= 64 columns: stem, 48 -> 48); from MT = 3 down the generic kernel's tile pairing is as fast or faster
  if (kp.MT < 4 && wpatch_mode() == 0) return 1;
  kp.tiles_w = (W + PW - 1) / PW;
  kp.tiles_h = (H + PH - 1) / PH;
  kp.ptiles = kp.tiles_w * kp.tiles_h * dy.N;
  if (make_box_map(&kp.tmDY, dy, PW, PH) || make_box_map(&kp.tmX, x, PW + 2, PH + 2)) return -1;
  const int base_items = kp.cboxes * kp.m_groups * kp.n_tiles;
  const int sms = wgrad_max_grid();
  int splits = std::max(1, (wgrad_waves() * sms) / base_items);
  splits = std::min(splits, std::max(1, kp.ptiles / 8));
  if (max_splits > 0) splits = std::min(splits, max_splits);
  const size_t per_split = (size_t)kp.Mpad * kp.Npad;
  splits = (int)std::min<size_t>(splits, partial_floats / per_split);
  if (splits < 1) return 1;
  {
    const int items = base_items * splits;
    const int waves = (items + sms - 1) / sms;
    if (waves > 1 && items < waves * sms * 0.85) splits = std::max(1, (waves - 1) * sms / base_items);
  }
  kp.splits = splits;
  kp.partial = partial;
  kp.fd_nt = make_fdiv(kp.n_tiles);
  kp.fd_mg = make_fdiv(kp.m_groups);
  kp.fd_cb = make_fdiv(kp.cboxes);
  kp.fd_tw = make_fdiv(kp.tiles_w);
  kp.fd_th = make_fdiv(kp.tiles_h);
  // fields the shared reduce kernel reads from pl.kp
  pl.kp.partial = partial;
  pl.kp.splits = splits;
  pl.kp.Mpad = kp.Mpad;
  pl.kp.Npad = kp.Npad;
  pl.kp.Cin = kp.Cin;
  pl.kp.Cin_pad = kp.Cin_pad;
  pl.kp.Cout = kp.Cout;
  pl.kp.ldo = 9 * kp.Cin;
  pl.kind = 1;
  pl.grid = std::min(base_items * splits, sms);
  pl.smem = (int)(1024 + kBarRegion + (size_t)kp.stages * kp.stage_bytes);
  pl.smem = std::max(pl.smem, 120 * 1024);
  return 0;
}

int wgrad_patch_launch(const WgradPlan& pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  YB_CHECK_CUDA(launch_pdl(conv_wgrad_patch_kernel, dim3(pl.grid), dim3(kThreads), pl.smem, st, pl.pp));
  YB_LAUNCHED();
  return 0;
}

}  // namespace yb
