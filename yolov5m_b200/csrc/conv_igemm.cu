// Implicit-GEMM convolution on tcgen05 + TMA (sm_100a).  See conv_igemm.cuh.
//
// Replaces, for the reference's hot path, every nn.Conv2d fprop/dgrad that cuDNN would run
// (reference model.py:16 CBL conv, model.py:162 head conv; autograd dgrad of the same).
//
// CTA = 512 threads, one CTA per SM, persistent over output tiles:
//   warp 0      TMA producer   (one elected lane): A = shifted NHWC patch [128 px x KC ch], B = weights [BLOCK_N x KC]
//   warp 1      MMA issuer     (one elected lane): tcgen05.mma kind::f16, M=128, N=BLOCK_N, K=16, fp32 accum in TMEM
//   warp 2      TMEM allocator (512 columns = 2 accumulator buffers of <=256 columns)
//   warps 4..15 epilogue       TMEM -> registers -> (BN batch-stat partials) -> scale/shift/SiLU/residual -> global
//               (three warps per TMEM lane quarter, interleaved over the 16-column chunks: the epilogue is a long
//               dependent chain -- tcgen05.ld, shuffle butterfly, stores -- and one warp per scheduler cannot hide it)
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers (MMA <-> epilogue).
#include "conv_igemm.cuh"
#include "conv_epilogue.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace yb {

static constexpr int kThreads = 512;
static constexpr int kEpiGroups = 3;                  // epilogue warps per TMEM lane quarter
static constexpr int kEpiThreads = 4 * kEpiGroups * 32;
static constexpr int kMaxStages = 12;
static constexpr int kBarRegion = 1024;

struct TileCoord {
  int g, nb, hb, wb, nt;
};

struct TileDec {
  FDiv c, w, h, n;
};
__device__ __forceinline__ TileDec load_tile_dec(const ConvKParams& p) {
  TileDec d{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
  keep_in_reg(d.c); keep_in_reg(d.w); keep_in_reg(d.h); keep_in_reg(d.n);
  return d;
}
__device__ __forceinline__ TileCoord decode_tile(const TileDec& d, int t) {
  TileCoord c;
  int m;
  fdivmod(t, d.c, m, c.nt);
  fdivmod(m, d.w, m, c.wb);
  fdivmod(m, d.h, m, c.hb);
  fdivmod(m, d.n, c.g, c.nb);
  return c;
}

template <int MODE, bool X32>
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  ConvTap* s_taps = reinterpret_cast<ConvTap*>(smem + 512);      // [16] shared copies: indexed constant-bank loads are slow
  ConvGroup* s_groups = reinterpret_cast<ConvGroup*>(smem + 640);  // [4]
  uint8_t* o_stage = smem + kBarRegion;               // TMA-store staging: ceil(BLOCK_N / 64) slabs of [128 px][128 B]
  uint8_t* a_smem = o_stage + p.stage_bytes;
  uint8_t* b_smem = a_smem + (size_t)p.stages * p.a_stage_bytes;
  float* s_stats = reinterpret_cast<float*>(b_smem + (size_t)p.stages * p.b_stage_bytes);  // [4][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.ngroups * p.tiles_n * p.tiles_h * p.tiles_w * p.tiles_c;

  if (threadIdx.x < 16) s_taps[threadIdx.x] = p.taps[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 36) s_groups[threadIdx.x - 32] = p.groups[threadIdx.x - 32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * kEpiGroups);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmA[i]);
    if (p.tma_store)
      for (int i = 0; i < p.ngroups; ++i) tma_prefetch_desc(&p.tmO[i]);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if ((MODE & 7) != EPI_DGRAD && (MODE & 7) != EPI_EVAL && (MODE & 7) != EPI_HEAD && p.stats != nullptr && warp >= 4) {
    for (int i = threadIdx.x - 128; i < 4 * 2 * p.Cout; i += kEpiThreads) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // whole warp, warp-uniform values, one elected lane issues (TMA instructions take uniform-register operands too)
    {
      const bool leader = elect_one();
      const TileDec td{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
      const int PW = p.PW, PH = p.PH, PN = p.PN, BN = p.BLOCK_N, chunks = p.chunks, KC = p.KC, stages = p.stages;
      const uint32_t tx = p.a_tx_bytes + p.b_tx_bytes, a_sb = p.a_stage_bytes, b_sb = p.b_stage_bytes;
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(td, t);
        const int tbeg = p.groups[tc.g].tap_begin, tend = p.groups[tc.g].tap_end;
        const int w0 = tc.wb * PW, h0 = tc.hb * PH, n0 = tc.nb * PN, c0 = tc.nt * BN;
        for (int tp = tbeg; tp < tend; ++tp) {
          const int tmap = p.taps[tp].map, tdw = p.taps[tp].dw, tdh = p.taps[tp].dh, tkb = p.taps[tp].kbase;
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (leader) {
              mbar_expect_tx(&full_bar[s], tx);
              tma_load_4d(&p.tmA[tmap], &full_bar[s], a_smem + (size_t)s * a_sb, ch * KC, w0 + tdw, h0 + tdh, n0);
              tma_load_2d(&p.tmB, &full_bar[s], b_smem + (size_t)s * b_sb, tkb + ch * KC, c0);
            }
            __syncwarp();
            if (++s == stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp runs this loop with warp-uniform control flow and values, and one elected lane issues the MMAs:
    // tcgen05.mma takes its descriptors from uniform registers, and in a single-lane (divergent) loop every MMA costs
    // four R2UR moves plus an ELECT on the issuing thread's critical path (~100 cycles per MMA, more than a small-N MMA
    // itself).  No asm-laundered values here: they would make the operands per-thread again.
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(128, p.BLOCK_N, 0, 0);
      const uint32_t row_bytes = 2u * p.KC;
      const uint32_t lt = swizzle_layout_type(row_bytes);
      const uint32_t sbo = 8u * row_bytes;
      const int kinner = p.KC / 16;
      const uint64_t desc_hi = make_smem_desc(0, 16, sbo, lt);
      const uint64_t a_desc0 = desc_hi | (uint64_t)(smem_u32(a_smem) >> 4);
      const uint64_t b_desc0 = desc_hi | (uint64_t)(smem_u32(b_smem) >> 4);
      const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
      const TileDec td{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
      const int chunks = p.chunks, stages = p.stages;
      const int klast = (p.Ck - (p.chunks - 1) * p.KC + 15) / 16;
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const TileCoord tc = decode_tile(td, t);
        const int nk = (p.groups[tc.g].tap_end - p.groups[tc.g].tap_begin) * chunks;
        const int ab = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * 256;
        uint32_t acc = 0;
        int chk = 0;
        for (int ks = 0; ks < nk; ++ks) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = a_desc0 + (uint64_t)(s * a_step), db = b_desc0 + (uint64_t)(s * b_step);
          const bool last_chunk = ++chk == chunks;  // the zero-padded tail of the last channel chunk needs no MMAs
          if (last_chunk) chk = 0;
          const int nmma = last_chunk ? klast : kinner;
          if (leader) {
            umma_bf16(d_tmem, da, db, idesc, acc);
            if (nmma > 1) umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
            if (nmma > 2) umma_bf16(d_tmem, da + 4, db + 4, idesc, 1u);
            if (nmma > 3) umma_bf16(d_tmem, da + 6, db + 6, idesc, 1u);
            umma_commit(&empty_bar[s]);
          }
          __syncwarp();
          acc = 1u;
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
        if (leader) umma_commit(&tfull_bar[ab]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int eg = (warp - 4) >> 2;  // which of the kEpiGroups warps of that quarter: takes chunks eg, eg + kEpiGroups, ...
    const int r = q * 32 + lane;
    const int pn = r / (p.PH * p.PW);
    const int phh = (r / p.PW) % p.PH;
    const int pw = r % p.PW;
    float* my_stats = s_stats + (size_t)q * 2 * p.Cout;
    // register copies of everything the per-tile loop reads (see keep_in_reg)
    EpiArgs ea = load_epi_args<MODE>(p);
    if (((MODE & 7) == EPI_FULL && p.out_kind == OUT_HEAD_F32) || (MODE & 7) == EPI_HEAD) ea.head_scratch = s_stats;  // the head has no BN statistics: the region is the transpose scratch
    const TileDec td = load_tile_dec(p);
    int PW = p.PW, PH = p.PH, PN = p.PN, NB = p.NB, BN = p.BLOCK_N;
    int64_t t_on = p.PN * p.os_n, t_oh = p.PH * p.os_h, t_ow = p.PW * p.os_w;  // element strides between tiles
    int64_t t_an = p.PN * p.as_n, t_ah = p.PH * p.as_h, t_aw = p.PW * p.as_w;
    int64_t o_thr = pn * p.os_n + phh * p.os_h + pw * p.os_w, a_thr = pn * p.as_n + phh * p.as_h + pw * p.as_w;
    keep_in_reg(PW); keep_in_reg(PH); keep_in_reg(PN); keep_in_reg(NB); keep_in_reg(BN);
    keep_in_reg(t_on); keep_in_reg(t_oh); keep_in_reg(t_ow); keep_in_reg(t_an); keep_in_reg(t_ah); keep_in_reg(t_aw);
    keep_in_reg(o_thr); keep_in_reg(a_thr);
    const bool row_ok = r < p.PW * p.PH * p.PN;
    const int nchunks = BN / 16;
    // TMA-store epilogue: the bf16 tile is staged in shared memory (swizzled box layout) and leaves as one bulk tensor
    // store per 64-channel slab, issued by one thread; per-thread st.global rows (32 sectors per warp instruction) were the
    // top stall of every small-K layer (store back-pressure)
    int tma_store = (MODE & 7) != EPI_FULL ? 0 : p.tma_store, x32 = (MODE & 7) == EPI_FULL ? p.epi_x32 : (X32 ? 1 : 0);  // compile-time in the role instantiations
    if ((MODE & 7) == EPI_FULL) keep_in_reg(tma_store);
    if ((MODE & 7) == EPI_FULL) keep_in_reg(x32);
    uint8_t* stage_row = tma_store ? o_stage + (size_t)r * 128 : nullptr;
    const bool issuer = threadIdx.x == 128;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const TileCoord tc = decode_tile(td, t);
      const ConvGroup grp = s_groups[tc.g];
      const int n = tc.nb * PN + pn, h = tc.hb * PH + phh, w = tc.wb * PW + pw;
      const bool valid = row_ok && n < NB && h < ea.H && w < ea.W;
      const int64_t opix = grp.out_off + tc.nb * t_on + tc.hb * t_oh + tc.wb * t_ow + o_thr;
      const int64_t apix = grp.add_off + tc.nb * t_an + tc.hb * t_ah + tc.wb * t_aw + a_thr;
      const int ab = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tfull_bar[ab], aph);
      tc_fence_after();
      if (tma_store) {  // the previous tile's stores have finished READING the staging area before anyone overwrites it
        if (issuer && it > 0) bulk_wait_group_read0();
        asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256;
      if (x32) {
        for (int cc = 2 * eg; cc < nchunks; cc += 2 * kEpiGroups) {
          const int col0 = tc.nt * BN + cc * 16;
          if (col0 >= ea.Cout) break;
          conv_epilogue_chunk2<true>(ea, t_addr + cc * 16, col0, valid, n, h, w, opix, apix, my_stats, lane, stage_row, cc, r);
        }
      } else {
        for (int cc = eg; cc < nchunks; cc += kEpiGroups) {
          const int col0 = tc.nt * BN + cc * 16;
          if (col0 >= ea.Cout) break;
          conv_epilogue_chunk(ea, t_addr + cc * 16, col0, valid, n, h, w, opix, apix, my_stats, lane, stage_row, cc, r);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[ab]);
      if (tma_store) {
        fence_proxy_async();  // generic-proxy writes of the staging area -> visible to the TMA (async proxy)
        asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
        if (issuer) {
          const int c0 = tc.nt * BN;
          for (int s = 0; s * 64 < BN && c0 + s * 64 < ea.Cout; ++s)
            tma_store_4d(&p.tmO[tc.g], o_stage + (size_t)s * 16384, c0 + s * 64, tc.wb * PW, tc.hb * PH, tc.nb * PN);
          bulk_commit_group();
        }
      }
    }
    if (tma_store && issuer) bulk_wait_group0();  // all stores complete before the CTA (and its shared memory) goes away
    if ((MODE & 7) != EPI_DGRAD && (MODE & 7) != EPI_EVAL && (MODE & 7) != EPI_HEAD && p.stats != nullptr) {
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      float* dst = p.stats + (size_t)blockIdx.x * 2 * p.Cout;
      for (int i = threadIdx.x - 128; i < 2 * p.Cout; i += kEpiThreads) {
        dst[i] = ((s_stats[i] + s_stats[2 * p.Cout + i]) + s_stats[4 * p.Cout + i]) + s_stats[6 * p.Cout + i];
      }
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
static int g_num_sms = 0;
int conv_max_grid() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0)
      g_num_sms = 148;
  }
  return g_num_sms;
}

static void choose_patch(int W, int H, int NB, int& PW, int& PH, int& PN) {
  long best = -1;
  for (int pw = 1; pw <= std::min(W, 128); ++pw) {
    for (int ph = 1; ph <= std::min(H, 128 / pw); ++ph) {
      const int pn = std::max(1, std::min(NB, 128 / (pw * ph)));
      const long tiles = (long)((W + pw - 1) / pw) * ((H + ph - 1) / ph) * ((NB + pn - 1) / pn);
      // fewer tiles first; then prefer wide patches (longer contiguous runs per TMA row)
      const long score = tiles * 1024 - pw;
      if (best < 0 || score < best) {
        best = score;
        PW = pw;
        PH = ph;
        PN = pn;
      }
    }
  }
}

// channels per K chunk (= TMA box width, 2*KC bytes = swizzle span).  From 48 channels up always 64: a channel count that
// is not a multiple of 64 is zero-padded by the TMA out-of-bounds fill (no bytes fetched), which keeps the pipeline at
// full-width 128-byte rows and a third of the stages (C = 48: 1 chunk of 64 instead of 3 chunks of 16).  The matching
// weight columns of the padded part belong to the next tap (or are out of bounds = 0) and meet zeros.
static int pick_kc(int C) { return C >= 48 ? 64 : (C % 32 == 0 ? 32 : 16); }

static int pick_block_n(int Cout) {
  const int c16 = (Cout + 15) / 16 * 16;
  if (c16 <= 256) return c16;
  const int parts = (c16 + 255) / 256;
  return ((c16 + parts - 1) / parts + 15) / 16 * 16;
}

// A tensor map over (a parity sub-grid of) an NHWC view: dims (C, W/sx, H/sy, N).
static int make_a_map(CUtensorMap* m, const TView& v, int KC, int PW, int PH, int PN, int py, int px, int sy, int sx) {
  const bf16* base = reinterpret_cast<const bf16*>(v.ptr) + (long)py * v.rowp() + (long)px * v.pitch;
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)(v.W / sx), (uint64_t)(v.H / sy), (uint64_t)v.N};
  uint64_t strides[3] = {(uint64_t)v.pitch * sx * 2, (uint64_t)v.rowp() * sy * 2,
                         (uint64_t)v.rowp() * v.H * 2};
  uint32_t box[4] = {(uint32_t)KC, (uint32_t)PW, (uint32_t)PH, (uint32_t)PN};
  return encode_tmap(m, base, 4, dims, strides, box, 2 * KC, 2);
}

// $YB_TMA_STORE=1 turns the TMA-store epilogue on (default off).  Measured on B200 (profiles/ab_tma_store_r2.json, all 29
// layer shapes at bs=64): correct everywhere, but slower overall -- forward 5.37 vs 4.78 ms, dgrad 4.62 vs 4.38 ms, step 28.8
// vs 27.8 ms.  The per-tile hand-off (two 384-thread barriers + one issuing thread + wait_group.read before the staging area
// is reused, and one or two operand stages given up for the staging slabs) costs more than the st.global back-pressure it
// removes on every layer with many small tiles per CTA (48/96-channel layers at 160x160 / 80x80: +25..57 %); it wins only on
// the 20x20 maps with <= 4 tiles per CTA (-3..-10 %, ~0.05 ms per step in total).  Kept as a knob, not as the default.
bool conv_epi_x32(int block_n) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YB_EPI_X32");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  const int nchunks = block_n / 16;
  return v != 0 && block_n % 32 == 0 && (nchunks % 6 == 0 || nchunks >= 12);
}

bool conv_tma_store_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YB_TMA_STORE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

// output map over (a parity sub-grid of) an NHWC bf16 view: dims (C, W/sx, H/sy, N), box [64 x bw x bh x bn], SWIZZLE_128B.
// The box may exceed C (64-channel slabs of a 48 / 96-channel tensor): the out-of-bounds part is clipped by the store.
bool conv_make_out_map(CUtensorMap* m, const TView& v, int bw, int bh, int bn, int py, int px, int sy, int sx) {
  const bf16* base = reinterpret_cast<const bf16*>(v.ptr) + (long)py * v.rowp() + (long)px * v.pitch;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (v.pitch * 2) % 16 != 0 || bw > 256 || bh > 256 || bn > 256) return false;
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)(v.W / sx), (uint64_t)(v.H / sy), (uint64_t)v.N};
  uint64_t strides[3] = {(uint64_t)v.pitch * sx * 2, (uint64_t)v.rowp() * sy * 2, (uint64_t)v.rowp() * v.H * 2};
  uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  return encode_tmap(m, base, 4, dims, strides, box, 128, 2) == 0;
}

// dgrad_stride: 0 = forward (one output group), 1 / 2 = dgrad with that stride (2: four output-parity groups)
static int finish_plan(ConvPlan& pl, const bf16* wmat, int wrows, long wcols, const TView& out,
                       const ConvEpilogue& ep, int dgrad_stride = 0) {
  ConvKParams& kp = pl.kp;
  kp.Cout = wrows;
  kp.BLOCK_N = pick_block_n(wrows);
  kp.tiles_c = (wrows + kp.BLOCK_N - 1) / kp.BLOCK_N;
  {
    uint64_t dims[2] = {(uint64_t)wcols, (uint64_t)wrows};
    uint64_t strides[1] = {(uint64_t)wcols * 2};
    uint32_t box[2] = {(uint32_t)kp.KC, (uint32_t)kp.BLOCK_N};
    if (encode_tmap(&kp.tmB, wmat, 2, dims, strides, box, 2 * kp.KC, 2)) return -1;
  }
  kp.tiles_w = (kp.W + kp.PW - 1) / kp.PW;
  kp.tiles_h = (kp.H + kp.PH - 1) / kp.PH;
  kp.tiles_n = (kp.NB + kp.PN - 1) / kp.PN;
  kp.fd_c = make_fdiv(kp.tiles_c);
  kp.fd_w = make_fdiv(kp.tiles_w);
  kp.fd_h = make_fdiv(kp.tiles_h);
  kp.fd_n = make_fdiv(kp.tiles_n);
  kp.a_stage_bytes = 128u * 2u * kp.KC;
  kp.b_stage_bytes = ((uint32_t)kp.BLOCK_N * 2u * kp.KC + 1023u) & ~1023u;
  kp.a_tx_bytes = (uint32_t)(kp.PW * kp.PH * kp.PN) * 2u * kp.KC;
  kp.b_tx_bytes = (uint32_t)kp.BLOCK_N * 2u * kp.KC;
  kp.out_kind = ep.out_kind;
  kp.out = out.ptr;
  kp.scale = ep.scale;
  kp.shift = ep.shift;
  kp.act = ep.act;
  kp.addend = ep.addend;
  kp.stats = ep.stats;
  kp.head_na = ep.head_na;
  kp.head_no = ep.head_no;
  YB_REQUIRE(ep.out_kind == OUT_HEAD_F32 || ep.out_kind == OUT_HEAD_F32_ACC || wrows % 16 == 0,
             "conv: Cout=%d must be a multiple of 16", wrows);
  YB_REQUIRE(ep.scale == nullptr || ep.shift != nullptr, "conv: scale without shift");
  YB_REQUIRE(!(ep.stats && ep.out_kind == OUT_HEAD_F32), "conv: head output with BN statistics");
  const size_t stats_bytes = ep.stats ? (size_t)4 * 2 * wrows * sizeof(float)
                                      : (ep.out_kind == OUT_HEAD_F32 ? (size_t)kEpiGroups * 4 * 32 * 17 * sizeof(float) : 0);
  size_t budget = 227 * 1024 - 1024 /*align slack*/ - kBarRegion - stats_bytes;
  // TMA-store epilogue (bf16 NHWC outputs): staging = one [128 px][128 B] slab per 64 output channels of the tile, taken
  // from the operand ring; kept only if at least three pipeline stages remain
  kp.tma_store = 0;
  kp.stage_bytes = 0;
  if (ep.out_kind == OUT_BF16 && conv_tma_store_enabled()) {
    const size_t st_bytes = (size_t)((kp.BLOCK_N + 63) / 64) * 16384;
    bool ok = budget > st_bytes && (budget - st_bytes) / (kp.a_stage_bytes + kp.b_stage_bytes) >= 3;
    const int step = dgrad_stride == 2 ? 2 : 1;
    for (int g = 0; ok && g < kp.ngroups; ++g) {
      const int py = step == 2 ? g / 2 : 0, px = step == 2 ? g % 2 : 0;
      ok = conv_make_out_map(&kp.tmO[g], out, kp.PW, kp.PH, kp.PN, py, px, step, step);
    }
    if (ok) {
      kp.tma_store = 1;
      kp.stage_bytes = (uint32_t)st_bytes;
      budget -= st_bytes;
    }
  }
  kp.epi_x32 = (!kp.tma_store && conv_epi_x32(kp.BLOCK_N)) ? 1 : 0;
  int stages = (int)(budget / (kp.a_stage_bytes + kp.b_stage_bytes));
  stages = std::min(stages, kMaxStages);
  YB_REQUIRE(stages >= 2, "conv: tile does not fit in shared memory");
  kp.stages = stages;
  pl.smem = (int)(1024 + kBarRegion + kp.stage_bytes + (size_t)stages * (kp.a_stage_bytes + kp.b_stage_bytes) + stats_bytes);
  pl.smem = std::max(pl.smem, 120 * 1024);  // one CTA per SM: each CTA allocates all 512 TMEM columns
  for (int g = 0; g < kp.ngroups; ++g)
    YB_REQUIRE(kp.groups[g].tap_end > kp.groups[g].tap_begin, "conv: output group %d has no taps", g);
  const long total = (long)kp.ngroups * kp.tiles_n * kp.tiles_h * kp.tiles_w * kp.tiles_c;
  pl.grid = (int)std::min<long>(total, conv_max_grid());
  return 0;
}

int conv_plan_fwd(ConvPlan& pl, const TView& in, const bf16* wp, int ks, int stride, const TView& out,
                  const ConvEpilogue& ep) {
  memset(&pl, 0, sizeof(pl));
  {  // 3x3 on maps that tile well into 16x8-pixel tiles: halo-patch kernel (conv_patch.cu)
    const int rc = conv_patch_plan_fwd(pl, in, wp, ks, stride, out, ep);
    if (rc <= 0) return rc;
    pl.kind = 0;
  }
  ConvKParams& kp = pl.kp;
  // ks = 31: three vertical taps (a 3 x 1 kernel, stride 1) -- the stem after horizontal tap gathering (yb_prep_input)
  YB_REQUIRE(((ks == 1 || ks == 3) && (stride == 1 || stride == 2)) || (ks == 31 && stride == 1),
             "conv fwd: ks=%d stride=%d unsupported", ks, stride);
  const int ksh = ks == 31 ? 3 : ks, ksw = ks == 31 ? 1 : ks;
  YB_REQUIRE(in.C % 16 == 0 && in.pitch % 8 == 0 && (ep.out_kind != OUT_BF16 || out.pitch % 8 == 0),
             "conv fwd: channel alignment");
  YB_REQUIRE(in.H % stride == 0 && in.W % stride == 0, "conv fwd: odd input for stride 2");
  YB_REQUIRE(out.H == in.H / stride && out.W == in.W / stride && out.N == in.N, "conv fwd: geometry");
  kp.KC = pick_kc(in.C);
  kp.chunks = (in.C + kp.KC - 1) / kp.KC;
  kp.Ck = in.C;
  kp.W = out.W;
  kp.H = out.H;
  kp.NB = out.N;
  choose_patch(kp.W, kp.H, kp.NB, kp.PW, kp.PH, kp.PN);
  const int pad = ks / 2;
  int nt = 0;
  if (stride == 1) {
    if (make_a_map(&kp.tmA[0], in, kp.KC, kp.PW, kp.PH, kp.PN, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    for (int kh = 0; kh < ksh; ++kh)
      for (int kw = 0; kw < ksw; ++kw) {
        kp.taps[nt] = ConvTap{0, (int8_t)(kw - ksw / 2), (int8_t)(kh - ksh / 2), 0, (int32_t)((kh * ksw + kw) * in.C)};
        ++nt;
      }
  } else {
    // input row 2*ho + kh - 1 :  kh=0 -> parity 1, block ho-1 ; kh=1 -> parity 0, block ho ; kh=2 -> parity 1, block ho
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        if (make_a_map(&kp.tmA[py * 2 + px], in, kp.KC, kp.PW, kp.PH, kp.PN, py, px, 2, 2)) return -1;
    for (int kh = 0; kh < ks; ++kh)
      for (int kw = 0; kw < ks; ++kw) {
        const int oy = kh - pad, ox = kw - pad;  // input offset relative to 2*ho
        const int py = oy & 1, px = ox & 1;
        const int dh = (oy - py) / 2, dw = (ox - px) / 2;
        kp.taps[nt] = ConvTap{(int8_t)(py * 2 + px), (int8_t)dw, (int8_t)dh, 0, (int32_t)((kh * ks + kw) * in.C)};
        ++nt;
      }
  }
  kp.ngroups = 1;
  kp.groups[0] = ConvGroup{0, nt, 0, 0};
  kp.os_n = (int64_t)out.pitch * out.W * out.H;
  kp.os_h = (int64_t)out.pitch * out.W;
  kp.os_w = out.pitch;
  if (ep.addend) {
    kp.as_n = (int64_t)ep.addend_pitch * out.W * out.H;
    kp.as_h = (int64_t)ep.addend_pitch * out.W;
    kp.as_w = ep.addend_pitch;
  }
  return finish_plan(pl, wp, out.C, (long)ksh * ksw * in.C, out, ep);
}

int conv_plan_dgrad(ConvPlan& pl, const TView& dy, const bf16* wt, int ks, int stride, const TView& dx,
                    const ConvEpilogue& ep) {
  memset(&pl, 0, sizeof(pl));
  {
    const int rc = conv_patch_plan_dgrad(pl, dy, wt, ks, stride, dx, ep);
    if (rc <= 0) return rc;
    pl.kind = 0;
  }
  ConvKParams& kp = pl.kp;
  YB_REQUIRE((ks == 1 || ks == 3) && (stride == 1 || stride == 2), "conv dgrad: ks=%d stride=%d unsupported", ks,
             stride);
  YB_REQUIRE(dy.C % 16 == 0 && dy.pitch % 8 == 0 && dx.pitch % 8 == 0, "conv dgrad: channel alignment");
  YB_REQUIRE(dy.H == dx.H / stride && dy.W == dx.W / stride && dy.N == dx.N, "conv dgrad: geometry");
  kp.KC = pick_kc(dy.C);
  kp.chunks = (dy.C + kp.KC - 1) / kp.KC;
  kp.Ck = dy.C;
  kp.NB = dx.N;
  const int pad = ks / 2;
  int nt = 0;
  if (stride == 1) {
    kp.W = dx.W;
    kp.H = dx.H;
    choose_patch(kp.W, kp.H, kp.NB, kp.PW, kp.PH, kp.PN);
    if (make_a_map(&kp.tmA[0], dy, kp.KC, kp.PW, kp.PH, kp.PN, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    // dx[h,w] = sum_{kh,kw} dy[h + pad - kh, w + pad - kw] * W[:, :, kh, kw]
    for (int kh = 0; kh < ks; ++kh)
      for (int kw = 0; kw < ks; ++kw) {
        kp.taps[nt] = ConvTap{0, (int8_t)(pad - kw), (int8_t)(pad - kh), 0, (int32_t)((kh * ks + kw) * dy.C)};
        ++nt;
      }
    kp.ngroups = 1;
    kp.groups[0] = ConvGroup{0, nt, 0, 0};
    kp.os_n = (int64_t)dx.pitch * dx.W * dx.H;
    kp.os_h = (int64_t)dx.pitch * dx.W;
    kp.os_w = dx.pitch;
  } else {
    // output parity classes: dx[2*hb+py, 2*wb+px]; contributing kh satisfy (2*hb+py+pad-kh) even, ho = that / 2.
    kp.W = dx.W / 2;
    kp.H = dx.H / 2;
    choose_patch(kp.W, kp.H, kp.NB, kp.PW, kp.PH, kp.PN);
    if (make_a_map(&kp.tmA[0], dy, kp.KC, kp.PW, kp.PH, kp.PN, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    kp.ngroups = 4;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const int g = py * 2 + px;
        kp.groups[g].tap_begin = nt;
        for (int kh = 0; kh < ks; ++kh) {
          if (((py + pad - kh) & 1) != 0) continue;
          for (int kw = 0; kw < ks; ++kw) {
            if (((px + pad - kw) & 1) != 0) continue;
            const int dh = (py + pad - kh) / 2, dw = (px + pad - kw) / 2;  // exact: numerators are even
            kp.taps[nt] = ConvTap{0, (int8_t)dw, (int8_t)dh, 0, (int32_t)((kh * ks + kw) * dy.C)};
            ++nt;
          }
        }
        kp.groups[g].tap_end = nt;
        kp.groups[g].out_off = ((int64_t)py * dx.W + px) * dx.pitch;
        kp.groups[g].add_off = ((int64_t)py * dx.W + px) * ep.addend_pitch;
      }
    kp.os_n = (int64_t)dx.pitch * dx.W * dx.H;
    kp.os_h = (int64_t)dx.pitch * dx.W * 2;
    kp.os_w = dx.pitch * 2;
  }
  if (ep.addend) {
    kp.as_n = (int64_t)ep.addend_pitch * dx.W * dx.H;
    kp.as_h = (int64_t)ep.addend_pitch * dx.W * stride;
    kp.as_w = ep.addend_pitch * stride;
  }
  return finish_plan(pl, wt, dx.C, (long)ks * ks * dy.C, dx, ep, stride);
}

int conv_stats_rows(const ConvPlan& pl) { return pl.grid; }

int conv_run(const ConvPlan& pl, cudaStream_t st) {
  if (pl.kind == 1) return conv_patch_run(pl, st);
  static bool attr_set = false;
  if (!attr_set) {
    const int big = 227 * 1024;
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_FULL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_TRAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_TRAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_DGRAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_DGRAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_EVAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_EVAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_DGRAD | EPI_NOADD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_DGRAD | EPI_NOADD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_EVAL | EPI_NOADD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_EVAL | EPI_NOADD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<EPI_HEAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    attr_set = true;
  }
  const auto& kq = pl.kp;
  int mode = EPI_FULL;
  if (kq.out_kind == OUT_BF16 && !kq.tma_store && conv_lean_enabled()) {
    const bool affine = kq.scale != nullptr || kq.shift != nullptr || kq.act != 0;
    if (kq.stats != nullptr && !affine && kq.addend == nullptr) mode = EPI_TRAIN;
    else if (kq.stats == nullptr && !affine) mode = EPI_DGRAD;
    else if (kq.stats == nullptr && kq.act == 1 && kq.scale != nullptr && kq.shift != nullptr && kq.Cout % 16 == 0 &&
             ((reinterpret_cast<uintptr_t>(kq.scale) | reinterpret_cast<uintptr_t>(kq.shift)) & 15) == 0)
      mode = EPI_EVAL;  // folded BatchNorm + SiLU with 16-byte readable parameters: what every inference CBL has
  }
  if (kq.out_kind == OUT_HEAD_F32 && kq.epi_x32 && conv_lean_enabled() && kq.stats == nullptr && kq.scale == nullptr &&
      kq.shift != nullptr && kq.act == 0 && kq.addend == nullptr)
    mode = EPI_HEAD;
#define YB_LAUNCH_ROLE(M, X) YB_CHECK_CUDA(launch_pdl(conv_igemm_kernel<M, X>, dim3(pl.grid), dim3(kThreads), pl.smem, st, pl.kp))
  const bool x32 = kq.epi_x32 != 0;
  if ((mode == EPI_DGRAD || mode == EPI_EVAL) && kq.addend == nullptr) mode |= EPI_NOADD;
  switch (mode) {
    case EPI_TRAIN: if (x32) YB_LAUNCH_ROLE(EPI_TRAIN, true); else YB_LAUNCH_ROLE(EPI_TRAIN, false); break;
    case EPI_DGRAD: if (x32) YB_LAUNCH_ROLE(EPI_DGRAD, true); else YB_LAUNCH_ROLE(EPI_DGRAD, false); break;
    case EPI_EVAL: if (x32) YB_LAUNCH_ROLE(EPI_EVAL, true); else YB_LAUNCH_ROLE(EPI_EVAL, false); break;
    case EPI_DGRAD | EPI_NOADD:
      if (x32) YB_LAUNCH_ROLE(EPI_DGRAD | EPI_NOADD, true); else YB_LAUNCH_ROLE(EPI_DGRAD | EPI_NOADD, false);
      break;
    case EPI_EVAL | EPI_NOADD:
      if (x32) YB_LAUNCH_ROLE(EPI_EVAL | EPI_NOADD, true); else YB_LAUNCH_ROLE(EPI_EVAL | EPI_NOADD, false);
      break;
    case EPI_HEAD: YB_LAUNCH_ROLE(EPI_HEAD, true); break;
    default: YB_LAUNCH_ROLE(EPI_FULL, false); break;
  }
#undef YB_LAUNCH_ROLE

  YB_LAUNCHED();
  return 0;
}

}  // namespace yb
