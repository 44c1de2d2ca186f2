// 3x3 implicit-GEMM convolution with tap reuse from ONE halo patch in shared memory (tcgen05 + TMA, sm_100a).
//
// Same math and epilogue as conv_igemm.cu (reference model.py:16 CBL conv fprop and its autograd dgrad), different
// operand traffic.  conv_igemm.cu fetches a shifted [128 px x 64 ch] A tile per tap: nine trips through L2 for the same
// activations, and on B200 the L2 -> SM path (not the tensor pipe) bounds that kernel (profiles/).  Here a
// work item is a super-tile of TH x TW tiles of 16 rows x 8 columns of output pixels.  Per 64-channel chunk ONE TMA box
// brings the halo patch [(16*TH + halo) x (8*TW + halo) px x 64 ch] into shared memory (SWIZZLE_128B, 128 bytes per
// pixel row), and every tap of every tile is an MMA whose A descriptor simply STARTS at a different 128-byte row of that
// patch:   start = patch + ((16*ty + oh) * pitch + 8*tx + ow) * 128,   SBO = pitch * 128   (8-row groups = image rows).
// tcgen05 applies the 128-byte swizzle on absolute shared-memory address bits, so a descriptor may start at any row of a
// TMA-written tile (verified on hardware: tools/umma_shift_probe.cu).  The weight slice of a (tap, chunk) is fetched once
// per super-tile and shared by its TH*TW accumulators (TH*TW*BLOCK_N <= 256 TMEM columns, double buffered).
// Stride 2: fprop = four input-parity patches (1/2/2/4 taps); dgrad = four output-parity groups with one patch each.
//
// CTA = 512 threads, one CTA per SM, persistent:
//   warp 0 patch producer (TMA), warp 3 weight producer (TMA), warp 1 MMA issuer, warp 2 TMEM allocator,
//   warps 4..15 epilogue (conv_epilogue.cuh).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "conv_epilogue.cuh"
#include "conv_igemm.cuh"

namespace yb {

static constexpr int kThreads = 512;
static constexpr int kEpiGroups = 3;
static constexpr int kEpiThreads = 4 * kEpiGroups * 32;
static constexpr int kMaxSA = 4, kMaxSB = 12;
static constexpr int kBarRegion = 1024;

struct STile {
  int g, n, hb, wb, nt;
};

struct STileDec {
  FDiv c, w, h, n;
};
__device__ __forceinline__ STileDec load_stile_dec(const PatchKParams& p) {
  STileDec d{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
  keep_in_reg(d.c); keep_in_reg(d.w); keep_in_reg(d.h); keep_in_reg(d.n);
  return d;
}
__device__ __forceinline__ STile decode_stile(const STileDec& d, int t) {
  STile c;
  int m;
  fdivmod(t, d.c, m, c.nt);
  fdivmod(m, d.w, m, c.wb);
  fdivmod(m, d.h, m, c.hb);
  fdivmod(m, d.n, c.g, c.n);
  return c;
}

template <int MODE, bool X32>
__global__ void __launch_bounds__(kThreads, 1) conv_patch_kernel(const __grid_constant__ PatchKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem);
  uint64_t* aempty = afull + kMaxSA;
  uint64_t* bfull = aempty + kMaxSA;
  uint64_t* bempty = bfull + kMaxSB;
  uint64_t* tfull_bar = bempty + kMaxSB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  PTap* s_taps = reinterpret_cast<PTap*>(smem + 512);          // [16] shared copies: indexed constant-bank loads are slow
  PPatch* s_patches = reinterpret_cast<PPatch*>(smem + 640);   // [4] x 28 B
  ConvGroup* s_groups = reinterpret_cast<ConvGroup*>(smem + 768);  // [4] x 24 B
  STileDec* s_td = reinterpret_cast<STileDec*>(smem + 896);        // tile-decode divisors: 32 B (constant-bank loads of these
                                                                   // were 12 % of the dgrad epilogue's stall samples)
  uint8_t* o_stage = smem + kBarRegion;  // TMA-store staging: ceil(BLOCK_N / 64) slabs of [128 px][128 B] (conv_igemm.cu)
  uint8_t* a_smem = o_stage + p.stage_bytes;
  uint8_t* b_smem = a_smem + (size_t)p.sa * p.a_stage_bytes;
  float* s_stats = reinterpret_cast<float*>(b_smem + (size_t)p.sb * p.b_stage_bytes);  // [4][2][Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.ngroups * p.NB * p.tiles_h * p.tiles_w * p.tiles_c;
  const int T = p.TH * p.TW;

  if (threadIdx.x < 16) s_taps[threadIdx.x] = p.taps[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 36) s_patches[threadIdx.x - 32] = p.patches[threadIdx.x - 32];
  if (threadIdx.x >= 64 && threadIdx.x < 68) s_groups[threadIdx.x - 64] = p.groups[threadIdx.x - 64];
  if (threadIdx.x == 96) *s_td = STileDec{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.sa; ++i) {
      mbar_init(&afull[i], 1);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < p.sb; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
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
  if ((MODE & 7) != EPI_DGRAD && (MODE & 7) != EPI_EVAL && p.stats != nullptr && warp >= 4) {
    for (int i = threadIdx.x - 128; i < 4 * 2 * p.Cout; i += kEpiThreads) s_stats[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // PDL: let the next kernel of the stream start its prologue ...
  pdl_wait();               // ... and do not touch global memory before the preceding kernels have completed

  if (warp == 0) {
    // ------------------------------------------------------------------ patch producer (whole warp, elected lane issues)
    {
      const bool leader = elect_one();
      const STileDec td{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
      const int SW = p.TW * 8, SH = p.TH * 16, chunks = p.chunks, nsa = p.sa;
      const uint32_t a_sb = p.a_stage_bytes;
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const STile tc = decode_stile(td, t);
        const int pbeg = p.groups[tc.g].tap_begin, pend = p.groups[tc.g].tap_end;
        const int w0 = tc.wb * SW, h0 = tc.hb * SH;
        for (int pi = pbeg; pi < pend; ++pi) {
          const int pmap = p.patches[pi].map, pox = p.patches[pi].ox, poy = p.patches[pi].oy;
          const uint32_t pbytes = p.patches[pi].bytes;
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(&aempty[s], ph ^ 1);
            if (leader) {
              mbar_expect_tx(&afull[s], pbytes);
              tma_load_4d(&p.tmA[pmap], &afull[s], a_smem + (size_t)s * a_sb, ch * 64, w0 + pox, h0 + poy, tc.n);
            }
            __syncwarp();
            if (++s == nsa) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ weight producer (whole warp, elected lane issues)
    {
      const bool leader = elect_one();
      const STileDec td{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
      const int BN = p.BLOCK_N, chunks = p.chunks, nsb = p.sb;
      const uint32_t b_sb = p.b_stage_bytes, b_tx = p.b_tx_bytes;
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const STile tc = decode_stile(td, t);
        const int pbeg = p.groups[tc.g].tap_begin, pend = p.groups[tc.g].tap_end;
        const int c0 = tc.nt * BN;
        for (int pi = pbeg; pi < pend; ++pi) {
          const int tbeg = p.patches[pi].tap_begin, tend = p.patches[pi].tap_end;
          for (int ch = 0; ch < chunks; ++ch) {
            for (int tp = tbeg; tp < tend; ++tp) {
              mbar_wait(&bempty[s], ph ^ 1);
              if (leader) {
                mbar_expect_tx(&bfull[s], b_tx);
                tma_load_2d(&p.tmB, &bfull[s], b_smem + (size_t)s * b_sb, p.taps[tp].kbase + ch * 64, c0);
              }
              __syncwarp();
              if (++s == nsb) {
                s = 0;
                ph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp runs this loop with warp-uniform control flow and values; one elected lane issues the MMAs (see
    // conv_igemm.cu: in a single-lane loop every tcgen05.mma costs four R2UR moves + an ELECT on the critical path).
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc_bf16(128, p.BLOCK_N, 0, 0);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(b_smem), 16, 1024, 2);
      const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
      const uint32_t a_base16 = smem_u32(a_smem) >> 4;
      const STileDec td{p.fd_c, p.fd_w, p.fd_h, p.fd_n};
      const int BN = p.BLOCK_N, chunks = p.chunks, nsa = p.sa, nsb = p.sb, TWl = p.TW;
      const int klast = (p.Ck - (p.chunks - 1) * 64 + 15) / 16;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const STile tc = decode_stile(td, t);
        const int pbeg = p.groups[tc.g].tap_begin, pend = p.groups[tc.g].tap_end;
        const int ab = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + ab * 256;
        uint32_t acc = 0;
        for (int pi = pbeg; pi < pend; ++pi) {
          const int pitch = p.patches[pi].pitch, tbeg = p.patches[pi].tap_begin, tend = p.patches[pi].tap_end;
          const uint64_t a_hi = make_smem_desc(0, 16, (uint32_t)pitch * 128u, 2);
          // tile (ty, tx) starts (16*ty*pitch + 8*tx) pixel rows (x 128 B = x 8 descriptor units) into the patch
          uint32_t tile_off[4];
          {
            int ty = 0, tx = 0;  // (ty, tx) = (tt / TW, tt % TW) without the runtime division
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              tile_off[tt] = (uint32_t)(ty * 16 * pitch + tx * 8) * 8u;
              if (++tx == TWl) {
                tx = 0;
                ++ty;
              }
            }
          }
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(&afull[sa], pha);
            tc_fence_after();
            const int nk = ch + 1 == chunks ? klast : 4;
            const uint64_t a_desc = a_hi | (uint64_t)(a_base16 + sa * a_step);
            for (int tp = tbeg; tp < tend; ++tp) {
              const uint64_t a_tap = a_desc + (uint64_t)((uint32_t)p.taps[tp].row_off * 8u);
              mbar_wait(&bfull[sb], phb);
              tc_fence_after();
              const uint64_t db = b_desc0 + (uint64_t)(sb * b_step);
              if (leader) {
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) {
                  if (tt >= T) break;
                  const uint64_t da = a_tap + tile_off[tt];
                  const uint32_t d_tmem = d_base + (uint32_t)tt * BN;
                  umma_bf16(d_tmem, da, db, idesc, acc);
                  if (nk > 1) umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);  // nk < 4: zero-padded tail of the last chunk
                  if (nk > 2) umma_bf16(d_tmem, da + 4, db + 4, idesc, 1u);
                  if (nk > 3) umma_bf16(d_tmem, da + 6, db + 6, idesc, 1u);
                }
                umma_commit(&bempty[sb]);
              }
              __syncwarp();
              acc = 1u;
              if (++sb == nsb) {
                sb = 0;
                phb ^= 1;
              }
            }
            if (leader) umma_commit(&aempty[sa]);
            __syncwarp();
            if (++sa == nsa) {
              sa = 0;
              pha ^= 1;
            }
          }
        }
        if (leader) umma_commit(&tfull_bar[ab]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int eg = (warp - 4) >> 2;
    const int r = q * 32 + lane;
    const int phh = r >> 3, pw = r & 7;
    float* my_stats = s_stats + (size_t)q * 2 * p.Cout;
    // register copies of everything the per-tile loop reads (see keep_in_reg in conv_epilogue.cuh)
    const EpiArgs ea = load_epi_args<MODE>(p);
    const STileDec td = *s_td;
    int BN = p.BLOCK_N, TWl = p.TW, THl = p.TH;
    int64_t os_n = p.os_n, os_h = p.os_h, os_w = p.os_w, as_n = p.as_n, as_h = p.as_h, as_w = p.as_w;
    keep_in_reg(BN); keep_in_reg(TWl); keep_in_reg(THl);
    keep_in_reg(os_n); keep_in_reg(os_h); keep_in_reg(os_w); keep_in_reg(as_n); keep_in_reg(as_h); keep_in_reg(as_w);
    const int nchunks = BN / 16;
    int tma_store = (MODE & 7) != EPI_FULL ? 0 : p.tma_store, x32 = (MODE & 7) == EPI_FULL ? p.epi_x32 : (X32 ? 1 : 0);  // compile-time in the role instantiations
    if ((MODE & 7) == EPI_FULL) keep_in_reg(tma_store);
    if ((MODE & 7) == EPI_FULL) keep_in_reg(x32);
    const int units = x32 ? nchunks / 2 : nchunks;
    uint8_t* stage_row = o_stage + (size_t)r * 128;
    const bool issuer = threadIdx.x == 128;
    int it = 0, stores = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const STile tc = decode_stile(td, t);
      const ConvGroup grp = s_groups[tc.g];
      const int64_t o_base = grp.out_off + tc.n * os_n, a_base = grp.add_off + tc.n * as_n;
      const int ab = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tfull_bar[ab], aph);
      tc_fence_after();
      if (tma_store) {
        // TMA-store epilogue (see conv_igemm.cu): tile by tile -- stage the bf16 tile in shared memory, one bulk tensor
        // store per 64-channel slab; box = [64 ch x 8 x 16 x 1] pixels, clipped at the tensor edges
        int ty = 0, tx = 0;
        for (int tt = 0; tt < T; ++tt) {
          const int h0 = (tc.hb * THl + ty) * 16, w0 = (tc.wb * TWl + tx) * 8;
          const int h = h0 + phh, w = w0 + pw;
          const bool valid = h < ea.H && w < ea.W;
          const int64_t opix = o_base + h * os_h + w * os_w;
          const int64_t apix = a_base + h * as_h + w * as_w;
          if (issuer && stores > 0) bulk_wait_group_read0();  // the previous tile's stores no longer read the staging area
          asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads) : "memory");
          for (int cc = eg; cc < nchunks; cc += kEpiGroups) {
            const int col0 = tc.nt * BN + cc * 16;
            if (col0 >= ea.Cout) break;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256 + tt * BN + cc * 16;
            conv_epilogue_chunk(ea, t_addr, col0, valid, tc.n, h, w, opix, apix, my_stats, lane, stage_row, cc, r);
          }
          if (tt + 1 == T) {  // all accumulators of this hand-off have been read: the MMA warp may reuse the TMEM buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[ab]);
          }
          fence_proxy_async();
          asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads) : "memory");
          if (issuer && h0 < ea.H && w0 < ea.W) {
            const int c0 = tc.nt * BN;
            for (int s = 0; s * 64 < BN && c0 + s * 64 < ea.Cout; ++s)
              tma_store_4d(&p.tmO[tc.g], o_stage + (size_t)s * 16384, c0 + s * 64, w0, h0, tc.n);
            bulk_commit_group();
            ++stores;
          }
          if (++tx == TWl) {
            tx = 0;
            ++ty;
          }
        }
        continue;
      }
      // (tile tt, unit cc) of the flat index idx = tt * units + cc and (ty, tx) of tt advance incrementally: two
      // runtime integer divisions per chunk were ~100 of the ~300 instructions of a chunk (ncu source page, r1).
      // A unit is one 16-column chunk, or a pair of them read with one 32-column TMEM load (epi_x32).
      int tt = 0, cc = eg, ty = 0, tx = 0;
      while (cc >= units) {
        cc -= units;
        ++tt;
        if (++tx == TWl) {
          tx = 0;
          ++ty;
        }
      }
      for (; tt < T; ) {
        const int ch0 = x32 ? 2 * cc : cc;  // first 16-column chunk of this unit
        const int col0 = tc.nt * BN + ch0 * 16;
        const int tt_c = tt, ty_c = ty, tx_c = tx;
        cc += kEpiGroups;
        while (cc >= units) {
          cc -= units;
          ++tt;
          if (++tx == TWl) {
            tx = 0;
            ++ty;
          }
        }
        if (col0 >= ea.Cout) continue;
        const int h = (tc.hb * THl + ty_c) * 16 + phh, w = (tc.wb * TWl + tx_c) * 8 + pw;
        const bool valid = h < ea.H && w < ea.W;
        const int64_t opix = o_base + h * os_h + w * os_w;
        const int64_t apix = a_base + h * as_h + w * as_w;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256 + tt_c * BN + ch0 * 16;
        if (x32)
          conv_epilogue_chunk2(ea, t_addr, col0, valid, tc.n, h, w, opix, apix, my_stats, lane, nullptr, 0, 0);
        else
          conv_epilogue_chunk<true>(ea, t_addr, col0, valid, tc.n, h, w, opix, apix, my_stats, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[ab]);
    }
    if (tma_store && issuer) bulk_wait_group0();  // all stores complete before the CTA (and its shared memory) goes away
    if ((MODE & 7) != EPI_DGRAD && (MODE & 7) != EPI_EVAL && p.stats != nullptr) {
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
static int pick_block_n(int Cout) {
  const int c16 = (Cout + 15) / 16 * 16;
  if (c16 <= 256) return c16;
  const int parts = (c16 + 255) / 256;
  return ((c16 + parts - 1) / parts + 15) / 16 * 16;
}

// tensor map over (a parity sub-grid of) an NHWC view with a [64 ch x bw x bh x 1] box
static int make_patch_map(CUtensorMap* m, const TView& v, int bw, int bh, int py, int px, int sy, int sx) {
  const bf16* base = reinterpret_cast<const bf16*>(v.ptr) + (long)py * v.rowp() + (long)px * v.pitch;
  uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)(v.W / sx), (uint64_t)(v.H / sy), (uint64_t)v.N};
  uint64_t strides[3] = {(uint64_t)v.pitch * sx * 2, (uint64_t)v.rowp() * sy * 2,
                         (uint64_t)v.rowp() * v.H * 2};
  uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, 1u};
  return encode_tmap(m, base, 4, dims, strides, box, 128, 2);
}

// 0 = auto (heuristic), 1 = use the patch kernel wherever it is legal, 2 = additionally allow multi-tile super-tiles on
// small problems (tests), -1 = never.  Initialised from $YB_CONV_PATCH, changed with yb_set_conv_patch_mode().
static int g_patch_mode = -2;
static int patch_mode() {
  if (g_patch_mode == -2) {
    const char* e = getenv("YB_CONV_PATCH");
    g_patch_mode = e ? atoi(e) : 0;
  }
  return g_patch_mode;
}
void set_patch_mode(int m) { g_patch_mode = m; }

// common tail: super-tile shape, pipeline depths, epilogue fields.  Wg/Hg: output grid of one group.
static int finish_patch_plan(ConvPlan& pl, const bf16* wmat, int wrows, long wcols, int Kc, const TView& out,
                             const ConvEpilogue& ep, int max_patch_w_halo, int max_patch_h_halo) {
  PatchKParams& kp = pl.pp;
  pl.kind = 1;
  kp.Cout = wrows;
  kp.BLOCK_N = pick_block_n(wrows);
  kp.tiles_c = (wrows + kp.BLOCK_N - 1) / kp.BLOCK_N;
  kp.chunks = (Kc + 63) / 64;
  kp.Ck = Kc;
  {
    uint64_t dims[2] = {(uint64_t)wcols, (uint64_t)wrows};
    uint64_t strides[1] = {(uint64_t)wcols * 2};
    uint32_t box[2] = {64u, (uint32_t)kp.BLOCK_N};
    if (encode_tmap(&kp.tmB, wmat, 2, dims, strides, box, 128, 2)) return -1;
  }
  kp.b_stage_bytes = ((uint32_t)kp.BLOCK_N * 128u + 1023u) & ~1023u;
  kp.b_tx_bytes = (uint32_t)kp.BLOCK_N * 128u;
  kp.out_kind = ep.out_kind;
  kp.out = out.ptr;
  kp.scale = ep.scale;
  kp.shift = ep.shift;
  kp.act = ep.act;
  kp.addend = ep.addend;
  kp.stats = ep.stats;
  kp.head_na = ep.head_na;
  kp.head_no = ep.head_no;
  (void)max_patch_w_halo;
  (void)max_patch_h_halo;
  return 0;
}

// choose TH x TW: minimise the TMA bytes per useful output tile (patch bytes + nine weight slices, shared by the
// TH*TW accumulators of a super-tile; padding to whole super-tiles counts as waste) under the TMEM (256 columns per
// buffer) and shared-memory budgets, keeping >= ~4 super-tiles per SM so the persistent loop stays balanced
static bool choose_supertile(PatchKParams& kp, int Wg, int Hg, int NB, int ngroups, int halo_w, int halo_h, size_t stats_bytes) {
  const size_t budget = 227 * 1024 - 1024 - kBarRegion - stats_bytes;
  const int cands[4][2] = {{2, 2}, {1, 2}, {2, 1}, {1, 1}};
  double best = -1;
  for (int i = 0; i < 4; ++i) {
    const int TH = cands[i][0], TW = cands[i][1], T = TH * TW;
    if (T * kp.BLOCK_N > 256) continue;
    const size_t a_bytes = (((size_t)(16 * TH + halo_h) * (8 * TW + halo_w) * 128) + 1023) & ~(size_t)1023;
    if (2 * a_bytes + 3 * (size_t)kp.b_stage_bytes > budget) continue;
    const int tw = (Wg + 8 * TW - 1) / (8 * TW), th = (Hg + 16 * TH - 1) / (16 * TH);
    const long stiles = (long)tw * th * NB * ngroups * kp.tiles_c;
    if (T > 1 && stiles < 4L * conv_max_grid() && patch_mode() < 2) continue;
    const double eff = (double)Wg * Hg / ((double)tw * 8 * TW * th * 16 * TH);
    const double cost = ((double)a_bytes + 9.0 * kp.b_stage_bytes) / T / eff;
    if (best < 0 || cost < best) {
      best = cost;
      kp.TH = TH;
      kp.TW = TW;
      kp.a_stage_bytes = (uint32_t)a_bytes;
      kp.tiles_w = tw;
      kp.tiles_h = th;
    }
  }
  if (best < 0) return false;
  // measured (profiles/conv_layers_r1f_*.json): with a single tile per weight stage the generic kernel (fuller 128-pixel
  // tiles, no 16x8 padding) is as fast or faster -- the patch kernel pays off where several accumulators share the weights
  if (kp.TH * kp.TW == 1 && patch_mode() == 0) return false;
  const size_t a_bytes = kp.a_stage_bytes;
  const int sa_fit = (int)std::min<size_t>(kMaxSA, (budget - 4 * (size_t)kp.b_stage_bytes) / a_bytes);
  kp.sa = std::max(2, std::min(sa_fit, 3));
  kp.sb = (int)std::min<size_t>(kMaxSB, (budget - (size_t)kp.sa * a_bytes) / kp.b_stage_bytes);
  return kp.sb >= 3;
}

static bool patch_eligible(int Wg, int Hg) {
  const int m = patch_mode();
  if (m < 0) return false;
  if (m > 0) return true;
  // tiles are 16 x 8 pixels: skip maps where the padding to whole tiles wastes more than ~20 % of the MMA work
  const double eff = (double)Wg * Hg / ((double)((Wg + 7) / 8 * 8) * ((Hg + 15) / 16 * 16));
  return eff >= 0.8;
}

static void set_out_strides(PatchKParams& kp, const TView& o, int step, const ConvEpilogue& ep) {
  kp.fd_c = make_fdiv(kp.tiles_c);
  kp.fd_w = make_fdiv(kp.tiles_w);
  kp.fd_h = make_fdiv(kp.tiles_h);
  kp.fd_n = make_fdiv(kp.NB);
  kp.os_n = (int64_t)o.pitch * o.W * o.H;
  kp.os_h = (int64_t)o.pitch * o.W * step;
  kp.os_w = o.pitch * step;
  if (ep.addend) {
    kp.as_n = (int64_t)ep.addend_pitch * o.W * o.H;
    kp.as_h = (int64_t)ep.addend_pitch * o.W * step;
    kp.as_w = ep.addend_pitch * step;
  }
}

int conv_patch_plan_fwd(ConvPlan& pl, const TView& in, const bf16* wp, int ks, int stride, const TView& out,
                        const ConvEpilogue& ep) {
  const bool nhwc_out = ep.out_kind == OUT_BF16 || ep.out_kind == OUT_F32 || ep.out_kind == OUT_F32_ACC;
  if ((ks != 3 && ks != 1 && ks != 31) || (stride != 1 && stride != 2) || !nhwc_out) return 1;
  if (ks != 3 && stride != 1) return 1;
  if (in.C % 8 || in.pitch % 8 || out.pitch % 8 || out.C % 16) return 1;
  if (in.H % stride || in.W % stride || out.H != in.H / stride || out.W != in.W / stride || out.N != in.N) return 1;
  if (!patch_eligible(out.W, out.H)) return 1;
  memset(&pl.pp, 0, sizeof(pl.pp));
  PatchKParams& kp = pl.pp;
  kp.W = out.W;
  kp.H = out.H;
  kp.NB = out.N;
  kp.ngroups = 1;
  const int ntap_total = ks == 31 ? 3 : ks * ks;
  if (finish_patch_plan(pl, wp, out.C, (long)ntap_total * in.C, in.C, out, ep, 0, 0)) return -1;
  const size_t stats_bytes = ep.stats ? (size_t)4 * 2 * out.C * sizeof(float) : 0;
  const int halo = ks == 1 ? 0 : (stride == 1 ? 2 : 1);
  // TMA-store epilogue: staging slabs come out of the operand budget; fall back to direct stores if the tiling does not fit
  size_t st_bytes = (ep.out_kind == OUT_BF16 && conv_tma_store_enabled()) ? (size_t)((kp.BLOCK_N + 63) / 64) * 16384 : 0;
  if (st_bytes && !(conv_make_out_map(&kp.tmO[0], out, 8, 16, 1, 0, 0, 1, 1) &&
                    choose_supertile(kp, out.W, out.H, out.N, 1, ks == 31 ? 0 : halo, halo, stats_bytes + st_bytes)))
    st_bytes = 0;
  if (!st_bytes && !choose_supertile(kp, out.W, out.H, out.N, 1, ks == 31 ? 0 : halo, halo, stats_bytes)) return 1;
  kp.tma_store = st_bytes ? 1 : 0;
  kp.stage_bytes = (uint32_t)st_bytes;
  kp.epi_x32 = (!kp.tma_store && conv_epi_x32(kp.BLOCK_N)) ? 1 : 0;
  const int SW = 8 * kp.TW, SH = 16 * kp.TH;
  int nt = 0, np = 0;
  if (ks == 1) {
    // 1x1: no halo, one tap; the point is several 128-pixel tiles per weight stage and per TMEM hand-off
    PPatch& pa = kp.patches[np++];
    pa.map = 0;
    pa.ox = 0;
    pa.oy = 0;
    pa.pitch = SW;
    pa.bytes = (uint32_t)SW * SH * 128u;
    pa.tap_begin = 0;
    if (make_patch_map(&kp.tmA[0], in, SW, SH, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    kp.taps[nt++] = PTap{0, 0};
    pa.tap_end = nt;
  } else if (ks == 31) {
    // 3 x 1: vertical halo only; tap kh starts kh patch rows down
    PPatch& pa = kp.patches[np++];
    pa.map = 0;
    pa.ox = 0;
    pa.oy = -1;
    pa.pitch = SW;
    pa.bytes = (uint32_t)SW * (SH + 2) * 128u;
    pa.tap_begin = 0;
    if (make_patch_map(&kp.tmA[0], in, SW, SH + 2, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    for (int kh = 0; kh < 3; ++kh) kp.taps[nt++] = PTap{kh * pa.pitch, kh * in.C};
    pa.tap_end = nt;
  } else if (stride == 1) {
    PPatch& pa = kp.patches[np++];
    pa.map = 0;
    pa.ox = -1;
    pa.oy = -1;
    pa.pitch = SW + 2;
    pa.bytes = (uint32_t)(SW + 2) * (SH + 2) * 128u;
    pa.tap_begin = 0;
    if (make_patch_map(&kp.tmA[0], in, SW + 2, SH + 2, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) kp.taps[nt++] = PTap{kh * pa.pitch + kw, (kh * 3 + kw) * in.C};
    pa.tap_end = nt;
  } else {
    // input pixel (2*ho + kh - 1, 2*wo + kw - 1): k = 0 -> parity 1 at block offset -1; k = 1 -> parity 0 at 0; k = 2 -> parity 1 at 0
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        PPatch& pa = kp.patches[np];
        pa.map = np;
        pa.ox = px ? -1 : 0;
        pa.oy = py ? -1 : 0;
        const int bw = SW + (px ? 1 : 0), bh = SH + (py ? 1 : 0);
        pa.pitch = bw;
        pa.bytes = (uint32_t)bw * bh * 128u;
        pa.tap_begin = nt;
        if (make_patch_map(&kp.tmA[np], in, bw, bh, py, px, 2, 2)) return -1;
        for (int kh = 0; kh < 3; ++kh) {
          if (((kh - 1) & 1) != py) continue;
          for (int kw = 0; kw < 3; ++kw) {
            if (((kw - 1) & 1) != px) continue;
            const int oh = py ? (kh == 0 ? 0 : 1) : 0, ow = px ? (kw == 0 ? 0 : 1) : 0;
            kp.taps[nt++] = PTap{oh * bw + ow, (kh * 3 + kw) * in.C};
          }
        }
        pa.tap_end = nt;
        ++np;
      }
  }
  kp.groups[0] = ConvGroup{0, np, 0, 0};
  set_out_strides(kp, out, 1, ep);
  const long total = (long)kp.NB * kp.tiles_h * kp.tiles_w * kp.tiles_c;
  pl.grid = (int)std::min<long>(total, conv_max_grid());
  pl.smem = (int)(1024 + kBarRegion + kp.stage_bytes + (size_t)kp.sa * kp.a_stage_bytes + (size_t)kp.sb * kp.b_stage_bytes +
                  stats_bytes);
  pl.smem = std::max(pl.smem, 120 * 1024);
  return 0;
}

int conv_patch_plan_dgrad(ConvPlan& pl, const TView& dy, const bf16* wt, int ks, int stride, const TView& dx,
                          const ConvEpilogue& ep) {
  const bool nhwc_out = ep.out_kind == OUT_BF16 || ep.out_kind == OUT_F32 || ep.out_kind == OUT_F32_ACC;
  if ((ks != 3 && ks != 1) || (stride != 1 && stride != 2) || !nhwc_out) return 1;
  if (ks == 1 && stride != 1) return 1;
  if (dy.C % 8 || dy.pitch % 8 || dx.pitch % 8 || dx.C % 16) return 1;
  if (dx.H % stride || dx.W % stride || dy.H != dx.H / stride || dy.W != dx.W / stride || dy.N != dx.N) return 1;
  const int Wg = dx.W / stride, Hg = dx.H / stride;
  if (!patch_eligible(Wg, Hg)) return 1;
  memset(&pl.pp, 0, sizeof(pl.pp));
  PatchKParams& kp = pl.pp;
  kp.W = Wg;
  kp.H = Hg;
  kp.NB = dx.N;
  kp.ngroups = stride == 1 ? 1 : 4;
  if (finish_patch_plan(pl, wt, dx.C, (long)ks * ks * dy.C, dy.C, dx, ep, 0, 0)) return -1;
  const int halo = ks == 1 ? 0 : (stride == 1 ? 2 : 1);
  size_t st_bytes = (ep.out_kind == OUT_BF16 && conv_tma_store_enabled()) ? (size_t)((kp.BLOCK_N + 63) / 64) * 16384 : 0;
  for (int g = 0; st_bytes && g < kp.ngroups; ++g)  // stride 2: one map per output-parity group
    if (!conv_make_out_map(&kp.tmO[g], dx, 8, 16, 1, stride == 2 ? g / 2 : 0, stride == 2 ? g % 2 : 0, stride, stride)) st_bytes = 0;
  if (st_bytes && !choose_supertile(kp, Wg, Hg, dx.N, kp.ngroups, halo, halo, st_bytes)) st_bytes = 0;
  if (!st_bytes && !choose_supertile(kp, Wg, Hg, dx.N, kp.ngroups, halo, halo, 0)) return 1;
  kp.tma_store = st_bytes ? 1 : 0;
  kp.stage_bytes = (uint32_t)st_bytes;
  kp.epi_x32 = (!kp.tma_store && conv_epi_x32(kp.BLOCK_N)) ? 1 : 0;
  const int SW = 8 * kp.TW, SH = 16 * kp.TH;
  int nt = 0;
  if (ks == 1) {
    PPatch& pa = kp.patches[0];
    pa.map = 0;
    pa.ox = 0;
    pa.oy = 0;
    pa.pitch = SW;
    pa.bytes = (uint32_t)SW * SH * 128u;
    pa.tap_begin = 0;
    if (make_patch_map(&kp.tmA[0], dy, SW, SH, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    kp.taps[nt++] = PTap{0, 0};
    pa.tap_end = nt;
    kp.groups[0] = ConvGroup{0, 1, 0, 0};
  } else if (stride == 1) {
    // dx[h,w] = sum_{kh,kw} dy[h + 1 - kh, w + 1 - kw] * W[:, :, kh, kw]
    PPatch& pa = kp.patches[0];
    pa.map = 0;
    pa.ox = -1;
    pa.oy = -1;
    pa.pitch = SW + 2;
    pa.bytes = (uint32_t)(SW + 2) * (SH + 2) * 128u;
    pa.tap_begin = 0;
    if (make_patch_map(&kp.tmA[0], dy, SW + 2, SH + 2, 0, 0, 1, 1)) return -1;
    for (int i = 1; i < 4; ++i) kp.tmA[i] = kp.tmA[0];
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) kp.taps[nt++] = PTap{(2 - kh) * pa.pitch + (2 - kw), (kh * 3 + kw) * dy.C};
    pa.tap_end = nt;
    kp.groups[0] = ConvGroup{0, 1, 0, 0};
  } else {
    // output parity classes dx[2*hb+py, 2*wb+px]; contributing kh: (py + 1 - kh) even, dy row = hb + (py + 1 - kh) / 2
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const int g = py * 2 + px;
        PPatch& pa = kp.patches[g];
        pa.map = g;
        pa.ox = 0;
        pa.oy = 0;
        const int bw = SW + px, bh = SH + py;
        pa.pitch = bw;
        pa.bytes = (uint32_t)bw * bh * 128u;
        pa.tap_begin = nt;
        if (make_patch_map(&kp.tmA[g], dy, bw, bh, 0, 0, 1, 1)) return -1;
        for (int kh = 0; kh < 3; ++kh) {
          if (((py + 1 - kh) & 1) != 0) continue;
          for (int kw = 0; kw < 3; ++kw) {
            if (((px + 1 - kw) & 1) != 0) continue;
            const int dh = (py + 1 - kh) / 2, dw = (px + 1 - kw) / 2;  // 0 or 1
            kp.taps[nt++] = PTap{dh * bw + dw, (kh * 3 + kw) * dy.C};
          }
        }
        pa.tap_end = nt;
        kp.groups[g].tap_begin = g;
        kp.groups[g].tap_end = g + 1;
        kp.groups[g].out_off = ((int64_t)py * dx.W + px) * dx.pitch;
        kp.groups[g].add_off = ((int64_t)py * dx.W + px) * ep.addend_pitch;
      }
  }
  set_out_strides(kp, dx, stride, ep);
  const long total = (long)kp.ngroups * kp.NB * kp.tiles_h * kp.tiles_w * kp.tiles_c;
  pl.grid = (int)std::min<long>(total, conv_max_grid());
  pl.smem = (int)(1024 + kBarRegion + kp.stage_bytes + (size_t)kp.sa * kp.a_stage_bytes + (size_t)kp.sb * kp.b_stage_bytes);
  pl.smem = std::max(pl.smem, 120 * 1024);
  return 0;
}

int conv_patch_run(const ConvPlan& pl, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    const int big = 227 * 1024;
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_FULL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_TRAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_TRAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_DGRAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_DGRAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_EVAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_EVAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_DGRAD | EPI_NOADD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_DGRAD | EPI_NOADD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_EVAL | EPI_NOADD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    YB_CHECK_CUDA(cudaFuncSetAttribute(conv_patch_kernel<EPI_EVAL | EPI_NOADD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    attr_set = true;
  }
  const auto& kq = pl.pp;
  int mode = EPI_FULL;
  if (kq.out_kind == OUT_BF16 && !kq.tma_store && conv_lean_enabled()) {
    const bool affine = kq.scale != nullptr || kq.shift != nullptr || kq.act != 0;
    if (kq.stats != nullptr && !affine && kq.addend == nullptr) mode = EPI_TRAIN;
    else if (kq.stats == nullptr && !affine) mode = EPI_DGRAD;
    else if (kq.stats == nullptr && kq.act == 1 && kq.scale != nullptr && kq.shift != nullptr && kq.Cout % 16 == 0 &&
             ((reinterpret_cast<uintptr_t>(kq.scale) | reinterpret_cast<uintptr_t>(kq.shift)) & 15) == 0)
      mode = EPI_EVAL;  // folded BatchNorm + SiLU with 16-byte readable parameters: what every inference CBL has
  }
#define YB_LAUNCH_ROLE(M, X) YB_CHECK_CUDA(launch_pdl(conv_patch_kernel<M, X>, dim3(pl.grid), dim3(kThreads), pl.smem, st, pl.pp))
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
    default: YB_LAUNCH_ROLE(EPI_FULL, false); break;
  }
#undef YB_LAUNCH_ROLE

  YB_LAUNCHED();
  return 0;
}

}  // namespace yb
