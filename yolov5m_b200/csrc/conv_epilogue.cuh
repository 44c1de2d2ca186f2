// Shared epilogue of the tcgen05 convolution kernels (conv_igemm.cu, conv_patch.cu): one 16-column chunk of one
// accumulator row per thread.  TMEM -> registers -> (BN batch-stat partials) -> scale/shift/SiLU/residual -> global.
#pragma once
#include "conv_igemm.cuh"

namespace yb {

// Kernel parameters live in constant bank 0 (the parameter struct is > 1.5 KB with its tensor maps); the compiler
// re-materialises every p.field use as an LDC inside the per-tile loops, and those constant loads miss often enough to
// show up as the top long-scoreboard stall of the epilogue (profiles/).  The hot loops therefore work on REGISTER copies:
// keep_in_reg() launders a value through an empty asm so it cannot be folded back into a constant-bank operand.
__device__ __forceinline__ void keep_in_reg(int& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void keep_in_reg(uint32_t& x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void keep_in_reg(int64_t& x) { asm volatile("" : "+l"(x)); }
template <class T>
__device__ __forceinline__ void keep_in_reg(T*& x) {
  unsigned long long v = reinterpret_cast<unsigned long long>(x);
  asm volatile("" : "+l"(v));
  x = reinterpret_cast<T*>(v);
}
__device__ __forceinline__ void keep_in_reg(FDiv& f) {
  keep_in_reg(f.mul);
  keep_in_reg(f.d);
}

struct EpiArgs {
  float* head_scratch;  // OUT_HEAD_F32: shared-memory transpose scratch, 12 epilogue warps x [32][17] floats (or null)
  float* stats;
  const float* scale;
  const float* shift;
  const bf16* addend;
  void* out;
  int act, out_kind, Cout, H, W, head_na, head_no;
  int ss_vec;  // scale / shift readable as aligned float4 runs of 16 (Cout % 16 == 0, 16-byte aligned pointers)
  int affine;  // 0 = none, 1 = bias only, 2 = scale + shift (see conv_epilogue_affine)
};
// MODE: kernel instantiations for bf16 NHWC outputs with direct stores (every conv of the production train step and of
// the inference forward except the three head convs).  What an instantiation cannot need is a compile-time constant, so the
// loop the common layers run does not carry the fp32 / head / accumulate / TMA-store code, nor the paths of the other two
// roles (measured: the extra code of an unused path costs every layer 1-3 %):
//   0 = everything (heads, fp32 parity outputs, TMA store, any epilogue combination)
//   1 = train fprop: raw bf16 output + BN statistics, no affine / activation / addend
//   2 = dgrad: bf16 output, optional addend, no statistics / affine / activation
//   3 = inference fprop: folded scale + shift, activation, optional addend, no statistics
//   4 = head conv (igemm kernel only): fp32 (B, na, H, W, no) output through the per-warp transpose scratch, bias only
//   + EPI_NOADD (dgrad / inference fprop): the launch has no addend either
enum { EPI_FULL = 0, EPI_TRAIN = 1, EPI_DGRAD = 2, EPI_EVAL = 3, EPI_HEAD = 4, EPI_NOADD = 8 };
template <int MODE = 0, class P>
__device__ __forceinline__ EpiArgs load_epi_args(const P& p) {
  EpiArgs e;
  e.head_scratch = nullptr;
  e.stats = ((MODE & 7) == EPI_DGRAD || (MODE & 7) == EPI_EVAL || (MODE & 7) == EPI_HEAD) ? nullptr : p.stats;
  e.scale = ((MODE & 7) == EPI_TRAIN || (MODE & 7) == EPI_DGRAD || (MODE & 7) == EPI_HEAD) ? nullptr : p.scale;
  e.shift = ((MODE & 7) == EPI_TRAIN || (MODE & 7) == EPI_DGRAD) ? nullptr : p.shift;
  e.addend = ((MODE & 7) == EPI_TRAIN || (MODE & 7) == EPI_HEAD || (MODE & EPI_NOADD)) ? nullptr : p.addend;
  e.out = p.out;
  e.act = ((MODE & 7) == EPI_TRAIN || (MODE & 7) == EPI_DGRAD || (MODE & 7) == EPI_HEAD) ? 0 : ((MODE & 7) == EPI_EVAL ? 1 : p.act);
  e.out_kind = (MODE & 7) == EPI_HEAD ? (int)OUT_HEAD_F32 : ((MODE & 7) != EPI_FULL ? (int)OUT_BF16 : p.out_kind);
  e.Cout = p.Cout; e.H = p.H; e.W = p.W;
  const bool heads = (MODE & 7) == EPI_FULL || (MODE & 7) == EPI_HEAD;
  e.head_na = heads ? p.head_na : 0; e.head_no = heads ? p.head_no : 1;
  keep_in_reg(e.out); keep_in_reg(e.Cout); keep_in_reg(e.H); keep_in_reg(e.W);
  if ((MODE & 7) == EPI_FULL || (MODE & 7) == EPI_TRAIN) keep_in_reg(e.stats);
  if ((MODE & 7) == EPI_FULL || (MODE & 7) == EPI_EVAL) keep_in_reg(e.scale);
  if ((MODE & 7) == EPI_FULL || (MODE & 7) == EPI_EVAL || (MODE & 7) == EPI_HEAD) keep_in_reg(e.shift);
  if ((MODE & 7) == EPI_FULL) keep_in_reg(e.act);
  if ((MODE & 7) != EPI_TRAIN && (MODE & 7) != EPI_HEAD && !(MODE & EPI_NOADD)) keep_in_reg(e.addend);
  if ((MODE & 7) == EPI_FULL) keep_in_reg(e.out_kind);
  if (heads) {
    keep_in_reg(e.head_na); keep_in_reg(e.head_no);
  }
  e.ss_vec = (MODE & 7) == EPI_EVAL ? 1 : (p.scale != nullptr && p.shift != nullptr && (p.Cout & 15) == 0 &&
              ((reinterpret_cast<uintptr_t>(p.scale) | reinterpret_cast<uintptr_t>(p.shift)) & 15) == 0) ? 1 : 0;
  if ((MODE & 7) != EPI_EVAL) keep_in_reg(e.ss_vec);
  e.affine = ((MODE & 7) == EPI_TRAIN || (MODE & 7) == EPI_DGRAD) ? 0
             : ((MODE & 7) == EPI_EVAL ? 2 : ((MODE & 7) == EPI_HEAD ? 1 : (p.scale != nullptr ? 2 : (p.shift != nullptr ? 1 : 0))));
  if ((MODE & 7) == EPI_FULL) keep_in_reg(e.affine);
  return e;
}

// P: EpiArgs (register copy of the epilogue fields of ConvKParams (stats, scale, shift, act, addend, out_kind, out,
// Cout, H, W, head_na, head_no).  t_addr: TMEM address of column 0 of this chunk for this warp's lane quarter.
// (n, h, w): output pixel of this thread's row; opix / apix: element offsets of that pixel in out / addend.
struct EpiAffine {
  float sc[16], sh[16];
};
// scale / shift of 16 columns.  mode: 0 = none (train fprop, dgrad), 1 = bias only (head convs; sc unused), 2 = scale + shift
// (folded BN of the inference path).  Called BEFORE the wait on the TMEM load where the registers allow it, so that the
// (L1-hit) latency overlaps the TMEM round trip: the first dependent FFMA / FADD was the top stall of the inference and
// head epilogues (ncu source pages, round 2).
__device__ __forceinline__ int conv_epilogue_affine_mode(const EpiArgs& p) { return p.affine; }
__device__ __forceinline__ void conv_epilogue_affine(const EpiArgs& p, int mode, int col0, EpiAffine& a) {
  if (mode == 2 && p.ss_vec) {
    // eight 16-byte loads instead of 32 predicated scalar ones (the per-column predicates and loads were ~60 of the ~280
    // instructions of a chunk, ncu source page of the stem at 1280x1280)
    const float4* s4 = reinterpret_cast<const float4*>(p.scale + col0);
    const float4* h4 = reinterpret_cast<const float4*>(p.shift + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 x = __ldg(s4 + j), y = __ldg(h4 + j);
      a.sc[4 * j] = x.x; a.sc[4 * j + 1] = x.y; a.sc[4 * j + 2] = x.z; a.sc[4 * j + 3] = x.w;
      a.sh[4 * j] = y.x; a.sh[4 * j + 1] = y.y; a.sh[4 * j + 2] = y.z; a.sh[4 * j + 3] = y.w;
    }
  } else if (mode == 2) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool in = col0 + j < p.Cout;
      a.sc[j] = in ? __ldg(p.scale + col0 + j) : 1.f;
      a.sh[j] = in ? __ldg(p.shift + col0 + j) : 0.f;
    }
  } else if (mode == 1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a.sh[j] = col0 + j < p.Cout ? __ldg(p.shift + col0 + j) : 0.f;
  }
}

__device__ __forceinline__ void conv_epilogue_process(const EpiArgs& p, const uint32_t* vr, int col0, bool valid, int n, int h, int w,
                                                      int64_t opix, int64_t apix, float* my_stats, int lane, uint8_t* stage_row,
                                                      int ccl, int row, int affine, const EpiAffine* pre);

// stage_row: when non-null, the bf16 result goes to shared memory instead of global memory: the address of this thread's
// 128-byte row in the FIRST 64-channel slab of the tile's staging area (slabs are 16 KB apart); ccl = index of this
// 16-channel chunk inside the tile.  The layout is the SWIZZLE_128B box layout the output tensor map stores from: 16-byte
// unit u of row r sits at unit u ^ (r & 7).
template <bool PRE = false>
__device__ __forceinline__ void conv_epilogue_chunk(const EpiArgs& p, uint32_t t_addr, int col0, bool valid, int n, int h, int w,
                                                    int64_t opix, int64_t apix, float* my_stats, int lane,
                                                    uint8_t* stage_row = nullptr, int ccl = 0, int row = 0) {
    uint32_t vr[16];
    tmem_ld16(t_addr, vr);
    EpiAffine aff;
    const int affine = conv_epilogue_affine_mode(p);
    if (PRE) conv_epilogue_affine(p, affine, col0, aff);
    tmem_ld_wait();
    conv_epilogue_process(p, vr, col0, valid, n, h, w, opix, apix, my_stats, lane, stage_row, ccl, row, affine, PRE ? &aff : nullptr);
}

// the per-chunk work on 16 accumulator columns already in registers
__device__ __forceinline__ void conv_epilogue_process(const EpiArgs& p, const uint32_t* vr, int col0, bool valid, int n, int h, int w,
                                                      int64_t opix, int64_t apix, float* my_stats, int lane, uint8_t* stage_row,
                                                      int ccl, int row, int affine, const EpiAffine* pre) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(vr[j]);

    if (p.stats != nullptr) {
      // per-column sum / sum-of-squares over the 32 rows of this warp: butterfly transpose-reduce
      float a8[8], b8[8];
      const bool u16 = lane & 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float lo = valid ? v[j] : 0.f, hi = valid ? v[j + 8] : 0.f;
        const float keep = u16 ? hi : lo, send = u16 ? lo : hi;
        const float rs = __shfl_xor_sync(0xffffffffu, send, 16);
        const float rq = __shfl_xor_sync(0xffffffffu, send * send, 16);
        a8[j] = keep + rs;
        b8[j] = keep * keep + rq;
      }
      float a4[4], b4[4];
      const bool u8 = lane & 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ka = u8 ? a8[j + 4] : a8[j], sa = u8 ? a8[j] : a8[j + 4];
        const float kb = u8 ? b8[j + 4] : b8[j], sb = u8 ? b8[j] : b8[j + 4];
        a4[j] = ka + __shfl_xor_sync(0xffffffffu, sa, 8);
        b4[j] = kb + __shfl_xor_sync(0xffffffffu, sb, 8);
      }
      float a2[2], b2[2];
      const bool u4 = lane & 4;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float ka = u4 ? a4[j + 2] : a4[j], sa = u4 ? a4[j] : a4[j + 2];
        const float kb = u4 ? b4[j + 2] : b4[j], sb = u4 ? b4[j] : b4[j + 2];
        a2[j] = ka + __shfl_xor_sync(0xffffffffu, sa, 4);
        b2[j] = kb + __shfl_xor_sync(0xffffffffu, sb, 4);
      }
      const bool u2 = lane & 2;
      float a1 = (u2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, u2 ? a2[0] : a2[1], 2);
      float b1 = (u2 ? b2[1] : b2[0]) + __shfl_xor_sync(0xffffffffu, u2 ? b2[0] : b2[1], 2);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
      b1 += __shfl_xor_sync(0xffffffffu, b1, 1);
      // lane l now holds column ((l>>4)&1)*8 + ((l>>3)&1)*4 + ((l>>2)&1)*2 + ((l>>1)&1)
      if ((lane & 1) == 0) {
        const int cj = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        const int col = col0 + cj;
        if (col < p.Cout) {
          my_stats[col] += a1;
          my_stats[p.Cout + col] += b1;
        }
      }
    }

    if (valid) {
      if (pre != nullptr) {  // parameters already in registers (loaded in front of the TMEM wait)
        if (affine == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], pre->sc[j], pre->sh[j]);
        } else if (affine == 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += pre->sh[j];
        }
      } else if (affine == 2 && p.ss_vec) {
        const float4* s4 = reinterpret_cast<const float4*>(p.scale + col0);
        const float4* h4 = reinterpret_cast<const float4*>(p.shift + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = __ldg(s4 + j), b = __ldg(h4 + j);
          v[4 * j] = fmaf(v[4 * j], a.x, b.x);
          v[4 * j + 1] = fmaf(v[4 * j + 1], a.y, b.y);
          v[4 * j + 2] = fmaf(v[4 * j + 2], a.z, b.z);
          v[4 * j + 3] = fmaf(v[4 * j + 3], a.w, b.w);
        }
      } else if (affine == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.Cout) v[j] = fmaf(v[j], __ldg(p.scale + col0 + j), __ldg(p.shift + col0 + j));
      } else if (affine == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.Cout) v[j] += __ldg(p.shift + col0 + j);
      }
      if (p.act) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
      }
      if (p.addend != nullptr) {
        const uint4* ap = reinterpret_cast<const uint4*>(p.addend + apix + col0);
        uint4 r0, r1;
        if ((reinterpret_cast<uintptr_t>(ap) & 31) == 0) {
          asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w), "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w)
                       : "l"(ap));
        } else {
          r0 = __ldg(ap);
          r1 = __ldg(ap + 1);
        }
        const bf16* e0 = reinterpret_cast<const bf16*>(&r0);
        const bf16* e1 = reinterpret_cast<const bf16*>(&r1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] += __bfloat162float(e0[j]);
          v[8 + j] += __bfloat162float(e1[j]);
        }
      }
      if (p.out_kind == OUT_BF16) {
        uint4 o0, o1;
        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          h1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
        }
        // one 32-byte store per thread (st.global.v8, sm_100): a full L2 sector instead of two half-sector writes
        // (every tensor of the network has a pitch that is a multiple of 16 channels; other pitches take the 2 x 16 B path)
        bf16* op = reinterpret_cast<bf16*>(p.out) + opix + col0;
        if (stage_row != nullptr) {
          uint8_t* slab = stage_row + (size_t)(ccl >> 2) * 16384;
          const int u0 = (ccl & 3) * 2;
          *reinterpret_cast<uint4*>(slab + ((u0 ^ (row & 7)) << 4)) = o0;
          *reinterpret_cast<uint4*>(slab + (((u0 + 1) ^ (row & 7)) << 4)) = o1;
        } else if ((reinterpret_cast<uintptr_t>(op) & 31) == 0) {
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(o0.x), "r"(o0.y),
                       "r"(o0.z), "r"(o0.w), "r"(o1.x), "r"(o1.y), "r"(o1.z), "r"(o1.w)
                       : "memory");
        } else {
          reinterpret_cast<uint4*>(op)[0] = o0;
          reinterpret_cast<uint4*>(op)[1] = o1;
        }
      } else if (p.out_kind == OUT_F32 || p.out_kind == OUT_F32_ACC) {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + opix + col0);
        if (p.out_kind == OUT_F32_ACC) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 o = op[j];
            v[4 * j] += o.x; v[4 * j + 1] += o.y; v[4 * j + 2] += o.z; v[4 * j + 3] += o.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else if (p.out_kind == OUT_HEAD_F32 && p.head_scratch != nullptr) {
        // handled below (warp-cooperative transposed store; needs every lane, valid or not)
      } else {  // OUT_HEAD_F32 / OUT_HEAD_F32_ACC
        float* ob = reinterpret_cast<float*>(p.out);
        const int64_t hw = (int64_t)p.H * p.W;
        const int64_t pix = (int64_t)h * p.W + w;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = col0 + j;
          if (col < p.Cout) {
            const int a = col / p.head_no, o = col - a * p.head_no;
            float* dst = ob + (((int64_t)n * p.head_na + a) * hw + pix) * p.head_no + o;
            *dst = p.out_kind == OUT_HEAD_F32_ACC ? *dst + v[j] : v[j];
          }
        }
      }
    }
    if (p.out_kind == OUT_HEAD_F32 && p.head_scratch != nullptr) {
      // Head layout (B, na, H, W, no) fp32 (model.py:173): a pixel's `no` floats are contiguous, pixels 4 * no bytes apart.
      // One thread = one pixel: a direct store instruction would touch 32 sectors for 128 bytes (at bs=128, 1280x1280 the
      // P3 head took 7.1 ms against a 0.7 ms HBM floor).  Transpose the 32 x 16 chunk through a per-warp scratch instead:
      // each instruction then writes two pixels x 16 consecutive floats (two 64-byte runs).
      float* sc = p.head_scratch + (size_t)((threadIdx.x >> 5) - 4) * (32 * 17);
      const int64_t hw = (int64_t)p.H * p.W;
      const int64_t pixbase = valid ? (((int64_t)n * p.head_na) * hw + (int64_t)h * p.W + w) * p.head_no : -1;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) sc[lane * 17 + j] = v[j];
      __syncwarp();
      const int cj = lane & 15, col = col0 + cj;
      const int a = col / p.head_no, o = col - a * p.head_no;
      const int64_t coff = (int64_t)a * hw * p.head_no + o;
      float* ob = reinterpret_cast<float*>(p.out);
#pragma unroll
      for (int st = 0; st < 16; ++st) {
        const int r = 2 * st + (lane >> 4);
        const int64_t pb = __shfl_sync(0xffffffffu, pixbase, r);
        const float val = sc[r * 17 + cj];
        if (pb >= 0 && col < p.Cout) ob[pb + coff] = val;
      }
    }
}

// two adjacent 16-column chunks with ONE 32-column TMEM load: half as many load -> wait round trips per tile (the
// epilogue of the small-K layers is bound by that latency: ncu long-scoreboard stalls on the first use of the loaded
// registers, three epilogue warps per scheduler cannot hide it)
template <bool PRE = false>
__device__ __forceinline__ void conv_epilogue_chunk2(const EpiArgs& p, uint32_t t_addr, int col0, bool valid, int n, int h, int w,
                                                     int64_t opix, int64_t apix, float* my_stats, int lane, uint8_t* stage_row,
                                                     int ccl, int row) {
    uint32_t vr[32];
    tmem_ld32(t_addr, vr);
    EpiAffine aff;
    const int affine = conv_epilogue_affine_mode(p);
    // 32 accumulator registers are live here: only the 16 bias values of the first half fit in front of the wait
    if (PRE && affine == 1) conv_epilogue_affine(p, 1, col0, aff);
    tmem_ld_wait();
    conv_epilogue_process(p, vr, col0, valid, n, h, w, opix, apix, my_stats, lane, stage_row, ccl, row, affine,
                          PRE && affine == 1 ? &aff : nullptr);
    if (col0 + 16 < p.Cout)
      conv_epilogue_process(p, vr + 16, col0 + 16, valid, n, h, w, opix, apix, my_stats, lane, stage_row, ccl + 1, row, affine,
                            nullptr);
}

}  // namespace yb
