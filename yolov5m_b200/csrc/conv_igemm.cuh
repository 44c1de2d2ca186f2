// Implicit-GEMM convolution on tcgen05 (sm_100a): shared declarations.
//
// One persistent, warp-specialised kernel covers
//   * forward conv 1x1 / 3x3, stride 1 or 2 (stride 2 through four input-parity tensor maps),
//   * the 6x6/s2 stem as a 3x3/s1 conv over the space-to-depth input,
//   * data-gradient (dgrad) of all of the above (stride-2 dgrad = four output-parity groups),
// as "sum over taps of shifted NHWC tiles x K-major weight slices".
#pragma once
#include "common.cuh"

namespace yb {

// NHWC bf16 view of (a channel slice of) a tensor.  pitch = elements between consecutive pixels.
struct TView {
  void* ptr;   // points at channel 0 of the slice of pixel (0,0,0)
  int N, H, W, C;
  long pitch;  // elements between consecutive pixels (>= C, except for the stem's overlapping-window view: 16 < 48)
  long row_pitch = 0;  // elements between image rows; 0 = pitch * W (dense rows)
  long rowp() const { return row_pitch ? row_pitch : pitch * W; }
};
// The stem's operand (ks code 31, Cin = 48, pitch = 16): the three horizontal taps of the 16-channel space-to-depth image
// are a VIEW of its row-padded storage (N, H, W + 2, 16) -- 48 contiguous values starting at padded column w
// YB_CONV_LEAN=0 runs every launch through the full kernel instantiation (A/B knob)
inline bool conv_lean_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YB_CONV_LEAN");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

inline void stem_view(TView& v, int ks) {
  if (ks == 31 && v.pitch < v.C) v.row_pitch = (long)(v.W + 2) * v.pitch;
}

struct ConvTap {
  int8_t map;  // which A tensor map (input parity for stride 2)
  int8_t dw, dh;
  int8_t pad;
  int32_t kbase;  // first column of this tap inside the B (weight) matrix
};

struct ConvGroup {
  int32_t tap_begin, tap_end;
  int64_t out_off;  // element offset of this group's (0,0,0) output pixel
  int64_t add_off;  // same for the addend tensor
};

// OUT_F32_ACC / OUT_HEAD_F32_ACC: out += value (fp32 read-modify-write) -- the parity mode runs one convolution as several
// bf16-split passes of the same kernel that accumulate into one fp32 tensor (model.py parity engine, csrc/parity.cu)
enum { OUT_BF16 = 0, OUT_HEAD_F32 = 1, OUT_F32 = 2, OUT_F32_ACC = 3, OUT_HEAD_F32_ACC = 4 };

struct ConvKParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmO[4];        // output maps (one per group) of the TMA-store epilogue: box [64 ch x PW x PH x PN], SWIZZLE_128B
  int32_t tma_store;         // 1: bf16 tiles leave through shared memory + cp.async.bulk.tensor stores (stage_bytes of smem)
  uint32_t stage_bytes;
  int32_t epi_x32;           // 1: the epilogue reads TMEM 32 columns at a time (pairs of 16-column chunks per warp)
  ConvTap taps[16];
  ConvGroup groups[4];
  int32_t ngroups;
  int32_t chunks, KC;        // channel chunks per tap, channels per chunk (16/32/64)
  int32_t Ck;                // true channels per tap: the last chunk issues only ceil((Ck - (chunks-1)*KC) / 16) MMAs
  int32_t PW, PH, PN;        // output-pixel patch of one 128-row tile
  int32_t tiles_w, tiles_h, tiles_n, tiles_c;
  int32_t W, H, NB;          // output pixel grid (per group)
  int32_t Cout, BLOCK_N;
  FDiv fd_c, fd_w, fd_h, fd_n;  // tile index decoding
  int32_t stages;
  uint32_t a_stage_bytes, b_stage_bytes, a_tx_bytes, b_tx_bytes;
  // epilogue
  int32_t out_kind;
  void* out;
  int64_t os_n, os_h, os_w;  // output pixel strides in elements
  const float* scale;        // per out channel or null
  const float* shift;        // per out channel or null (bias when scale == null)
  int32_t act;               // 1 = SiLU
  const bf16* addend;        // optional tensor added after the activation (residual / grad accumulation)
  int64_t as_n, as_h, as_w;
  float* stats;              // [gridDim.x][2][Cout] per-CTA partial sum / sum of squares of the RAW accumulator
  int32_t head_na, head_no;  // OUT_HEAD_F32: out[((n*na+a)*H*W + h*W + w)*no + o], column = a*no+o
};

// ---- "patch" variant (conv_patch.cu): 3x3 convolutions with tap reuse from one halo patch in shared memory ----------
struct PTap {
  int32_t row_off;  // (oh * pitch + ow): first patch row of this tap for tile (0,0)
  int32_t kbase;    // first column of this tap inside the B (weight) matrix
};
struct PPatch {
  int32_t map;                 // A tensor map (input parity for stride 2)
  int32_t ox, oy;              // patch origin relative to the super-tile origin (w0, h0) in that map's pixel grid
  int32_t pitch;               // patch width in pixels (= TMA box width; rows of the patch are `pitch` pixels apart)
  int32_t tap_begin, tap_end;  // taps served by this patch
  uint32_t bytes;              // TMA box bytes (pitch * height * 128)
};
struct PatchKParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmO[4];        // output maps (one per group) of the TMA-store epilogue: box [64 ch x 8 x 16 x 1], SWIZZLE_128B
  int32_t tma_store;
  uint32_t stage_bytes;
  int32_t epi_x32;
  PTap taps[16];
  PPatch patches[4];
  ConvGroup groups[4];  // tap_begin/tap_end index PATCHES here
  int32_t ngroups;
  int32_t chunks;            // 64-channel K chunks (the last one zero-padded by TMA OOB fill)
  int32_t Ck;                // true channels per tap: the last chunk issues only ceil((Ck - 64*(chunks-1)) / 16) MMAs
  int32_t TH, TW;            // 128-pixel tiles (16 rows x 8 cols) per super-tile, vertically / horizontally
  int32_t tiles_w, tiles_h, tiles_c;  // super-tiles per image row / column, N tiles
  int32_t W, H, NB;          // output pixel grid (per group)
  int32_t Cout, BLOCK_N;
  FDiv fd_c, fd_w, fd_h, fd_n;  // super-tile index decoding
  int32_t sa, sb;            // pipeline depth of the patch ring / weight ring
  uint32_t a_stage_bytes, b_stage_bytes, b_tx_bytes;
  // epilogue (same meaning as ConvKParams)
  int32_t out_kind;
  void* out;
  int64_t os_n, os_h, os_w;
  const float* scale;
  const float* shift;
  int32_t act;
  const bf16* addend;
  int64_t as_n, as_h, as_w;
  float* stats;
  int32_t head_na, head_no;
};

struct ConvPlan {
  ConvKParams kp;
  PatchKParams pp;
  int kind;  // 0 = conv_igemm_kernel (kp), 1 = conv_patch_kernel (pp)
  int grid;
  int smem;
};

struct ConvEpilogue {
  int out_kind = OUT_BF16;
  const float* scale = nullptr;
  const float* shift = nullptr;
  int act = 0;
  const bf16* addend = nullptr;  // same geometry as the output view
  long addend_pitch = 0;
  float* stats = nullptr;
  int head_na = 3, head_no = 85;
};

// forward: in (N,Hin,Win,Cin) -> out (N,Hout,Wout,Cout); wp = [Cout][ks*ks][Cin] bf16.
int conv_plan_fwd(ConvPlan& pl, const TView& in, const bf16* wp, int ks, int stride, const TView& out,
                  const ConvEpilogue& ep);
// dgrad: dy (N,Hout,Wout,Cout) -> dx (N,Hin,Win,Cin); wt = [Cin][ks*ks][Cout] bf16 (tap = kh*ks+kw).
int conv_plan_dgrad(ConvPlan& pl, const TView& dy, const bf16* wt, int ks, int stride, const TView& dx,
                    const ConvEpilogue& ep);
int conv_run(const ConvPlan& pl, cudaStream_t st);
// patch variant: returns 1 when the shape is not eligible (caller falls back to the generic kernel), 0 ok, <0 error
int conv_patch_plan_fwd(ConvPlan& pl, const TView& in, const bf16* wp, int ks, int stride, const TView& out,
                        const ConvEpilogue& ep);
int conv_patch_plan_dgrad(ConvPlan& pl, const TView& dy, const bf16* wt, int ks, int stride, const TView& dx,
                          const ConvEpilogue& ep);
int conv_patch_run(const ConvPlan& pl, cudaStream_t st);
void set_patch_mode(int m);
int conv_stats_rows(const ConvPlan& pl);  // number of per-CTA partial rows written to ep.stats (= grid)
int conv_max_grid();
// output tensor map of the TMA-store epilogue over (a parity sub-grid of) an NHWC bf16 view; false -> not encodable
bool conv_make_out_map(CUtensorMap* m, const TView& v, int bw, int bh, int bn, int py, int px, int sy, int sx);
bool conv_epi_x32(int block_n);  // $YB_EPI_X32: 32-column TMEM loads in the epilogue where the chunk pairs balance over the warps
bool conv_tma_store_enabled();  // $YB_TMA_STORE=1 (default off: measured slower, see conv_igemm.cu)

}  // namespace yb
