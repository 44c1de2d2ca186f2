// Weight-gradient convolution on tcgen05 (sm_100a): shared declarations.  See conv_wgrad.cu.
#pragma once
#include "conv_igemm.cuh"

namespace yb {

struct WgradKParams {
  CUtensorMap tmDY;    // dy  (Cout, Wo, Ho, N), box (64, PW, PH, PN)
  CUtensorMap tmX[4];  // x   (Cin, W/s, H/s, N) per input parity, box (64, PW, PH, PN)
  ConvTap taps[16];    // kbase = tap * Cin (column of the tap inside one dW row)
  int32_t ntaps;
  int32_t cboxes;            // 64-channel boxes per tap = ceil(Cin / 64)
  int32_t boxes_total;       // ntaps * cboxes; M tile t = boxes 2t, 2t+1
  int32_t m_tiles, MT, m_groups;  // 128-row M tiles, tiles per work item (TMEM accumulators), items along M
  int32_t nb, BLOCK_N, n_tiles;   // dy boxes per N tile, N tile width (multiple of 64, <= 256), N tiles
  int32_t KP, PW, PH, PN;    // pixels per pipeline stage (GEMM-K chunk) and its patch shape
  int32_t tiles_w, tiles_h, tiles_n, ptiles;
  int32_t splits;
  int32_t exact_n;           // 1: MMA N = real channels of the N tile (multiple of 16) instead of BLOCK_N
  FDiv fd_nt, fd_mg, fd_tw, fd_th, fd_cb;  // item / pixel-tile / box index decoding
  int32_t Cout, Cin, Cin_pad, ldo;  // ldo = ntaps * Cin = row length of dW
  int32_t Mpad, Npad;        // partial tile: [Mpad = ntaps*Cin_pad][Npad = n_tiles*BLOCK_N]
  int32_t stages;
  uint32_t box_bytes, stage_bytes;
  float* partial;            // [splits][Mpad][Npad]
};

// ---- "patch" variant (conv_wgrad_patch.cu): 3x3 / stride 1 with the nine shifted x operands read from ONE halo patch --
struct WPatchKParams {
  CUtensorMap tmDY;        // dy (Cout, W, H, N), box (64, PW, PH, 1)
  CUtensorMap tmX;         // x  (Cin,  W, H, N), box (64, PW + 2, PH + 2, 1)
  int32_t tap_off16[9];    // (kh * pitch + kw) * 128 / 16: start of tap (kh, kw) inside the patch, descriptor units
  int32_t cboxes;          // 64-channel chunks of Cin (one patch each)
  int32_t MT, m_groups;    // M tiles (pairs of taps) per work item, items per chunk; 5 tiles per chunk: (0,1)(2,3)(4,5)(6,7)(8,-)
  int32_t nb, BLOCK_N, n_tiles;
  int32_t PW, PH, KP, pitch;
  int32_t a_step16, a_sbo;  // A start advance per K = 16 step (descriptor units) and byte stride between its two 8-pixel groups
  int32_t tiles_w, tiles_h, ptiles, splits;
  int32_t exact_n;
  FDiv fd_nt, fd_mg, fd_cb, fd_tw, fd_th;
  int32_t Cout, Cin, Cin_pad, Mpad, Npad;
  int32_t stages;
  uint32_t patch_tx, patch_bytes, box_bytes, stage_bytes;  // patch_tx = TMA box bytes, patch_bytes = aligned slot
  float* partial;          // [splits][Mpad = 9 * Cin_pad][Npad]  (same layout as WgradKParams::partial)
};

struct WgradPlan {
  WgradKParams kp;
  WPatchKParams pp;
  int kind;  // 0 = conv_wgrad_kernel (kp), 1 = conv_wgrad_patch_kernel (pp)
  int grid;
  int smem;
};
// patch variant: 1 = shape not eligible (use the generic kernel), 0 = planned, < 0 error
int wgrad_patch_plan(WgradPlan& pl, const TView& x, const TView& dy, int ks, int stride, float* partial,
                     size_t partial_floats, int max_splits);
int wgrad_patch_launch(const WgradPlan& pl, cudaStream_t st);
int wgrad_max_grid();
// Shared-memory budget of the operand ring of the weight-gradient kernels.  Default: the whole 227 KB.  YB_WGRAD_SMEM_KB
// (128..227) shrinks it so that the CTAs of the HBM-bound BN/SiLU-backward passes (<= 17 KB each) fit on the same SM while
// a weight-gradient kernel runs on the side stream (yolov5m_b200/model.py, _Engine._wgrad_async).
size_t wgrad_smem_budget();
int wgrad_waves();    // YB_WGRAD_WAVES: split-K work items per SM (waves of the persistent grid), default 1
int wgrad_exact_n();  // YB_WGRAD_EXACT_N (default 1)

// x: conv input (N,H,W,Cin); dy: grad of the conv output (N,H/s,W/s,Cout).
int wgrad_plan(WgradPlan& pl, const TView& x, const TView& dy, int ks, int stride, float* partial,
               size_t partial_floats, int max_splits);
// runs the split-K GEMM and the ordered reduction + transpose:
//   out[map ? map[i] : i] (+)= sum_s partial[s][tap*Cin_pad+ci][co],  i = co*ldo + tap*Cin + ci,  co < out_rows
int wgrad_run(const WgradPlan& pl, float* out, int out_rows, const int* map, int accumulate, cudaStream_t st, int phase = 0);
// minimum workspace (one split) for a layer
size_t wgrad_min_workspace_floats(int Cin, int Cout, int ks);

}  // namespace yb
