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
  FDiv fd_nt, fd_mg, fd_tw, fd_th, fd_cb;  // item / pixel-tile / box index decoding
  int32_t Cout, Cin, Cin_pad, ldo;  // ldo = ntaps * Cin = row length of dW
  int32_t Mpad, Npad;        // partial tile: [Mpad = ntaps*Cin_pad][Npad = n_tiles*BLOCK_N]
  int32_t stages;
  uint32_t box_bytes, stage_bytes;
  float* partial;            // [splits][Mpad][Npad]
};

struct WgradPlan {
  WgradKParams kp;
  int grid;
  int smem;
};

// x: conv input (N,H,W,Cin); dy: grad of the conv output (N,H/s,W/s,Cout).
int wgrad_plan(WgradPlan& pl, const TView& x, const TView& dy, int ks, int stride, float* partial,
               size_t partial_floats, int max_splits);
// runs the split-K GEMM and the ordered reduction + transpose:
//   out[map ? map[i] : i] (+)= sum_s partial[s][tap*Cin_pad+ci][co],  i = co*ldo + tap*Cin + ci,  co < out_rows
int wgrad_run(const WgradPlan& pl, float* out, int out_rows, const int* map, int accumulate, cudaStream_t st);
// minimum workspace (one split) for a layer
size_t wgrad_min_workspace_floats(int Cin, int Cout, int ks);

}  // namespace yb
