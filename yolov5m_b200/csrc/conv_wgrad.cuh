// Weight-gradient convolution on tcgen05 (sm_100a): shared declarations.  See conv_wgrad.cu.
#pragma once
#include "conv_igemm.cuh"

namespace yb {

struct WgradKParams {
  CUtensorMap tmA;     // dy  (Cout, Wo, Ho, N), box (KCA, PW, PH, PN)
  CUtensorMap tmB[4];  // x   (Cin, W/s, H/s, N) per input parity, box (KCB, PW, PH, PN)
  ConvTap taps[16];    // kbase = tap * Cin (column of the tap inside one dW row)
  int32_t ntaps;
  int32_t KCA, KCB;          // channels per TMA box (16/32/64 <-> 32/64/128-byte swizzle)
  int32_t a_boxes, b_boxes;  // boxes per 128-row M tile / per N tile
  int32_t BLOCK_N;
  int32_t KP, PW, PH, PN;    // pixels per pipeline stage (GEMM-K chunk) and its patch shape
  int32_t tiles_w, tiles_h, tiles_n, ptiles;
  int32_t m_tiles, n_tiles, splits;
  int32_t Cout, Cin, ldo;    // ldo = ntaps * Cin = row length of dW
  int32_t stages;
  uint32_t a_stage_bytes, b_stage_bytes;
  float* partial;            // [splits][Cout][ldo]
};

struct WgradPlan {
  WgradKParams kp;
  int grid;
  int smem;
};

// x: conv input (N,H,W,Cin); dy: grad of the conv output (N,H/s,W/s,Cout).
int wgrad_plan(WgradPlan& pl, const TView& x, const TView& dy, int ks, int stride, float* partial,
               size_t partial_floats, int max_splits);
// runs the split-K GEMM and the ordered reduction: out[map ? map[i] : i] (+)= sum_s partial[s][i], i over [Cout][ldo]
// (i over the first out_rows rows of [Cout][ldo])
int wgrad_run(const WgradPlan& pl, float* out, int out_rows, const int* map, int accumulate, cudaStream_t st);

}  // namespace yb
