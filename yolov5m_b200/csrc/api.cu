// extern "C" entry points (include/yolov5m_b200.h).
#include "../../include/yolov5m_b200.h"

#include "common.cuh"
#include "conv_igemm.cuh"

namespace yb {
const char* last_error();
}
using namespace yb;

extern "C" {

const char* yb_last_error(void) { return yb::last_error(); }
int yb_version(void) { return 1; }
int yb_conv_max_partials(void) { return conv_max_grid(); }

int yb_conv2d_fwd(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* w_packed, int Cout,
                  int ks, int stride, void* y, int64_t y_pitch, int out_kind, const float* scale,
                  const float* shift, int act, const void* addend, int64_t addend_pitch, float* stats,
                  int* stats_rows, int head_na, int head_no, void* stream) {
  TView in{const_cast<void*>(x), N, H, W, Cin, (long)x_pitch};
  TView out{y, N, H / stride, W / stride, Cout, (long)y_pitch};
  ConvEpilogue ep;
  ep.out_kind = out_kind;
  ep.scale = scale;
  ep.shift = shift;
  ep.act = act;
  ep.addend = reinterpret_cast<const bf16*>(addend);
  ep.addend_pitch = (long)addend_pitch;
  ep.stats = stats;
  ep.head_na = head_na;
  ep.head_no = head_no;
  ConvPlan pl;
  int rc = conv_plan_fwd(pl, in, reinterpret_cast<const bf16*>(w_packed), ks, stride, out, ep);
  if (rc) return rc;
  if (stats_rows) *stats_rows = conv_stats_rows(pl);
  return conv_run(pl, reinterpret_cast<cudaStream_t>(stream));
}

int yb_conv2d_dgrad(const void* dy, int N, int H, int W, int Cout, int64_t dy_pitch, const void* wt_packed, int Cin,
                    int ks, int stride, void* dx, int64_t dx_pitch, const void* addend, int64_t addend_pitch,
                    void* stream) {
  TView g{const_cast<void*>(dy), N, H / stride, W / stride, Cout, (long)dy_pitch};
  TView o{dx, N, H, W, Cin, (long)dx_pitch};
  ConvEpilogue ep;
  ep.addend = reinterpret_cast<const bf16*>(addend);
  ep.addend_pitch = (long)addend_pitch;
  ConvPlan pl;
  int rc = conv_plan_dgrad(pl, g, reinterpret_cast<const bf16*>(wt_packed), ks, stride, o, ep);
  if (rc) return rc;
  return conv_run(pl, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
