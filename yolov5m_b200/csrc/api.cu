// extern "C" entry points (include/yolov5m_b200.h).
#include "../../include/yolov5m_b200.h"

#include "common.cuh"
#include "conv_igemm.cuh"
#include "conv_wgrad.cuh"

namespace yb {
void set_wgrad_patch_mode(int m);
const char* last_error();
long long launch_count();
}
using namespace yb;

extern "C" {

const char* yb_last_error(void) { return yb::last_error(); }
int yb_version(void) { return 1; }
int64_t yb_launch_count(void) { return (int64_t)yb::launch_count(); }
int yb_conv_max_partials(void) { return conv_max_grid(); }
void yb_set_conv_patch_mode(int mode) { set_patch_mode(mode); }
void yb_set_wgrad_patch_mode(int mode) { set_wgrad_patch_mode(mode); }

int yb_conv2d_fwd(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* w_packed, int Cout,
                  int ks, int stride, void* y, int64_t y_pitch, int out_kind, const float* scale,
                  const float* shift, int act, const void* addend, int64_t addend_pitch, float* stats,
                  int* stats_rows, int head_na, int head_no, void* stream) {
  TView in{const_cast<void*>(x), N, H, W, Cin, (long)x_pitch};
  stem_view(in, ks);
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
                    int out_kind, void* stream) {
  TView g{const_cast<void*>(dy), N, H / stride, W / stride, Cout, (long)dy_pitch};
  TView o{dx, N, H, W, Cin, (long)dx_pitch};
  ConvEpilogue ep;
  ep.out_kind = out_kind;
  ep.addend = reinterpret_cast<const bf16*>(addend);
  ep.addend_pitch = (long)addend_pitch;
  ConvPlan pl;
  int rc = conv_plan_dgrad(pl, g, reinterpret_cast<const bf16*>(wt_packed), ks, stride, o, ep);
  if (rc) return rc;
  return conv_run(pl, reinterpret_cast<cudaStream_t>(stream));
}


struct YbPlan {
  int kind;  // 0 conv (fwd/dgrad), 1 wgrad
  ConvPlan conv;
  WgradPlan wgrad;
};

void* yb_conv_fwd_plan(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* w_packed, int Cout,
                       int ks, int stride, void* y, int64_t y_pitch, int out_kind, const float* scale,
                       const float* shift, int act, const void* addend, int64_t addend_pitch, float* stats,
                       int* stats_rows, int head_na, int head_no) {
  TView in{const_cast<void*>(x), N, H, W, Cin, (long)x_pitch};
  stem_view(in, ks);
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
  YbPlan* pl = new YbPlan();
  pl->kind = 0;
  if (conv_plan_fwd(pl->conv, in, reinterpret_cast<const bf16*>(w_packed), ks, stride, out, ep)) {
    delete pl;
    return nullptr;
  }
  if (stats_rows) *stats_rows = conv_stats_rows(pl->conv);
  return pl;
}

void* yb_conv_dgrad_plan(const void* dy, int N, int H, int W, int Cout, int64_t dy_pitch, const void* wt_packed, int Cin,
                         int ks, int stride, void* dx, int64_t dx_pitch, const void* addend, int64_t addend_pitch,
                         int out_kind) {
  TView g{const_cast<void*>(dy), N, H / stride, W / stride, Cout, (long)dy_pitch};
  TView o{dx, N, H, W, Cin, (long)dx_pitch};
  ConvEpilogue ep;
  ep.out_kind = out_kind;
  ep.addend = reinterpret_cast<const bf16*>(addend);
  ep.addend_pitch = (long)addend_pitch;
  YbPlan* pl = new YbPlan();
  pl->kind = 0;
  if (conv_plan_dgrad(pl->conv, g, reinterpret_cast<const bf16*>(wt_packed), ks, stride, o, ep)) {
    delete pl;
    return nullptr;
  }
  return pl;
}

int yb_plan_run(void* plan, void* stream) {
  YbPlan* pl = reinterpret_cast<YbPlan*>(plan);
  YB_REQUIRE(pl != nullptr && pl->kind == 0, "yb_plan_run: not a conv plan");
  return conv_run(pl->conv, reinterpret_cast<cudaStream_t>(stream));
}

void yb_plan_destroy(void* plan) { delete reinterpret_cast<YbPlan*>(plan); }

int64_t yb_conv_wgrad_workspace_floats(int Cin, int Cout, int ks) {
  return (int64_t)wgrad_min_workspace_floats(Cin, Cout, ks);
}

void* yb_conv_wgrad_plan(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* dy, int Cout,
                         int64_t dy_pitch, int ks, int stride, float* workspace, int64_t workspace_floats,
                         int max_splits) {
  TView xv{const_cast<void*>(x), N, H, W, Cin, (long)x_pitch};
  stem_view(xv, ks);
  TView gv{const_cast<void*>(dy), N, H / stride, W / stride, Cout, (long)dy_pitch};
  YbPlan* pl = new YbPlan();
  pl->kind = 1;
  if (wgrad_plan(pl->wgrad, xv, gv, ks, stride, workspace, (size_t)workspace_floats, max_splits)) {
    delete pl;
    return nullptr;
  }
  return pl;
}

int yb_wgrad_plan_run(void* plan, float* dw, int out_rows, const int* index_map, int accumulate, void* stream) {
  YbPlan* pl = reinterpret_cast<YbPlan*>(plan);
  YB_REQUIRE(pl != nullptr && pl->kind == 1, "yb_wgrad_plan_run: not a wgrad plan");
  return wgrad_run(pl->wgrad, dw, out_rows, index_map, accumulate, reinterpret_cast<cudaStream_t>(stream));
}

int yb_wgrad_plan_run_phase(void* plan, float* dw, int out_rows, const int* index_map, int accumulate, int phase, void* stream) {
  YbPlan* pl = reinterpret_cast<YbPlan*>(plan);
  YB_REQUIRE(pl != nullptr && pl->kind == 1, "yb_wgrad_plan_run_phase: not a wgrad plan");
  YB_REQUIRE(phase >= 0 && phase <= 2, "yb_wgrad_plan_run_phase: phase=%d", phase);
  return wgrad_run(pl->wgrad, dw, out_rows, index_map, accumulate, reinterpret_cast<cudaStream_t>(stream), phase);
}

int yb_conv2d_wgrad(const void* x, int N, int H, int W, int Cin, int64_t x_pitch, const void* dy, int Cout,
                    int64_t dy_pitch, int ks, int stride, float* workspace, int64_t workspace_floats, int max_splits,
                    float* dw, int out_rows, const int* index_map, int accumulate, void* stream) {
  TView xv{const_cast<void*>(x), N, H, W, Cin, (long)x_pitch};
  stem_view(xv, ks);
  TView gv{const_cast<void*>(dy), N, H / stride, W / stride, Cout, (long)dy_pitch};
  WgradPlan pl;
  int rc = wgrad_plan(pl, xv, gv, ks, stride, workspace, (size_t)workspace_floats, max_splits);
  if (rc) return rc;
  return wgrad_run(pl, dw, out_rows, index_map, accumulate, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
