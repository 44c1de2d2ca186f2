// Box decode + per-image NMS on the GPU (sm_100a; HBM/latency-bound fp32 + integer work, no tensor cores).
//
// Replaces, for the detect / eval path of the reference:
//   cells_to_bboxes(is_pred=True)   utils/plot_utils.py:10-40   sigmoid, grid/anchor decode, class argmax
//   non_max_suppression             utils/bboxes_utils.py:175-209  confidence filter, xywh->xyxy, +cls offset,
//                                   torchvision.ops.nms (stable descending sort + greedy IoU suppression), first max_det rows
//
// decode_kernel     one warp per cell: coalesced 85-float row, warp arg-max (first maximum wins, like torch.argmax)
// nms_image_kernel  ONE CTA PER IMAGE (B = 128 images ~ one per SM), three phases, no host round trip:
//   1. ordered compaction of the candidates with score > threshold (block prefix sum -> original order preserved)
//   2. stable LSD radix sort (4 x 8 bit) on the descending-score key; every warp owns a contiguous segment so its
//      per-digit running offsets live in shared memory and ties keep their original order (= stable sort of
//      torchvision: ties -> lower index first)
//   3. greedy suppression in tiles of 512 sorted candidates: test against the kept list (<= max_det boxes in shared
//      memory), then a 512x512 bit matrix + warp-serial resolve inside the tile; stops at max_det keeps, which
//      yields exactly the first max_det rows the reference keeps after its full N^2 pass (bboxes_utils.py:202-203).
// IoU arithmetic uses explicitly rounded fp32 ops (no FMA contraction) in torchvision's operation order so keep
// sets are bit-exact.
#include "../../include/yolov5m_b200.h"
#include "common.cuh"

#include <algorithm>

namespace yb {

__device__ __forceinline__ float sigmoid_t(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------ decode
__global__ void __launch_bounds__(256) decode_rows_kernel(const float* __restrict__ p, long cells, int na, int H, int W, int no,
                                                     float stride, const float* __restrict__ anchors_px, int is_pred,
                                                     float* __restrict__ out, long rows_per_image, long level_off) {
  const int lane = threadIdx.x & 31;
  const long wglobal = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const long per_img = (long)na * H * W;
  const int nc = no - 5;
  for (long c = wglobal; c < cells; c += nwarps) {
    const float* ps = p + c * no;
    if (!is_pred) {  // target tensors (plot_utils.py:29-34): no sigmoid, class id stored in channel 5
      if (lane < 6) {
        const long b = c / per_img, rem = c - b * per_img;
        const long sp = rem % ((long)H * W);
        const int gy = (int)(sp / W), gx = (int)(sp - (long)gy * W);
        float v;
        if (lane == 0) v = ps[5];
        else if (lane == 1) v = ps[4];
        else if (lane < 4) v = __fmul_rn(__fadd_rn(ps[lane - 2], lane == 2 ? (float)gx : (float)gy), stride);
        else v = __fmul_rn(ps[lane - 2], stride);
        out[(b * rows_per_image + level_off + rem) * 6 + lane] = v;
      }
      continue;
    }
    // class arg-max over sigmoid(logit): first maximum wins (torch.argmax), plot_utils.py:27.  sigmoid is monotonic, so
    // only classes whose logit is within rounding reach of the largest one can hold the maximum sigmoid: find the max
    // LOGIT first (no special-function work), then evaluate sigmoid only for those candidates (ties after rounding --
    // close logits, or saturation to 1.0f above ~16.6 -- still resolve to the first index, bit-exactly).
    float lv[3];
    float lmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = lane + 32 * i;
      lv[i] = k < nc ? ps[5 + k] : -INFINITY;
      lmax = fmaxf(lmax, lv[i]);
    }
    for (int k = lane + 96; k < nc; k += 32) lmax = fmaxf(lmax, ps[5 + k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    const float cut = fminf(lmax, 16.0f) - fmaxf(1e-3f * fmaxf(1.0f, fabsf(lmax)), 2.4e-7f * __expf(fminf(lmax, 17.0f)));
    float bv = -1.f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = lane + 32 * i;
      if (k < nc && lv[i] >= cut) {
        const float s = sigmoid_t(lv[i]);
        if (s > bv) {
          bv = s;
          bi = k;
        }
      }
    }
    for (int k = lane + 96; k < nc; k += 32) {
      const float l = ps[5 + k];
      if (l >= cut) {
        const float s = sigmoid_t(l);
        if (s > bv) {
          bv = s;
          bi = k;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane < 6) {
      // 32-bit index math (cells < 2^31 is checked on the host): 64-bit divisions cost more than the rest of the cell
      const unsigned cu = (unsigned)c, hw = (unsigned)(H * W);
      const unsigned bu = cu / (unsigned)per_img, remu = cu - bu * (unsigned)per_img;
      const long b = bu, rem = remu;
      const int a = (int)(remu / hw);
      const unsigned sp = remu - (unsigned)a * hw;
      const int gy = (int)(sp / (unsigned)W), gx = (int)(sp - (unsigned)gy * (unsigned)W);
      float v;
      if (lane == 0) {
        v = (float)bi;
      } else if (lane == 1) {
        v = sigmoid_t(ps[4]);  // :24
      } else if (lane < 4) {
        const float s = sigmoid_t(ps[lane - 2]);
        const float g = lane == 2 ? (float)gx : (float)gy;
        v = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(2.f, s), g), 0.5f), stride);  // (2*s + grid - 0.5) * stride, :25
      } else {
        const float s = sigmoid_t(ps[lane - 2]);
        const float t = __fmul_rn(2.f, s);
        const float an = anchors_px[a * 2 + (lane - 4)];
        v = __fmul_rn(__fmul_rn(t, t), an);  // (2*s)**2 * anchor_grid, :26
      }
      out[(b * rows_per_image + level_off + rem) * 6 + lane] = v;
    }
  }
}


// ------------------------------------------------------------------------------------------------ decode of predictions
// cells_to_bboxes(is_pred=True) is HBM-bound: 340 bytes of logits in, 24 bytes out per cell (1280x1280, bs=128: 4.39 GB).
// A warp owns a chunk of 32 consecutive cells = 10,880 contiguous, 16-byte-aligned bytes: ONE bulk asynchronous copy
// (cp.async.bulk, the TMA engine) brings the chunk into the warp's shared-memory buffer (16 warps per SM: while some wait
// for their copy the others decode), then every lane decodes one cell from shared memory (row stride 85 words is odd ->
// conflict-free), and the 32 x 6 results leave through a shared-memory transpose as coalesced stores.
// Arithmetic is unchanged from the row kernel: exact expf sigmoid for the five box / objectness channels, class arg-max =
// first maximum of sigmoid(logit) (torch.argmax, plot_utils.py:27) found on the logits, with the sigmoid evaluated only for
// the classes whose logit is within rounding reach of the largest one.  The reach: sigmoid is flat to one fp32 ulp over
// dx ~ 6e-8 * e^x for large x (and saturates to 1.0f above ~16.6), so the candidate cut is
// min(lmax, 16) - max(1e-3 * max(1,|lmax|), 2.4e-7 * e^min(lmax,17)).
static constexpr int kDecWarps = 8;           // warps per CTA (two CTAs per SM: 16 chunks = 170 KB in flight / being decoded)
static constexpr int kDecCells = 32;          // cells per chunk (one per lane)
static constexpr int kDecMaxNo = 96;          // row length limit (5 + nc <= 96)

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct DecGeom {
  unsigned per_img, hw, w;  // cells per image (na*H*W), per anchor plane (H*W), per row (W)
};

__global__ void __launch_bounds__(kDecWarps * 32, 2) decode_pred_kernel(const float* __restrict__ p, long cells, int no,
                                                                        float stride, const float* __restrict__ anchors_px,
                                                                        float* __restrict__ out, long rows_per_image,
                                                                        long level_off, const DecGeom g) {
  extern __shared__ __align__(128) uint8_t dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t buf_bytes = ((uint32_t)kDecCells * no * 4u + 127u) & ~127u;
  float* buf = reinterpret_cast<float*>(dsm + (size_t)warp * (buf_bytes + 1024));
  float* stage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(buf) + buf_bytes);  // [32][6] output transpose
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + 32 * 6);
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const long nchunks = (cells + kDecCells - 1) / kDecCells;
  const long wglobal = (long)blockIdx.x * kDecWarps + warp, wstride = (long)gridDim.x * kDecWarps;
  const int nc = no - 5;
  const float a0w = anchors_px[0], a0h = anchors_px[1], a1w = anchors_px[2], a1h = anchors_px[3];
  uint32_t phase = 0;
  for (long ch = wglobal; ch < nchunks; ch += wstride) {
    const long c0 = ch * kDecCells;
    const int n = (int)min((long)kDecCells, cells - c0);
    {  // a chunk whose byte count and start are multiples of 16 goes through the bulk-copy engine; the (rare) unaligned
       // tail chunk of odd-sized maps is copied with plain loads
      const uint32_t bytes = (uint32_t)n * no * 4u;
      const float* src = p + c0 * no;
      if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        if (lane == 0) {
          mbar_expect_tx(bar, bytes);
          bulk_g2s(buf, src, bytes, bar);
        }
      } else {
        for (uint32_t i = lane; i < (uint32_t)n * no; i += 32) buf[i] = src[i];
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
      }
    }
    // index decode of this lane's cell while the copy is in flight
    // (32-bit unsigned divisions, three per lane and chunk: cells < 2^31 is checked on the host)
    const unsigned cu = (unsigned)(c0 + lane);
    const unsigned remu = cu % g.per_img;
    const int a = (int)(remu / g.hw);
    const unsigned sp = remu - (unsigned)a * g.hw;
    const int gy = (int)(sp / g.w), gx = (int)(sp - (unsigned)gy * g.w);
    mbar_wait(bar, phase);
    phase ^= 1;
    if (lane < n) {
      const float* ps = buf + lane * no;
      // one pass: largest and second-largest class logit (first index of the largest)
      float m1 = -INFINITY, m2 = -INFINITY;
      int i1 = 0;
#pragma unroll 8
      for (int k = 0; k < nc; ++k) {
        const float l = ps[5 + k];
        const bool gt = l > m1;
        m2 = gt ? m1 : fmaxf(m2, l);
        i1 = gt ? k : i1;
        m1 = gt ? l : m1;
      }
      int bi = i1;
      const float reach = fmaxf(1e-3f * fmaxf(1.0f, fabsf(m1)), 2.4e-7f * __expf(fminf(m1, 17.0f)));
      const float cut = fminf(m1, 16.0f) - reach;
      if (m2 >= cut || !(m1 == m1)) {  // another class within rounding reach (or NaN): resolve on the sigmoids, first maximum wins
        float bv = -1.f;
        bi = 0;
        for (int k = 0; k < nc; ++k) {
          const float l = ps[5 + k];
          if (l >= cut || l != l) {
            const float sg = sigmoid_t(l);
            if (sg > bv || (sg != sg && bv == bv)) {
              bv = sg;
              bi = k;
            }
          }
        }
      }
      const float sx = sigmoid_t(ps[0]), sy = sigmoid_t(ps[1]), sw = sigmoid_t(ps[2]), sh = sigmoid_t(ps[3]);
      const float tw = __fmul_rn(2.f, sw), th = __fmul_rn(2.f, sh);
      const float aw = a == 0 ? a0w : (a == 1 ? a1w : anchors_px[a * 2]);
      const float ah = a == 0 ? a0h : (a == 1 ? a1h : anchors_px[a * 2 + 1]);
      float* so = stage + lane * 6;
      so[0] = (float)bi;
      so[1] = sigmoid_t(ps[4]);                                                                  // plot_utils.py:24
      so[2] = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(2.f, sx), (float)gx), 0.5f), stride);      // :25
      so[3] = __fmul_rn(__fsub_rn(__fadd_rn(__fmul_rn(2.f, sy), (float)gy), 0.5f), stride);
      so[4] = __fmul_rn(__fmul_rn(tw, tw), aw);                                                  // :26
      so[5] = __fmul_rn(__fmul_rn(th, th), ah);
    }
    __syncwarp();
    // coalesced output: the chunk's 32 x 6 floats are consecutive in `out` unless the chunk straddles an image boundary
    const unsigned q0 = (unsigned)c0 / g.per_img, r0 = (unsigned)c0 - q0 * g.per_img;
    const unsigned q1 = (unsigned)(c0 + n - 1) / g.per_img;
    if (q0 == q1) {
      float* dst = out + ((long)q0 * rows_per_image + level_off + r0) * 6;
      for (int i = lane; i < n * 6; i += 32) dst[i] = stage[i];
    } else {
      for (int i = lane; i < n * 6; i += 32) {
        const int cell = i / 6, f = i - cell * 6;
        const unsigned qq = (unsigned)(c0 + cell) / g.per_img, rr = (unsigned)(c0 + cell) - qq * g.per_img;
        out[((long)qq * rows_per_image + level_off + rr) * 6 + f] = stage[i];
      }
    }
    __syncwarp();  // the buffer and the stage are free for the next chunk
  }
}

// ------------------------------------------------------------------------------------------------ NMS
struct NmsParams {
  const float* boxes;  // (B,N,6)
  long N;
  float iou_thr, thr;
  int max_det;
  uint32_t *keys0, *keys1, *idx0, *idx1;  // [B][N] scratch
  float* out;                              // [B][max_det][6]
  int* out_count;                          // [B]
  int* out_index;                          // [B][max_det] or null
  int* cand_count;                         // [B] or null
};

static constexpr int kNmsThreads = 1024;
static constexpr int kTile = 512;
static constexpr int kTileWords = kTile / 32;

__device__ __forceinline__ bool iou_gt(float ix1, float iy1, float ix2, float iy2, float iarea, float jx1, float jy1,
                                       float jx2, float jy2, float jarea, float thr) {
  const float xx1 = fmaxf(ix1, jx1), yy1 = fmaxf(iy1, jy1), xx2 = fminf(ix2, jx2), yy2 = fminf(iy2, jy2);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
  return ovr > thr;
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_image_kernel(const __grid_constant__ NmsParams P) {
  extern __shared__ uint8_t smem_raw[];
  // layout: whist[32][256] ints (32 KB, reused as the tile bit matrix), digit_total[256], tile arrays, kept arrays
  int* whist = reinterpret_cast<int*>(smem_raw);
  uint32_t* tmask = reinterpret_cast<uint32_t*>(smem_raw);  // [kTile][kTileWords] = 32 KB (phase 3)
  int* digit_total = whist + 32 * 256;
  float* tbox = reinterpret_cast<float*>(digit_total + 256);  // [5][kTile]: ox1, oy1, ox2, oy2, area
  float* trow = tbox + 5 * kTile;                              // [kTile][6]  output rows
  int* tidx = reinterpret_cast<int*>(trow + 6 * kTile);        // [kTile]
  uint32_t* talive = reinterpret_cast<uint32_t*>(tidx + kTile);  // [kTileWords]
  float* kbox = reinterpret_cast<float*>(talive + kTileWords);   // [5][max_det]
  __shared__ int s_warp_tot[32];
  __shared__ int s_running;
  __shared__ int s_skip;
  __shared__ int s_kept;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long img = blockIdx.x;
  const float* boxes = P.boxes + img * P.N * 6;
  uint32_t* kA = P.keys0 + img * P.N;
  uint32_t* kB = P.keys1 + img * P.N;
  uint32_t* vA = P.idx0 + img * P.N;
  uint32_t* vB = P.idx1 + img * P.N;

  // ---------------------------------------------------------------- phase 1: filter (ordered compaction)
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (long base = 0; base < P.N; base += kNmsThreads) {
    const long i = base + tid;
    float sc = 0.f;
    bool take = false;
    if (i < P.N) {
      sc = boxes[i * 6 + 1];
      take = sc > P.thr;  // bboxes_utils.py:186
    }
    const unsigned bal = __ballot_sync(0xffffffffu, take);
    if (lane == 0) s_warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = s_running;
    for (int w2 = 0; w2 < warp; ++w2) before += s_warp_tot[w2];
    if (take) {
      const int pos = before + __popc(bal & ((1u << lane) - 1u));
      uint32_t u = __float_as_uint(sc);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
      kA[pos] = ~u;                                     // descending score = ascending key
      vA[pos] = (uint32_t)i;
    }
    __syncthreads();
    if (tid == 0) {
      int s = s_running;
      for (int w2 = 0; w2 < 32; ++w2) s += s_warp_tot[w2];
      s_running = s;
    }
    __syncthreads();
  }
  const int n = s_running;
  if (tid == 0 && P.cand_count != nullptr) P.cand_count[img] = n;

  // ---------------------------------------------------------------- phase 2: stable LSD radix sort, 4 x 8 bits
  const int seg = ((((n + 31) / 32) + 31) / 32) * 32;  // per-warp contiguous segment, multiple of 32
  const int s0 = min(n, warp * seg), s1 = min(n, (warp + 1) * seg);
  for (int pass = 0; pass < 4 && n > 1; ++pass) {
    const int shift = pass * 8;
    for (int i = tid; i < 32 * 256; i += kNmsThreads) whist[i] = 0;
    __syncthreads();
    int* myh = whist + warp * 256;
    for (int base = s0; base < s1; base += 32) {
      const int i = base + lane;
      const bool act = i < s1;
      const uint32_t d = act ? ((kA[i] >> shift) & 255u) : (256u + lane);
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (act && (peers & ((1u << lane) - 1u)) == 0) myh[d] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    if (tid < 256) {  // column prefix over warps
      int run = 0;
      for (int w2 = 0; w2 < 32; ++w2) {
        const int c = whist[w2 * 256 + tid];
        whist[w2 * 256 + tid] = run;
        run += c;
      }
      digit_total[tid] = run;
    }
    __syncthreads();
    if (tid == 0) {
      int run = 0, skip = 0;
      for (int d = 0; d < 256; ++d) {
        const int c = digit_total[d];
        if (c == n) skip = 1;  // every key has the same digit: the pass is the identity
        digit_total[d] = run;
        run += c;
      }
      s_skip = skip;
    }
    __syncthreads();
    if (s_skip) continue;
    for (int i = tid; i < 32 * 256; i += kNmsThreads) whist[i] += digit_total[i & 255];
    __syncthreads();
    for (int base = s0; base < s1; base += 32) {
      const int i = base + lane;
      const bool act = i < s1;
      uint32_t key = 0, val = 0;
      if (act) {
        key = kA[i];
        val = vA[i];
      }
      const uint32_t d = act ? ((key >> shift) & 255u) : (256u + lane);
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      int off = 0;
      if (act) off = myh[d];
      __syncwarp();
      if (act) {
        kB[off + rank] = key;
        vB[off + rank] = val;
        if (rank == 0) myh[d] = off + __popc(peers);
      }
      __syncwarp();
    }
    __syncthreads();
    uint32_t* t = kA; kA = kB; kB = t;
    t = vA; vA = vB; vB = t;
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 3: greedy suppression, early exit at max_det
  if (tid == 0) s_kept = 0;
  __syncthreads();
  float* out = P.out + img * (long)P.max_det * 6;
  for (int base = 0; base < n; base += kTile) {
    const int kept0 = s_kept;
    if (kept0 >= P.max_det) break;
    const int tn = min(kTile, n - base);
    if (tid < kTileWords) talive[tid] = 0;
    __syncthreads();
    if (tid < kTile) {
      bool alive = false;
      if (tid < tn) {
        const uint32_t ci = vA[base + tid];
        const float* r = boxes + (long)ci * 6;
        const float cls = r[0], sc = r[1], cx = r[2], cy = r[3], w = r[4], h = r[5];
        const float x1 = __fsub_rn(cx, __fdiv_rn(w, 2.f));  // :190
        const float y1 = __fsub_rn(cy, __fdiv_rn(h, 2.f));  // :191
        const float y2 = __fadd_rn(h, y1);                  // :192
        const float x2 = __fadd_rn(w, x1);                  // :193
        const float ox1 = __fadd_rn(x1, cls), oy1 = __fadd_rn(y1, cls), ox2 = __fadd_rn(x2, cls), oy2 = __fadd_rn(y2, cls);  // :195
        const float area = __fmul_rn(__fsub_rn(ox2, ox1), __fsub_rn(oy2, oy1));
        tbox[0 * kTile + tid] = ox1; tbox[1 * kTile + tid] = oy1; tbox[2 * kTile + tid] = ox2; tbox[3 * kTile + tid] = oy2;
        tbox[4 * kTile + tid] = area;
        float* tr = trow + tid * 6;
        tr[0] = cls; tr[1] = sc; tr[2] = x1; tr[3] = y1; tr[4] = x2; tr[5] = y2;
        tidx[tid] = (int)ci;
        alive = true;
        for (int k = 0; k < kept0; ++k) {
          if (iou_gt(kbox[k], kbox[P.max_det + k], kbox[2 * P.max_det + k], kbox[3 * P.max_det + k], kbox[4 * P.max_det + k],
                     ox1, oy1, ox2, oy2, area, P.iou_thr)) {
            alive = false;
            break;
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) talive[tid >> 5] = bal;
    }
    __syncthreads();
    {  // bit matrix among the tile's survivors: two threads per row
      const int r = tid >> 1, half = tid & 1;
      const bool ra = r < tn && ((talive[r >> 5] >> (r & 31)) & 1u);
      const float rx1 = tbox[r], ry1 = tbox[kTile + r], rx2 = tbox[2 * kTile + r], ry2 = tbox[3 * kTile + r], rar = tbox[4 * kTile + r];
      for (int wd = half * (kTileWords / 2); wd < (half + 1) * (kTileWords / 2); ++wd) {
        uint32_t bits = 0;
        if (ra && wd * 32 + 31 > r) {
          const uint32_t al = talive[wd];
          for (int bbit = 0; bbit < 32; ++bbit) {
            const int j = wd * 32 + bbit;
            if (j > r && ((al >> bbit) & 1u) &&
                iou_gt(rx1, ry1, rx2, ry2, rar, tbox[j], tbox[kTile + j], tbox[2 * kTile + j], tbox[3 * kTile + j],
                       tbox[4 * kTile + j], P.iou_thr))
              bits |= 1u << bbit;
          }
        }
        tmask[r * kTileWords + wd] = bits;
      }
    }
    __syncthreads();
    if (warp == 0) {  // serial resolve over the survivors (in score order)
      uint32_t aw = lane < kTileWords ? talive[lane] : 0u;
      int kept = kept0;
      while (kept < P.max_det) {
        int pos = aw ? (lane * 32 + __ffs(aw) - 1) : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pos = min(pos, __shfl_xor_sync(0xffffffffu, pos, o));
        if (pos == 0x7fffffff) break;
        if (lane < kTileWords) {
          aw &= ~tmask[pos * kTileWords + lane];
          if ((pos >> 5) == lane) aw &= ~(1u << (pos & 31));
        }
        if (lane < 5) kbox[lane * P.max_det + kept] = tbox[lane * kTile + pos];
        if (lane < 6) out[(long)kept * 6 + lane] = trow[pos * 6 + lane];
        if (lane == 6 && P.out_index != nullptr) P.out_index[img * P.max_det + kept] = tidx[pos];
        ++kept;
      }
      if (lane == 0) s_kept = kept;
    }
    __syncthreads();
  }
  if (tid == 0) P.out_count[img] = s_kept;
}

// ------------------------------------------------------------------------------------------------ eval counters
// YOLO_EVAL.check_class_accuracy (utils/validation_utils.py:58-69) for one level: over the cells the label tensor marks as
// objects (y[...,4] == 1):  class hit = argmax(logits[5:]) == y[...,5] (first maximum, torch.argmax);
// "obj" hit = (sigmoid(logit[0]) > conf) == y[...,4] -- the reference thresholds channel 0, not the objectness (App. B5);
// mirrored.  counters[0] += objects, [1] += class hits, [2] += obj hits.
__global__ void class_accuracy_kernel(const float* __restrict__ p, const float* __restrict__ y, long cells, int no, int ny,
                                      float conf, unsigned long long* __restrict__ counters) {
  unsigned long long tot = 0, cc = 0, co = 0;
  for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long)gridDim.x * blockDim.x) {
    const float* yr = y + c * ny;
    if (yr[4] != 1.0f) continue;
    const float* ps = p + c * no;
    float best = ps[5];
    int bi = 0;
    for (int k = 1; k < no - 5; ++k) {
      const float v = ps[5 + k];
      if (v > best || (v != v && best == best)) {  // torch.argmax: first maximum; NaN counts as the maximum
        best = v;
        bi = k;
      }
    }
    ++tot;
    if ((float)bi == yr[5]) ++cc;
    if (sigmoid_t(ps[0]) > conf) ++co;  // True == 1.0 (the label's objectness)
  }
  if (tot) {
    atomicAdd(&counters[0], tot);
    atomicAdd(&counters[1], cc);
    atomicAdd(&counters[2], co);
  }
}

static int nms_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

}  // namespace yb

using namespace yb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int yb_decode_level(const float* p, int B, int na, int H, int W, int no, float stride, const float* anchors_px, int is_pred,
                    float* out, int64_t rows_per_image, int64_t level_off, void* stream) {
  YB_REQUIRE(no >= 6 && na >= 1, "decode: no=%d na=%d", no, na);
  const long cells = (long)B * na * H * W;
  if (cells == 0) return 0;
  YB_REQUIRE(cells < (1L << 31), "decode: %ld cells (limit 2^31)", cells);
  if (is_pred && no <= kDecMaxNo) {
    const size_t buf = ((size_t)kDecCells * no * 4 + 127) & ~(size_t)127;
    const size_t smem = (size_t)kDecWarps * (buf + 1024);
    static size_t attr = 0;
    if (smem > attr) {
      YB_CHECK_CUDA(cudaFuncSetAttribute(decode_pred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    const long nchunks = (cells + kDecCells - 1) / kDecCells;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (size_t)(220 * 1024) / (smem + 1024)));
    const int blocks = (int)std::max<long>(1, std::min<long>((nchunks + kDecWarps - 1) / kDecWarps, (long)nms_sm_count() * per_sm));
    DecGeom g;
    g.per_img = (unsigned)(na * H * W);
    g.hw = (unsigned)(H * W);
    g.w = (unsigned)W;
    decode_pred_kernel<<<blocks, kDecWarps * 32, smem, ST(stream)>>>(p, cells, no, stride, anchors_px, out, rows_per_image,
                                                                     level_off, g);
    YB_LAUNCHED();
    return 0;
  }
  const int blocks = (int)std::max<long>(1, std::min<long>((cells + 7) / 8, (long)nms_sm_count() * 32));
  decode_rows_kernel<<<blocks, 256, 0, ST(stream)>>>(p, cells, na, H, W, no, stride, anchors_px, is_pred, out, rows_per_image,
                                                     level_off);
  YB_LAUNCHED();
  return 0;
}

int yb_class_accuracy(const float* p, const float* y, int64_t cells, int no, int ny, float conf, uint64_t* counters,
                      void* stream) {
  YB_REQUIRE(no >= 6 && ny >= 6, "class_accuracy: no=%d ny=%d", no, ny);
  if (cells == 0) return 0;
  const int blocks = (int)std::max<long>(1, std::min<long>((cells + 255) / 256, (long)nms_sm_count() * 8));
  class_accuracy_kernel<<<blocks, 256, 0, ST(stream)>>>(p, y, cells, no, ny, conf,
                                                        reinterpret_cast<unsigned long long*>(counters));
  YB_LAUNCHED();
  return 0;
}

int64_t yb_nms_scratch_bytes(int B, int64_t N) { return (int64_t)4 * sizeof(uint32_t) * B * N; }

int yb_nms_batched(const float* boxes, int B, int64_t N, float iou_threshold, float threshold, int max_det, void* scratch,
                   float* out, int* out_count, int* out_index, int* cand_count, void* stream) {
  YB_REQUIRE(max_det >= 1 && max_det <= 4096, "nms: max_det=%d (1..4096)", max_det);
  YB_REQUIRE(N < (1L << 31), "nms: N too large");
  if (B == 0) return 0;
  if (N == 0) {
    YB_CHECK_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int) * B, ST(stream)));
    if (cand_count) YB_CHECK_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * B, ST(stream)));
    return 0;
  }
  NmsParams P;
  P.boxes = boxes;
  P.N = N;
  P.iou_thr = iou_threshold;
  P.thr = threshold;
  P.max_det = max_det;
  uint32_t* s = reinterpret_cast<uint32_t*>(scratch);
  const size_t bn = (size_t)B * N;
  P.keys0 = s;
  P.keys1 = s + bn;
  P.idx0 = s + 2 * bn;
  P.idx1 = s + 3 * bn;
  P.out = out;
  P.out_count = out_count;
  P.out_index = out_index;
  P.cand_count = cand_count;
  const size_t smem = (size_t)(32 * 256 + 256) * 4 + (size_t)(5 * kTile + 6 * kTile + kTile + kTileWords) * 4 +
                      (size_t)5 * max_det * 4;
  static size_t attr = 0;
  if (smem > attr) {
    YB_CHECK_CUDA(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  nms_image_kernel<<<B, kNmsThreads, smem, ST(stream)>>>(P);
  YB_LAUNCHED();
  return 0;
}

}  // extern "C"
