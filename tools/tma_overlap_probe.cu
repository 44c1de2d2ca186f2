// Hardware probe (not part of the library): does a TILED tensor map accept a stride SMALLER than the extent of the dimension
// below it, i.e. overlapping windows?  The stem needs, per output pixel, the 16 channels of three horizontally adjacent
// space-to-depth pixels = 48 contiguous bf16 values starting one pixel to the left.  With dims (48, W, H) and strides
// (32 B, (W + 2) * 32 B) over a row-padded 16-channel tensor, that "tap-gathered" operand becomes a VIEW and the 48-channel
// staging tensor (3x the bytes) never has to exist.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o /tmp/tma_overlap_probe tools/tma_overlap_probe.cu \
//        yolov5m_b200/csrc/runtime.cu -lcuda
#include <cstdio>
#include <vector>

#include "../yolov5m_b200/csrc/common.cuh"

using namespace yb;

struct Params {
  CUtensorMap tm;
  bf16* out;  // [32 px][64] (box 64 channels wide, 48 in bounds)
};

__global__ void probe_kernel(const __grid_constant__ Params p, int w0, int h0) {
  __shared__ __align__(1024) uint8_t tile[32 * 128];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 32 * 128);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(tile)),
                 "l"(reinterpret_cast<uint64_t>(&p.tm)), "r"(smem_u32(&bar)), "r"(0), "r"(w0), "r"(h0)
                 : "memory");
  }
  mbar_wait(&bar, 0);
  // SWIZZLE_128B: 16-byte unit u of row r sits at u ^ (r & 7)
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int u = (c / 8) ^ (r & 7);
    p.out[i] = reinterpret_cast<bf16*>(tile + r * 128 + u * 16)[c % 8];
  }
}

int main() {
  const int W = 40, H = 6, WP = W + 2;  // padded rows
  std::vector<bf16> h((size_t)H * WP * 16);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < WP; ++x)
      for (int c = 0; c < 16; ++c) {
        const bool pad = x == 0 || x == WP - 1;
        h[((size_t)y * WP + x) * 16 + c] = __float2bfloat16(pad ? 0.f : (float)(y * 1000 + (x - 1) * 16 + c));
      }
  bf16 *d, *dout;
  cudaMalloc(&d, h.size() * 2);
  cudaMalloc(&dout, 32 * 64 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  Params p;
  uint64_t dims[3] = {48, (uint64_t)W, (uint64_t)H};
  uint64_t strides[2] = {32, (uint64_t)WP * 32};
  uint32_t box[3] = {64, 32, 1};
  const int rc = encode_tmap(&p.tm, d, 3, dims, strides, box, 128, 2);
  printf("encode rc=%d %s\n", rc, rc ? "(overlapping strides rejected)" : "(accepted)");
  if (rc) return 1;
  p.out = dout;
  int bad = 0;
  for (int trial = 0; trial < 2; ++trial) {
    const int w0 = trial == 0 ? 0 : 16, h0 = trial == 0 ? 2 : 5;  // second trial: box runs past W = 40 (8 pixels out of bounds)
    probe_kernel<<<1, 128>>>(p, w0, h0);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      return 2;
    }
    std::vector<bf16> o(32 * 64);
    cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
    for (int r = 0; r < 32; ++r)
      for (int c = 0; c < 64; ++c) {
        const int w = w0 + r;  // output pixel; window = original pixels w-1, w, w+1
        float want = 0.f;
        if (c < 48 && w < W) {
          const int xo = w - 1 + c / 16;  // original pixel
          want = (xo < 0 || xo >= W) ? 0.f : (float)(h0 * 1000 + xo * 16 + c % 16);
        }
        const float got = __bfloat162float(o[r * 64 + c]);
        if (got != __bfloat162float(__float2bfloat16(want))) {
          if (bad < 8) printf("trial %d mismatch r=%d c=%d got %g want %g\n", trial, r, c, got, want);
          ++bad;
        }
      }
  }
  printf(bad ? "OVERLAP VIEW: %d mismatches\n" : "OVERLAP VIEW OK (%d mismatches)\n", bad);
  return bad ? 3 : 0;
}
