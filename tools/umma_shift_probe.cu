// Hardware probe (not part of the library): can a tcgen05 shared-memory descriptor address a SWIZZLE_128B tile that
// starts at an arbitrary 128-byte row (not 1024-byte aligned)?  This decides whether a 3x3 convolution can keep ONE halo
// patch per channel chunk in shared memory and feed all nine taps from it by shifting the descriptor start address
// ("halo reuse"), instead of fetching nine shifted copies through TMA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o umma_shift_probe tools/umma_shift_probe.cu \
//        yolov5m_b200/csrc/runtime.cu -lcuda
//
// Tests, each for row shifts j = 0..9 and base_offset mode in {0: field = 0, 1: field = (addr >> 7) & 7}:
//   K-major A  (fwd/dgrad form): D[m][n] = sum_k A[row(m) + j][k] * I[n][k]   with row(m) = m            (SBO = 1024)
//                                                                             or row(m) = (m/8)*16 + m%8 (SBO = 2048)
//                                                                             or row(m) = (m/8)*10 + m%8 (SBO = 1280)
//   MN-major A (wgrad form):     D[m][n] = sum_{k<16} X[j + k][m] * Y[k][n],  Y[k][n] = (n == k)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../yolov5m_b200/csrc/common.cuh"

using namespace yb;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt, int mode) {
  uint64_t d = make_smem_desc(saddr, lbo, sbo, lt);
  if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  return d;
}

struct Params {
  CUtensorMap tmA;  // [288 rows][64] bf16, box (64, 144)
  CUtensorMap tmB;  // [64 rows][64] bf16, box (64, 64)
  float* out;       // [variants][128][64]
  long long* cycles;
};

// variant v = ((test * 10 + j) * 2 + mode); test 0: K-major SBO 1024, 1: K-major SBO 2048, 2: MN-major
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* mbar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  uint8_t* a_smem = smem + 1024;                // 288 rows x 128 B = 36 KB
  uint8_t* b_smem = a_smem + 288 * 128;         // 64 rows x 128 B = 8 KB (36 KB is a multiple of 1024)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_base2 = tmem_base;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 288 * 128 + 64 * 128);
    tma_load_2d(&p.tmA, bar, a_smem, 0, 0);
    tma_load_2d(&p.tmA, bar, a_smem + 144 * 128, 0, 144);
    tma_load_2d(&p.tmB, bar, b_smem, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t mph = 0;
  for (int test = 0; test < 4; ++test)
    for (int j = 0; j < 10; ++j)
      for (int mode = 0; mode < 2; ++mode) {
        const int v = (test * 10 + j) * 2 + mode;
        if (threadIdx.x == 0) {
          const uint32_t a0 = smem_u32(a_smem) + j * 128, b0 = smem_u32(b_smem);
          if (test != 2) {
            const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
            const uint32_t sbo = test == 0 ? 1024 : (test == 1 ? 2048 : 1280);
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, desc_bo(a0 + k * 32, 16, sbo, 2, mode), make_smem_desc(b0 + k * 32, 16, 1024, 2), idesc,
                        k != 0);
          } else {
            // MN-major A: [K = pixel rows][M = 64 channels]; M = 128 needs two 64-channel boxes: reuse the same box twice
            // (LBO = 0 is not allowed to be meaningful here, so point LBO at the tile 144 rows further down)
            const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
            umma_bf16(tmem_base, desc_bo(a0, 144 * 128, 1024, 2, mode), make_smem_desc(b0, 16 * 128, 1024, 2), idesc, 0);
          }
          umma_commit(mbar);
        }
        mbar_wait(mbar, mph);
        mph ^= 1;
        tc_fence_after();
        float* o = p.out + ((size_t)v * 128 + warp * 32 + lane) * 64;
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t vr[16];
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 16, vr);
          tmem_ld_wait();
          for (int i = 0; i < 16; ++i) o[cc * 16 + i] = __uint_as_float(vr[i]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
      }
  // ---- timing: 512 back-to-back MMAs (M=128, N=64, K=16) per variant; cycles per MMA (ideal 32)
  for (int v = 0; v < 8; ++v) {
    const int j = v & 1;
    const uint32_t sbo = (v >> 1) == 0 ? 1024u : ((v >> 1) == 1 ? 1280u : ((v >> 1) == 2 ? 2048u : 2304u));
    long long t0 = 0;
    if (threadIdx.x == 0) {
      const uint32_t a0 = smem_u32(a_smem) + j * 128, b0 = smem_u32(b_smem);
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      t0 = clock64();
      for (int i = 0; i < 512; ++i)
        umma_bf16(tmem_base, make_smem_desc(a0 + (i & 3) * 32, 16, sbo, 2), make_smem_desc(b0 + (i & 3) * 32, 16, 1024, 2),
                  idesc, 1);
      umma_commit(mbar);
    }
    mbar_wait(mbar, mph);
    mph ^= 1;
    if (threadIdx.x == 0) p.cycles[v] = clock64() - t0;
    __syncthreads();
  }
  // ---- timing 2: accumulator placement.  N = 96 / 48, `nacc` accumulators `stride` columns apart, 4 MMAs (K = 64) per
  //      accumulator visit, round robin -- the issue pattern of the multi-tile conv kernels
  for (int v = 0; v < 8; ++v) {
    const int N = v < 4 ? 96 : 48;
    const int nacc = (v & 3) == 0 ? 1 : (v < 4 ? 2 : 4);
    const int stride = (v & 3) == 1 ? N : ((v & 3) == 2 ? 128 : ((v & 3) == 3 ? 64 : 0));
    long long t0 = 0;
    if (threadIdx.x == 0) {
      const uint32_t a0 = smem_u32(a_smem), b0 = smem_u32(b_smem);
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      t0 = clock64();
      for (int i = 0; i < 128; ++i) {
        const uint32_t d = tmem_base2 + (uint32_t)((i % nacc) * stride);
        for (int k = 0; k < 4; ++k)
          umma_bf16(d, make_smem_desc(a0 + k * 32, 16, 1024, 2), make_smem_desc(b0 + k * 32, 16, 1024, 2), idesc, 1);
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, mph);
    mph ^= 1;
    if (threadIdx.x == 0) p.cycles[8 + v] = clock64() - t0;
    __syncthreads();
  }
  // ---- timing 3: what slows the MMA stream down inside a real kernel?  N = 96 MMAs by thread 0 while warps 1..3
  //      (a) idle, (b) stream tcgen05.ld from another TMEM column range, (c) hammer shared memory with 16-byte stores,
  //      (d) (thread 32) keeps TMA loads of the A tile in flight into a scratch buffer
  __shared__ volatile int s_stop;
  for (int v = 0; v < 4; ++v) {
    if (threadIdx.x == 0) s_stop = 0;
    __syncthreads();
    long long t0 = 0;
    if (threadIdx.x == 0) {
      const uint32_t a0 = smem_u32(a_smem), b0 = smem_u32(b_smem);
      const uint32_t idesc = make_idesc_bf16(128, 96, 0, 0);
      t0 = clock64();
      for (int i = 0; i < 256; ++i)
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base2 + (i & 1) * 96, make_smem_desc(a0 + k * 32, 16, 1024, 2), make_smem_desc(b0 + k * 32, 16, 1024, 2), idesc, 1);
      umma_commit(mbar);
      mbar_wait(mbar, mph);
      p.cycles[16 + v] = clock64() - t0;
      s_stop = 1;
    } else if (warp >= 1) {
      uint32_t sink = 0;
      if (v == 1) {
        while (!s_stop) {
          uint32_t vr[16];
          tmem_ld16(tmem_base2 + ((uint32_t)(warp * 32) << 16) + 256 + (sink & 63), vr);
          tmem_ld_wait();
          sink += vr[0] & 1;
        }
      } else if (v == 2) {
        uint4* scratch = reinterpret_cast<uint4*>(b_smem + 16 * 1024);  // 16 KB further down: unused shared memory
        while (!s_stop) {
#pragma unroll
          for (int u = 0; u < 8; ++u) scratch[(threadIdx.x - 32) + 96 * u] = make_uint4(sink, u, 0, 0);
          sink++;
        }
      } else if (v == 3 && threadIdx.x == 32) {
        uint64_t* bar2 = bar + 8;
        uint32_t ph2 = 0;
        mbar_init(bar2, 1);
        fence_mbar_init();
        while (!s_stop) {
          mbar_expect_tx(bar2, 144 * 128);
          tma_load_2d(&p.tmA, bar2, b_smem + 16 * 1024, 0, 0);
          mbar_wait(bar2, ph2);
          ph2 ^= 1;
        }
      }
      if (sink == 0xffffffffu) p.cycles[31] = sink;
    }
    if (threadIdx.x == 0) mph ^= 1;
    __syncthreads();
    mph = __shfl_sync(0xffffffffu, mph, 0);
  }
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static float aval(int r, int c) { return (float)(((r * 7 + c * 3) % 23) - 11); }

int main() {
  const int R = 288, C = 64;
  std::vector<__nv_bfloat16> hA(R * C), hB(64 * 64);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) hA[r * C + c] = __float2bfloat16(aval(r, c));
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  const int NV = 80;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dO, (size_t)NV * 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, (size_t)NV * 128 * 64 * 4);
  Params p;
  uint64_t dimsA[2] = {64, (uint64_t)R}, strA[1] = {128};
  uint32_t boxA[2] = {64, 144};
  uint64_t dimsB[2] = {64, 64};
  uint32_t boxB[2] = {64, 64};
  if (encode_tmap(&p.tmA, dA, 2, dimsA, strA, boxA, 128, 2) || encode_tmap(&p.tmB, dB, 2, dimsB, strA, boxB, 128, 2)) {
    printf("encode failed\n");
    return 1;
  }
  p.out = dO;
  long long* dC;
  cudaMalloc(&dC, 32 * sizeof(long long));
  p.cycles = dC;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe_kernel<<<1, 128, 100 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hO((size_t)NV * 128 * 64);
  cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
  const char* names[4] = {"K-major SBO=1024 (dense rows)", "K-major SBO=2048 (pitch-16 patch)", "MN-major (K=16, shift along K)",
                          "K-major SBO=1280 (pitch-10 patch)"};
  for (int test = 0; test < 4; ++test)
    for (int mode = 0; mode < 2; ++mode) {
      printf("%-36s base_offset=%s :", names[test], mode ? "(addr>>7)&7" : "0");
      for (int j = 0; j < 10; ++j) {
        const int v = (test * 10 + j) * 2 + mode;
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float want;
            if (test == 0) want = aval(m + j, n);
            else if (test == 1) want = aval((m / 8) * 16 + (m % 8) + j, n);
            else if (test == 3) want = aval((m / 8) * 10 + (m % 8) + j, n);
            else want = n < 16 ? aval(j + n + (m >= 64 ? 144 : 0), m % 64) : 0.f;
            if (hO[((size_t)v * 128 + m) * 64 + n] != want) ++bad;
          }
        printf(" j=%d:%s", j, bad ? "FAIL" : "ok");
        if (bad) printf("(%d)", bad);
      }
      printf("\n");
    }
  long long hC[32];
  cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost);
  const int sbos[4] = {1024, 1280, 2048, 2304};
  for (int v = 0; v < 8; ++v)
    printf("timing: start row %d, SBO %4d : %.1f cycles per MMA (M=128 N=64 K=16; ideal 32)\n", v & 1, sbos[v >> 1], hC[v] / 512.0);
  const char* an[8] = {"N=96 one accumulator", "N=96 two accumulators 96 columns apart", "N=96 two accumulators 128 apart",
                       "N=96 two accumulators 64 apart (overlapping, timing only)", "N=48 one accumulator",
                       "N=48 four accumulators 48 apart", "N=48 four accumulators 128 apart", "N=48 four accumulators 64 apart"};
  const char* cn[4] = {"alone", "+ 3 warps streaming tcgen05.ld", "+ 3 warps storing to shared memory", "+ TMA loads kept in flight"};
  for (int v = 0; v < 4; ++v) printf("timing3: N=96 MMA stream %-40s : %.1f cycles per MMA\n", cn[v], hC[16 + v] / 1024.0);
  for (int v = 0; v < 8; ++v) printf("timing2: %-58s : %.1f cycles per MMA\n", an[v], hC[8 + v] / 512.0);
  return 0;
}
