// Experiment (round 2): can ONE TMA-loaded halo tile serve the 9 taps of a 3x3 convolution as tcgen05.mma A operands?
// Tile = 16 rows x 8 pixels (128 output pixels), halo = 18 x 10 pixel rows of 64 channels (128 bytes each, 128B-swizzled
// by TMA).  For tap (dy, dx) the A descriptor starts at halo row (1+dy)*10 + (1+dx) and steps 10 rows (1280 B = SBO) per
// 8-pixel tile row: start addresses are not 1024-byte aligned and SBO is not a multiple of 1024.  The question is whether
// the hardware's 128B swizzle (a function of absolute shared-memory address bits) reads back what TMA wrote, and what the
// descriptor's base-offset field has to hold.  Prints the max error against a host reference for several settings.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../fastgen_b200/csrc halo_umma_test.cu -o halo_test -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"

using namespace fg;

constexpr int IMG = 32, CH = 64, NOUT = 64;
constexpr int HALO_ROWS = 18 * 10;
constexpr int A_BYTES = ((HALO_ROWS * 128 + 1023) / 1024) * 1024;
constexpr int B_BYTES = NOUT * 128;  // one tap: 64 rows x 128 B

__device__ __forceinline__ uint64_t desc_sw128_ex(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                                  // LBO (unused for K-major SW128)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7u) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmW, int h0, int w0, int mode,
                                                      float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;             // 9 taps x B_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 9 * B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], HALO_ROWS * 128 + 9 * B_BYTES);
    tma_load_4d(sA, &tmA, &bars[0], 0, w0 - 1, h0 - 1, 0);                  // one halo box, zero-filled outside the image
    for (int tap = 0; tap < 9; ++tap) tma_load_3d(sB + tap * B_BYTES, &tmW, &bars[0], 0, tap, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(128, NOUT);
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const uint32_t row0 = (1 + dy) * 10 + (1 + dx);
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_addr = smem_u32(sA) + row0 * 128 + k * 32;
        uint32_t boff = 0;
        if (mode == 1) boff = (a_addr >> 7) & 7u;
        if (mode == 2) boff = row0 & 7u;
        const uint64_t da = desc_sw128_ex(a_addr, 1280, boff);
        const uint64_t db = umma_desc_sw128(smem_u32(sB + tap * B_BYTES) + k * 32);
        umma_bf16_ss(tmem, da, db, idesc, (tap | k) != 0 ? 1u : 0u);
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < NOUT; c += 32) {
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * NOUT + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  std::vector<__nv_bfloat16> hA(IMG * IMG * CH), hW(NOUT * 9 * CH);
  srand(1);
  for (auto& x : hA) x = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& x : hW) x = __float2bfloat16((rand() % 13 - 6) / 16.0f);
  __nv_bfloat16 *dA, *dW;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2), cudaMalloc(&dW, hW.size() * 2), cudaMalloc(&dO, 128 * NOUT * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = reinterpret_cast<EncodeFn>(fp);
  CUtensorMap tmA, tmW;
  {
    cuuint64_t dims[4] = {CH, IMG, IMG, 1}, str[3] = {CH * 2, CH * 2 * IMG, CH * 2 * IMG * IMG};
    cuuint32_t box[4] = {64, 10, 18, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) return printf("encode A failed %d\n", (int)r), 1;
  }
  {
    cuuint64_t dims[3] = {CH, 9, NOUT}, str[2] = {CH * 2, CH * 2 * 9};
    cuuint32_t box[3] = {64, 1, NOUT}, es[3] = {1, 1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dW, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) return printf("encode W failed %d\n", (int)r), 1;
  }
  const int smem = A_BYTES + 9 * B_BYTES + 2048;
  cudaFuncSetAttribute(halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hO(128 * NOUT);
  const int tiles[3][2] = {{8, 8}, {0, 0}, {16, 24}};   // interior, top-left corner (zero padding), bottom-right edge
  for (int mode = 0; mode < 3; ++mode)
    for (int t = 0; t < 3; ++t) {
      const int h0 = tiles[t][0], w0 = tiles[t][1];
      cudaMemset(dO, 0, 128 * NOUT * 4);
      halo_kernel<<<1, 128, smem>>>(tmA, tmW, h0, w0, mode, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) return printf("mode %d tile %d: CUDA error %s\n", mode, t, cudaGetErrorString(e)), 1;
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int p = 0; p < 128; ++p) {
        const int y = h0 + p / 8, x = w0 + p % 8;
        for (int n = 0; n < NOUT; ++n) {
          double acc = 0;
          for (int tap = 0; tap < 9; ++tap) {
            const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
            if (yy < 0 || yy >= IMG || xx < 0 || xx >= IMG) continue;
            for (int c = 0; c < CH; ++c)
              acc += (double)__bfloat162float(hA[(yy * IMG + xx) * CH + c]) * (double)__bfloat162float(hW[(n * 9 + tap) * CH + c]);
          }
          maxerr = fmax(maxerr, fabs(acc - hO[p * NOUT + n]));
          maxref = fmax(maxref, fabs(acc));
        }
      }
      printf("base_offset mode %d (0: zero, 1: (addr>>7)&7, 2: row0&7), tile (%2d,%2d): max |err| %.4g (max |ref| %.3g)\n", mode, h0,
             w0, maxerr, maxref);
    }
  return 0;
}
