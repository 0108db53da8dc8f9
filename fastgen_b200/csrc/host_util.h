// Host-side helpers shared by the launchers: error string, launch counter, TMA descriptor encoding.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/fastgen_b200.h"

namespace fg {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency, so
// the library also loads on a machine without a driver for the symbol check).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1; bf16 elements, 128B swizzle.
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box);
// same for an FG_F32 / FG_BF16 tensor (fg_dtype)
int encode_tmap(CUtensorMap* map, const void* base, int fg_dtype_, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box);

// One-time per-DEVICE set-up of a kernel (cudaFuncSetAttribute is per function and device): a process may drive several
// GPUs (tests on cuda:1, single-process multi-device use) and several host threads.  Setting an attribute twice is
// harmless, so a relaxed bit set is enough.
struct PerDeviceFlag {
  std::atomic<uint64_t> bits{0};
  bool test(int dev) const { return dev >= 0 && dev < 64 && ((bits.load(std::memory_order_acquire) >> dev) & 1ull); }
  void set(int dev) {
    if (dev >= 0 && dev < 64) bits.fetch_or(1ull << dev, std::memory_order_release);
  }
};
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
// multiprocessor count of a device, queried once per device
int device_sm_count(int dev);

#define FG_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      fg::set_error(__VA_ARGS__);    \
      return FG_ERR_INVALID;         \
    }                                \
  } while (0)

#define FG_CHECK_LAUNCH(name)                                                              \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      fg::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));              \
      return FG_ERR_CUDA;                                                                  \
    }                                                                                      \
    fg::count_launch();                                                                    \
  } while (0)

}  // namespace fg
