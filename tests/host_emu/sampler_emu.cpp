// HOST EMULATION OF THE VECTORISED SAMPLER-STEP KERNEL BODIES -- TEST INFRASTRUCTURE ONLY.
//
// Compiles fastgen_b200/csrc/sampler_vec.cuh (the very source the __global__ wrappers in sampler.cu call) with g++ and
// runs the thread body for every global thread id of a (grid, block) launch, sequentially.  tests/test_sampler_vec_host_cpu.py
// builds this file into tests/host_emu/build/ and compares the outputs bit for bit with the reference formulas evaluated by
// torch on the CPU.  Every vector access is checked: naturally aligned, inside one of the registered buffers, and no
// output element written twice.  Nothing here is linked into libfastgen_b200.so; the product has no CPU path.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

struct Region {
  const char* lo;
  const char* hi;
  int writable;
  std::vector<uint8_t> written;
};
static std::vector<Region> g_regions;
static int g_errors = 0;
static char g_msg[256] = "";

static void emu_access(const void* ptr, size_t bytes, int g_writing) {
  const char* p = static_cast<const char*>(ptr);
  if (reinterpret_cast<uintptr_t>(p) % bytes != 0) {
    if (!g_errors++) snprintf(g_msg, sizeof(g_msg), "misaligned %zu-byte access at %p", bytes, ptr);
    return;
  }
  for (auto& r : g_regions) {
    if (p >= r.lo && p + bytes <= r.hi) {
      if (g_writing) {
        if (!r.writable) {
          if (!g_errors++) snprintf(g_msg, sizeof(g_msg), "store into a read-only buffer at %p", ptr);
          return;
        }
        for (size_t i = 0; i < bytes; ++i) {
          uint8_t& w = r.written[(p - r.lo) + i];
          if (w && !g_errors++) snprintf(g_msg, sizeof(g_msg), "byte %zu of an output written twice", (size_t)(p - r.lo) + i);
          w = 1;
        }
      }
      return;
    }
  }
  if (!g_errors++) snprintf(g_msg, sizeof(g_msg), "%zu-byte access at %p outside every registered buffer", bytes, ptr);
}

#define FG_EMU_ACCESS(ptr, bytes, is_store) emu_access(ptr, bytes, is_store)
#include "../../fastgen_b200/csrc/sampler_vec.cuh"

extern "C" {

void emu_reset() {
  g_regions.clear();
  g_errors = 0;
  g_msg[0] = 0;
}
void emu_region(const void* base, int64_t bytes, int writable) {
  if (!base) return;
  Region r;
  r.lo = static_cast<const char*>(base), r.hi = r.lo + bytes, r.writable = writable;
  if (writable) r.written.assign(bytes, 0);
  g_regions.push_back(r);
}
int emu_errors() { return g_errors; }
const char* emu_message() { return g_msg; }
// fraction of the bytes of writable region `idx` (in registration order over writable regions) that were written
double emu_coverage(int idx) {
  int w = -1;
  for (auto& r : g_regions) {
    if (!r.writable) continue;
    if (++w == idx) {
      size_t n = 0;
      for (uint8_t b : r.written) n += b;
      return r.written.empty() ? 0.0 : static_cast<double>(n) / r.written.size();
    }
  }
  return -1.0;
}

// dtype: 0 fp32, 1 bf16 (fg_dtype); threads = grid * block of the emulated launch
void emu_edm_step(const fg::vec::EdmStepV* p, int dtype, int64_t threads) {
  for (int64_t t = 0; t < threads; ++t) {
    if (dtype == 0)
      fg::vec::edm_step_vec_body<float>(*p, t, threads);
    else
      fg::vec::edm_step_vec_body<__nv_bfloat16>(*p, t, threads);
  }
}

void emu_coef(int mode, int dtype, const void* a, const void* b, int64_t total, int64_t outer, int64_t mid, int64_t inner,
              const double* c0, const double* c1, void* out, int64_t threads) {
  for (int64_t t = 0; t < threads; ++t) {
#define RUN(T, M)                                                                                                  \
  fg::vec::coef_vec_body<T, M>(static_cast<const T*>(a), static_cast<const T*>(b), total, outer, mid, inner, c0, c1, \
                               static_cast<T*>(out), t, threads)
    if (dtype == 0) {
      if (mode == 0) RUN(float, 0); else RUN(float, 1);
    } else if (dtype == 1) {
      if (mode == 0) RUN(__nv_bfloat16, 0); else RUN(__nv_bfloat16, 1);
    } else {
      if (mode == 0) RUN(double, 0); else RUN(double, 1);
    }
#undef RUN
  }
}
}
