// Copy part of fg_peer_put as a plain function of (params, destination, thread id, thread count): the same source runs in
// peer_put_kernel and, compiled with g++, in tests/host_emu/peer_emu.cpp, which checks the index arithmetic against the
// descriptor reference of fastgen_b200/parallel.py on the CPU (tests/test_parallel_cpu.py).  The flag protocol around it
// (system-scope fences, epochs) only exists on the device.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define FG_PEER_HD __host__ __device__ __forceinline__
#else
#define FG_PEER_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define FG_PEER_UNROLL _Pragma("unroll")
#else
#define FG_PEER_UNROLL
#endif

#ifndef FG_MAX_PEERS
#define FG_MAX_PEERS 8
#endif

namespace fg {

struct alignas(16) Vec16 {  // one 16-byte vector (uint4 on the device)
  uint32_t x, y, z, w;
};

struct PeerPutK {
  const Vec16* src;
  Vec16* dst[FG_MAX_PEERS];
  unsigned long long* flags[FG_MAX_PEERS];
  unsigned long long* state;  // [0] put epoch, [1] CTA counter, [2] wait epoch, [3] timeout flag
  int n[4];
  int64_t ss[4], ds[4];       // strides in 16-byte units
  int64_t src_peer_stride;    // 16-byte units
  int vec;                    // 16-byte vectors per item
  int peers, rank;
  unsigned total_ctas;
};

// One warp per item (an item = `vec` consecutive 16-byte vectors, e.g. one head slice of one token row): the 4-D index is
// decomposed once per item with 32-bit arithmetic (the first version did four 64-bit divisions per 16 bytes), the lanes
// stride over the item's vectors, and a warp keeps TWO items in flight (all loads issued before the first store) so that
// the NVLink stores of one item overlap the local loads of the next.  `nthreads` must be a multiple of 32.
template <typename LoadFn>
FG_PEER_HD void peer_put_copy(const PeerPutK& k, int dst, int64_t tid, int64_t nthreads, LoadFn load) {
  const Vec16* src = k.src + dst * k.src_peer_stride;
  Vec16* out = k.dst[dst];
  const uint32_t items = static_cast<uint32_t>(k.n[0]) * k.n[1] * k.n[2] * k.n[3];
  const uint32_t warps = static_cast<uint32_t>(nthreads >> 5), warp = static_cast<uint32_t>(tid >> 5);
  const int lane = static_cast<int>(tid & 31);
  const uint32_t n3 = k.n[3], n2 = k.n[2], n1 = k.n[1];
  auto offsets = [&](uint32_t item, int64_t& so, int64_t& dof) {
    const uint32_t i3 = item % n3;
    item /= n3;
    const uint32_t i2 = item % n2;
    item /= n2;
    const uint32_t i1 = item % n1, i0 = item / n1;
    so = i0 * k.ss[0] + i1 * k.ss[1] + i2 * k.ss[2] + i3 * k.ss[3];
    dof = i0 * k.ds[0] + i1 * k.ds[1] + i2 * k.ds[2] + i3 * k.ds[3];
  };
  constexpr int kMaxPerLane = 4;   // vectors of one item a lane holds at once (items of up to 2 KB in one go)
  for (uint32_t a = warp; a < items; a += 2 * warps) {
    const uint32_t b = a + warps;
    const bool two = b < items;
    int64_t sa, da, sb = 0, db = 0;
    offsets(a, sa, da);
    if (two) offsets(b, sb, db);
    for (int v0 = 0; v0 < k.vec; v0 += 32 * kMaxPerLane) {
      Vec16 ra[kMaxPerLane], rb[kMaxPerLane];
FG_PEER_UNROLL
      for (int u = 0; u < kMaxPerLane; ++u) {
        const int v = v0 + u * 32 + lane;
        if (v < k.vec) {
          ra[u] = load(src + sa + v);
          if (two) rb[u] = load(src + sb + v);
        }
      }
FG_PEER_UNROLL
      for (int u = 0; u < kMaxPerLane; ++u) {
        const int v = v0 + u * 32 + lane;
        if (v < k.vec) {
          out[da + v] = ra[u];
          if (two) out[db + v] = rb[u];
        }
      }
    }
  }
}

}  // namespace fg
