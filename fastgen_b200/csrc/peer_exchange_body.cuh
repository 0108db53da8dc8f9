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

template <typename LoadFn>
FG_PEER_HD void peer_put_copy(const PeerPutK& k, int dst, int64_t tid, int64_t nthreads, LoadFn load) {
  const Vec16* src = k.src + dst * k.src_peer_stride;
  Vec16* out = k.dst[dst];
  const int64_t total = static_cast<int64_t>(k.n[0]) * k.n[1] * k.n[2] * k.n[3] * k.vec;
  for (int64_t i = tid; i < total; i += nthreads) {
    const int v = static_cast<int>(i % k.vec);
    int64_t item = i / k.vec;
    const int i3 = static_cast<int>(item % k.n[3]);
    item /= k.n[3];
    const int i2 = static_cast<int>(item % k.n[2]);
    item /= k.n[2];
    const int i1 = static_cast<int>(item % k.n[1]);
    const int i0 = static_cast<int>(item / k.n[1]);
    out[i0 * k.ds[0] + i1 * k.ds[1] + i2 * k.ds[2] + i3 * k.ds[3] + v] =
        load(src + i0 * k.ss[0] + i1 * k.ss[1] + i2 * k.ss[2] + i3 * k.ss[3] + v);
  }
}

}  // namespace fg
