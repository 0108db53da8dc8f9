// HOST EMULATION OF THE COPY PART OF fg_peer_put -- TEST INFRASTRUCTURE ONLY (see sampler_emu.cpp).
// Runs fastgen_b200/csrc/peer_exchange_body.cuh (the source peer_put_kernel calls) for every (destination, thread id) of an
// emulated launch; the "peers" are host buffers.  tests/test_parallel_cpu.py compares the result with the descriptor
// reference (fastgen_b200.parallel.apply_peer_put_cpu), i.e. with pack -> all_to_all -> unpack.
#include "../../fastgen_b200/csrc/peer_exchange_body.cuh"

extern "C" int emu_peer_put(const void* src, void* const* dst, int peers, const int* n, const int64_t* src_strides,
                            const int64_t* dst_strides, int64_t src_peer_stride, int64_t dst_offset, int item_bytes,
                            int64_t threads_per_dst) {
  // same conversions as the fg_peer_put launcher (bytes -> 16-byte units)
  if (item_bytes % 16 || src_peer_stride % 16 || dst_offset % 16) return -1;
  if (threads_per_dst <= 0 || threads_per_dst % 32) return -1;   // the copy is warp-granular (one warp per item)
  fg::PeerPutK k;
  k.src = static_cast<const fg::Vec16*>(src);
  k.vec = item_bytes / 16, k.peers = peers, k.rank = 0, k.total_ctas = 0, k.state = nullptr;
  k.src_peer_stride = src_peer_stride / 16;
  for (int i = 0; i < 4; ++i) {
    if (src_strides[i] % 16 || dst_strides[i] % 16) return -1;
    k.n[i] = n[i], k.ss[i] = src_strides[i] / 16, k.ds[i] = dst_strides[i] / 16;
  }
  for (int q = 0; q < peers; ++q) k.dst[q] = reinterpret_cast<fg::Vec16*>(static_cast<char*>(dst[q]) + dst_offset);
  for (int q = 0; q < peers; ++q)
    for (int64_t t = 0; t < threads_per_dst; ++t)
      fg::peer_put_copy(k, q, t, threads_per_dst, [](const fg::Vec16* p) { return *p; });
  return 0;
}
