// Ulysses all-to-all as ONE pass over the data: the pack permutation and the exchange are the same kernel, which stores
// this rank's slices straight into the peers' buffers over NVLink (peer-mapped / symmetric memory), then raises a flag on
// every peer; a one-CTA wait kernel, stream-ordered before the consumer, spins on the local flags.  No NCCL call, no
// host round trip, no staging buffers (the NCCL path needs pack -> send buffer -> all_to_all -> recv buffer -> unpack),
// and nothing in it changes from call to call, so a launch list containing it can be replayed or graph-captured.
//
// Replaces the sequence-parallel exchange around the self-attention of the Wan blocks (the reference reaches it through
// diffusers' context-parallel attention dispatch / torch all_to_all_single; call sites fastgen/networks/Wan/network.py:
// 106-153, network_causal.py:199-464).
//
// Protocol (per channel = one direction of the exchange; every rank runs the same sequence of put / wait pairs):
//   put : CTA (x, dst) copies 16-byte vectors of the 4-D item set of destination `dst` to dst_base[dst] + dst_offset;
//         every thread fences at system scope, the CTA counts itself done; the last CTA of the launch advances the local
//         put epoch e and writes e with release.sys to flags[dst][rank] of every peer (its own included).
//   wait: advances the local wait epoch e' (equal to the peers' put epoch, all ranks being in lockstep) and spins with
//         ld.acquire.sys until every local flag >= e'.  A timeout (no peer progress) sets state[3] and returns instead of
//         hanging the GPU; the host reads state[3] after synchronising and raises.
// Epochs only grow, so flags are never reset and a fast peer that is one put ahead is harmless (>= comparison).
// Write-after-read safety of the destination buffers comes from the two channels alternating: a peer can only issue
// its put of exchange k+1 after its wait of exchange k on the other channel, i.e. after this rank finished the kernel
// that consumed exchange k-1's data.
#include "common.cuh"
#include "host_util.h"
#include "peer_exchange_body.cuh"

namespace fg {

constexpr int kPeerThreads = 256;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// spin (one thread per peer) until every local flag has reached the next wait epoch; a timeout sets state[3]
__device__ __forceinline__ void peer_wait_body(const unsigned long long* flags, int peers, unsigned long long* state,
                                               long long timeout_ns, unsigned long long* epoch_smem) {
  if (threadIdx.x == 0) {
    *epoch_smem = state[2] + 1ull;
    state[2] = *epoch_smem;
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < peers) {
    const unsigned long long e = *epoch_smem;
    const uint64_t t0 = globaltimer_ns();
    while (ld_acquire_sys(flags + threadIdx.x) < e) {
      __nanosleep(32);
      if (timeout_ns > 0 && static_cast<long long>(globaltimer_ns() - t0) > timeout_ns) {
        state[3] = 1ull;  // the host raises on it (next forward / end of the sample); the consumer reads stale data
        break;
      }
    }
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(kPeerThreads) peer_put_kernel(const PeerPutK k, const unsigned long long* wait_flags,
                                                                long long wait_timeout_ns) {
  __shared__ unsigned long long epoch;
  __shared__ int is_last;
  peer_put_copy(k, blockIdx.y, blockIdx.x * static_cast<int64_t>(kPeerThreads) + threadIdx.x,
                static_cast<int64_t>(gridDim.x) * kPeerThreads, [](const Vec16* p) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
                  return Vec16{u.x, u.y, u.z, u.w};
                });
  __threadfence_system();  // this thread's peer stores are ordered before the counter / flag writes below
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(k.state + 1), 1u);
    is_last = done == k.total_ctas - 1;
    if (is_last) {  // every other CTA fenced before it counted itself
      __threadfence_system();
      *reinterpret_cast<unsigned*>(k.state + 1) = 0u;
      const unsigned long long e = k.state[0] + 1ull;
      k.state[0] = e;
      for (int p = 0; p < k.peers; ++p) st_release_sys(k.flags[p] + k.rank, e);
    }
  }
  if (wait_flags == nullptr) return;
  // fused wait: the CTA that published this rank's flags stays and waits for everybody else's, so the kernel's end means
  // "the exchange is complete here" and no separate one-CTA wait launch sits between the put and the consumer
  __syncthreads();
  if (!is_last) return;
  peer_wait_body(wait_flags, k.peers, k.state, wait_timeout_ns, &epoch);
}

__global__ void peer_wait_kernel(const unsigned long long* __restrict__ flags, int peers, unsigned long long* state,
                                 long long timeout_ns) {
  __shared__ unsigned long long epoch;
  peer_wait_body(flags, peers, state, timeout_ns, &epoch);
}

}  // namespace fg

extern "C" int fg_peer_put(const fg_peer_put_params* p, fg_stream_t stream_) {
  using namespace fg;
  set_error("");
  FG_CHECK_ARG(p && p->src && p->state, "peer_put: null pointer");
  FG_CHECK_ARG(p->peers >= 1 && p->peers <= FG_MAX_PEERS && p->rank >= 0 && p->rank < p->peers,
               "peer_put: peers %d (max %d), rank %d", p->peers, FG_MAX_PEERS, p->rank);
  FG_CHECK_ARG(p->item_bytes > 0 && p->item_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(p->src) & 15) == 0 &&
                   p->src_peer_stride % 16 == 0 && p->dst_offset % 16 == 0 && p->dst_offset >= 0,
               "peer_put: items, offsets and the source must be 16-byte multiples / aligned");
  PeerPutK k;
  k.src = static_cast<const Vec16*>(p->src);
  k.state = reinterpret_cast<unsigned long long*>(p->state);
  k.vec = p->item_bytes / 16, k.peers = p->peers, k.rank = p->rank;
  k.src_peer_stride = p->src_peer_stride / 16;
  int64_t items = 1;
  for (int i = 0; i < 4; ++i) {
    FG_CHECK_ARG(p->n[i] > 0 && p->src_strides[i] % 16 == 0 && p->dst_strides[i] % 16 == 0,
                 "peer_put: extents must be positive, strides (bytes) multiples of 16");
    k.n[i] = p->n[i], k.ss[i] = p->src_strides[i] / 16, k.ds[i] = p->dst_strides[i] / 16;
    items *= p->n[i];
  }
  for (int q = 0; q < FG_MAX_PEERS; ++q) {
    k.dst[q] = nullptr, k.flags[q] = nullptr;
    if (q >= p->peers) continue;
    FG_CHECK_ARG(p->dst[q] && p->flags[q] && (reinterpret_cast<uintptr_t>(p->dst[q]) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(p->flags[q]) & 7) == 0,
                 "peer_put: destination / flag pointer of peer %d missing or misaligned", q);
    k.dst[q] = reinterpret_cast<Vec16*>(static_cast<char*>(p->dst[q]) + p->dst_offset);
    k.flags[q] = reinterpret_cast<unsigned long long*>(p->flags[q]);
  }
  FG_CHECK_ARG(items > 0 && items < (1ll << 31), "peer_put: too many items");
  if (p->wait_flags)
    FG_CHECK_ARG((reinterpret_cast<uintptr_t>(p->wait_flags) & 7) == 0, "peer_put: wait_flags must be 8-byte aligned");
  // one warp per item, two items in flight per warp: enough CTAs per destination that all destinations together fill the
  // machine (8 warps per CTA), at least two items per warp
  int64_t bx = (items + 2 * (kPeerThreads / 32) - 1) / (2 * (kPeerThreads / 32));
  const int64_t cap = (148ll * 4 + p->peers - 1) / p->peers;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  k.total_ctas = static_cast<unsigned>(bx * p->peers);
  peer_put_kernel<<<dim3(static_cast<unsigned>(bx), p->peers), kPeerThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      k, reinterpret_cast<const unsigned long long*>(p->wait_flags), p->wait_timeout_ns);
  FG_CHECK_LAUNCH("peer_put");
  return FG_OK;
}

extern "C" int fg_peer_wait(const uint64_t* flags, int32_t peers, uint64_t* state, int64_t timeout_ns,
                            fg_stream_t stream_) {
  using namespace fg;
  set_error("");
  FG_CHECK_ARG(flags && state && peers >= 1 && peers <= FG_MAX_PEERS, "peer_wait: bad arguments (peers %d)", peers);
  FG_CHECK_ARG((reinterpret_cast<uintptr_t>(flags) & 7) == 0 && (reinterpret_cast<uintptr_t>(state) & 7) == 0,
               "peer_wait: flags / state must be 8-byte aligned");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const unsigned long long*>(flags), peers, reinterpret_cast<unsigned long long*>(state), timeout_ns);
  FG_CHECK_LAUNCH("peer_wait");
  return FG_OK;
}
