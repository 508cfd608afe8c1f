// Device side of the peer-memory exchange protocol (see p2p.cu for the design): the region header, the system-scope flag
// operations and the three steps a kernel wraps around its own staging / combining code:
//     P2pCall k = p2p_begin(regions, rank, slot_bytes, timeout_ns);   // call number, parity, data-slot offset
//     ... write this CTA's values into regions[rank] + k.slot_off ...
//     p2p_signal_wait(regions, rank, world, blockIdx.x, k);   // flag to every peer, wait for every peer's flag (bounded)
//     ... read regions[p] + k.slot_off for p = 0 .. world-1 (ld_peer*), combine in rank order ...
//     if (p2p_close_call(k, gridDim.x)) { ... exactly one thread of the grid, after every CTA is done ... }
// Used by the all-reduce kernel (p2p.cu) and by kernels that fuse their own reduction tail with the exchange (norm_loss.cu).
#pragma once
#include <stdlib.h>
#include "common.cuh"

namespace ercg {

constexpr int P2P_MAX_CTAS = 128;
constexpr int P2P_MAX_WORLD = 16;
constexpr int P2P_THREADS = 128;
// Bound of every wait for a peer: 30 s (ranks may be seconds apart at start-up: lazy library initialisation); the environment
// variable ERCG_P2P_TIMEOUT_MS (read once per process) overrides it -- the test of the failure path uses 300 ms.
inline unsigned long long p2p_timeout_ns() {
  static const unsigned long long v = [] {
    const char* e = getenv("ERCG_P2P_TIMEOUT_MS");
    long long ms = e ? atoll(e) : 30000;
    if (ms < 1) ms = 1;
    return (unsigned long long)ms * 1000000ull;
  }();
  return v;
}

struct P2pHeader {
  unsigned long long call;                                   // calls completed by this rank
  unsigned int done;                                         // CTAs of the running call that have finished
  int status;                                                // 0, or ERCG_P2P_ETIMEOUT (sticky)
  unsigned int pad[60];                                     // header = 256 bytes + the flag words
  unsigned int flag[2][P2P_MAX_WORLD][P2P_MAX_CTAS];         // [parity][writer rank][CTA] = call number (low 32 bits)
};
static_assert(sizeof(P2pHeader) % 256 == 0, "data slots start 256-byte aligned");

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer16(const void* p) {   // peer memory: never through L1
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
template <typename T>
__device__ __forceinline__ T ld_peer(const T* p) {
  return *reinterpret_cast<const volatile T*>(p);
}
__device__ __forceinline__ void add16(float4& a, const float4& b, float) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void add16(float4& a, const float4& b, double) {
  double2& x = reinterpret_cast<double2&>(a);
  const double2& y = reinterpret_cast<const double2&>(b);
  x.x += y.x; x.y += y.y;
}


struct P2pCall {
  P2pHeader* hdr;
  unsigned long long call;
  size_t slot_off;      // byte offset of this call's data slot inside every region
  unsigned long long timeout_ns;
  unsigned int tag;
  int par;
};
// every thread of every CTA of the grid; the value is the same on every rank (each rank has completed the same calls)
__device__ __forceinline__ P2pCall p2p_begin(unsigned char* const* __restrict__ regions, int rank, size_t slot_bytes,
                                             unsigned long long timeout_ns) {
  P2pCall k;
  k.timeout_ns = timeout_ns;
  k.hdr = reinterpret_cast<P2pHeader*>(regions[rank]);
  k.call = *reinterpret_cast<volatile unsigned long long*>(&k.hdr->call) + 1;
  k.par = (int)(k.call & 1);
  k.tag = (unsigned int)k.call;
  k.slot_off = sizeof(P2pHeader) + (size_t)k.par * slot_bytes;
  return k;
}
// CTA c has staged its values: tell every peer (and ourselves), then wait until every rank's CTA c has done the same.
// Whole CTA (two barriers); blockDim.x >= world.
__device__ __forceinline__ void p2p_signal_wait(unsigned char* const* __restrict__ regions, int rank, int world, int c, const P2pCall& k) {
  __syncthreads();                                            // the release stores below are cumulative over the CTA's writes
  const int tid = threadIdx.x;
  if (tid < world) {
    P2pHeader* peer = reinterpret_cast<P2pHeader*>(regions[tid]);
    st_release_sys(&peer->flag[k.par][rank][c], k.tag);
    const unsigned int* mine = &k.hdr->flag[k.par][tid][c];
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(mine) != k.tag) {
      if (++spins < (1u << 14)) continue;                     // the usual case: the peers are a few microseconds apart
      __nanosleep(500);
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > k.timeout_ns) {                          // the peer is not coming
        k.hdr->status = ERCG_P2P_ETIMEOUT;
        break;
      }
    }
  }
  __syncthreads();
}
// Whole CTA, after its last read of peer data.  True in exactly one thread of the grid (thread 0 of the last CTA to get
// here), which has then closed the call.
__device__ __forceinline__ bool p2p_close_call(const P2pCall& k, int G) {
  __syncthreads();
  if (threadIdx.x != 0) return false;
  __threadfence();
  if (atomicAdd(&k.hdr->done, 1u) != (unsigned int)G - 1) return false;
  k.hdr->done = 0;
  __threadfence();
  *reinterpret_cast<volatile unsigned long long*>(&k.hdr->call) = k.call;
  return true;
}

}  // namespace ercg
