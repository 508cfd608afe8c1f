// Shared helpers for the libercgraph kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include "../../include/ercgraph.h"

namespace ercg {

extern std::atomic<unsigned long long> g_launches;   // counted on the host side, one per kernel launch (any thread)

inline int finish_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaPeekAtLastError() == cudaSuccess ? ERCG_OK : ERCG_ECUDA;
}

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming variants: data touched once, keep it out of L1
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st4_stream(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void fma4(float4& a, float s, const float4& b) {
  a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y); a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
}
// Packed variants.  sm_100a has a packed fp32 FMA (fma.rn.f32x2 -> SASS FFMA2, one operand may be a broadcast scalar): two
// IEEE FMAs per issue slot, bit-identical to two fmaf.  It does not raise the FMA pipe's peak, it halves the issue slots of
// a weighted-sum / dot-product loop.  Measured on the graph kernels (tools/bench_graph.py, 2^20 nodes): window
// attention forward 0.558 -> 0.539 ms, by-source backward and both gathers unchanged (they wait on loads, not on issue
// slots), the 256-thread by-destination backward SLOWER (0.531 -> 0.596 ms) -- so only the forward kernel uses these.
__device__ __forceinline__ void fma2_p(float& a0, float& a1, float s, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rs;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\t"
      "fma.rn.f32x2 ra, rs, rb, ra;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1), "f"(s));
}
__device__ __forceinline__ void fma4_p(float4& a, float s, const float4& b) {
  fma2_p(a.x, a.y, s, b.x, b.y);
  fma2_p(a.z, a.w, s, b.z, b.w);
}
// acc (two partial sums: even / odd elements) += a . b as two packed FMAs; the caller adds acc.x + acc.y at the end
__device__ __forceinline__ void dot4_acc2(float2& acc, const float4& a, const float4& b) {
  asm("{\n\t.reg .b64 rc, ra, rb;\n\tmov.b64 rc, {%0, %1};\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 ra, {%6, %7};\n\tmov.b64 rb, {%8, %9};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(acc.x), "+f"(acc.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(a.z), "f"(a.w), "f"(b.z), "f"(b.w));
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Second level of the two-level column reductions (bias gradients, BatchNorm statistics and backward sums, classifier-tail
// gradients): column c of out = sum over the nb block partials, fixed order, fp64.  A block is FIN_COLS columns x FIN_ROWS
// row-lanes; a thread adds every FIN_ROWS-th partial of its column with four independent accumulators (four loads in
// flight), then the FIN_ROWS lane sums are added in lane order.  The first version (32 columns x 8 row-lanes, one
// accumulator) walked 600 partials in 75 load-latency-bound iterations: 11-19 us per launch, nine such launches per COGMEN
// step -- nothing at 2^20 utterances per GPU, ~5 % of the step when the same batch is sharded over 8 GPUs.
constexpr int FIN_COLS = 8, FIN_ROWS = 32;           // FIN_COLS * FIN_ROWS == 256 threads
__device__ __forceinline__ double fin_reduce(const float* __restrict__ partial, int nb, long long ld, int c, bool ok,
                                             double (*sm)[FIN_COLS + 1]) {
  const int tx = threadIdx.x & (FIN_COLS - 1), ty = threadIdx.x / FIN_COLS;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (ok) {
    const float* p = partial + c;
    int b = ty;
    for (; b + 3 * FIN_ROWS < nb; b += 4 * FIN_ROWS) {
      const float v0 = p[(long long)b * ld], v1 = p[(long long)(b + FIN_ROWS) * ld], v2 = p[(long long)(b + 2 * FIN_ROWS) * ld],
                  v3 = p[(long long)(b + 3 * FIN_ROWS) * ld];
      s0 += (double)v0; s1 += (double)v1; s2 += (double)v2; s3 += (double)v3;
    }
    for (; b < nb; b += FIN_ROWS) s0 += (double)p[(long long)b * ld];
  }
  __syncthreads();                                   // sm may still be read by a previous call
  sm[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int y = 0; y < FIN_ROWS; ++y) tot += sm[y][tx];
  return tot;
}
static inline unsigned fin_blocks(int ncols) { return (unsigned)((ncols + FIN_COLS - 1) / FIN_COLS); }

// counter-based hash RNG for dropout: uniform in [0,1) from (seed, element index)
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// Dropout of the GEMM epilogues (ERCG_ACT_RELU_DROPOUT): ONE 64-bit hash per group of four consecutive columns of a row,
// 16 bits per element (drop iff bits < p * 65536).  The per-element splitmix64 above costs ~25 integer instructions;
// in the tcgen05 GEMM epilogue that was 2.5x the whole tile (ncu: 181 us vs 73 us for the same product without it).
__device__ __forceinline__ uint64_t hash64(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline unsigned dropout_thr16(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (unsigned)t);
}
// element (m, n) of an [M, N] matrix: group index m * ceil(N/4) + n/4, lane n%4
__device__ __forceinline__ uint64_t dropout_group_hash(uint64_t seed, long long m, int n, int N) {
  return hash64(seed, (uint64_t)m * (uint64_t)((N + 3) >> 2) + (uint64_t)(n >> 2));
}
__device__ __forceinline__ bool dropout_drop(uint64_t h, int lane4, unsigned thr16) {
  return (unsigned)((h >> (16 * lane4)) & 0xffffull) < thr16;
}

constexpr int kNumSMs = 148;   // B200
constexpr int kMaxDevices = 64; // per-device caches of launch attributes are indexed by the CUDA device id

// Per-device one-time setup (cudaFuncSetAttribute is per device context) and per-device attribute caches.  Function-local
// `static bool done` flags made the second GPU of a process launch with the first GPU's setup; these are indexed by the
// CUDA device id and safe to race on (worst case: the setup runs twice).
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}
struct DeviceOnce {
  std::atomic<unsigned char> done[kMaxDevices];
  bool need() const {
    const int d = current_device();
    return d < 0 || d >= kMaxDevices || !done[d].load(std::memory_order_acquire);
  }
  void mark() {
    const int d = current_device();
    if (d >= 0 && d < kMaxDevices) done[d].store(1, std::memory_order_release);
  }
};
inline int device_sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  const int d = current_device();
  if (d >= 0 && d < kMaxDevices) {
    const int c = cache[d].load(std::memory_order_acquire);
    if (c > 0) return c;
  }
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
  if (n < 1) n = kNumSMs;
  if (d >= 0 && d < kMaxDevices) cache[d].store(n, std::memory_order_release);
  return n;
}

}  // namespace ercg
