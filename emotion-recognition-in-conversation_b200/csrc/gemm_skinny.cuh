// K2 (skinny shapes): dense transforms whose narrow side is <= 16 wide -- the last classifier layer Linear(100, C)
// (track_mm/cogmen.py:121, dgcn_models.py:161; C = 4..7 classes), its input gradient and its weight gradient.
// A 128x128 tile wastes >90 % of its lanes on those; here they are streaming kernels bound by the one wide operand:
//   skinny_nn_small_n : C[M,N] = A[M,K] @ B[K,N] + bias,   N <= 16   (one warp per row, K across the lanes)
//   skinny_nn_small_k : C[M,N] = A[M,K] @ B[K,N] + bias,   K <= 16   (one thread per 4 output columns)
//   skinny_tn_small_n : P[s][K1,N1] = A[rows of slab s]^T @ B[rows of slab s], N1 <= 16 (thread per k1, fixed order)
// exact fp32 FMA.  Dispatched from ercg_gemm_nn / ercg_gemm_tn (gemm_simt.cu); not separate entry points.
#pragma once
#include "common.cuh"

namespace ercg {

constexpr int SK_MAXN = 16;

template <int NN>
__global__ void __launch_bounds__(256)
skinny_nn_small_n_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                         const float* __restrict__ bias, float* __restrict__ C, long long ldc, long long M, int N, int K) {
  extern __shared__ float sk_sm[];            // B transposed: [NN][KP] (zero padded), so that lanes read consecutive k
  const int KP = (K + 3) & ~3;
  for (int i = threadIdx.x; i < NN * KP; i += 256) {
    const int n = i / KP, k = i % KP;
    sk_sm[i] = (n < N && k < K) ? B[(long long)k * ldb + n] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = ((lda & 3) == 0) && aligned16(A) && ((K & 3) == 0);
  for (long long m = (long long)blockIdx.x * 8 + warp; m < M; m += (long long)gridDim.x * 8) {
    const float* a = A + m * lda;
    float acc[NN];
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = 0.f;
    if (vec) {
      for (int c = lane; c < (K >> 2); c += 32) {
        const float4 v = ld4_stream(a + 4 * c);
#pragma unroll
        for (int n = 0; n < NN; ++n) acc[n] += dot4(v, *reinterpret_cast<const float4*>(sk_sm + n * KP + 4 * c));
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float v = a[k];
#pragma unroll
        for (int n = 0; n < NN; ++n) acc[n] = fmaf(v, sk_sm[n * KP + k], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < NN; ++n)
        if (n < N) C[m * ldc + n] = acc[n] + (bias ? bias[n] : 0.f);
    }
  }
}

// same product for 16-byte aligned rows with K <= 128 (the classifier: K = 100): EIGHT lanes per row, four rows per warp
// iteration -- all of a lane's (<= 4) 16-byte loads are issued before the first use, and the reduction costs 3 shuffles
// per output instead of 5 (the one-row-per-warp version spent its time in the 30 shuffles per row and had 400 bytes in
// flight per warp: 18 % of the HBM peak).
template <int NN>
__global__ void __launch_bounds__(256)
skinny_nn_small_n_q8_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                            const float* __restrict__ bias, float* __restrict__ C, long long ldc, long long M, int N, int K) {
  extern __shared__ float sk_sm[];            // B transposed: [NN][KP] (zero padded)
  const int KP = (K + 3) & ~3;
  for (int i = threadIdx.x; i < NN * KP; i += 256) {
    const int n = i / KP, k = i % KP;
    sk_sm[i] = (n < N && k < K) ? B[(long long)k * ldb + n] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, q = lane >> 3;
  const int nch = K >> 2;
  float bn = 0.f;                              // lane `sub` keeps bias[sub] / bias[sub + 8]
  float bn2 = 0.f;
  if (bias) { if (sub < N) bn = bias[sub]; if (NN > 8 && sub + 8 < N) bn2 = bias[sub + 8]; }
  for (long long m0 = ((long long)blockIdx.x * 8 + warp) * 4; m0 < M; m0 += (long long)gridDim.x * 32) {
    const long long m = m0 + q;
    const bool live = m < M;
    const float* a = A + (live ? m : M - 1) * lda;
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (sub + 8 * i < nch) ? ld4_stream(a + 4 * (sub + 8 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float acc[NN];
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (sub + 8 * i < nch) {
#pragma unroll
        for (int n = 0; n < NN; ++n) acc[n] += dot4(v[i], *reinterpret_cast<const float4*>(sk_sm + n * KP + 4 * (sub + 8 * i)));
      }
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 4);
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 2);
      acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 1);
    }
    // lane `sub` writes column sub (and sub + 8): pick acc[sub] without dynamic register indexing
    float o1 = 0.f, o2 = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) if (sub == n) { o1 = acc[n]; if (NN > 8) o2 = acc[n + 8]; }
    if (live) {
      if (sub < N) C[m * ldc + sub] = o1 + bn;
      if (NN > 8 && sub + 8 < N) C[m * ldc + sub + 8] = o2 + bn2;
    }
  }
}

__global__ void __launch_bounds__(256)
skinny_nn_small_k_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                         const float* __restrict__ bias, float* __restrict__ C, long long ldc, long long M, int N, int K) {
  extern __shared__ float sk_sm[];            // B [K][N4*4] zero padded, then bias
  const int N4 = (N + 3) >> 2, NP = N4 * 4;
  for (int i = threadIdx.x; i < K * NP; i += 256) {
    const int k = i / NP, n = i % NP;
    sk_sm[i] = n < N ? B[(long long)k * ldb + n] : 0.f;
  }
  for (int n = threadIdx.x; n < NP; n += 256) sk_sm[K * NP + n] = (bias && n < N) ? bias[n] : 0.f;
  __syncthreads();
  const bool vec_c = ((ldc & 3) == 0) && aligned16(C);
  const long long total = M * N4;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const long long m = idx / N4;
    const int c = (int)(idx % N4);
    const float* a = A + m * lda;
    float4 acc = *reinterpret_cast<const float4*>(sk_sm + K * NP + 4 * c);
    for (int k = 0; k < K; ++k) fma4(acc, __ldg(a + k), *reinterpret_cast<const float4*>(sk_sm + k * NP + 4 * c));
    float* o = C + m * ldc + 4 * c;
    if (vec_c && 4 * c + 4 <= N) {
      st4_stream(o, acc);
    } else {
      if (4 * c + 0 < N) o[0] = acc.x;
      if (4 * c + 1 < N) o[1] = acc.y;
      if (4 * c + 2 < N) o[2] = acc.z;
      if (4 * c + 3 < N) o[3] = acc.w;
    }
  }
}

// thread t owns k1 = t % KP of row group t / KP; the block's row groups take rows round-robin; partial sums of the
// groups are reduced through shared memory in group order, and the slabs by reduce_splits_kernel (fixed order).
template <int NN>
__global__ void __launch_bounds__(512)
skinny_tn_small_n_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                         float* __restrict__ P, long long M, int K1, int N1, long long rows_per_slab, int groups, int KP) {
  extern __shared__ float sk_sm[];            // [groups][KP][NN]
  const int g = threadIdx.x / KP, k1 = threadIdx.x % KP;
  const long long mbeg = (long long)blockIdx.x * rows_per_slab;
  long long mend = mbeg + rows_per_slab;
  if (mend > M) mend = M;
  float acc[NN];
#pragma unroll
  for (int n = 0; n < NN; ++n) acc[n] = 0.f;
  if (g < groups && k1 < K1) {
    long long m = mbeg + g;
    for (; m + 3LL * groups < mend; m += 4LL * groups) {
      float a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = __ldg(A + (m + (long long)u * groups) * lda + k1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* b = B + (m + (long long)u * groups) * ldb;
#pragma unroll
        for (int n = 0; n < NN; ++n)
          if (n < N1) acc[n] = fmaf(a[u], __ldg(b + n), acc[n]);
      }
    }
    for (; m < mend; m += groups) {
      const float a = __ldg(A + m * lda + k1);
      const float* b = B + m * ldb;
#pragma unroll
      for (int n = 0; n < NN; ++n)
        if (n < N1) acc[n] = fmaf(a, __ldg(b + n), acc[n]);
    }
  }
  if (g < groups) {
#pragma unroll
    for (int n = 0; n < NN; ++n) sk_sm[(g * KP + k1) * NN + n] = acc[n];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K1 * N1; i += blockDim.x) {
    const int kk = i / N1, n = i % N1;
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += sk_sm[(gg * KP + kk) * NN + n];
    P[(long long)blockIdx.x * K1 * N1 + i] = s;
  }
}

// float4 variant for 16-byte aligned A rows with K1 % 4 == 0: thread = (row group, four consecutive k1), so one 16-byte
// load of A feeds 4*N1 FMAs (the scalar version issued 7 loads per 6 FMAs and ran at 15 % of the HBM peak).
// Shared memory: [groups][K1][N1] partials, reduced in group order.
template <int NN>
__global__ void __launch_bounds__(512)
skinny_tn_small_n_v4_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                            float* __restrict__ P, long long M, int K1, int N1, long long rows_per_slab, int groups) {
  extern __shared__ float sk_sm[];
  const int KQ = K1 >> 2;
  const int g = threadIdx.x / KQ, kq = threadIdx.x % KQ;
  const long long mbeg = (long long)blockIdx.x * rows_per_slab;
  long long mend = mbeg + rows_per_slab;
  if (mend > M) mend = M;
  float acc[4][NN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[i][n] = 0.f;
  if (g < groups) {
    long long m = mbeg + g;
    for (; m + 3LL * groups < mend; m += 4LL * groups) {
      float4 a[4];
      float b[4][NN];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = ld4_stream(A + (m + (long long)u * groups) * lda + 4 * kq);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* bp = B + (m + (long long)u * groups) * ldb;
#pragma unroll
        for (int n = 0; n < NN; ++n) b[u][n] = n < N1 ? __ldg(bp + n) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int n = 0; n < NN; ++n) {
          acc[0][n] = fmaf(a[u].x, b[u][n], acc[0][n]);
          acc[1][n] = fmaf(a[u].y, b[u][n], acc[1][n]);
          acc[2][n] = fmaf(a[u].z, b[u][n], acc[2][n]);
          acc[3][n] = fmaf(a[u].w, b[u][n], acc[3][n]);
        }
    }
    for (; m < mend; m += groups) {
      const float4 a = ld4_stream(A + m * lda + 4 * kq);
      const float* bp = B + m * ldb;
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        const float bv = n < N1 ? __ldg(bp + n) : 0.f;
        acc[0][n] = fmaf(a.x, bv, acc[0][n]);
        acc[1][n] = fmaf(a.y, bv, acc[1][n]);
        acc[2][n] = fmaf(a.z, bv, acc[2][n]);
        acc[3][n] = fmaf(a.w, bv, acc[3][n]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int n = 0; n < NN; ++n)
        if (n < N1) sk_sm[((long long)g * K1 + 4 * kq + i) * N1 + n] = acc[i][n];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K1 * N1; i += blockDim.x) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += sk_sm[(long long)gg * K1 * N1 + i];
    P[(long long)blockIdx.x * K1 * N1 + i] = s;
  }
}

}  // namespace ercg
