// Stand-alone DAG-ERC attention step: GAT_dialoggcn_v1.forward(Q, K, V, adj, s_mask) of the reference
// (track_mm/dagerc_models.py:326-365, mask_logic :83-90), for callers that use the class outside DAGERCModule's
// fused layer kernel (K10, dagerc.cu).
//
//   e[b,n]    = w_q . Q[b] + w_k . K[b,n] + bias - (1 - adj[b,n]) * 1e30          (Linear(2D,1) on [Q | K_n], mask_logic)
//   alpha[b,:] = softmax_n e[b,:]
//   S0[b] = sum_n alpha[b,n] * s[b,n] * V[b,n],  S1[b] = sum_n alpha[b,n] * (1 - s[b,n]) * V[b,n]
// and, by linearity of Wr0 / Wr1,  attn_sum[b] = Wr0 S0[b] + Wr1 S1[b]  =  [S0 | S1] @ [Wr0 | Wr1]^T -- that last product is
// an ordinary dense transform and goes through ercg_gemm_nn (the caller's job; the reference applies Wr0 AND Wr1 to all
// N context rows and then mixes, 2N matrix-vector products per query instead of 2).
// One CTA per query b; N (context length, <= a few hundred) lives in shared memory.
#include "common.cuh"

namespace ercg {

constexpr int GAT_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < GAT_THREADS / 32; ++w) t += red[w];        // fixed order: bit-reproducible
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[0];
#pragma unroll
  for (int w = 1; w < GAT_THREADS / 32; ++w) t = fmaxf(t, red[w]);
  return t;
}

__global__ void __launch_bounds__(GAT_THREADS)
dag_gat_fwd_kernel(const float* __restrict__ Q, long long ldq, const float* __restrict__ K, long long ldk_b, long long ldk_n,
                   const float* __restrict__ V, long long ldv_b, long long ldv_n, const float* __restrict__ adj,
                   const float* __restrict__ smask, long long ldm, const float* __restrict__ wlin, const float* __restrict__ blin,
                   float* __restrict__ alpha, float* __restrict__ S01, int N, int D) {
  extern __shared__ float sm[];             // e / alpha [N], s [N]
  __shared__ float red[GAT_THREADS / 32];
  float* e = sm;
  float* sv = sm + N;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* q = Q + (long long)b * ldq;
  float part = 0.f;
  for (int c = threadIdx.x; c < D; c += GAT_THREADS) part = fmaf(wlin[c], q[c], part);
  const float qdot = block_sum(part, red) + blin[0];
  for (int n = warp; n < N; n += GAT_THREADS / 32) {
    const float* k = K + (long long)b * ldk_b + (long long)n * ldk_n;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s = fmaf(wlin[D + c], k[c], s);
    s = warp_sum(s);
    if (lane == 0) {
      e[n] = (qdot + s) - (1.0f - adj[(long long)b * ldm + n]) * 1e30f;
      sv[n] = smask[(long long)b * ldm + n];
    }
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += GAT_THREADS) mx = fmaxf(mx, e[n]);
  mx = block_max(mx, red);
  float den = 0.f;
  for (int n = threadIdx.x; n < N; n += GAT_THREADS) { const float x = expf(e[n] - mx); e[n] = x; den += x; }
  den = block_sum(den, red);
  const float inv = 1.0f / den;
  for (int n = threadIdx.x; n < N; n += GAT_THREADS) { const float a = e[n] * inv; e[n] = a; alpha[(long long)b * N + n] = a; }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += GAT_THREADS) {
    float s0 = 0.f, s1 = 0.f;
    const float* v = V + (long long)b * ldv_b + c;
    for (int n = 0; n < N; ++n) {
      const float x = v[(long long)n * ldv_n], a = e[n], s = sv[n];
      s0 = fmaf(a * s, x, s0);
      s1 = fmaf(a * (1.0f - s), x, s1);
    }
    S01[(long long)b * 2 * D + c] = s0;
    S01[(long long)b * 2 * D + D + c] = s1;
  }
}

// backward: given dalpha_ext (may be NULL) and dS01 -> de [B,N] (gradient of the pre-softmax logits), dQ, dK, dV
__global__ void __launch_bounds__(GAT_THREADS)
dag_gat_bwd_kernel(const float* __restrict__ V, long long ldv_b, long long ldv_n, const float* __restrict__ smask, long long ldm,
                   const float* __restrict__ wlin, const float* __restrict__ alpha, const float* __restrict__ dalpha_ext,
                   const float* __restrict__ dS01, float* __restrict__ de, float* __restrict__ dQ, float* __restrict__ dK,
                   float* __restrict__ dV, int N, int D) {
  extern __shared__ float sm[];             // dalpha -> de [N], alpha [N], s [N]
  __shared__ float red[GAT_THREADS / 32];
  float* da = sm;
  float* al = sm + N;
  float* sv = sm + 2 * N;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* d0 = dS01 + (long long)b * 2 * D;
  const float* d1 = d0 + D;
  for (int n = warp; n < N; n += GAT_THREADS / 32) {
    const float* v = V + (long long)b * ldv_b + (long long)n * ldv_n;
    const float s = smask[(long long)b * ldm + n];
    float acc = 0.f;
    for (int c = lane; c < D; c += 32) acc = fmaf(fmaf(s, d0[c], (1.0f - s) * d1[c]), v[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      da[n] = acc + (dalpha_ext ? dalpha_ext[(long long)b * N + n] : 0.f);
      al[n] = alpha[(long long)b * N + n];
      sv[n] = s;
    }
  }
  __syncthreads();
  float part = 0.f;
  for (int n = threadIdx.x; n < N; n += GAT_THREADS) part = fmaf(al[n], da[n], part);
  const float dot = block_sum(part, red);
  float rs = 0.f;
  for (int n = threadIdx.x; n < N; n += GAT_THREADS) {
    const float g = al[n] * (da[n] - dot);
    da[n] = g;
    de[(long long)b * N + n] = g;
    rs += g;
  }
  rs = block_sum(rs, red);                  // (block_sum's barriers also publish da[] = de)
  for (int c = threadIdx.x; c < D; c += GAT_THREADS) {
    dQ[(long long)b * D + c] = rs * wlin[c];
    const float wk = wlin[D + c], g0 = d0[c], g1 = d1[c];
    for (int n = 0; n < N; ++n) {
      const long long o = ((long long)b * N + n) * D + c;
      dK[o] = da[n] * wk;
      dV[o] = al[n] * fmaf(sv[n], g0, (1.0f - sv[n]) * g1);
    }
  }
}

}  // namespace ercg

using namespace ercg;

extern "C" int ercg_dag_gat_fwd(const float* Q, int64_t ldq, const float* K, int64_t ldk_b, int64_t ldk_n, const float* V,
                                int64_t ldv_b, int64_t ldv_n, const float* adj, const float* s_mask, int64_t ldm,
                                const float* w_linear, const float* b_linear, float* alpha, float* S01, int B, int N, int D,
                                void* stream) {
  if (B < 0 || N < 1 || D < 1 || N > 8192) return ERCG_EINVAL;
  if (B == 0) return ERCG_OK;
  if (!Q || !K || !V || !adj || !s_mask || !w_linear || !b_linear || !alpha || !S01) return ERCG_EINVAL;
  dag_gat_fwd_kernel<<<B, GAT_THREADS, 2 * N * sizeof(float), (cudaStream_t)stream>>>(
      Q, ldq, K, ldk_b, ldk_n, V, ldv_b, ldv_n, adj, s_mask, ldm, w_linear, b_linear, alpha, S01, N, D);
  return finish_launch();
}

extern "C" int ercg_dag_gat_bwd(const float* V, int64_t ldv_b, int64_t ldv_n, const float* s_mask, int64_t ldm,
                                const float* w_linear, const float* alpha, const float* dalpha, const float* dS01, float* de,
                                float* dQ, float* dK, float* dV, int B, int N, int D, void* stream) {
  if (B < 0 || N < 1 || D < 1 || N > 4096) return ERCG_EINVAL;
  if (B == 0) return ERCG_OK;
  if (!V || !s_mask || !w_linear || !alpha || !dS01 || !de || !dQ || !dK || !dV) return ERCG_EINVAL;
  dag_gat_bwd_kernel<<<B, GAT_THREADS, 3 * N * sizeof(float), (cudaStream_t)stream>>>(
      V, ldv_b, ldv_n, s_mask, ldm, w_linear, alpha, dalpha, dS01, de, dQ, dK, dV, N, D);
  return finish_launch();
}
