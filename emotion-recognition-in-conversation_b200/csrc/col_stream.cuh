// Column reductions over a CONTIGUOUS row-major [M, H] fp32 matrix (row pitch == H, H % 4 == 0), used by bias gradients
// (colsum), BatchNorm statistics and the BatchNorm backward sums (track_mm/cogmen.py:67,72,116-122).
//
// The matrix is read as one flat stream of float4: thread t of the grid takes elements t, t + T, t + 2T, ... where the
// grid-wide stride T = gridDim.x * 256 is a multiple of H/4, so a thread always sees the SAME four columns and keeps its
// sums in registers.  Every lane is busy and every warp load is 512 contiguous bytes (the row-per-warp kernels used 25 of
// 32 lanes for H = 100 and ran at 44-49 % of the HBM peak); four loads are in flight per thread.  Per block the threads
// that share a column chunk are summed in thread order, the G <= 1024 block partials by the callers' fixed-order fp64
// second level => bit-reproducible.
#pragma once
#include "common.cuh"

namespace ercg {

constexpr int CS_MAX_BLOCKS = 1024;

// number of blocks: ~4 per SM, rounded to a multiple of nch / gcd(256, nch) so that gridDim.x * 256 % nch == 0
static inline int col_stream_blocks(int nch, long long total4) {
  int g = 256, n = nch;
  while (n) { const int t = g % n; g = n; n = t; }          // g = gcd(256, nch)
  const int q = nch / g;
  long long want = 4LL * kNumSMs;
  const long long need = (total4 + 255) / 256;
  if (want > need) want = need;
  long long blocks = (want + q - 1) / q * q;
  if (blocks < q) blocks = q;
  if (blocks > CS_MAX_BLOCKS) blocks = (long long)CS_MAX_BLOCKS / q * q;
  return blocks >= q ? (int)blocks : 0;                       // 0: shape not supported (q > CS_MAX_BLOCKS)
}
static inline bool col_stream_ok(const void* p, long long ld, int H, long long M) {
  return aligned16(p) && ld == H && (H & 3) == 0 && H >= 4 && (H >> 2) <= 256 && M >= 1024 &&
         col_stream_blocks(H >> 2, M * (long long)(H >> 2)) > 0;
}

// MODE 0: (sum (x - x[0,c]), sum (x - x[0,c])^2)     BatchNorm statistics, shifted by row 0
// MODE 1: (sum dy, sum dy * xhat), dy = dout * lrelu'(gamma * xhat + beta), xhat = (x - mean) * istd
// MODE 2: (sum x)                                     plain column sums
// partial layout: MODE 0/1 [gridDim.x][2H] (f0 | f1), MODE 2 [gridDim.x][H]
template <int MODE>
__global__ void __launch_bounds__(256)
col_stream_kernel(const float* __restrict__ x, const float* __restrict__ dout, const float* __restrict__ mean,
                  const float* __restrict__ var, float eps, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float slope, long long total4, int nch, float* __restrict__ partial) {
  __shared__ float4 red0[256], red1[256];
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long T = (long long)gridDim.x * 256;
  const int c = (int)(t % nch) * 4;                          // this thread's four columns, the same in every iteration
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  float4 mu = s0, istd = s0, g = s0, b = s0, shift = s0;
  if (MODE == 1) {
    mu = ld4(mean + c); g = ld4(gamma + c); b = ld4(beta + c);
    const float4 vv = ld4(var + c);
    istd = make_float4(1.0f / sqrtf(vv.x + eps), 1.0f / sqrtf(vv.y + eps), 1.0f / sqrtf(vv.z + eps), 1.0f / sqrtf(vv.w + eps));
  } else if (MODE == 0) {
    shift = ld4(x + c);
  }
  auto acc = [&](const float4& xv, const float4& dv) {
    if (MODE == 0) {
      const float4 xs = make_float4(xv.x - shift.x, xv.y - shift.y, xv.z - shift.z, xv.w - shift.w);
      s0.x += xs.x; s0.y += xs.y; s0.z += xs.z; s0.w += xs.w;
      s1.x = fmaf(xs.x, xs.x, s1.x); s1.y = fmaf(xs.y, xs.y, s1.y); s1.z = fmaf(xs.z, xs.z, s1.z); s1.w = fmaf(xs.w, xs.w, s1.w);
    } else if (MODE == 1) {
      const float4 xh = make_float4((xv.x - mu.x) * istd.x, (xv.y - mu.y) * istd.y, (xv.z - mu.z) * istd.z, (xv.w - mu.w) * istd.w);
      const float4 dy = make_float4(dv.x * (fmaf(g.x, xh.x, b.x) > 0.f ? 1.f : slope), dv.y * (fmaf(g.y, xh.y, b.y) > 0.f ? 1.f : slope),
                                    dv.z * (fmaf(g.z, xh.z, b.z) > 0.f ? 1.f : slope), dv.w * (fmaf(g.w, xh.w, b.w) > 0.f ? 1.f : slope));
      s0.x += dy.x; s0.y += dy.y; s0.z += dy.z; s0.w += dy.w;
      s1.x = fmaf(dy.x, xh.x, s1.x); s1.y = fmaf(dy.y, xh.y, s1.y); s1.z = fmaf(dy.z, xh.z, s1.z); s1.w = fmaf(dy.w, xh.w, s1.w);
    } else {
      s0.x += xv.x; s0.y += xv.y; s0.z += xv.z; s0.w += xv.w;
    }
  };
  long long i = t;
  for (; i + 3 * T < total4; i += 4 * T) {
    float4 xv[4], dv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xv[u] = ld4_stream(x + 4 * (i + u * T));
      dv[u] = MODE == 1 ? ld4_stream(dout + 4 * (i + u * T)) : xv[u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc(xv[u], dv[u]);
  }
  for (; i < total4; i += T) {
    const float4 xv = ld4_stream(x + 4 * i);
    acc(xv, MODE == 1 ? ld4_stream(dout + 4 * i) : xv);
  }
  red0[threadIdx.x] = s0;
  if (MODE != 2) red1[threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.x < nch) {                                   // threads tid, tid + nch, tid + 2 nch, ... share a column chunk
    float4 a = red0[threadIdx.x], q = MODE != 2 ? red1[threadIdx.x] : a;
    for (int k = threadIdx.x + nch; k < 256; k += nch) {
      const float4 v = red0[k];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      if (MODE != 2) {
        const float4 w = red1[k];
        q.x += w.x; q.y += w.y; q.z += w.z; q.w += w.w;
      }
    }
    const int H = nch * 4;
    if (MODE == 2) {
      st4(partial + (long long)blockIdx.x * H + c, a);
    } else {
      st4(partial + (long long)blockIdx.x * 2 * H + c, a);
      st4(partial + (long long)blockIdx.x * 2 * H + H + c, q);
    }
  }
}

// Plain column sums for narrow contiguous matrices whose width is even but not a multiple of 4 (the classifier's [N, 6]
// logit gradients): the same flat stream in float2 units.  partial layout [gridDim.x][H].
static inline bool col_stream2_ok(const void* p, long long ld, int H, long long M) {
  return (reinterpret_cast<uintptr_t>(p) & 7u) == 0 && ld == H && (H & 1) == 0 && H >= 2 && (H >> 1) <= 256 && M >= 1024 &&
         col_stream_blocks(H >> 1, M * (long long)(H >> 1)) > 0;
}
static __global__ void __launch_bounds__(256)
col_stream2_kernel(const float* __restrict__ x, long long total2, int nch, float* __restrict__ partial) {
  __shared__ float2 red[256];
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long T = (long long)gridDim.x * 256;
  const int c = (int)(t % nch) * 2;
  float2 s = make_float2(0.f, 0.f);
  const float2* x2 = reinterpret_cast<const float2*>(x);
  long long i = t;
  for (; i + 3 * T < total2; i += 4 * T) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x2 + i + u * T);
#pragma unroll
    for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; }
  }
  for (; i < total2; i += T) { const float2 v = __ldg(x2 + i); s.x += v.x; s.y += v.y; }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < nch) {
    float2 a = red[threadIdx.x];
    for (int k = threadIdx.x + nch; k < 256; k += nch) { a.x += red[k].x; a.y += red[k].y; }
    partial[(long long)blockIdx.x * nch * 2 + c] = a.x;
    partial[(long long)blockIdx.x * nch * 2 + c + 1] = a.y;
  }
}

}  // namespace ercg
