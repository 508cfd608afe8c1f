// K3: deterministic warp-segmented gather-reduce over the packed CSR (no atomics) and its backward.
//
//   fwd : out[k,:] = sum_{e in row k} w[e] * Y[col[e], etype[e]*H : +H] (+ Y[k, root_off:+H]) (+ bias)
//   bwd : dY[j, r*H:+H] = sum_{e in out(j), type(e)=r} w[e] * dout[dst(e)],  dY[j, root_off:+H] = dout[j],
//         dw[eid] = <dout[dst], Y[src, type]>
// Replaces MessagePassing.propagate / scatter_add (models/rgcn.py:188-221,15-46; edge_norm at :343),
// the per-relation mean aggregation of PyG RGCNConv (track_mm/cogmen.py:65,71; w = 1/|N_r(k)|) and the
// neighbour sum of PyG GraphConv (track_mm/dgcn_models.py:42,46; etype = NULL, w = NULL).
//
// One warp per node; lane c owns float4 chunk(s) c, c+32, ... of the H-wide row (H = 100 -> 25 lanes,
// 400-byte rows, 16-byte aligned).  Edge metadata is read warp-uniformly (broadcast); neighbour rows
// are 16-byte vector loads; the edge loop is unrolled x4 so four independent rows are in flight per
// lane.  The summation order is the CSR order => bit-reproducible, independent of the grid.
// Algorithmic HBM bytes per node: 4H*P (distinct Y slots referenced) + 4H (root) +
// 4H (out) + 4 (rowptr) + 9*deg (col, etype, w).
#include "common.cuh"
#include <cstdlib>

namespace ercg {

constexpr int GW = 8;      // warps per block
constexpr int MAXC = 2;    // float4 chunks per lane => H <= 256

template <int NC>
__global__ void __launch_bounds__(GW * 32)
gather_fwd_kernel(const float* __restrict__ Y, long long ldy, const int* __restrict__ rowptr,
                  const int* __restrict__ col, const uint8_t* __restrict__ etype, const int* __restrict__ rel_slot,
                  const float* __restrict__ w, int root_off, const float* __restrict__ bias, float* __restrict__ out, long long ldo,
                  long long N, int H) {
  const int lane = threadIdx.x & 31;
  const long long node = (long long)blockIdx.x * GW + (threadIdx.x >> 5);
  if (node >= N) return;
  const int nch = H >> 2;
  float4 acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int beg = rowptr[node], end = rowptr[node + 1];
  int e = beg;
  for (; e + 4 <= end; e += 4) {
    const float* p[4];
    float ww[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = col[e + u];
      int t = etype ? (int)etype[e + u] : 0;
      if (rel_slot) t = __ldg(rel_slot + t);
      ww[u] = w ? w[e + u] : 1.f;
      if (t < 0) { t = 0; ww[u] = 0.f; }                  // relation id without a slot in Y (layer has fewer relations): no message
      p[u] = Y + (long long)s * ldy + (long long)t * H;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld4(p[u] + 4 * ch);
#pragma unroll
        for (int u = 0; u < 4; ++u) fma4(acc[c], ww[u], v[u]);
      }
    }
  }
  for (; e < end; ++e) {
    const int s = col[e];
    int t = etype ? (int)etype[e] : 0;
    if (rel_slot) t = __ldg(rel_slot + t);
    float ww = w ? w[e] : 1.f;
    if (t < 0) { t = 0; ww = 0.f; }
    const float* p = Y + (long long)s * ldy + (long long)t * H;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) fma4(acc[c], ww, ld4(p + 4 * ch));
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nch) {
      float4 r = acc[c];
      if (root_off >= 0) {
        const float4 y = ld4(Y + node * ldy + root_off + 4 * ch);
        r.x += y.x; r.y += y.y; r.z += y.z; r.w += y.w;
      }
      if (bias) {
        const float4 b = ld4(bias + 4 * ch);
        r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
      }
      st4(out + node * ldo + 4 * ch, r);
    }
  }
}

// Forward with a FLAT thread-per-chunk mapping: thread c owns float4 chunk c % (H/4) of node c / (H/4).  With H = 100 a warp
// of the kernel above has 25 busy lanes of 32; here every lane is busy (a warp straddles two or three nodes: their edge
// metadata is read per lane, 2-3 distinct addresses per warp).  Same edges in the same order with the same 4-way grouping,
// so the result is bit-identical to gather_fwd_kernel.
__global__ void __launch_bounds__(256)
gather_fwd_flat_kernel(const float* __restrict__ Y, long long ldy, const int* __restrict__ rowptr,
                       const int* __restrict__ col, const uint8_t* __restrict__ etype, const int* __restrict__ rel_slot,
                       const float* __restrict__ w, int root_off, const float* __restrict__ bias, float* __restrict__ out,
                       long long ldo, long long N, int H) {
  const int nch = H >> 2;
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= N * nch) return;
  const long long node = c / nch;
  const int ch = (int)(c - node * nch);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int beg = rowptr[node], end = rowptr[node + 1];
  int e = beg;
  for (; e + 4 <= end; e += 4) {
    const float* p[4];
    float ww[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = col[e + u];
      int t = etype ? (int)etype[e + u] : 0;
      if (rel_slot) t = __ldg(rel_slot + t);
      ww[u] = w ? w[e + u] : 1.f;
      if (t < 0) { t = 0; ww[u] = 0.f; }                  // relation id without a slot in Y (layer has fewer relations): no message
      p[u] = Y + (long long)s * ldy + (long long)t * H + 4 * ch;
    }
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld4(p[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) fma4(acc, ww[u], v[u]);
  }
  for (; e < end; ++e) {
    const int s = col[e];
    int t = etype ? (int)etype[e] : 0;
    if (rel_slot) t = __ldg(rel_slot + t);
    float ww = w ? w[e] : 1.f;
    if (t < 0) { t = 0; ww = 0.f; }
    fma4(acc, ww, ld4(Y + (long long)s * ldy + (long long)t * H + 4 * ch));
  }
  if (root_off >= 0) {
    const float4 y = ld4(Y + node * ldy + root_off + 4 * ch);
    acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
  }
  if (bias) {
    const float4 b = ld4(bias + 4 * ch);
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  st4(out + node * ldo + 4 * ch, acc);
}

// by-source backward.  For every relation slot r the warp re-walks the out-edges of j (the dout rows
// of a window are L1-resident) and writes the full slot, so dY needs no zero-fill.
template <int NC>
__global__ void __launch_bounds__(GW * 32)
gather_bwd_kernel(const float* __restrict__ dout, long long ldo, const float* __restrict__ Y, long long ldy,
                  const int* __restrict__ t_rowptr, const int* __restrict__ t_col,
                  const uint8_t* __restrict__ t_etype, const int* __restrict__ t_eid, const int* __restrict__ rel_slot,
                  const float* __restrict__ w, int R, int root_off, float* __restrict__ dY, long long lddy, float* __restrict__ dw,
                  long long N, int H) {
  const int lane = threadIdx.x & 31;
  const long long node = (long long)blockIdx.x * GW + (threadIdx.x >> 5);
  if (node >= N) return;
  const int nch = H >> 2;
  const int beg = t_rowptr[node], end = t_rowptr[node + 1];
  if (end - beg <= 32) {
    // common case (window graphs: <= 11 / 21 out-edges): one lane-parallel metadata fetch, then per relation slot a
    // ballot selects the edges of that type; empty slots are written as zeros without touching memory
    const int deg = end - beg;
    int md = 0, mt = -1;
    float mw = 0.f;
    if (lane < deg) {
      md = t_col[beg + lane];
      mt = t_etype ? (int)t_etype[beg + lane] : 0;
      mw = w ? w[t_eid[beg + lane]] : 1.f;
    }
    for (int r = 0; r < R; ++r) {
      const int slot = rel_slot ? __ldg(rel_slot + r) : r;       // relation ids without a slot occur on no edge
      if (slot < 0) continue;
      unsigned m = __ballot_sync(0xffffffffu, mt == r);
      float4 acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      while (m) {
        const float* p[4];
        float ww[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ok[q] = m != 0;
          const int l = ok[q] ? __ffs(m) - 1 : 0;
          if (ok[q]) m &= m - 1;
          p[q] = dout + (long long)__shfl_sync(0xffffffffu, md, l) * ldo;
          ww[q] = __shfl_sync(0xffffffffu, mw, l);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int ch = lane + 32 * c;
          if (ch < nch) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = ok[q] ? ld4(p[q] + 4 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);   // warp-uniform
#pragma unroll
            for (int q = 0; q < 4; ++q) fma4(acc[c], ww[q], v[q]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nch) st4_stream(dY + node * lddy + (long long)slot * H + 4 * ch, acc[c]);
      }
    }
  } else {
    for (int r = 0; r < R; ++r) {
      const int slot = rel_slot ? __ldg(rel_slot + r) : r;
      if (slot < 0) continue;
      float4 acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = beg; e < end; ++e) {
        const int t = t_etype ? (int)t_etype[e] : 0;
        if (t != r) continue;                              // warp-uniform
        const int d = t_col[e];
        const float ww = w ? w[t_eid[e]] : 1.f;
        const float* p = dout + (long long)d * ldo;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int ch = lane + 32 * c;
          if (ch < nch) fma4(acc[c], ww, ld4(p + 4 * ch));
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nch) st4_stream(dY + node * lddy + (long long)slot * H + 4 * ch, acc[c]);
      }
    }
  }
  if (root_off >= 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nch) st4_stream(dY + node * lddy + root_off + 4 * ch, ld4(dout + node * ldo + 4 * ch));
    }
  }
  if (dw) {
    for (int e = beg; e < end; ++e) {
      int t = t_etype ? (int)t_etype[e] : 0;
      if (rel_slot) t = __ldg(rel_slot + t);
      const int d = t_col[e];
      const float* pd = dout + (long long)d * ldo;
      const float* py = Y + node * ldy + (long long)t * H;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int ch = lane + 32 * c;
        if (ch < nch) s += dot4(ld4(pd + 4 * ch), ld4(py + 4 * ch));
      }
      s = warp_sum(s);
      if (lane == 0) dw[t_eid[e]] = s;
    }
  }
}

// ---------------------------------------------------------------------------------- K3 backward on WINDOW graphs: CTA tiles
// For the graphs K1 builds, the destinations of source j are the contiguous rows [j - wlo, j + whi] (wlo = wp, whi = wf of
// batch_graphify).  A CTA owns 32 consecutive sources and stages the 32 + wlo + whi rows of dout around them in shared
// memory with cp.async (every row is needed by up to wlo + whi + 1 sources); edge metadata is fetched lane-per-edge while
// the copies are in flight.  Per relation slot a ballot selects the edges, rows come from shared memory.  Same summation
// order as gather_bwd_kernel (ascending by-source edge order inside a slot) => bit-identical results.
#ifndef GT_TILE_
#define GT_TILE_ 32
#endif
constexpr int GT_TILE = GT_TILE_;

__global__ void __launch_bounds__(256)
gather_bwd_tile_kernel(const float* __restrict__ dout, long long ldo, const int* __restrict__ t_rowptr,
                       const int* __restrict__ t_col, const uint8_t* __restrict__ t_etype, const int* __restrict__ t_eid,
                       const int* __restrict__ rel_slot, const float* __restrict__ w, int n_slots, int root_off,
                       float* __restrict__ dY, long long lddy, long long N, int H, int wlo, int whi) {
  extern __shared__ float4 gt_sm[];
  const int nch = H >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long t0 = (long long)blockIdx.x * GT_TILE;
  const int tn = (int)min((long long)GT_TILE, N - t0);
  const long long r0 = max(0LL, t0 - wlo), r1 = min(N, t0 + tn + whi);
  const int R = (int)(r1 - r0);
  if (lane < nch)
    for (int r = warp; r < R; r += 8) {
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(gt_sm + r * nch + lane);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(dout + (r0 + r) * ldo + 4 * lane) : "memory");
    }
  constexpr int NPW = GT_TILE / 8;
  int degs[NPW], rows[NPW], slots[NPW];
  float ws[NPW];
#pragma unroll
  for (int p = 0; p < NPW; ++p) {
    const int nl = warp * NPW + p;
    const bool nok = nl < tn;
    const int beg = nok ? t_rowptr[t0 + nl] : 0;
    degs[p] = nok ? t_rowptr[t0 + nl + 1] - beg : 0;
    rows[p] = 0;
    slots[p] = -1;
    ws[p] = 0.f;
    if (lane < degs[p]) {
      const long long dst = t_col[beg + lane];
      if (dst < r0 || dst >= r1) __trap();               // the caller promised a window graph
      rows[p] = (int)(dst - r0);
      int t = t_etype ? (int)t_etype[beg + lane] : 0;
      if (rel_slot) t = __ldg(rel_slot + t);
      slots[p] = t;
      ws[p] = w ? w[t_eid[beg + lane]] : 1.f;
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
#pragma unroll
  for (int p = 0; p < NPW; ++p) {
    const int nl = warp * NPW + p;
    if (nl >= tn) break;                                   // warp-uniform
    const long long node = t0 + nl;
    for (int sidx = 0; sidx < n_slots; ++sidx) {
      unsigned m = __ballot_sync(0xffffffffu, slots[p] == sidx);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const float wu = __shfl_sync(0xffffffffu, ws[p], l);
        const int rw = __shfl_sync(0xffffffffu, rows[p], l);
        if (lane < nch) fma4(acc, wu, gt_sm[rw * nch + lane]);
      }
      if (lane < nch) st4_stream(dY + node * lddy + (long long)sidx * H + 4 * lane, acc);
    }
    if (root_off >= 0 && lane < nch) st4_stream(dY + node * lddy + root_off + 4 * lane, gt_sm[(int)(node - r0) * nch + lane]);
  }
}

static int check_rows(const void* p, long long ld, int H) {
  if ((H & 3) || H <= 0 || H > 128 * MAXC) return ERCG_EINVAL;
  if ((ld & 3) || !aligned16(p)) return ERCG_EALIGN;
  return ERCG_OK;
}

}  // namespace ercg

using namespace ercg;

extern "C" int ercg_gather_fwd(const float* Y, int64_t ldy, const int32_t* rowptr, const int32_t* col,
                               const uint8_t* etype, const int32_t* rel_slot, const float* w, int root_off,
                               const float* bias, float* out, int64_t ldo, int64_t N, int H, void* stream) {
  if (N < 0) return ERCG_EINVAL;
  if (N == 0) return ERCG_OK;
  if (!Y || !rowptr || !col || !out) return ERCG_EINVAL;
  int rc = check_rows(Y, ldy, H);
  if (rc) return rc;
  rc = check_rows(out, ldo, H);
  if (rc) return rc;
  if ((root_off >= 0 && (root_off & 3)) || (bias && !aligned16(bias))) return ERCG_EALIGN;
  const unsigned blocks = (unsigned)((N + GW - 1) / GW);
  cudaStream_t st = (cudaStream_t)stream;
  const char* flat_env = getenv("ERCG_GATHER_FLAT");      // "0" forces the warp-per-node kernel (A/B tests; read per call)
  const int flat = flat_env ? atoi(flat_env) : 1;
  const long long chunks = N * (long long)(H >> 2);
  if (flat && (H >> 2) % 32 != 0 && chunks < 2147483647LL * 256) {      // widths that leave lanes idle in the warp-per-node form
    gather_fwd_flat_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(Y, ldy, rowptr, col, etype, etype ? rel_slot : nullptr,
                                                                            w, root_off, bias, out, ldo, N, H);
    return finish_launch();
  }
  if (H <= 128) gather_fwd_kernel<1><<<blocks, GW * 32, 0, st>>>(Y, ldy, rowptr, col, etype, etype ? rel_slot : nullptr, w, root_off, bias, out, ldo, N, H);
  else gather_fwd_kernel<2><<<blocks, GW * 32, 0, st>>>(Y, ldy, rowptr, col, etype, etype ? rel_slot : nullptr, w, root_off, bias, out, ldo, N, H);
  return finish_launch();
}

extern "C" int ercg_gather_bwd(const float* dout, int64_t ldo, const float* Y, int64_t ldy,
                               const int32_t* t_rowptr, const int32_t* t_col, const uint8_t* t_etype,
                               const int32_t* t_eid, const int32_t* rel_slot, const float* w, int R, int root_off,
                               float* dY, int64_t lddy, float* dw, int64_t N, int H, void* stream) {
  if (N < 0 || R < 1) return ERCG_EINVAL;
  if (N == 0) return ERCG_OK;
  if (!dout || !t_rowptr || !t_col || !dY) return ERCG_EINVAL;
  if ((w || dw) && !t_eid) return ERCG_EINVAL;
  if (dw && !Y) return ERCG_EINVAL;
  int rc = check_rows(dout, ldo, H);
  if (rc) return rc;
  rc = check_rows(dY, lddy, H);
  if (rc) return rc;
  if (dw && (rc = check_rows(Y, ldy, H))) return rc;
  if (root_off >= 0 && (root_off & 3)) return ERCG_EALIGN;
  const unsigned blocks = (unsigned)((N + GW - 1) / GW);
  cudaStream_t st = (cudaStream_t)stream;
  if (H <= 128) gather_bwd_kernel<1><<<blocks, GW * 32, 0, st>>>(dout, ldo, Y, ldy, t_rowptr, t_col, t_etype, t_eid, t_etype ? rel_slot : nullptr, w, R, root_off, dY, lddy, dw, N, H);
  else gather_bwd_kernel<2><<<blocks, GW * 32, 0, st>>>(dout, ldo, Y, ldy, t_rowptr, t_col, t_etype, t_eid, t_etype ? rel_slot : nullptr, w, R, root_off, dY, lddy, dw, N, H);
  return finish_launch();
}

// window-graph variant of ercg_gather_bwd without the per-edge weight gradient (contract as for ercg_attn_window_*)
extern "C" int ercg_gather_window_bwd(const float* dout, int64_t ldo, const int32_t* t_rowptr, const int32_t* t_col,
                                      const uint8_t* t_etype, const int32_t* t_eid, const int32_t* rel_slot,
                                      const float* w, int n_slots, int root_off, float* dY, int64_t lddy, int64_t N,
                                      int H, int wlo, int whi, void* stream) {
  if (N < 0 || n_slots < 1 || H > 128 || wlo < 0 || whi < 0 || wlo + whi + 1 > 32) return ERCG_EINVAL;
  if (N == 0) return ERCG_OK;
  if (!dout || !t_rowptr || !t_col || !dY || (w && !t_eid)) return ERCG_EINVAL;
  int rc = check_rows(dout, ldo, H);
  if (rc) return rc;
  rc = check_rows(dY, lddy, H);
  if (rc) return rc;
  if (root_off >= 0 && (root_off & 3)) return ERCG_EALIGN;
  const size_t sm = (size_t)(H >> 2) * 16 * (GT_TILE + wlo + whi);
  const unsigned blocks = (unsigned)((N + GT_TILE - 1) / GT_TILE);
  gather_bwd_tile_kernel<<<blocks, 256, sm, (cudaStream_t)stream>>>(dout, ldo, t_rowptr, t_col, t_etype, t_eid,
                                                                     t_etype ? rel_slot : nullptr, w, n_slots, root_off, dY,
                                                                     lddy, N, H, wlo, whi);
  return finish_launch();
}
