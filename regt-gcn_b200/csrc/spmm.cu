// F-wide CSR SpMM over the x[B,N,F,T] layout of the reference (load_dataset.py:456).
//
// Because T is the innermost dimension of x, one neighbour row x[b,j,:,:] is F*T
// contiguous floats (96 floats = 384 B at T=12): a single gather serves all T periods,
// so  S[b,n,:,:] = sum_e A_hat[e] * x[b,col[e],:,:]  is one SpMM on 384-byte rows instead
// of 3*T scatter-add passes on 1 KB rows (GCNConv.propagate, models/utils.py:169,175,181).
// HBM-bound: algorithmic bytes = read x once + write y once + CSR (SURVEY 8(d) SpMMBytes).
//
// One warp per output row; lane c owns float4 #c of the row (128-bit loads/stores, fully
// coalesced 384 B per neighbour); edge metadata is read once per warp (uniform address).
#include "common.cuh"

namespace regt {

__device__ __forceinline__ float4 ld_nc4(const float4* p) { return __ldg(p); }

template <int UNROLL>
__global__ void __launch_bounds__(256) k_spmm_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                   const float* __restrict__ val, const float4* __restrict__ x,
                                                   float4* __restrict__ y, int B, int n_out, int n_in, int W4) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)B * n_out) return;
  const int b = (int)(warp / n_out), r = (int)(warp % n_out);
  const float4* xb = x + (size_t)b * n_in * W4;
  float4* yr = y + ((size_t)b * n_out + r) * W4;
  const int e0 = rowptr[r], e1 = rowptr[r + 1];
  for (int c = lane; c < W4; c += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = e0;
    for (; e + UNROLL <= e1; e += UNROLL) {
      float4 v[UNROLL];
      float w[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        w[u] = __ldg(val + e + u);
        v[u] = ld_nc4(xb + (size_t)__ldg(col + e + u) * W4 + c);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {  // sequential CSR order: deterministic sums
        acc.x = fmaf(w[u], v[u].x, acc.x);
        acc.y = fmaf(w[u], v[u].y, acc.y);
        acc.z = fmaf(w[u], v[u].z, acc.z);
        acc.w = fmaf(w[u], v[u].w, acc.w);
      }
    }
    for (; e < e1; ++e) {
      float w = __ldg(val + e);
      float4 v = ld_nc4(xb + (size_t)__ldg(col + e) * W4 + c);
      acc.x = fmaf(w, v.x, acc.x);
      acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z);
      acc.w = fmaf(w, v.w, acc.w);
    }
    yr[c] = acc;
  }
}

// NB neighbour rows of one batch: lane metadata (column, weight) is broadcast by shuffle; slots past
// the end of the row re-read the row's last neighbour with weight 0 (acc + 0*v leaves acc unchanged),
// so the loads are unconditional and issue back to back.  Sums stay in sequential CSR order.
template <int NB>
__device__ __forceinline__ void gather_batch(const float4* __restrict__ xb, int W4, int c, int mycol, float myval, int j,
                                             float4& acc) {
  float4 v[NB];
  float wv[NB];
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    const int cu = __shfl_sync(0xffffffffu, mycol, (j + u) & 31);
    wv[u] = __shfl_sync(0xffffffffu, myval, (j + u) & 31);
    v[u] = __ldg(xb + (size_t)cu * W4 + c);
  }
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    acc.x = fmaf(wv[u], v[u].x, acc.x); acc.y = fmaf(wv[u], v[u].y, acc.y);
    acc.z = fmaf(wv[u], v[u].z, acc.z); acc.w = fmaf(wv[u], v[u].w, acc.w);
  }
}

// Feature builder of the tensor-core path: the same warp-per-row gather, but the results are written
// period-major so that one (tile, period) of the cell kernels reads contiguous 32-byte rows:
//   Xt[t][q][F] = x[q][:, t]      St[t][q][F] = (A_hat x)[q][:, t]      Ut[t][b*nseg+s][F] = (L_hat_r x) per segment
// warps [0, B*N) build Xt/St for row q; warps [B*N, B*N + B*nseg) build Ut for segment (b, s).
__global__ void __launch_bounds__(256) k_feat_tc(const int32_t* __restrict__ g_rowptr, const int32_t* __restrict__ g_col,
                                                 const float* __restrict__ g_val, const int32_t* __restrict__ seg_eptr,
                                                 const int32_t* __restrict__ c_col, const float* __restrict__ c_val,
                                                 const float* __restrict__ x, int B, int N, int xN, int nseg, int T,
                                                 float* __restrict__ Xt, float* __restrict__ St, float* __restrict__ Ut) {
  extern __shared__ __align__(16) float feat_stage[];   // [warps per block][2][F*T]
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long BN = (long long)B * N, BS = (long long)B * nseg;
  if (warp >= BN + BS) return;
  float* stage = feat_stage + (threadIdx.x >> 5) * 2 * REGT_F * T;
  const bool is_seg = warp >= BN;
  const long long w = is_seg ? warp - BN : warp;
  const int b = (int)(w / (is_seg ? nseg : N)), r = (int)(w % (is_seg ? nseg : N));
  const int* rowptr = is_seg ? seg_eptr : g_rowptr;
  const int* col = is_seg ? c_col : g_col;
  const float* val = is_seg ? c_val : g_val;
  const int W = REGT_F * T, W4 = W >> 2;  // floats / float4 per row (W % 4 == 0 since F = 8)
  const float4* xb = reinterpret_cast<const float4*>(x) + (size_t)b * xN * W4;   // xN >= N: halo rows follow the owned ones
  const int e0 = rowptr[r], e1 = rowptr[r + 1];
  const long long plane_rows = is_seg ? BS : BN;
  float* dst = is_seg ? Ut : St;
  // lane c owns float4 #c of the F*T-wide row (coalesced 128-bit gathers).  The row's edge metadata is
  // read ONCE, 32 entries per coalesced load, and broadcast by shuffle, so that all neighbour gathers of
  // a batch are in flight together (two dependent memory rounds per row instead of one per 4 edges).
  for (int cb = 0; cb < W4; cb += 32) {   // warp-uniform trip count (the shuffles need every lane)
    const bool live = cb + lane < W4;
    const int c = live ? cb + lane : W4 - 1;          // idle lanes shadow the last float4 (loads stay unconditional)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 self = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!is_seg) self = __ldg(xb + (size_t)r * W4 + c);
    for (int eb = e0; eb < e1; eb += 32) {
      const int n = min(32, e1 - eb);
      const int mycol = __ldg(col + eb + min(lane, n - 1));
      const float myval = lane < n ? __ldg(val + eb + lane) : 0.f;   // weight 0 past the end of the row
      int j = 0;
      while (j < n) {   // batches of 8 (or a final 4): every gather of a batch is in flight at once
        if (n - j > 4) {
          gather_batch<8>(xb, W4, c, mycol, myval, j, acc);
          j += 8;
        } else {
          gather_batch<4>(xb, W4, c, mycol, myval, j, acc);
          j += 4;
        }
      }
    }
    // stage the row in shared memory so that the period-major write is one 32-byte row per lane
    if (live) {
      reinterpret_cast<float4*>(stage)[c] = acc;
      if (!is_seg) reinterpret_cast<float4*>(stage + W)[c] = self;
    }
  }
  __syncwarp();
  for (int t = lane; t < T; t += 32) {   // lane t gathers its 8 features (stride T) and writes 2 x float4
    float4 lo, hi;
    lo.x = stage[0 * T + t]; lo.y = stage[1 * T + t]; lo.z = stage[2 * T + t]; lo.w = stage[3 * T + t];
    hi.x = stage[4 * T + t]; hi.y = stage[5 * T + t]; hi.z = stage[6 * T + t]; hi.w = stage[7 * T + t];
    float4* d = reinterpret_cast<float4*>(dst + ((size_t)t * plane_rows + w) * REGT_F);
    d[0] = lo;
    d[1] = hi;
    if (!is_seg) {
      const float* sx = stage + W;
      lo.x = sx[0 * T + t]; lo.y = sx[1 * T + t]; lo.z = sx[2 * T + t]; lo.w = sx[3 * T + t];
      hi.x = sx[4 * T + t]; hi.y = sx[5 * T + t]; hi.z = sx[6 * T + t]; hi.w = sx[7 * T + t];
      float4* dx = reinterpret_cast<float4*>(Xt + ((size_t)t * BN + w) * REGT_F);
      dx[0] = lo;
      dx[1] = hi;
    }
  }
}

int launch_feat_tc(const regt_graph_plan& p, const float* x, int B, int xN, int T, float* Xt, float* St, float* Ut, cudaStream_t st) {
  const long long warps = (long long)B * p.N + (long long)B * p.nseg;
  k_feat_tc<<<cdiv(warps * 32, 256), 256, 8 * 2 * REGT_F * T * sizeof(float), st>>>(p.g_rowptr, p.g_col, p.g_val, p.seg_eptr, p.c_col, p.c_val, x, B, p.N,
                                                  xN, p.nseg, T, Xt, St, Ut);
  REGT_LAUNCHED("k_feat_tc", st);
  return 0;
}

int launch_spmm_rows(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, float* y, int B,
                     int n_out, int n_in, int width, cudaStream_t st) {
  REGT_CHECK(width % 4 == 0 && width > 0, "spmm: width %d must be a positive multiple of 4", width);
  if (B == 0 || n_out == 0) return 0;
  long long warps = (long long)B * n_out;
  k_spmm_rows<4><<<cdiv(warps * 32, 256), 256, 0, st>>>(rowptr, col, val, (const float4*)x, (float4*)y, B, n_out, n_in,
                                                       width / 4);
  REGT_LAUNCHED("k_spmm_rows", st);
  return 0;
}

// K4 regional gather / scatter of node rows (region shards: owned + halo rows in, owned rows out)
//   gather : dst[b][i][:] = src[b][idx[i]][:]   i < n_idx   (src has n_src rows, dst n_idx rows)
//   scatter: dst[b][idx[i]][:] = src[b][i][:]   i < n_idx   (src has n_idx rows, dst n_dst rows)
template <bool SCATTER, typename V>
__global__ void __launch_bounds__(256) k_move_rows(const V* __restrict__ src, const int64_t* __restrict__ idx,
                                                   V* __restrict__ dst, int n_idx, int n_other, int WV, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % WV);
  const long long br = i / WV;
  const int r = (int)(br % n_idx), b = (int)(br / n_idx);
  const size_t packed = ((size_t)b * n_idx + r) * WV + c;
  const size_t strided = ((size_t)b * n_other + (size_t)__ldg(idx + r)) * WV + c;
  if (SCATTER) dst[strided] = src[packed];
  else dst[packed] = __ldg(src + strided);
}

template <bool SCATTER>
static int launch_move_rows(const float* src, const int64_t* idx, float* dst, int B, int n_idx, int n_other, int width,
                            cudaStream_t st) {
  REGT_CHECK(src && idx && dst, "gather/scatter_rows: NULL pointer");
  REGT_CHECK(B >= 0 && n_idx >= 0 && n_other >= 0 && width > 0, "gather/scatter_rows: bad sizes");
  if (B == 0 || n_idx == 0) return 0;
  const bool v4 = width % 4 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  const int WV = v4 ? width / 4 : width;
  const long long total = (long long)B * n_idx * WV;
  if (v4) k_move_rows<SCATTER, float4><<<cdiv(total, 256), 256, 0, st>>>((const float4*)src, idx, (float4*)dst, n_idx, n_other, WV, total);
  else k_move_rows<SCATTER, float><<<cdiv(total, 256), 256, 0, st>>>(src, idx, dst, n_idx, n_other, WV, total);
  REGT_LAUNCHED(SCATTER ? "k_scatter_rows" : "k_gather_rows", st);
  return 0;
}

}  // namespace regt

extern "C" int regt_gather_rows(const float* src, const int64_t* idx, float* dst, int32_t B, int32_t n_src, int32_t n_idx,
                                int32_t width, regt_stream_t stream) {
  return regt::launch_move_rows<false>(src, idx, dst, B, n_idx, n_src, width, (cudaStream_t)stream);
}
extern "C" int regt_scatter_rows(const float* src, const int64_t* idx, float* dst, int32_t B, int32_t n_idx, int32_t n_dst,
                                 int32_t width, regt_stream_t stream) {
  return regt::launch_move_rows<true>(src, idx, dst, B, n_idx, n_dst, width, (cudaStream_t)stream);
}

extern "C" int regt_spmm_f8(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, float* y,
                            int32_t B, int32_t N, int32_t width, regt_stream_t stream) {
  return regt::launch_spmm_rows(rowptr, col, val, x, y, B, N, N, width, (cudaStream_t)stream);
}
