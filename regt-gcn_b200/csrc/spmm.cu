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

}  // namespace regt

extern "C" int regt_spmm_f8(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, float* y,
                            int32_t B, int32_t N, int32_t width, regt_stream_t stream) {
  return regt::launch_spmm_rows(rowptr, col, val, x, y, B, N, N, width, (cudaStream_t)stream);
}
