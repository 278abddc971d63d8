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
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

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

// Shared-memory staged variant (north_star: "warp-per-row CSR gather ... shared-memory staging of neighbour feature
// tiles").  A block of NB consecutive nodes of one snapshot is ONE contiguous run of NB * W floats in x[b]: it is pulled
// into shared memory with the bulk-copy engine (DRAM -> smem, no register round trip), then the block's rows are
// gathered from shared memory whenever the neighbour lies in the block (road graphs: regions are contiguous id ranges, so
// most neighbours do) and from global memory (L2) otherwise.  The warp-per-row kernel above reads every neighbour row
// through L2 -- 7.1 x 384 B per output row at config 5 against 768 B of DRAM traffic -- and runs L2-gather-bound; here
// the L2 side shrinks to the block load plus the out-of-block neighbours.  Same sequential CSR order and the same fmaf
// chain as k_spmm_rows: bit-identical results.
// v2 (this round): the block's CSR slice (rowptr, col, val -- contiguous, because the rows are) is staged too, so the
// edge loop has no dependent global loads (v1 read col/val per edge through L2: 2.67 ms at config 5, latency-bound, 0.28
// of the HBM peak), and two CTAs share an SM so that one's block load overlaps the other's gather.
constexpr int SPMM_BLK_THREADS = 512;
constexpr int SPMM_EMAX = 2304;          // staged edges per block (more: the tail is read from global memory)
// v3: the kernel was ISSUE-bound (ncu: 70 % issue-active, integer compares and address arithmetic around four FMAs per
// edge and lane).  The staging threads now resolve every edge ONCE per block into (float4 index of the neighbour row inside
// the staged block, or -1 - its global row) + weight, one 8-byte word; the gather loop is a broadcast LDS.64, a warp-uniform
// branch, one 128-bit load and four FMAs.
__global__ void __launch_bounds__(SPMM_BLK_THREADS, 2) k_spmm_blk(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                  const float* __restrict__ val, const float4* __restrict__ x,
                                                                  float4* __restrict__ y, int B, int n_out, int n_in, int W4, int NB,
                                                                  int nblk) {
  extern __shared__ __align__(128) uint8_t spmm_smem[];
  __shared__ uint64_t bar;
  float4* xs = reinterpret_cast<float4*>(spmm_smem);
  int32_t* s_ptr = reinterpret_cast<int32_t*>(spmm_smem + (size_t)NB * W4 * 16);   // [NB + 1] edge offsets relative to the block's first edge
  int2* s_edge = reinterpret_cast<int2*>(s_ptr + ((NB + 1 + 3) & ~3));              // [SPMM_EMAX] (resolved neighbour, weight bits)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  for (long long item = blockIdx.x; item < (long long)B * nblk; item += gridDim.x) {
    // items are ordered snapshot-major: the CTAs running side by side work on neighbouring node blocks of ONE snapshot, so
    // an out-of-block neighbour row is (still) in L2 because the CTA next door staged it
    const int b = (int)(item / nblk), blk = (int)(item - (long long)b * nblk);
    const int r0 = blk * NB, r1 = min(n_out, r0 + NB);
    const int s1 = min(n_in, r0 + NB);                       // staged node range [r0, s1)
    const float4* xb = x + (size_t)b * n_in * W4;
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)(s1 - r0) * W4 * 16;
      tc::mbar_arrive_expect_tx(&bar, bytes);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(xb + (size_t)r0 * W4);
      for (uint32_t o = 0; o < bytes; o += 32768) tc::bulk_g2s(spmm_smem + o, src + o, min(32768u, bytes - o), &bar);
    }
    const int eb = __ldg(rowptr + r0), ee = __ldg(rowptr + r1);
    for (int i = threadIdx.x; i <= r1 - r0; i += blockDim.x) s_ptr[i] = __ldg(rowptr + r0 + i) - eb;
    for (int i = threadIdx.x; i < min(ee - eb, SPMM_EMAX); i += blockDim.x) {
      const int j = __ldg(col + eb + i);
      s_edge[i] = make_int2((j >= r0 && j < s1) ? (j - r0) * W4 : -1 - j * W4, __float_as_int(__ldg(val + eb + i)));
    }
    __syncthreads();
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    for (int r = r0 + warp; r < r1; r += nwarp) {
      const int e0 = s_ptr[r - r0], e1 = s_ptr[r - r0 + 1];
      float4* yr = y + ((size_t)b * n_out + r) * W4;
      const int e1s = min(e1, SPMM_EMAX);
      for (int c = lane; c < W4; c += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* xc = xs + c;
        const float4* gc = xb + c;
#pragma unroll 2
        for (int e = e0; e < e1s; ++e) {      // sequential CSR order: deterministic sums
          const int2 ed = s_edge[e];          // one broadcast 8-byte load: resolved neighbour + weight
          const float w = __int_as_float(ed.y);
          float4 v;
          if (ed.x >= 0) v = xc[ed.x];        // warp-uniform: the neighbour row is staged
          else v = __ldg(gc + (-1 - ed.x));   // ... or lives in another block (L2)
          acc.x = fmaf(w, v.x, acc.x);
          acc.y = fmaf(w, v.y, acc.y);
          acc.z = fmaf(w, v.z, acc.z);
          acc.w = fmaf(w, v.w, acc.w);
        }
        for (int e = max(e0, SPMM_EMAX); e < e1; ++e) {   // edges beyond the staged slice (blocks with very many edges)
          const int j = __ldg(col + eb + e);
          const float w = __ldg(val + eb + e);
          const float4 v = (j >= r0 && j < s1) ? xc[(j - r0) * W4] : __ldg(gc + (size_t)j * W4);
          acc.x = fmaf(w, v.x, acc.x);
          acc.y = fmaf(w, v.y, acc.y);
          acc.z = fmaf(w, v.z, acc.z);
          acc.w = fmaf(w, v.w, acc.w);
        }
        yr[c] = acc;
      }
    }
    __syncthreads();     // every warp is done with the staged block before the next one lands
  }
}

// Feature builder of the tensor-core path: the results are written period-major so that one (tile, period) of the
// cell kernels reads contiguous 32-byte rows:
//   Xt[t][q][F] = x[q][:, t]      St[t][q][F] = (A_hat x)[q][:, t]      Ut[t][b*nseg+s][F] = (L_hat_r x) per segment
// One THREAD per (row, period), period fastest: the 12 period-threads of a row read a neighbour's whole 384-byte
// feature row between them (48-byte runs per feature), the edge metadata of a row is a broadcast load, and each
// thread writes its own 32-byte output row -- no shuffles, no staging, ~4x fewer instructions than the warp-per-row
// gather (which spent its time on issue slots and three dependent memory rounds at 37 % occupancy).  Sums stay in
// sequential CSR order (bit-identical to the stand-alone SpMM).
#ifndef REGT_FEAT_MINB
#define REGT_FEAT_MINB 4
#endif
__global__ void __launch_bounds__(256, REGT_FEAT_MINB) k_feat_tc(const int32_t* __restrict__ g_rowptr, const int32_t* __restrict__ g_col,
                                                 const float* __restrict__ g_val, const int32_t* __restrict__ seg_eptr,
                                                 const int32_t* __restrict__ c_col, const float* __restrict__ c_val,
                                                 const float* __restrict__ x, int B, int N, int xN, int nseg, int T,
                                                 float* __restrict__ Xt, float* __restrict__ St, float* __restrict__ Ut) {
  constexpr int F = REGT_F;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long BN = (long long)B * N, BS = (long long)B * nseg;
  const long long nA = BN * T;
  if (i >= nA + BS * T) return;
  const bool is_seg = i >= nA;
  const long long j = is_seg ? i - nA : i;
  const long long w = j / T;                       // b * N + r   (or b * nseg + s)
  const int t = (int)(j - w * T);
  const int per = is_seg ? nseg : N;
  const int b = (int)(w / per), r = (int)(w - (long long)b * per);
  const int* rowptr = is_seg ? seg_eptr : g_rowptr;
  const int* col = is_seg ? c_col : g_col;
  const float* val = is_seg ? c_val : g_val;
  const int W = F * T;
  const float* xb = x + (size_t)b * xN * W + t;    // element (node, f) of this sample and period: xb[node * W + f * T]
  const int e0 = __ldg(rowptr + r), e1 = __ldg(rowptr + r + 1);
  float acc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) acc[f] = 0.f;
  int e = e0;
  for (; e + 4 <= e1; e += 4) {   // four edges' loads in flight, accumulated in CSR order
    int c4[4];
    float v4[4], xv[4][F];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      c4[u] = __ldg(col + e + u);
      v4[u] = __ldg(val + e + u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* xr = xb + (size_t)c4[u] * W;
#pragma unroll
      for (int f = 0; f < F; ++f) xv[u][f] = __ldg(xr + f * T);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int f = 0; f < F; ++f) acc[f] = fmaf(v4[u], xv[u][f], acc[f]);
  }
  for (; e < e1; ++e) {
    const float v = __ldg(val + e);
    const float* xr = xb + (size_t)__ldg(col + e) * W;
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = fmaf(v, __ldg(xr + f * T), acc[f]);
  }
  float4* d = reinterpret_cast<float4*>((is_seg ? Ut : St) + ((size_t)t * (is_seg ? BS : BN) + w) * F);
  d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  if (!is_seg) {
    const float* xr = xb + (size_t)r * W;
    float sx[F];
#pragma unroll
    for (int f = 0; f < F; ++f) sx[f] = __ldg(xr + f * T);
    float4* dx = reinterpret_cast<float4*>(Xt + ((size_t)t * BN + w) * F);
    dx[0] = make_float4(sx[0], sx[1], sx[2], sx[3]);
    dx[1] = make_float4(sx[4], sx[5], sx[6], sx[7]);
  }
}

int launch_feat_tc(const regt_graph_plan& p, const float* x, int B, int xN, int T, float* Xt, float* St, float* Ut, cudaStream_t st) {
  const long long threads = ((long long)B * p.N + (long long)B * p.nseg) * T;
  k_feat_tc<<<cdiv(threads, 256), 256, 0, st>>>(p.g_rowptr, p.g_col, p.g_val, p.seg_eptr, p.c_col, p.c_val, x, B, p.N, xN, p.nseg, T, Xt,
                                                St, Ut);
  REGT_LAUNCHED("k_feat_tc", st);
  return 0;
}

int launch_spmm_rows(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, float* y, int B,
                     int n_out, int n_in, int width, cudaStream_t st) {
  REGT_CHECK(width % 4 == 0 && width > 0, "spmm: width %d must be a positive multiple of 4", width);
  if (B == 0 || n_out == 0) return 0;
  // staged variant: output row r <-> node r (square operator, possibly with halo columns behind the rows), blocks of
  // NB nodes = up to 192 KB of shared memory; small problems keep the plain warp-per-row kernel (less than one wave)
  static const bool no_blk = getenv("REGT_SPMM_PLAIN") && getenv("REGT_SPMM_PLAIN")[0] == '1';
  const int row_bytes = width * 4;
  const int meta_bytes = 2 * SPMM_EMAX * 4 + 64;
  // two CTAs per SM: each gets half of the 227 KB, minus the staged CSR slice
  int NB = ((227 * 1024) / 2 - 2048 - meta_bytes) / (row_bytes + 4) / 8 * 8;
  if (!no_blk && n_out <= n_in && NB >= 64 && (long long)B * n_out >= 4096 && ((uintptr_t)x % 16) == 0) {
    NB = min(NB, (n_out + 7) / 8 * 8);
    // balance the blocks of a snapshot (a short last block would idle an SM for most of a wave)
    const int nblk = cdiv(n_out, NB);
    NB = (cdiv(n_out, nblk) + 7) / 8 * 8;
    const size_t smem = (size_t)NB * row_bytes + (size_t)((NB + 1 + 3) & ~3) * 4 + 2 * SPMM_EMAX * 4;
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    REGT_CUDA(cudaFuncSetAttribute(k_spmm_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = (long long)B * nblk;
    k_spmm_blk<<<(int)min(items, 2ll * sms), SPMM_BLK_THREADS, smem, st>>>(rowptr, col, val, (const float4*)x, (float4*)y, B, n_out,
                                                                         n_in, width / 4, NB, nblk);
    REGT_LAUNCHED("k_spmm_blk", st);
    return 0;
  }
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
