// Generic fp32 GEMMs on the sm_100a tensor cores with fp32-equivalent accuracy (3xTF32: every contraction is
// evaluated as hi*hi + lo*hi + hi*lo with a = hi + lo, hi = tf32(a)), for hidden widths the fused cell
// kernels are not built for.  Operands live in global memory as plain fp32 row-major arrays; the loader
// warps split them into tf32 (hi, lo) pairs and write the UMMA shared-memory layouts by hand
// (tc_common.cuh); accumulators live in TMEM.
//
//   gemm_nt : C[M][N]  = A[M][K] . Bt[N][K]^T             (K-major operands, SWIZZLE_128B)
//             used for the H x H gate contractions  [rows, H] x [H, 2H | H]  and the data gradients.
//   gemm_tn : Cp[s][M][N] = sum_{r in split s} A[r][M] . B[r][N]   (contraction over rows: the loaders transpose
//             into the same K-major tiles), split over the rows; the partials are summed by
//             k_reduce_splits.  Used for the weight gradients  D^T . [h | h*R | S | X | 1].
#include "cell_tc.cuh"

namespace regt {
using namespace tc;

namespace {
constexpr int GT_ROWS = 128;         // M tile = UMMA M = TMEM lanes
constexpr int GT_KC = 32;            // K chunk: 32 fp32 = one 128-byte swizzle row
constexpr int GT_NS = 3;             // pipeline stages
constexpr int GT_TILE = GT_ROWS * 128;               // one [128][128 B] operand tile (16 KB)
constexpr int GT_STAGE = 4 * GT_TILE;                // A hi | A lo | B hi | B lo
constexpr int GT_BTILE2 = (128 + 32) * 128;          // tn: B tile with the 32 extra feature columns (20 KB)
constexpr int GT_STAGE2 = 2 * GT_TILE + 2 * GT_BTILE2;
constexpr int GT_LOADERS = 256;                      // loader / epilogue threads: 8 warps, two per TMEM lane quarter
constexpr int GT_LW = GT_LOADERS / 32;               // loader warps; warp GT_LW issues the MMAs
constexpr int GT_THREADS = GT_LOADERS + 32;          // + the MMA issuer warp

__device__ __forceinline__ uint32_t tf32_rn(float a) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a));
  return u;
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(tf32_rn(v.x)); hi.y = __uint_as_float(tf32_rn(v.y));
  hi.z = __uint_as_float(tf32_rn(v.z)); hi.w = __uint_as_float(tf32_rn(v.w));
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}

struct GemmArgs {
  const float *A, *B;
  float* C;
  long long M;          // nt: rows of A / C.   tn: rows contracted over
  int N, K;             // nt: C is [M][N], K = contraction.   tn: C is [K][N] (K = columns of A used as output rows)
  long long lda, ldb, ldc;
  long long chunk;      // tn: rows per split
  // tn only: a second, 32-column B operand (the F-wide feature plane) contracted in the same pass over A;
  // its output goes to C2[split][K][32].  Attached to N tile 0.
  const float* B2;
  float* C2;
  long long ldb2;
  int N2;
  long long c2_split;   // elements between the splits of C2 (its rows may be a window of a taller array)
  int relu_b;           // tn: B is read through max(., 0)  (the head's relu(out_hidden))
};
}  // namespace

// ------------------------------------------------------------------------------------------
// C[M][N] = A[M][K] . Bt[N][K]^T      persistent: grid = min(#tiles, #SMs), tile = 128 rows x (<=128) columns.
// Warps 0-7 load (and split) operands, warp 8 issues the MMAs, warps 9-12 drain the accumulators: two TMEM
// accumulators alternate, so the epilogue of tile i runs under the main loop of tile i+1, and the operand
// pipeline never drains between tiles (K is only H here: a tile is 2-8 chunks).
// ------------------------------------------------------------------------------------------
constexpr int GT_NT_THREADS = GT_LOADERS + 32 + 128;
__global__ void __launch_bounds__(GT_NT_THREADS, 1) k_gemm_nt_tf32x3(GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_full[GT_NS], bar_empty[GT_NS], bar_acc_full[2], bar_acc_free[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntn = (a.N + 127) / 128;
  const long long ntiles = ((a.M + GT_ROWS - 1) / GT_ROWS) * ntn;
  const int nchunks = (a.K + GT_KC - 1) / GT_KC;
  if (tid == 0) {
    for (int s = 0; s < GT_NS; ++s) {
      mbar_init(&bar_full[s], GT_LOADERS);
      mbar_init(&bar_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_acc_full[b], 1);
      mbar_init(&bar_acc_free[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == GT_LW) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp < GT_LW) {
    // ---- loaders: thread (r, half) owns 64 bytes (4 float4) of row r of the A tile and of the Bt tile.
    //      The global loads of chunk kc+1 are issued before chunk kc is written to shared memory. ----
    const int r = tid & 127, half = tid >> 7;
    long long gc = 0;   // chunks issued so far (stage ring position, continues across tiles)
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long m0 = (tile / ntn) * GT_ROWS;
      const int n0 = (int)(tile % ntn) * 128;
      const int nt = min(128, a.N - n0);
      const long long arow = m0 + r;
      const bool a_ok = arow < a.M, b_ok = r < nt;
      const float* ap = a.A + (a_ok ? arow : 0) * a.lda + half * 16;
      const float* bp = a.B + (size_t)(b_ok ? n0 + r : 0) * a.ldb + half * 16;
      float4 av[4], bv[4], an[4], bn[4];
      auto load = [&](int kc, float4 (&x)[4], float4 (&y)[4]) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = kc * GT_KC + half * 16 + 4 * c;
          x[c] = (a_ok && k < a.K) ? __ldg(reinterpret_cast<const float4*>(ap + kc * GT_KC + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          y[c] = (b_ok && k < a.K) ? __ldg(reinterpret_cast<const float4*>(bp + kc * GT_KC + 4 * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load(0, av, bv);
      for (int kc = 0; kc < nchunks; ++kc, ++gc) {
        const int s = (int)(gc % GT_NS);
        if (kc + 1 < nchunks) load(kc + 1, an, bn);
        if (gc >= GT_NS) mbar_wait(&bar_empty[s], (uint32_t)((gc / GT_NS - 1) & 1));
        uint8_t* st = sm + s * GT_STAGE;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 hi, lo;
          const uint32_t off = sw128_off(r, (half * 4 + c) * 16, GT_ROWS);
          split4(av[c], hi, lo);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + GT_TILE + off) = lo;
          split4(bv[c], hi, lo);
          *reinterpret_cast<float4*>(st + 2 * GT_TILE + off) = hi;
          *reinterpret_cast<float4*>(st + 3 * GT_TILE + off) = lo;
        }
        fence_proxy_async();
        mbar_arrive(&bar_full[s]);
#pragma unroll
        for (int c = 0; c < 4; ++c) { av[c] = an[c]; bv[c] = bn[c]; }
      }
    }
  } else if (warp == GT_LW) {
    // ---- MMA issuer ----
    const uint32_t base = smem_u32(sm);
    long long gc = 0, li = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++li) {
      const int nt = min(128, a.N - (int)(tile % ntn) * 128);
      const uint32_t idesc = make_idesc(FMT_TF32, 128, nt, 0, 0);
      const int buf = (int)(li & 1);
      if (li >= 2) {   // the epilogue has drained this accumulator (its use number li/2 - 1)
        mbar_wait(&bar_acc_free[buf], (uint32_t)((li / 2 - 1) & 1));
        tc_fence_after();
      }
      for (int kc = 0; kc < nchunks; ++kc, ++gc) {
        const int s = (int)(gc % GT_NS);
        mbar_wait(&bar_full[s], (uint32_t)((gc / GT_NS) & 1));
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = base + s * GT_STAGE;
#pragma unroll
          for (int p = 0; p < 3; ++p) {   // hi*hi, lo*hi, hi*lo
            const uint32_t at = st + (p == 1 ? GT_TILE : 0), bt = st + 2 * GT_TILE + (p == 2 ? GT_TILE : 0);
#pragma unroll
            for (int k = 0; k < GT_KC / 8; ++k)
              umma<FMT_TF32>(tmem + buf * 128, make_desc(at + k * 32, 16, 1024, LAYOUT_SW128),
                             make_desc(bt + k * 32, 16, 1024, LAYOUT_SW128), idesc, (kc > 0 || p > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&bar_empty[s]);
          if (kc + 1 == nchunks) umma_commit(&bar_acc_full[buf]);
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  } else {
    // ---- epilogue warps: TMEM -> registers -> C (thread = row of the tile) ----
    const int r = (warp & 3) * 32 + lane;     // a warp may only touch TMEM lanes 32*(warp%4) ..
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    long long li = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++li) {
      const long long m0 = (tile / ntn) * GT_ROWS;
      const int n0 = (int)(tile % ntn) * 128;
      const int nt = min(128, a.N - n0);
      const int buf = (int)(li & 1);
      const long long arow = m0 + r;
      const bool a_ok = arow < a.M;
      float* cp = a.C + (a_ok ? arow : 0) * a.ldc + n0;
      mbar_wait(&bar_acc_full[buf], (uint32_t)((li / 2) & 1));
      tc_fence_after();
      for (int c0 = 0; c0 < nt; c0 += 32) {
        float v[32];
        if (c0 + 32 <= nt) {
          tmem_ld32(tlane + buf * 128 + c0, v);
        } else {   // nt is a multiple of 16: a final half group
          float u[16];
          tmem_ld16(tlane + buf * 128 + c0, u);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = u[j];
        }
        if (a_ok) {
          const int w = min(32, nt - c0);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < w) *reinterpret_cast<float4*>(cp + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_acc_free[buf]);
    }
  }
  __syncthreads();
  if (warp == GT_LW) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------
// Cp[z][K][N] = sum_{r in rows of split z} A[r][k0 .. k0+128) ^T . B[r][n0 .. n0+128)
// grid (ceil(K/128), ceil(N/128), splits)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GT_THREADS, 1) k_gemm_tn_tf32x3(GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_full[GT_NS], bar_empty[GT_NS], bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int kt = min(128, a.K - k0), nt = max(0, min(128, a.N - n0));     // output tile: kt rows (multiple of 32) x nt cols
  const int n2 = (blockIdx.y == 0) ? a.N2 : 0;                           // + the 32 feature columns on N tile 0
  const int ntot = nt + n2;
  const long long r0 = (long long)blockIdx.z * a.chunk, r1 = min(a.M, r0 + a.chunk);
  const int nchunks = (int)((r1 - r0 + GT_KC - 1) / GT_KC);
  if (tid == 0) {
    for (int s = 0; s < GT_NS; ++s) {
      mbar_init(&bar_full[s], GT_LOADERS);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == GT_LW) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // The contraction runs over the ROWS of A and B, so the loaders transpose on the way into shared memory:
  // element (row r of the chunk, column m) goes to K-major tile row m, K position r.  The tiles are then
  // ordinary K-major SWIZZLE_128B operands ([128 columns][32 rows of the chunk = 128 B]), exactly as in gemm_nt.
  if (warp < GT_LW) {
    // ---- loaders: warp (w & 3) owns column block w & 3 (32 floats = 128 B of every row), its half w >> 2 the
    //      float4s 4*half .. 4*half+3 of that block; lane = row of the chunk.  A lane's 4-byte stores of one
    //      (c, e) hit 32 different banks (the swizzle spreads lane >> 2).  Loads run one chunk ahead. ----
    const int blk = warp & 3, half = warp >> 2;
    const bool a_ok = blk * 32 < kt, b_ok = blk * 32 < nt;
    float4 av[4], bv[4], an[4], bn[4];
    float4 fv = make_float4(0.f, 0.f, 0.f, 0.f), fn = fv;   // this warp's float4 of the feature row (cols 4*warp ..)
    auto load = [&](int kc, float4 (&x)[4], float4 (&y)[4], float4& f) {
      const long long r = r0 + (long long)kc * GT_KC + lane;
      const bool r_ok = r < r1;
      const float4* ap = reinterpret_cast<const float4*>(a.A + (r_ok ? r : 0) * a.lda + k0 + blk * 32) + half * 4;
      const float4* bp = reinterpret_cast<const float4*>(a.B + (r_ok ? r : 0) * a.ldb + n0 + blk * 32) + half * 4;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        x[c] = (r_ok && a_ok) ? __ldg(ap + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        y[c] = (r_ok && b_ok && blk * 32 + 16 * half + 4 * c < nt) ? __ldg(bp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.relu_b) y[c] = make_float4(fmaxf(y[c].x, 0.f), fmaxf(y[c].y, 0.f), fmaxf(y[c].z, 0.f), fmaxf(y[c].w, 0.f));
      }
      if (n2) f = r_ok ? __ldg(reinterpret_cast<const float4*>(a.B2 + r * a.ldb2) + warp) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (nchunks > 0) load(0, av, bv, fv);
    for (int kc = 0; kc < nchunks; ++kc) {
      const int s = kc % GT_NS;
      if (kc + 1 < nchunks) load(kc + 1, an, bn, fn);
      if (kc >= GT_NS) mbar_wait(&bar_empty[s], (uint32_t)((kc / GT_NS - 1) & 1));
      uint8_t* st = sm + s * GT_STAGE2;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float4 hi, lo;
        split4(av[c], hi, lo);
        const float ah[4] = {hi.x, hi.y, hi.z, hi.w}, al[4] = {lo.x, lo.y, lo.z, lo.w};
        split4(bv[c], hi, lo);
        const float bh[4] = {hi.x, hi.y, hi.z, hi.w}, bl[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t off = sw128_off(blk * 32 + 16 * half + 4 * c + e, lane * 4, GT_ROWS);
          *reinterpret_cast<float*>(st + off) = ah[e];
          *reinterpret_cast<float*>(st + GT_TILE + off) = al[e];
          if (blk * 32 + 16 * half + 4 * c + e < nt) {   // rows nt.. of the B tile belong to the feature columns
            *reinterpret_cast<float*>(st + 2 * GT_TILE + off) = bh[e];
            *reinterpret_cast<float*>(st + 2 * GT_TILE + GT_BTILE2 + off) = bl[e];
          }
        }
      }
      if (n2) {   // feature columns: B tile rows nt + 4*warp + e
        float4 hi, lo;
        split4(fv, hi, lo);
        const float fh[4] = {hi.x, hi.y, hi.z, hi.w}, fl[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t off = sw128_off(nt + 4 * warp + e, lane * 4, GT_ROWS);
          *reinterpret_cast<float*>(st + 2 * GT_TILE + off) = fh[e];
          *reinterpret_cast<float*>(st + 2 * GT_TILE + GT_BTILE2 + off) = fl[e];
        }
      }
      fence_proxy_async();
      mbar_arrive(&bar_full[s]);
#pragma unroll
      for (int c = 0; c < 4; ++c) { av[c] = an[c]; bv[c] = bn[c]; }
      fv = fn;
    }
    // ---- epilogue: partial tile -> Cp[z] (thread = output row; the two warps of a lane quarter alternate 16-column groups) ----
    if (nchunks > 0) {   // an empty split (more splits than row chunks) contributes zeros
      mbar_wait(&bar_done, 0);
      tc_fence_after();
    }
    const int orow = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const bool row_ok = orow < kt;
    float* cp = a.C + ((size_t)blockIdx.z * a.K + k0 + (row_ok ? orow : 0)) * a.ldc + n0;
    float* cp2 = n2 ? a.C2 + (size_t)blockIdx.z * a.c2_split + (size_t)(k0 + (row_ok ? orow : 0)) * 32 : nullptr;
    for (int c0 = half * 16; c0 < ntot; c0 += 32) {
      float v[16];
      if (nchunks > 0) {
        tmem_ld16(tlane + c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
      if (row_ok) {
        float* o = c0 < nt ? cp + c0 : cp2 + (c0 - nt);
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    tc_fence_before();
  } else {
    const uint32_t idesc = make_idesc(FMT_TF32, 128, ntot, 0, 0);
    const uint32_t base = smem_u32(sm);
    for (int kc = 0; kc < nchunks; ++kc) {
      const int s = kc % GT_NS;
      mbar_wait(&bar_full[s], (uint32_t)((kc / GT_NS) & 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = base + s * GT_STAGE2;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint32_t at = st + (p == 1 ? GT_TILE : 0), bt = st + 2 * GT_TILE + (p == 2 ? GT_BTILE2 : 0);
#pragma unroll
          for (int k = 0; k < GT_KC / 8; ++k)
            umma<FMT_TF32>(tmem, make_desc(at + k * 32, 16, 1024, LAYOUT_SW128), make_desc(bt + k * 32, 16, 1024, LAYOUT_SW128),
                           idesc, (kc > 0 || p > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&bar_empty[s]);
        if (kc + 1 == nchunks) umma_commit(&bar_done);
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == GT_LW) tmem_dealloc(tmem, 256);
}

static int gemm_check(const float* A, const float* B, const float* C, long long lda, long long ldb, long long ldc, const char* who) {
  REGT_CHECK(A && B && C, "%s: NULL operand", who);
  REGT_CHECK(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "%s: leading dimensions must be multiples of 4 floats", who);
  REGT_CHECK(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0), "%s: operands must be 16-byte aligned", who);
  return 0;
}

// C[M][N] = A[M][K] . Bt[N][K]^T ; N multiple of 16, K multiple of 4
int launch_gemm_nt_tf32x3(const float* A, long long lda, const float* Bt, long long ldb, float* C, long long ldc, long long M,
                          int N, int K, cudaStream_t st) {
  if (gemm_check(A, Bt, C, lda, ldb, ldc, "gemm_nt")) return -1;
  REGT_CHECK(N % 16 == 0 && K % 4 == 0 && N > 0 && K > 0, "gemm_nt: N=%d must be a multiple of 16 and K=%d of 4", N, K);
  if (M == 0) return 0;
  GemmArgs a{A, Bt, C, M, N, K, lda, ldb, ldc, 0, nullptr, nullptr, 0, 0, 0, 0};
  const size_t smem = (size_t)GT_NS * GT_STAGE + 1024;
  REGT_CUDA(cudaFuncSetAttribute(k_gemm_nt_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  const long long ntiles = (long long)cdiv(M, GT_ROWS) * cdiv(N, 128);
  k_gemm_nt_tf32x3<<<(int)min(ntiles, (long long)sms), GT_NT_THREADS, smem, st>>>(a);
  REGT_LAUNCHED("k_gemm_nt_tf32x3", st);
  return 0;
}

// Cp[z][K][N] (z < splits) = partial sums over row chunks of A[M][K]^T . B[M][N] ; K multiple of 32, N of 16.
// Optional second operand B2 [M][32] (ldb2): Cp2[z][K][32] = partials of A^T . B2 in the same pass (N may then be 0).
int launch_gemm_tn_tf32x3(const float* A, long long lda, const float* B, long long ldb, float* Cp, long long M, int K, int N,
                          int splits, cudaStream_t st, const float* B2 = nullptr, long long ldb2 = 0, float* Cp2 = nullptr,
                          long long c2_split = 0, int relu_b = 0) {
  if (N > 0 && gemm_check(A, B, Cp, lda, ldb, N, "gemm_tn")) return -1;
  REGT_CHECK(K % 32 == 0 && N % 16 == 0 && K > 0 && N >= 0 && splits > 0, "gemm_tn: K=%d must be a multiple of 32 and N=%d of 16", K, N);
  REGT_CHECK(N > 0 || B2, "gemm_tn: nothing to contract with");
  REGT_CHECK(!B2 || (Cp2 && ldb2 % 4 == 0 && (uintptr_t)B2 % 16 == 0 && (uintptr_t)Cp2 % 16 == 0), "gemm_tn: bad second operand");
  long long chunk = (M + splits - 1) / splits;
  chunk = (chunk + GT_KC - 1) / GT_KC * GT_KC;
  GemmArgs a{A, N > 0 ? B : A, Cp, M, N, K, lda, N > 0 ? ldb : lda, N, chunk, B2, Cp2, ldb2, B2 ? 32 : 0, c2_split > 0 ? c2_split : (long long)K * 32, relu_b};
  const size_t smem = (size_t)GT_NS * GT_STAGE2 + 1024;
  REGT_CUDA(cudaFuncSetAttribute(k_gemm_tn_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gemm_tn_tf32x3<<<dim3(cdiv(K, 128), max(1, cdiv(N, 128)), splits), GT_THREADS, smem, st>>>(a);
  REGT_LAUNCHED("k_gemm_tn_tf32x3", st);
  return 0;
}

}  // namespace regt

// debug entry points (not part of the reference-facing ABI): tests/test_gpu_gemm.py
extern "C" int regt_debug_gemm_nt(const float* A, int64_t lda, const float* Bt, int64_t ldb, float* C, int64_t ldc, int64_t M,
                                  int32_t N, int32_t K, regt_stream_t stream) {
  return regt::launch_gemm_nt_tf32x3(A, lda, Bt, ldb, C, ldc, M, N, K, (cudaStream_t)stream);
}
extern "C" int regt_debug_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* Cp, int64_t M, int32_t K,
                                  int32_t N, int32_t splits, regt_stream_t stream) {
  return regt::launch_gemm_tn_tf32x3(A, lda, B, ldb, Cp, M, K, N, splits, (cudaStream_t)stream);
}
extern "C" int regt_debug_gemm_tn2(const float* A, int64_t lda, const float* B, int64_t ldb, float* Cp, int64_t M, int32_t K,
                                   int32_t N, int32_t splits, const float* B2, int64_t ldb2, float* Cp2, regt_stream_t stream) {
  return regt::launch_gemm_tn_tf32x3(A, lda, B, ldb, Cp, M, K, N, splits, (cudaStream_t)stream, B2, ldb2, Cp2);
}
