// Unit test of the hand-built UMMA operand layouts / descriptors (tc_common.cuh): one CTA computes
// D[128 x N] from operands that CUDA-core threads wrote into shared memory, in each of the four
// layout roles the fused cell kernels use.  Driven by tests/test_gpu_umma.py through the C-ABI.
#include "common.cuh"
#include "tc_common.cuh"

namespace regt {
using namespace tc;

// variant 0: K-major SW128   A[128][K], B[N][K]
// variant 1: K-major chunk   A[128][K], B[N][K]          (K = one MMA k-step: 8 tf32 / 16 bf16)
// variant 2: MN-major SW128  A[K][128], B[K][N]          (contraction over the tile rows)
// variant 3: MN-major: A SW128 [K][128], B chunk tile [K][N]
// variant 4: MN-major SW128 with 32-byte swizzle atoms (descriptor layout 1)  A[K][128], B[K][N]   (32-bit elements), SBO = 1024
// variant 5: the same with SBO = 512: the 32-byte-atom swizzle repeats every FOUR 128-byte rows (CuTe's
//            Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o (1024 bits, 4 rows)), so consecutive K groups are 512 bytes apart
// variant 6 / 7: MN-major without swizzle ("interleaved"): tile [MN/16 B][K rows][16 B], i.e. the 16 bytes of four fp32
//            (eight bf16) consecutive MN indices of one K row are contiguous and the K rows of such a group are 16 bytes
//            apart -- one 8 x 16 B core matrix per k-step; the groups along MN are K*16 bytes apart.
//            6: LBO = K*16 (MN), SBO = 128 (K);  7: LBO = 128, SBO = K*16
template <int FMT>
__global__ void __launch_bounds__(128) k_umma_selftest(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ D, int variant, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int ES = (FMT == FMT_TF32) ? 4 : 2;
  constexpr int UK = (FMT == FMT_TF32) ? 8 : 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool mn = variant >= 2;
  const int a_rows = mn ? K : 128, a_cols = mn ? 128 : K;   // tile rows / extent along the 128-byte direction
  const int b_rows = mn ? K : N, b_cols = mn ? N : K;
  const bool a_chunk = (variant == 1), b_chunk = (variant == 1 || variant == 3);
  const bool b32 = (variant == 4 || variant == 5);
  const bool inter = (variant == 6 || variant == 7);
  uint8_t* As = smem;
  uint8_t* Bs = smem + 64 * 1024;

  auto put = [&](uint8_t* base, bool chunk, int rows, int r, int c, float v) {
    uint32_t off = chunk ? chunk_off(r, (c * ES) >> 4, rows) + ((c * ES) & 15)
                         : (b32 ? sw128b32_off(r, c * ES, rows) : sw128_off(r, c * ES, rows));
    if (inter) off = (uint32_t)(((c * ES) >> 4) * rows * 16 + r * 16 + ((c * ES) & 15));
    if constexpr (FMT == FMT_TF32) *reinterpret_cast<float*>(base + off) = v;
    else *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
  };
  for (int i = tid; i < a_rows * a_cols; i += 128) put(As, a_chunk, a_rows, i / a_cols, i % a_cols, A[i]);
  for (int i = tid; i < b_rows * b_cols; i += 128) put(Bs, b_chunk, b_rows, i / b_cols, i % b_cols, B[i]);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = make_idesc(FMT, 128, N, mn ? 1 : 0, mn ? 1 : 0);
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    const int ksteps = K / UK;
    for (int s = 0; s < ksteps; ++s) {
      uint64_t da, db;
      if (!mn) {
        const int kb = s * UK * ES;  // byte position along K
        da = a_chunk ? make_desc(a0 + s * 2 * a_rows * 16, a_rows * 16, 128, LAYOUT_NONE)
                     : make_desc(a0 + (kb >> 7) * a_rows * 128 + (kb & 127), 16, 1024, LAYOUT_SW128);
        db = b_chunk ? make_desc(b0 + s * 2 * b_rows * 16, b_rows * 16, 128, LAYOUT_NONE)
                     : make_desc(b0 + (kb >> 7) * b_rows * 128 + (kb & 127), 16, 1024, LAYOUT_SW128);
      } else if (inter) {
        const uint32_t mn_stride = (uint32_t)K * 16u, k_stride = 128u;
        const uint32_t lbo = variant == 6 ? mn_stride : k_stride, sbo = variant == 6 ? k_stride : mn_stride;
        da = make_desc(a0 + s * UK * 16, lbo, sbo, LAYOUT_NONE);
        db = make_desc(b0 + s * UK * 16, lbo, sbo, LAYOUT_NONE);
      } else {
        const uint32_t lay = b32 ? LAYOUT_SW128_B32 : LAYOUT_SW128;
        const uint32_t sbo = (variant == 5) ? 512 : 1024;
        da = make_desc(a0 + s * UK * 128, a_rows * 128, sbo, lay);
        db = b_chunk ? make_desc(b0 + s * UK * 16, 128, b_rows * 16, LAYOUT_NONE)
                     : make_desc(b0 + s * UK * 128, b_rows * 128, sbo, lay);
      }
      umma<FMT>(tmem, da, db, idesc, s > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}
}  // namespace regt

// debug entry point (not part of the reference-facing ABI): fmt 1 = bf16, 2 = tf32
extern "C" int regt_debug_umma_selftest(int fmt, int variant, const float* A, const float* B, float* D, int N, int K,
                                        regt_stream_t stream) {
  using namespace regt;
  cudaStream_t st = (cudaStream_t)stream;
  REGT_CHECK(fmt == 1 || fmt == 2, "selftest: fmt must be 1 (bf16) or 2 (tf32)");
  REGT_CHECK(variant >= 0 && variant <= 7 && N % 32 == 0 && N >= 32 && N <= 128, "selftest: bad variant/N");
  const size_t smem = 129 * 1024;
  if (fmt == 2) {
    REGT_CUDA(cudaFuncSetAttribute(k_umma_selftest<tc::FMT_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest<tc::FMT_TF32><<<1, 128, smem, st>>>(A, B, D, variant, N, K);
  } else {
    REGT_CUDA(cudaFuncSetAttribute(k_umma_selftest<tc::FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma_selftest<tc::FMT_BF16><<<1, 128, smem, st>>>(A, B, D, variant, N, K);
  }
  REGT_LAUNCH_CHECK();
  return 0;
}
