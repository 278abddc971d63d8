// tcgen05 (sm_100a tensor core) kernels of the regional temporal GCN cell.
//
// One persistent CTA per SM.  A work item is 128 rows (b,n) x a chunk of `tp` periods; thread r of
// the 256 epilogue threads owns row r (= TMEM lane r) and one half of the H columns, and walks the
// periods of its item, so the period-attention sum  sum_t probs[t] * H'_t  stays in registers.
// Per period:
//   P   CUDA cores : h = act(X_t M0 + U_t M1[r] + c0)  -> bf16 / tf32(hi,lo) SW128 operand tile in smem
//   M1  tcgen05.mma: [S_t | h] x Wzr  -> TMEM cols [0,2H)          (weights resident in smem)
//   E1  CUDA cores : Z,R = sigmoid(. + czr) (tcgen05.ld), save Z,R; h*R -> operand tile
//   M2  tcgen05.mma: [S_t | h*R] x Wc -> TMEM cols [2H,3H)
//   E2  CUDA cores : H~ = tanh(. + cc), save H~;  acc += probs[t] * (Z h + (1-Z) H~)
// Precision: REGT_PREC_BF16 = bf16 operands, fp32 accumulate; REGT_PREC_TF32X3 = three tf32 products
// (hi*hi + lo*hi + hi*lo) per contraction, fp32-equivalent accuracy.
// Reference arithmetic replaced: models/utils.py:168-188 + models/RegionalTemporalGCN.py:134-148.
#include "cell_tc.cuh"

namespace regt {
using namespace tc;
constexpr int F = REGT_F;
constexpr int NEPI = 256;       // epilogue threads (8 warps); warp 8 issues the MMAs
constexpr int NTHREADS = NEPI + 32;

__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float fast_tanh(float v) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v)); }
__device__ __forceinline__ uint32_t tf32_rn_bits(float a) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a));
  return u;
}

// ------------------------------------------------------------------------------------------
// weight image: fp32 collapsed weights (k_prep) -> swizzled operand tiles
// ------------------------------------------------------------------------------------------
template <int FMT, int HH>
__device__ __forceinline__ void put_w(uint8_t* tile0, int split_stride, bool chunk_tile, int rows, int r, int c, float v) {
  using Cfg = TcCfg<FMT, HH>;
  const uint32_t off = chunk_tile ? chunk_off(r, (c * Cfg::ES) >> 4, rows) + ((c * Cfg::ES) & 15)
                                  : sw128_off(r, c * Cfg::ES, rows);
  if constexpr (FMT == FMT_TF32) {
    const float hi = __uint_as_float(tf32_rn_bits(v));
    *reinterpret_cast<float*>(tile0 + off) = hi;
    *reinterpret_cast<float*>(tile0 + split_stride + off) = v - hi;
  } else {
    *reinterpret_cast<__nv_bfloat16*>(tile0 + off) = __float2bfloat16(v);
  }
}

// Wzr [F+H][2H], Wc [F+H][H] (rows 0..F-1: S part, rows F..: h part), lin_w[g] [H][2H]
template <int FMT, int HH>
__global__ void k_pack_tc(const float* __restrict__ Wzr, const float* __restrict__ Wc, const float* __restrict__ czr,
                          const float* __restrict__ cc, const float* __restrict__ c0, const float* __restrict__ M0t,
                          const float* __restrict__ M1t, const float* __restrict__ probs, int T,
                          const float* __restrict__ lw0, const float* __restrict__ lw1, const float* __restrict__ lw2,
                          uint8_t* __restrict__ img_f, uint8_t* __restrict__ img_b) {
  using Cfg = TcCfg<FMT, HH>;
  constexpr int SKE = 16 * 2 / Cfg::ES;  // elements of the padded S part (8 tf32 / 16 bf16)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int n_zr_h = 2 * HH * HH, n_zr_s = 2 * HH * SKE, n_c_h = HH * HH, n_c_s = HH * SKE;
  int j = i;
  if (j < n_zr_h) {
    const int n = j / HH, k = j % HH;
    put_w<FMT, HH>(img_f, Cfg::FWD_SPLIT, false, 2 * HH, n, k, Wzr[(size_t)(F + k) * 2 * HH + n]);
    return;
  }
  j -= n_zr_h;
  if (j < n_zr_s) {
    const int n = j / SKE, f = j % SKE;
    put_w<FMT, HH>(img_f + Cfg::WZR_H, Cfg::FWD_SPLIT, true, 2 * HH, n, f, f < F ? Wzr[(size_t)f * 2 * HH + n] : 0.f);
    return;
  }
  j -= n_zr_s;
  if (j < n_c_h) {
    const int n = j / HH, k = j % HH;
    put_w<FMT, HH>(img_f + Cfg::WZR_H + Cfg::WZR_S, Cfg::FWD_SPLIT, false, HH, n, k, Wc[(size_t)(F + k) * HH + n]);
    return;
  }
  j -= n_c_h;
  if (j < n_c_s) {
    const int n = j / SKE, f = j % SKE;
    put_w<FMT, HH>(img_f + Cfg::WZR_H + Cfg::WZR_S + Cfg::WC_H, Cfg::FWD_SPLIT, true, HH, n, f,
                   f < F ? Wc[(size_t)f * HH + n] : 0.f);
    return;
  }
  j -= n_c_s;
  if (j < Cfg::C_FLOATS) {
    float v = 0.f;
    if (j < Cfg::C_CC) v = czr[j];
    else if (j < Cfg::C_C0) v = cc[j - Cfg::C_CC];
    else if (j < Cfg::C_M0) v = c0 ? c0[j - Cfg::C_C0] : 0.f;
    else if (j < Cfg::C_M1) v = M0t ? M0t[j - Cfg::C_M0] : 0.f;
    else if (j < Cfg::C_PROBS) v = M1t ? M1t[j - Cfg::C_M1] : 0.f;
    else v = (j - Cfg::C_PROBS) < T ? probs[j - Cfg::C_PROBS] : 0.f;
    reinterpret_cast<float*>(img_f + Cfg::FWD_W)[j] = v;
    reinterpret_cast<float*>(img_b + Cfg::BWD_W)[j] = v;
    return;
  }
  j -= Cfg::C_FLOATS;
  if (j < 3 * HH * HH) {  // backward: Bt_g[k][n] = linear_g.weight[n][HH + k]
    const int g = j / (HH * HH), rem = j % (HH * HH);
    const int k = rem / HH, n = rem % HH;
    const float* lw = g == 0 ? lw0 : (g == 1 ? lw1 : lw2);
    put_w<FMT, HH>(img_b + g * Cfg::BT, 3 * Cfg::BT, false, HH, k, n, lw[(size_t)n * 2 * HH + HH + k]);
  }
}

// ------------------------------------------------------------------------------------------
// shared device helpers
// ------------------------------------------------------------------------------------------
// write 32 consecutive columns [c0, c0+32) of row r into the [128 x HH] SW128 operand tile(s)
template <int FMT, int HH>
__device__ __forceinline__ void store_operand32(uint8_t* tile, int r, int c0, const float (&v)[32]) {
  using Cfg = TcCfg<FMT, HH>;
  if constexpr (FMT == FMT_TF32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = __uint_as_float(tf32_rn_bits(v[j + e]));
        lo[e] = v[j + e] - hi[e];
      }
      const uint32_t off = sw128_off(r, (c0 + j) * 4, TC_ROWS);
      *reinterpret_cast<float4*>(tile + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(tile + Cfg::A_TILE + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 p;
      p.x = pack_bf16(v[j], v[j + 1]);
      p.y = pack_bf16(v[j + 2], v[j + 3]);
      p.z = pack_bf16(v[j + 4], v[j + 5]);
      p.w = pack_bf16(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(tile + sw128_off(r, (c0 + j) * 2, TC_ROWS)) = p;
    }
  }
}
// write the 8 F-wide values of row r into chunk tile(s): chunk c holds 16 bytes
template <int FMT, int HH>
__device__ __forceinline__ void store_small8(uint8_t* tile, int split_stride, int r, int elem0, const float (&v)[8]) {
  if constexpr (FMT == FMT_TF32) {  // 8 fp32 = 2 chunks, starting at chunk elem0/4
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      hi[e] = __uint_as_float(tf32_rn_bits(v[e]));
      lo[e] = v[e] - hi[e];
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t off = chunk_off(r, elem0 / 4 + c, TC_ROWS);
      *reinterpret_cast<float4*>(tile + off) = make_float4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
      *reinterpret_cast<float4*>(tile + split_stride + off) = make_float4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
    }
  } else {  // 8 bf16 = 1 chunk
    uint4 p;
    p.x = pack_bf16(v[0], v[1]);
    p.y = pack_bf16(v[2], v[3]);
    p.z = pack_bf16(v[4], v[5]);
    p.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + chunk_off(r, elem0 / 8, TC_ROWS)) = p;
  }
}

// one contraction D[128 x N] (+)= [h-part | S-part] x W  with all precision products
template <int FMT, int HH>
__device__ __forceinline__ void issue_gate_mma(uint32_t tmem_d, uint32_t a_h, uint32_t a_s, uint32_t w_h, uint32_t w_s,
                                               int n_rows_w, uint32_t idesc) {
  using Cfg = TcCfg<FMT, HH>;
  uint32_t acc = 0;
  constexpr int NPROD = (FMT == FMT_TF32) ? 3 : 1;
#pragma unroll
  for (int p = 0; p < NPROD; ++p) {
    const int pa = (p == 1) ? 1 : 0, pb = (p == 2) ? 1 : 0;  // hi*hi, lo*hi, hi*lo
    const uint32_t ah = a_h + pa * Cfg::A_TILE, as = a_s + pa * Cfg::AS_TILE;
    const uint32_t wh = w_h + pb * Cfg::FWD_SPLIT, ws = w_s + pb * Cfg::FWD_SPLIT;
#pragma unroll
    for (int s = 0; s < HH / Cfg::UK; ++s) {
      const int kb = s * 32;
      const uint64_t da = make_desc(ah + (kb >> 7) * TC_ROWS * 128 + (kb & 127), 16, 1024, LAYOUT_SW128);
      const uint64_t db = make_desc(wh + (kb >> 7) * n_rows_w * 128 + (kb & 127), 16, 1024, LAYOUT_SW128);
      umma<FMT>(tmem_d, da, db, idesc, acc);
      acc = 1;
    }
    const uint64_t da = make_desc(as, TC_ROWS * 16, 128, LAYOUT_NONE);
    const uint64_t db = make_desc(ws, n_rows_w * 16, 128, LAYOUT_NONE);
    umma<FMT>(tmem_d, da, db, idesc, 1);
  }
}

// h[32] for columns [c0, c0+32) of row q at period t (regional combine on F-wide features)
template <int HH>
__device__ __forceinline__ void compute_h32(const TcArgs& a, const float* consts_s, bool valid, long long q, int b, int s0,
                                            int s1, int t, int c0, float (&h)[32], float (&sv)[8]) {
  using C = TcCfg<FMT_BF16, HH>;  // constant offsets do not depend on FMT
  float xv[8];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    xv[f] = valid ? __ldg(a.x + q * F * a.T + f * a.T + t) : 0.f;
    sv[f] = valid ? __ldg(a.S + q * F * a.T + f * a.T + t) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) h[j] = consts_s[C::C_C0 + c0 + j];
#pragma unroll
  for (int f = 0; f < F; ++f) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 w = *reinterpret_cast<const float4*>(consts_s + C::C_M0 + f * HH + c0 + j);
      h[j] = fmaf(xv[f], w.x, h[j]);
      h[j + 1] = fmaf(xv[f], w.y, h[j + 1]);
      h[j + 2] = fmaf(xv[f], w.z, h[j + 2]);
      h[j + 3] = fmaf(xv[f], w.w, h[j + 3]);
    }
  }
  for (int s = s0; s < s1; ++s) {
    const int reg = a.seg_reg[s];
    const float* ur = a.U + ((size_t)b * a.nseg + s) * F * a.T + t;
    float uv[8];
#pragma unroll
    for (int f = 0; f < F; ++f) uv[f] = __ldg(ur + f * a.T);
    if (reg == 0) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 w = *reinterpret_cast<const float4*>(consts_s + C::C_M1 + f * HH + c0 + j);
          h[j] = fmaf(uv[f], w.x, h[j]);
          h[j + 1] = fmaf(uv[f], w.y, h[j + 1]);
          h[j + 2] = fmaf(uv[f], w.z, h[j + 2]);
          h[j + 3] = fmaf(uv[f], w.w, h[j + 3]);
        }
      }
    } else {
      const float* m = a.M1t + (size_t)reg * F * HH + c0;
#pragma unroll
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(m + f * HH + j));
          h[j] = fmaf(uv[f], w.x, h[j]);
          h[j + 1] = fmaf(uv[f], w.y, h[j + 1]);
          h[j + 2] = fmaf(uv[f], w.z, h[j + 2]);
          h[j + 3] = fmaf(uv[f], w.w, h[j + 3]);
        }
      }
    }
  }
  if (a.mode == REGT_MODE_REGIONAL) {
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] = h[j] > 0.f ? h[j] : 0.01f * h[j];
  }
}

// saved-plane tile layout [T][nqt][HH/4][128][4]: thread r stores its 32 columns as 8 float4
template <int HH>
__device__ __forceinline__ float* plane_ptr(float* plane, int nqt, int t, int qt, int c4, int r) {
  return plane + ((((size_t)t * nqt + qt) * (HH / 4) + c4) * TC_ROWS + r) * 4;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int FMT, int HH>
__global__ void __launch_bounds__(NTHREADS, 1) k_cell_fwd_tc(TcArgs a) {
  using Cfg = TcCfg<FMT, HH>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* W = smem;                                   // weight image (tiles + consts)
  uint8_t* Ah = W + ((Cfg::FWD_IMG + 1023) & ~1023);   // [NSPLIT][128 x HH]
  uint8_t* As = Ah + Cfg::NSPLIT * Cfg::A_TILE;        // [NSPLIT] chunk tiles
  __shared__ uint64_t bar_a, bar_zr, bar_a2, bar_c;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < Cfg::FWD_IMG / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(W)[i] = __ldg(reinterpret_cast<const uint4*>(a.img) + i);
  if (tid == 0) {
    mbar_init(&bar_a, NEPI);
    mbar_init(&bar_zr, 1);
    mbar_init(&bar_a2, NEPI);
    mbar_init(&bar_c, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&tmem_base_s, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const float* consts = reinterpret_cast<const float*>(W + Cfg::FWD_W);
  uint32_t ph = 0;

  if (warp < 8) {
    // ================= epilogue / prologue threads =================
    const int r = (warp & 3) * 32 + lane;  // row = TMEM lane
    const int ch = warp >> 2;              // column half
    const int c0 = ch * 32;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
      const int qt = item / a.ntc, tc_i = item % a.ntc;
      const long long q = (long long)qt * TC_ROWS + r;
      const bool valid = q < a.BN;
      const int b = valid ? (int)(q / a.N) : 0, n = valid ? (int)(q % a.N) : 0;
      int s0 = 0, s1 = 0;
      if (valid && a.mode != REGT_MODE_TGCN) {
        s0 = a.seg_ptr[n];
        s1 = a.seg_ptr[n + 1];
      }
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
      for (int t = tc_i * a.tp; t < (tc_i + 1) * a.tp; ++t) {
        float h[32], sv[8];
        compute_h32<HH>(a, consts, valid, q, b, s0, s1, t, c0, h, sv);
        store_operand32<FMT, HH>(Ah, r, c0, h);
        if (ch == 0) {
          store_small8<FMT, HH>(As, Cfg::AS_TILE, r, 0, sv);
          if constexpr (FMT == FMT_BF16)  // zero padding chunk of the 16-wide k-step
            *reinterpret_cast<uint4*>(As + chunk_off(r, 1, TC_ROWS)) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bar_a);

        // ---- E1: gates ----
        mbar_wait(&bar_zr, ph);
        tc_fence_after();
        float z[32];
        {
          float raw[32];
          tmem_ld32(tlane + c0, raw);
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = fast_sigmoid(raw[j] + consts[Cfg::C_CZR + c0 + j]);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(plane_ptr<HH>(a.Zp, a.nqt, t, qt, (c0 + j) / 4, r)) =
                make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]);
          tmem_ld32(tlane + HH + c0, raw);
          float hr[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float rg = fast_sigmoid(raw[j] + consts[Cfg::C_CZR + HH + c0 + j]);
            raw[j] = rg;
            hr[j] = h[j] * rg;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(plane_ptr<HH>(a.Rp, a.nqt, t, qt, (c0 + j) / 4, r)) =
                make_float4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
          store_operand32<FMT, HH>(Ah, r, c0, hr);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bar_a2);

        // ---- E2: candidate, blend, attention accumulation ----
        mbar_wait(&bar_c, ph);
        tc_fence_after();
        {
          float raw[32];
          tmem_ld32(tlane + 2 * HH + c0, raw);
          const float pt = consts[Cfg::C_PROBS + t];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float hc = fast_tanh(raw[j] + consts[Cfg::C_CC + c0 + j]);
            raw[j] = hc;
            acc[j] = fmaf(pt, z[j] * h[j] + (1.0f - z[j]) * hc, acc[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(plane_ptr<HH>(a.Hcp, a.nqt, t, qt, (c0 + j) / 4, r)) =
                make_float4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
        }
        ph ^= 1;
      }
      if (valid) {
        float* o = a.hid_part + ((size_t)tc_i * a.BN + q) * HH + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
    }
    tc_fence_before();
  } else {
    // ================= MMA issuer (warp 8, one elected lane) =================
    const uint32_t idesc_zr = make_idesc(FMT, 128, 2 * HH, 0, 0), idesc_c = make_idesc(FMT, 128, HH, 0, 0);
    const uint32_t ah = smem_u32(Ah), as = smem_u32(As), w = smem_u32(W);
    for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
      for (int t = 0; t < a.tp; ++t) {
        mbar_wait(&bar_a, ph);
        tc_fence_after();
        if (lane == 0) {
          issue_gate_mma<FMT, HH>(tmem, ah, as, w, w + Cfg::WZR_H, 2 * HH, idesc_zr);
          umma_commit(&bar_zr);
        }
        __syncwarp();
        mbar_wait(&bar_a2, ph);
        tc_fence_after();
        if (lane == 0) {
          issue_gate_mma<FMT, HH>(tmem + 2 * HH, ah, as, w + Cfg::WZR_H + Cfg::WZR_S,
                                  w + Cfg::WZR_H + Cfg::WZR_S + Cfg::WC_H, HH, idesc_c);
          umma_commit(&bar_c);
        }
        __syncwarp();
        ph ^= 1;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

// out_hidden[q][j] = sum over the t-chunks of the per-item partial attention sums
__global__ void k_hid_reduce(const float* __restrict__ part, int ntc, long long count, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i * 4 >= count) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < ntc; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)c * count) + i);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = s;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int launch_spmm_rows(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, float* y, int B,
                     int n_out, int n_in, int width, cudaStream_t st);
int launch_prep(const regt_args* a, const Layout& L, cudaStream_t st);

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// periods per work item: maximise SM utilisation of the last wave, prefer longer chunks
static int choose_tp(int nqt, int T, int slots) {
  int best = 1;
  double best_eff = -1.0;
  for (int tp = T; tp >= 1; --tp) {
    if (T % tp) continue;
    const long long items = (long long)nqt * (T / tp);
    const long long waves = (items + slots - 1) / slots;
    const double eff = (double)items / (double)(waves * slots);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = tp;
    }
  }
  return best;
}

static TcArgs make_tcargs(const regt_args* a, const Layout& L, int slots) {
  TcArgs k{};
  k.BN = a->B * a->N; k.N = a->N; k.T = a->T; k.nseg = a->plan.nseg; k.mode = a->mode;
  k.nqt = (k.BN + TC_ROWS - 1) / TC_ROWS;
  k.tp = choose_tp(k.nqt, a->T, slots);
  k.ntc = a->T / k.tp;
  k.items = k.nqt * k.ntc;
  k.x = a->x; k.S = L.S; k.U = L.U;
  k.seg_ptr = a->plan.seg_ptr; k.seg_reg = a->plan.seg_reg;
  k.M1t = L.M1t;
  k.img = L.tc_img_f;
  k.Zp = L.Zp; k.Rp = L.Rp; k.Hcp = L.Hcp; k.hid_part = L.hid_part;
  k.G = L.G; k.dhp = L.dhp_p; k.wpart = L.tc_wpart; k.dprobs_part = L.tc_dpp;
  return k;
}

template <int FMT, int HH>
static int run_fwd_tc(const regt_args* a, const Layout& L, cudaStream_t st) {
  using Cfg = TcCfg<FMT, HH>;
  static_assert(Cfg::FWD_IMG <= TC_IMG_BYTES && Cfg::BWD_IMG <= TC_IMG_BYTES, "weight image too large");
  const int n_pack = 2 * HH * HH + 2 * HH * 16 + HH * HH + HH * 16 + Cfg::C_FLOATS + 3 * HH * HH;
  k_pack_tc<FMT, HH><<<cdiv(n_pack, 256), 256, 0, st>>>(L.Wzr, L.Wc, L.czr, L.cc, L.c0, L.M0t, L.M1t, L.probs, a->T,
                                                       a->p.lin_w[0], a->p.lin_w[1], a->p.lin_w[2], L.tc_img_f,
                                                       L.tc_img_b);
  REGT_LAUNCHED("k_pack_tc", st);
  const int slots = num_sms();
  TcArgs k = make_tcargs(a, L, slots);
  const size_t smem = 1024 + ((Cfg::FWD_IMG + 1023) & ~1023) + Cfg::NSPLIT * (Cfg::A_TILE + Cfg::AS_TILE);
  REGT_CUDA(cudaFuncSetAttribute(k_cell_fwd_tc<FMT, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(slots, k.items);
  k_cell_fwd_tc<FMT, HH><<<grid, NTHREADS, smem, st>>>(k);
  REGT_LAUNCHED("k_cell_fwd_tc", st);
  const long long count = (long long)k.BN * HH;
  k_hid_reduce<<<cdiv(count / 4, 256), 256, 0, st>>>(L.hid_part, k.ntc, count, a->out_hidden);
  REGT_LAUNCHED("k_hid_reduce", st);
  return 0;
}

int cell_forward_tc(const regt_args* a, const Layout& L, cudaStream_t st) {
  REGT_CHECK(a->H == 64, "tensor-core precisions are built for hidden=64 (got %d); use precision fp32", a->H);
  REGT_CHECK(a->mode != REGT_MODE_TGCN, "tensor-core precisions do not cover the bare TGCN cell; use precision fp32");
  REGT_CHECK(a->T <= 64, "tensor-core path supports up to 64 periods");
  if (launch_prep(a, L, st)) return -1;
  if (launch_spmm_rows(a->plan.g_rowptr, a->plan.g_col, a->plan.g_val, a->x, L.S, a->B, a->N, a->N, F * a->T, st)) return -1;
  if (a->plan.nseg > 0) {
    if (launch_spmm_rows(a->plan.seg_eptr, a->plan.c_col, a->plan.c_val, a->x, L.U, a->B, a->plan.nseg, a->N, F * a->T, st))
      return -1;
  }
  if (a->precision == REGT_PREC_TF32X3) return run_fwd_tc<FMT_TF32, 64>(a, L, st);
  return run_fwd_tc<FMT_BF16, 64>(a, L, st);
}

int cell_backward_tc(const regt_args*, const Layout&, cudaStream_t) {
  set_error("tensor-core backward is not built yet; use precision fp32 for training");
  return -10;
}

}  // namespace regt
