// Decoder head of RegionalTemporalGCN / TemporalGCN and the call-site loss:
//   out = linear2(relu(linear1(relu(out_hidden))))      models/RegionalTemporalGCN.py:35-38
//   loss = sum_b mean_{n,o} (out - y)^2                 run.py:180 (one backward per snapshot)
// plus their autograd (run.py:190).  One row per (b, n); tiny next to the cell.
#include "gemm_simt.cuh"

namespace regt {

constexpr int TMH = 64;

struct HeadK {
  long long BN;
  int N, H, O;
  const float *hid, *y, *W1t, *W2t, *b1, *b2;  // W1t [H][128], W2t [128][O]
  const float *w1, *w2;                        // natural layouts for the data gradients
  const float* d_hidden;
  float *a1, *out, *d_out, *loss_part, *d_a1, *G;
};

__global__ void k_head_transpose(const float* __restrict__ w1, const float* __restrict__ w2, int H, int O,
                                 float* __restrict__ W1t, float* __restrict__ W2t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * HEAD_HID) {
    int k = i / HEAD_HID, m = i % HEAD_HID;  // W1t[k][m] = w1[m][k]
    W1t[i] = w1[(size_t)m * H + k];
  } else if (i < H * HEAD_HID + HEAD_HID * O) {
    int j = i - H * HEAD_HID;
    int m = j / O, o = j % O;  // W2t[m][o] = w2[o][m]
    W2t[j] = w2[(size_t)o * HEAD_HID + m];
  }
}

__global__ void __launch_bounds__(TMH * 4) k_head_fwd(HeadK a) {
  constexpr int NT = TMH * 4;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, O = a.O, lda = H + 1, ld1 = HEAD_HID + 1;
  float* A0 = smem;                 // relu(hid) [TMH][H+1]
  float* A1 = A0 + TMH * lda;       // a1 [TMH][129]
  float* Ws = A1 + TMH * ld1;       // offsets are multiples of 64 floats
  __shared__ float red[32];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long q0 = (long long)blockIdx.x * TMH;
  for (int idx = tid; idx < TMH * H; idx += NT) {
    const int rl = idx / H, j = idx - rl * H;
    const long long q = q0 + rl;
    A0[rl * lda + j] = q < a.BN ? fmaxf(a.hid[q * H + j], 0.f) : 0.f;
  }
  for (int n0 = 0; n0 < HEAD_HID; n0 += TN) {
    float acc[4][4];
    zero_acc(acc);
    tile_gemm<NT>(A0, lda, H, a.W1t, HEAD_HID, n0, HEAD_HID, Ws, acc, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = ty * 4 + i;
      const long long q = q0 + rl;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        const float v = fmaxf(acc[i][j] + __ldg(a.b1 + n), 0.f);
        A1[rl * ld1 + n] = v;
        if (q < a.BN) a.a1[q * HEAD_HID + n] = v;
      }
    }
  }
  float lsum = 0.f;
  const float scale = 1.0f / ((float)a.N * (float)O);
  for (int n0 = 0; n0 < O; n0 += TN) {
    float acc[4][4];
    zero_acc(acc);
    tile_gemm<NT>(A1, ld1, HEAD_HID, a.W2t, O, n0, O, Ws, acc, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long q = q0 + ty * 4 + i;
      if (q >= a.BN) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n >= O) continue;
        const float v = acc[i][j] + __ldg(a.b2 + n);
        a.out[q * O + n] = v;
        if (a.y) {
          const float diff = v - __ldg(a.y + q * O + n);
          lsum = fmaf(diff, diff, lsum);
          a.d_out[q * O + n] = 2.0f * diff * scale;
        }
      }
    }
  }
  if (a.y) {
    lsum = block_sum(lsum, red);
    if (tid == 0) a.loss_part[blockIdx.x] = lsum * scale;
  }
}

__global__ void __launch_bounds__(TMH * 4) k_head_bwd(HeadK a) {
  constexpr int NT = TMH * 4;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, O = a.O, ldo = O + 1, ld1 = HEAD_HID + 1;
  float* Do = smem;                 // d_out [TMH][O+1]
  float* D1 = Do + TMH * ldo;       // d_a1 [TMH][129]
  float* Ws = D1 + TMH * ld1;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long q0 = (long long)blockIdx.x * TMH;
  for (int idx = tid; idx < TMH * O; idx += NT) {
    const int rl = idx / O, j = idx - rl * O;
    const long long q = q0 + rl;
    Do[rl * ldo + j] = q < a.BN ? a.d_out[q * O + j] : 0.f;
  }
  // d_a1 = (d_out . W2) * (a1 > 0)
  for (int n0 = 0; n0 < HEAD_HID; n0 += TN) {
    float acc[4][4];
    zero_acc(acc);
    tile_gemm<NT>(Do, ldo, O, a.w2, HEAD_HID, n0, HEAD_HID, Ws, acc, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = ty * 4 + i;
      const long long q = q0 + rl;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        float v = 0.f;
        if (q < a.BN) {
          v = a.a1[q * HEAD_HID + n] > 0.f ? acc[i][j] : 0.f;
          a.d_a1[q * HEAD_HID + n] = v;
        }
        D1[rl * ld1 + n] = v;
      }
    }
  }
  // G = (d_a1 . W1) * (hid > 0) + d_hidden
  for (int n0 = 0; n0 < H; n0 += TN) {
    float acc[4][4];
    zero_acc(acc);
    tile_gemm<NT>(D1, ld1, HEAD_HID, a.w1, H, n0, H, Ws, acc, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long q = q0 + ty * 4 + i;
      if (q >= a.BN) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n >= H) continue;
        float v = a.hid[q * H + n] > 0.f ? acc[i][j] : 0.f;
        if (a.d_hidden) v += a.d_hidden[q * H + n];
        a.G[q * H + n] = v;
      }
    }
  }
}

// gradient flowing only through out_hidden (no head gradient): G = d_hidden or 0
__global__ void k_copy_or_zero(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src ? src[i] : 0.f;
}

static HeadK make_headk(const regt_args* a, const Layout& L, float* W1t, float* W2t) {
  HeadK k{};
  k.BN = (long long)a->B * a->N;
  k.N = a->N; k.H = a->H; k.O = a->O;
  k.hid = a->out_hidden; k.y = a->y; k.W1t = W1t; k.W2t = W2t; k.b1 = a->p.head_b1; k.b2 = a->p.head_b2;
  k.w1 = a->p.head_w1; k.w2 = a->p.head_w2; k.d_hidden = a->d_hidden;
  k.a1 = L.a1; k.out = a->out; k.d_out = a->d_out; k.loss_part = L.part; k.d_a1 = L.d_a1; k.G = L.G;
  return k;
}

int head_forward_fp32(const regt_args* a, const Layout& L, cudaStream_t st) {
  const int H = a->H, O = a->O;
  // transposed head weights live at the tail of the split-K scratch's first page
  float* W1t = L.part + L.part_floats - ((size_t)H * HEAD_HID + (size_t)HEAD_HID * O);
  float* W2t = W1t + (size_t)H * HEAD_HID;
  k_head_transpose<<<cdiv((long long)H * HEAD_HID + HEAD_HID * O, 256), 256, 0, st>>>(a->p.head_w1, a->p.head_w2, H, O,
                                                                                  W1t, W2t);
  REGT_LAUNCHED("k_head_transpose", st);
  HeadK k = make_headk(a, L, W1t, W2t);
  const int nblk = cdiv(k.BN, TMH);
  const size_t smem = ((size_t)TMH * (H + 1) + TMH * (HEAD_HID + 1) + KT * TN) * sizeof(float);
  REGT_CUDA(cudaFuncSetAttribute(k_head_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_head_fwd<<<nblk, TMH * 4, smem, st>>>(k);
  REGT_LAUNCHED("k_head_fwd", st);
  if (a->y) {
    REGT_CHECK(a->loss && a->d_out, "head_forward: y given but loss/d_out is NULL");
    if (launch_reduce_splits(L.part, a->loss, 1, nblk, 0, st)) return -1;
  }
  return 0;
}

int head_backward_fp32(const regt_args* a, const Layout& L, cudaStream_t st) {
  const int H = a->H, O = a->O;
  const long long BN = (long long)a->B * a->N;
  if (!a->d_out) {  // only out_hidden carries gradient
    k_copy_or_zero<<<cdiv(BN * H, 256), 256, 0, st>>>(a->d_hidden, L.G, BN * H);
    REGT_LAUNCHED("k_copy_or_zero", st);
    return 0;
  }
  HeadK k = make_headk(a, L, nullptr, nullptr);
  const int nblk = cdiv(BN, TMH);
  const size_t smem = ((size_t)TMH * (O + 1) + TMH * (HEAD_HID + 1) + KT * TN + 4) * sizeof(float);
  REGT_CUDA(cudaFuncSetAttribute(k_head_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_head_bwd<<<nblk, TMH * 4, smem, st>>>(k);
  REGT_LAUNCHED("k_head_bwd", st);
  // weight gradients: dW2 = d_out^T a1, dW1 = d_a1^T relu(hid); biases = column sums
  const int splits = (int)max(1ll, min(32ll, BN / 128));
  float* part = L.part;
  TNBatch tb{};
  tb.nprob = 2;
  tb.p[0] = TNProb{a->d_out, L.a1, part, O, HEAD_HID, O, HEAD_HID, 0};
  tb.p[1] = TNProb{L.d_a1, a->out_hidden, part + (size_t)splits * O * HEAD_HID, HEAD_HID, H, HEAD_HID, H, 1};
  if (launch_wgrad_tn(tb, BN, splits, st)) return -1;
  if (launch_reduce_splits(tb.p[0].part, a->g.head_w2, (long long)O * HEAD_HID, splits, a->accumulate, st)) return -1;
  if (launch_reduce_splits(tb.p[1].part, a->g.head_w1, (long long)HEAD_HID * H, splits, a->accumulate, st)) return -1;
  if (launch_colsum(a->d_out, O, O, BN, splits, part, st)) return -1;
  if (launch_reduce_splits(part, a->g.head_b2, O, splits, a->accumulate, st)) return -1;
  if (launch_colsum(L.d_a1, HEAD_HID, HEAD_HID, BN, splits, part, st)) return -1;
  if (launch_reduce_splits(part, a->g.head_b1, HEAD_HID, splits, a->accumulate, st)) return -1;
  return 0;
}

}  // namespace regt
