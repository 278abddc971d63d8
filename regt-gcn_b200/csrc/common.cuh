// Shared helpers for libregt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "regt_b200.h"

namespace regt {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void prof_mark(const char* name, cudaStream_t st);
// fork: returns a library-owned side stream that waits for everything enqueued on `main` so far (or
// nullptr: run serially); join: `main` waits for everything enqueued on the side stream
cudaStream_t fork_side(cudaStream_t main);
int join_side(cudaStream_t main);

#define REGT_CHECK(cond, ...)         \
  do {                                \
    if (!(cond)) {                    \
      regt::set_error(__VA_ARGS__);   \
      return -1;                      \
    }                                 \
  } while (0)

#define REGT_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      regt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                         \
    }                                                                                    \
  } while (0)

#define REGT_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    regt::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      regt::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -3;                                                                         \
    }                                                                                    \
  } while (0)

// like REGT_LAUNCH_CHECK, and (when regt_profile(1) is active) records a CUDA event on the
// launching stream so that consecutive marks bracket every kernel of the step
#define REGT_LAUNCHED(name, st)       \
  do {                                \
    REGT_LAUNCH_CHECK();              \
    regt::prof_mark(name, st);        \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// bump allocator over the caller's workspace
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base((char*)p), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = base ? (T*)(base + off) : (T*)nullptr;
    off += n * sizeof(T);
    return r;
  }
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// ---- workspace layout shared by forward / backward (cell.cu, head.cu, api.cu) -------
struct Layout {
  // collapsed weights (rebuilt every forward: parameters change between steps)
  float* Wzr;    // [F+H][2H]   rows 0..F-1: (A_g W_g)^T, rows F..: B_g^T ; cols 0..H-1 z, H..2H-1 r
  float* Wc;     // [F+H][H]
  float* czr;    // [2H]        A_g b_g + linear_g.bias
  float* cc;     // [H]
  float* M0t;    // [F][H]      (sum_r L_r W0)^T          (A3TGCN: W0^T)
  float* M1t;    // [R][F][H]   (L_r W1)^T                (A3TGCN: W1^T)
  float* c0;     // [H]         (sum_r L_r) b + b_lin     (A3TGCN: b)
  float* Lsum;   // [H][H]      sum_r L_r                 (regional only)
  float* probs;  // [T]
  // F-wide features
  float* S;      // [B*N][F][T]      A_hat X
  float* U;      // [B][nseg][F][T]  L_hat_r X per (node, region) segment
  // saved planes, row = (b*N+n)*T + t
  float* h;      // [rows][H]
  float* Z;
  float* Rg;
  float* Hc;     // candidate H~
  float* hR;     // h * R
  float* Hn;     // H' (cell output per period)
  // head
  float* a1;     // [B*N][128]  relu(linear1(relu(hid)))
  float* G;      // [B*N][H]    gradient wrt out_hidden
  float* d_a1;   // [B*N][128]
  // backward planes
  float* D;      // [rows][4H]  d_pre_z | d_pre_r | d_pre_h | d_hpre
  float* Feat;   // [rows][32]  S_t | X_t | 1 | 0  (tf32x3: B operand of the F-wide weight-gradient GEMM)
  // tensor-core head (head_f.cu, precision tf32x3)
  float* hf_rh;      // [BN][H]   relu(out_hidden) forward, then the unmasked gradient wrt out_hidden backward
  float* hf_do32;    // [BN][32]  d_out | 1 (column 16): feature-plane operand of the head's weight-gradient contractions
  float* hf_split;   // hi | lo images of the weight operand of the head GEMMs
  float* hf_loss;    // per-block loss partials
  size_t hf_loss_floats;
  float* FeatT;  // fused tf32x3 cell: transposed feature tiles [T*nqt][32][128] (cell_f.cu)
  float* bsplit; // tf32x3: hi | lo images of the weight operand of the current gate GEMM (gemm_tma.cu)
  // collapsed-weight gradients
  float* dB;     // [3][H][H]   dB_z, dB_r, dB_h
  float* dP;     // [3][H][F]
  float* dcg;    // [3][H]
  float* dM0;    // [H][F]
  float* dM1;    // [R][H][F]
  float* dc0;    // [H]
  float* dprobs; // [T]
  float* hpart;  // fused-head per-CTA partials
  float* part;   // split-K partials
  size_t part_floats;
  int32_t* m1cp; // [R+1] chunk table of the per-region dM1 reduction (cell.cu k_m1_chunks)
  // tensor-core path (precision != FP32): weight images, tile-layout planes, per-CTA partials
  unsigned char* tc_img_f;   // forward weight image
  unsigned char* tc_img_b;   // backward weight image
  float *Zp, *Rp, *Hcp, *dhp_p;  // [T][nqt][H/4][128][4]
  float* Xt;                 // [T][BN][F] period-major copy of x (S and U are period-major too)
  float* hid_part;           // [T (max t-chunks)][nqt][H/4][128][4]
  float* tc_wpart;           // [TC_MAX_CTAS][TC_WPART_FLOATS]
  float* tc_dpp;             // [T * nqt] attention-gradient partials
  size_t total;
};
constexpr int TC_MAX_CTAS = 160;
constexpr int TC_IMG_BYTES = 160 * 1024;
constexpr int F_IMG_BYTES = 432 * 1024;   // fused 3xTF32 cell (cell_f.cu): 12 ring stages of 32 KB + F-wide weights + constants at H = 128
bool cell_f_usable(const regt_args* a);
bool head_f_usable(const regt_args* a);   // precision tf32x3, hidden % 32 == 0, output_dim <= 16, >= 128 rows   // precision tf32x3, hidden 128 / 64, not the bare TGCN cell

Layout make_layout(const regt_args* a, void* base);
size_t gemm_nt_scratch_floats(int N, int K);

constexpr int HEAD_HID = 128;  // hidden_dim of the decoder MLP (models/RegionalTemporalGCN.py:19)
constexpr int WGRAD_SPLITS = 64;

}  // namespace regt
