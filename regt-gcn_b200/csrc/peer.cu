// The exchange step of a sharded job as ONE kernel over NVLink peer memory: a one-shot sum all-reduce of the flat
// shared-weight gradient buffer (P*4 bytes once per step; P = 37 k .. 4.4 M floats).
//
// Every rank owns one cudaMalloc'ed communication region  [ flags | data | scratch ]  that the other ranks of the node map
// through CUDA IPC.  The weight-gradient kernels write straight into `data` (the .grad tensors are views of it).  The
// all-reduce kernel then, per thread block b and without any grid-wide or host synchronisation:
//   1. reduce-scatter by PUSH: stores its copy of chunk r into slot `rank` of rank r's scratch (128-bit NVLink writes);
//   2. barrier of block b across the ranks (one flag word per (phase, block, peer), stored with release.sys into the
//      peer's region, polled with acquire.sys in the own region): all copies of the own chunk have landed;
//   3. sums them in rank order 0..W-1 (each element is summed exactly once, by its owner) and pushes the sums into chunk
//      `rank` of every rank's data -- every rank ends up with the same bits;
//   4. second barrier: all chunks have landed before the kernel ends.
// Small buffers (<= PEER_LL_MAX floats) take a push path instead (k_peer_allreduce_ll, NCCL's "LL" idea): every rank
// STORES its values into a receive slot of every peer as 8-byte (value, epoch) words -- data and flag arrive in one
// atomic store, one NVLink traversal, no barrier -- and sums what lands in its own slots, again in rank order.  Slots are
// double-buffered by epoch parity (a peer can be at most one call ahead).
// The barrier epoch lives in device memory (one counter per block), so the launch can be captured in a CUDA graph and
// replayed.  NCCL needs ~35-50 us for this buffer at 2-8 GPUs (latency-bound ring/tree protocol); the one-shot kernel
// is two NVLink round trips plus (W-1) * P * 4 bytes of reads.
#include "common.cuh"

namespace regt {
namespace {
constexpr int PEER_MAXW = 8;          // ranks of one NVSwitch node
constexpr int PEER_MAXB = 296;        // thread blocks of the all-reduce kernel (2 per SM)
constexpr int PEER_THREADS = 256;
constexpr size_t PEER_FLAG_BYTES = 64 * 1024;   // 2 phases x MAXB x MAXW flag words + MAXB counters, rounded up
static_assert((2 * PEER_MAXB * PEER_MAXW + PEER_MAXB + 1024) * 4 + 64 <= PEER_FLAG_BYTES, "flag area too small");   // + push-kernel epochs + error word

constexpr long long PEER_LL_MAX = 1 << 18;     // floats (<= 512 blocks); above this the pull kernel (bandwidth-bound) is used

struct PeerArgs {
  float* data[PEER_MAXW];
  uint32_t* flags[PEER_MAXW];
  uint32_t* counter;   // own region, after the flag words
  float* scratch[PEER_MAXW];   // every rank's scratch (the reduce-scatter results are gathered from their owners)
  int* error;          // own region: set when a peer never arrived
  int rank, world;
  long long n4;        // float4 elements
  // push path
  unsigned long long* slots[PEER_MAXW];   // [parity][source rank][n] (value, epoch) words in every rank's region
  long long n;
  const float* last_in;   // optional: added to this rank's element n - 4 (the loss slot) before the sum
  float* last_out;        // optional: receives the reduced element n - 4, which is then cleared in data
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// thread p < world signals rank p and waits for rank p's signal (slot of this block and phase)
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, int phase, uint32_t epoch) {
  __syncthreads();   // every thread of the block has finished the preceding reads / writes
  if ((int)threadIdx.x < a.world) {
    const int p = threadIdx.x;
    const size_t slot = ((size_t)phase * PEER_MAXB + blockIdx.x) * PEER_MAXW;
    __threadfence_system();
    st_release_sys(a.flags[p] + slot + a.rank, epoch);
    const uint32_t* mine = a.flags[a.rank] + slot + p;
    long long spins = 0;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > (1ll << 27)) {   // ~ seconds: a peer died or the ranks disagree about the call sequence
        *a.error = 1;
        break;
      }
    }
  }
  __syncthreads();
}

// large buffers, two stages, every NVLink transfer a WRITE (peer reads are request/response bound: the read-based version
// of this kernel reached 285 GB/s per GPU).  Rank r owns chunk r of the buffer.
//   stage 1 (reduce-scatter by push): every rank stores its copy of chunk r into slot `rank` of rank r's scratch;
//   barrier; rank r sums the W copies in rank order (its own straight from data, the others from its scratch: local reads);
//   stage 2 (all-gather by push): rank r stores the sums into chunk r of EVERY rank's data;
//   barrier: all chunks have landed here before the kernel ends (and nobody still reads a scratch the next call refills).
// Block b of every rank works on the same sub-slices of every chunk, so both barriers are per block.
__global__ void __launch_bounds__(PEER_THREADS) k_peer_allreduce(PeerArgs a) {
  __shared__ uint32_t epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = a.counter[blockIdx.x] + 1;
    a.counter[blockIdx.x] = epoch_s;
  }
  __syncthreads();
  const uint32_t epoch = epoch_s;
  const long long c4 = (a.n4 + a.world - 1) / a.world;           // float4 elements per chunk
  const long long stride = (long long)gridDim.x * PEER_THREADS;
  const float4* mine = reinterpret_cast<const float4*>(a.data[a.rank]);
  // stage 1: my copy of chunk r -> slot `rank` of rank r's scratch (scratch = W slots of c4 elements)
  for (long long j = blockIdx.x * (long long)PEER_THREADS + threadIdx.x; j < c4; j += stride) {
    float4 v[PEER_MAXW];
#pragma unroll
    for (int r = 0; r < PEER_MAXW; ++r) {
      const long long i = (long long)r * c4 + j;
      if (r < a.world && r != a.rank && i < a.n4) v[r] = mine[i];
    }
#pragma unroll
    for (int r = 0; r < PEER_MAXW; ++r) {
      const long long i = (long long)r * c4 + j;
      if (r < a.world && r != a.rank && i < a.n4) reinterpret_cast<float4*>(a.scratch[r])[(long long)a.rank * c4 + j] = v[r];
    }
  }
  peer_barrier(a, 0, epoch);   // every rank's copies of my chunk have landed in my scratch
  {
    const long long lo = (long long)a.rank * c4, hi = min(a.n4, lo + c4);
    const float4* slots = reinterpret_cast<const float4*>(a.scratch[a.rank]);
    for (long long j = blockIdx.x * (long long)PEER_THREADS + threadIdx.x; lo + j < hi; j += stride) {
      float4 v[PEER_MAXW];
#pragma unroll
      for (int r = 0; r < PEER_MAXW; ++r)
        if (r < a.world) v[r] = (r == a.rank) ? mine[lo + j] : ld_volatile4(slots + (long long)r * c4 + j);
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < PEER_MAXW; ++r)
        if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
      // stage 2: the sum goes to chunk `rank` of every rank's data (the own copy last: it was an input above)
#pragma unroll
      for (int r = 0; r < PEER_MAXW; ++r)
        if (r < a.world) reinterpret_cast<float4*>(a.data[r])[lo + j] = s;
    }
  }
  peer_barrier(a, 1, epoch);   // every chunk has landed in my data; every rank is done with its scratch
}

// push path: thread = 2 consecutive floats
__global__ void __launch_bounds__(PEER_THREADS) k_peer_allreduce_ll(PeerArgs a) {
  __shared__ uint32_t epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = a.counter[blockIdx.x] + 1;
    a.counter[blockIdx.x] = epoch_s;
  }
  __syncthreads();
  const uint32_t epoch = epoch_s;
  const long long i = 2 * (blockIdx.x * (long long)PEER_THREADS + threadIdx.x);
  if (i >= a.n) return;
  float* mine = a.data[a.rank];
  float2 v = *reinterpret_cast<const float2*>(mine + i);
  const long long li = a.n - 4;   // loss slot (even index)
  if (a.last_in && i == li) v.x += __ldg(a.last_in);
  const unsigned long long w0 = ((unsigned long long)epoch << 32) | __float_as_uint(v.x);
  const unsigned long long w1 = ((unsigned long long)epoch << 32) | __float_as_uint(v.y);
  const size_t base = ((size_t)(epoch & 1) * PEER_MAXW) * a.n;
#pragma unroll
  for (int p = 0; p < PEER_MAXW; ++p) {
    if (p < a.world && p != a.rank) {
      unsigned long long* dst = a.slots[p] + base + (size_t)a.rank * a.n + i;
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
    }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < PEER_MAXW; ++r) {
    if (r >= a.world) continue;
    if (r == a.rank) {
      s.x += v.x; s.y += v.y;
      continue;
    }
    const unsigned long long* src = a.slots[a.rank] + base + (size_t)r * a.n + i;
    unsigned long long x0, x1;
    long long spins = 0;
    do {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(x0), "=l"(x1) : "l"(src) : "memory");
      if (++spins > (1ll << 26)) {
        *a.error = 1;
        break;
      }
    } while ((uint32_t)(x0 >> 32) != epoch || (uint32_t)(x1 >> 32) != epoch);
    s.x += __uint_as_float((uint32_t)x0);
    s.y += __uint_as_float((uint32_t)x1);
  }
  if (a.last_out && i == li) {
    *a.last_out = s.x;
    s.x = 0.f;
  }
  *reinterpret_cast<float2*>(mine + i) = s;
}
}  // namespace
}  // namespace regt

using namespace regt;

// ---- communication region: [ 64 KiB flags + counters | data: n floats | scratch: n floats ] ---------------------------
extern "C" size_t regt_comm_region_bytes(int64_t n_floats) {
  const size_t n = ((size_t)n_floats + 3) / 4 * 4;
  // scratch holds world * ceil(n4 / world) float4 slots: up to (world - 1) slots more than n4 when world does not divide n4
  // (world sizes 3, 5, 6, 7), so it is sized for the worst case of PEER_MAXW ranks
  size_t bytes = PEER_FLAG_BYTES + align_up(n * sizeof(float), 256) + align_up((n + 4 * PEER_MAXW) * sizeof(float), 256);
  if ((long long)n <= PEER_LL_MAX) bytes += 2 * PEER_MAXW * n * sizeof(unsigned long long);   // push-path receive slots
  return bytes;
}
extern "C" size_t regt_comm_data_offset(void) { return PEER_FLAG_BYTES; }

// explicit allocation of a peer-mappable region (cudaMalloc: the IPC handle must cover a whole allocation); zero-filled
extern "C" int regt_comm_alloc(size_t bytes, void** ptr) {
  REGT_CHECK(ptr && bytes >= PEER_FLAG_BYTES, "regt_comm_alloc: bad arguments");
  REGT_CUDA(cudaMalloc(ptr, bytes));
  REGT_CUDA(cudaMemset(*ptr, 0, bytes));
  REGT_CUDA(cudaDeviceSynchronize());
  return 0;
}
extern "C" int regt_comm_free(void* ptr) {
  REGT_CUDA(cudaFree(ptr));
  return 0;
}
// handle: 64 bytes (cudaIpcMemHandle_t), host
extern "C" int regt_comm_export(void* ptr, unsigned char* handle) {
  REGT_CHECK(ptr && handle, "regt_comm_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  REGT_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof(h));
  return 0;
}
extern "C" int regt_comm_import(const unsigned char* handle, void** peer_ptr) {
  REGT_CHECK(handle && peer_ptr, "regt_comm_import: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  REGT_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int regt_comm_unimport(void* peer_ptr) {
  REGT_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

// Sum all-reduce (in place) of the `n_floats` data floats of every rank's region.  regions[r] = base of rank r's region as
// mapped in THIS process (regions[rank] = the own allocation).  Every rank must call it the same number of times with the
// same n_floats.  Enqueued on `stream`; capturable in a CUDA graph.
extern "C" int regt_peer_allreduce_f32(void* const* regions, int32_t rank, int32_t world, int64_t n_floats, const float* last_in,
                                       float* last_out, regt_stream_t stream) {
  REGT_CHECK(regions && world >= 1 && world <= PEER_MAXW && rank >= 0 && rank < world && n_floats > 0,
             "regt_peer_allreduce_f32: bad arguments (world=%d rank=%d n=%lld)", world, rank, (long long)n_floats);
  PeerArgs a{};
  const size_t n = ((size_t)n_floats + 3) / 4 * 4;
  for (int r = 0; r < world; ++r) {
    REGT_CHECK(regions[r], "regt_peer_allreduce_f32: region %d is NULL", r);
    a.flags[r] = (uint32_t*)regions[r];
    a.data[r] = (float*)((char*)regions[r] + PEER_FLAG_BYTES);
  }
  a.counter = a.flags[rank] + 2 * PEER_MAXB * PEER_MAXW;
  a.error = (int*)((char*)regions[rank] + PEER_FLAG_BYTES - 64);
  for (int r = 0; r < world; ++r) a.scratch[r] = (float*)((char*)regions[r] + PEER_FLAG_BYTES + align_up(n * sizeof(float), 256));
  a.rank = rank;
  a.world = world;
  a.n4 = (long long)(n / 4);
  a.n = (long long)n;
  a.last_in = last_in;
  a.last_out = last_out;
  static int force_pull = -1;
  if (force_pull < 0) {
    const char* e = getenv("REGT_PEER_PULL");
    force_pull = (e && e[0] == '1') ? 1 : 0;
  }
  if ((long long)n <= PEER_LL_MAX && !force_pull) {
    for (int r = 0; r < world; ++r)
      a.slots[r] = (unsigned long long*)((char*)regions[r] + PEER_FLAG_BYTES + align_up(n * sizeof(float), 256) +
                                         align_up((n + 4 * PEER_MAXW) * sizeof(float), 256));
    a.counter += PEER_MAXB;   // the push kernel's blocks keep their own epochs (its grid differs from the pull kernel's)
    const int blocks = (int)((n / 2 + PEER_THREADS - 1) / PEER_THREADS);
    k_peer_allreduce_ll<<<blocks, PEER_THREADS, 0, (cudaStream_t)stream>>>(a);
    REGT_LAUNCHED("k_peer_allreduce_ll", (cudaStream_t)stream);
    return 0;
  }
  REGT_CHECK(!last_in && !last_out, "regt_peer_allreduce_f32: last_in / last_out need n_floats <= %lld", PEER_LL_MAX);
  const int blocks = (int)max(1ll, min((long long)PEER_MAXB, (a.n4 + PEER_THREADS - 1) / PEER_THREADS));
  k_peer_allreduce<<<blocks, PEER_THREADS, 0, (cudaStream_t)stream>>>(a);
  REGT_LAUNCHED("k_peer_allreduce", (cudaStream_t)stream);
  return 0;
}
extern "C" int64_t regt_peer_push_max_floats(void) { return PEER_LL_MAX; }
// 1 if any all-reduce of this region gave up waiting for a peer (host read; synchronises the device)
extern "C" int regt_comm_error(void* region) {
  int e = 0;
  if (cudaMemcpy(&e, (char*)region + PEER_FLAG_BYTES - 64, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return e;
}
