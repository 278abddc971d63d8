// error string, version, launch counter and the per-kernel event profiler
// (all thread-local; the library has no other globals)
#include <stdarg.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace regt {
static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;
static thread_local bool g_prof = false;
struct Mark {
  const char* name;
  cudaEvent_t ev;
};
static thread_local std::vector<Mark> g_marks;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
void prof_mark(const char* name, cudaStream_t st) {
  if (!g_prof) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  g_marks.push_back(Mark{name, ev});
}

// ---- fork / join onto a library-owned side stream (independent kernels of one step run concurrently;
// legal under stream capture: the side stream joins the caller's capture through the fork event) ----
struct Side {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static thread_local Side g_side;

cudaStream_t fork_side(cudaStream_t main) {
  if (g_prof) return nullptr;   // the per-kernel profiler brackets kernels on ONE stream
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (g_side.device != dev) {
    if (g_side.stream) {
      cudaStreamDestroy(g_side.stream);
      cudaEventDestroy(g_side.fork);
      cudaEventDestroy(g_side.join);
      g_side = Side{};
    }
    if (cudaStreamCreateWithFlags(&g_side.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&g_side.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g_side.join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      g_side = Side{};
      return nullptr;
    }
    g_side.device = dev;
  }
  if (cudaEventRecord(g_side.fork, main) != cudaSuccess || cudaStreamWaitEvent(g_side.stream, g_side.fork, 0) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return g_side.stream;
}
int join_side(cudaStream_t main) {
  REGT_CUDA(cudaEventRecord(g_side.join, g_side.stream));
  REGT_CUDA(cudaStreamWaitEvent(main, g_side.join, 0));
  return 0;
}
}  // namespace regt

extern "C" int regt_version(void) { return REGT_VERSION; }
extern "C" const char* regt_last_error(void) { return regt::g_err; }
extern "C" int64_t regt_launch_count(int reset) {
  long long v = regt::g_launches;
  if (reset) regt::g_launches = 0;
  return v;
}

extern "C" int regt_profile(int enable, regt_stream_t stream) {
  for (auto& m : regt::g_marks) cudaEventDestroy(m.ev);
  regt::g_marks.clear();
  regt::g_prof = enable != 0;
  if (regt::g_prof) regt::prof_mark("<begin>", (cudaStream_t)stream);
  return 0;
}

// Synchronises on the recorded events.  names: '\n'-separated kernel names, ms[i] = time between
// mark i and mark i+1 on the launching stream (kernel i+1 incl. its launch gap).  Returns the
// number of intervals written (<= max_n), or <0 on error.
extern "C" int regt_profile_read(char* names, size_t names_len, float* ms, int max_n) {
  auto& M = regt::g_marks;
  int n = 0;
  std::string s;
  for (size_t i = 1; i < M.size() && n < max_n; ++i) {
    if (strcmp(M[i].name, "<begin>") == 0) continue;
    if (cudaEventSynchronize(M[i].ev) != cudaSuccess) return -1;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, M[i - 1].ev, M[i].ev) != cudaSuccess) return -1;
    ms[n++] = t;
    s += M[i].name;
    s += '\n';
  }
  if (names && names_len) {
    strncpy(names, s.c_str(), names_len - 1);
    names[names_len - 1] = 0;
  }
  return n;
}

// re-arm the interval origin (call right before a profiled API call)
extern "C" int regt_profile_begin(regt_stream_t stream) {
  regt::prof_mark("<begin>", (cudaStream_t)stream);
  return 0;
}
