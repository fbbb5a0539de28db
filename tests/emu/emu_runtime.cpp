// Warp-lockstep CPU runner for generated kernels (see cuda_shim.h).  Test infrastructure only.
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>
#include "cuda_shim.h"

thread_local emu_uint3 threadIdx, blockIdx;
emu_uint3 blockDim, gridDim;

namespace {
thread_local int t_lane = 0;
unsigned long long g_slots[32];
bool g_votes[32];
int g_active = 32;
std::barrier<>* g_bar = nullptr;
}  // namespace

namespace emu {
int lane() { return t_lane; }
unsigned long long exchange(unsigned long long bits, int src_lane) {
  g_slots[t_lane] = bits;
  g_bar->arrive_and_wait();
  unsigned long long r = g_slots[src_lane < g_active ? src_lane : t_lane];
  g_bar->arrive_and_wait();
  return r;
}
bool any(bool pred) {
  g_votes[t_lane] = pred;
  g_bar->arrive_and_wait();
  bool r = false;
  for (int i = 0; i < g_active; ++i) r = r || g_votes[i];
  g_bar->arrive_and_wait();
  return r;
}
void fail(const char* what) {
  fprintf(stderr, "b200 emulator: %s\n", what);
  fflush(stderr);
  abort();
}
static long long g_trace[8];
void trace(int slot) { __atomic_fetch_add(&g_trace[slot & 7], 1LL, __ATOMIC_RELAXED); }
}  // namespace emu

extern "C" long long emu_trace_read(int slot, int reset) {
  long long v = __atomic_load_n(&emu::g_trace[slot & 7], __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&emu::g_trace[slot & 7], 0LL, __ATOMIC_RELAXED);
  return v;
}

extern "C" void emu_launch(void (*kernel)(const void*), const void* args, const unsigned grid[3], const unsigned block[3],
                           int lockstep) {
  gridDim = {grid[0], grid[1], grid[2]};
  blockDim = {block[0], block[1], block[2]};
  const unsigned nthreads = block[0] * block[1] * block[2];
  if (!lockstep) {  // kernels without warp-level primitives: plain sequential execution of every thread
    for (unsigned bz = 0; bz < grid[2]; ++bz)
      for (unsigned by = 0; by < grid[1]; ++by)
        for (unsigned bx = 0; bx < grid[0]; ++bx)
          for (unsigned tid = 0; tid < nthreads; ++tid) {
            threadIdx = {tid % block[0], (tid / block[0]) % block[1], tid / (block[0] * block[1])};
            blockIdx = {bx, by, bz};
            kernel(args);
          }
    return;
  }
  for (unsigned bz = 0; bz < grid[2]; ++bz)
    for (unsigned by = 0; by < grid[1]; ++by)
      for (unsigned bx = 0; bx < grid[0]; ++bx)
        for (unsigned w0 = 0; w0 < nthreads; w0 += 32) {
          const int active = (int)std::min(32u, nthreads - w0);
          g_active = active;
          std::barrier<> bar(active);
          g_bar = &bar;
          std::vector<std::thread> lanes;
          for (int l = 0; l < active; ++l)
            lanes.emplace_back([&, l] {
              const unsigned tid = w0 + l;
              t_lane = l;
              threadIdx = {tid % block[0], (tid / block[0]) % block[1], tid / (block[0] * block[1])};
              blockIdx = {bx, by, bz};
              kernel(args);
              // a lane that leaves early must not strand the others at a barrier: kernels only
              // return warp-uniformly, so nothing to do here
            });
          for (auto& t : lanes) t.join();
        }
}
