// common.cu - error plumbing and small process-wide helpers of libdpi_b200.
#include "common.cuh"

#include <string.h>

#include <mutex>
#include <unordered_map>

namespace dpi {

static thread_local char t_error[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

namespace {
std::mutex g_handle_mu;
std::unordered_map<const void*, int>& handle_table() {
  static std::unordered_map<const void*, int> t;
  return t;
}
}  // namespace
void handle_register(const void* h, int kind) {
  std::lock_guard<std::mutex> lk(g_handle_mu);
  handle_table()[h] = kind;
}
bool handle_unregister(const void* h, int kind) {
  std::lock_guard<std::mutex> lk(g_handle_mu);
  auto it = handle_table().find(h);
  if (it == handle_table().end() || it->second != kind) return false;
  handle_table().erase(it);
  return true;
}
bool handle_live(const void* h, int kind) {
  if (!h) return false;
  std::lock_guard<std::mutex> lk(g_handle_mu);
  auto it = handle_table().find(h);
  return it != handle_table().end() && it->second == kind;
}

int sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int current_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 148; }
  if (dev < 0 || dev >= 64) return sm_count(dev);
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) { n = sm_count(dev); if (n <= 0) n = 148; cache[dev].store(n, std::memory_order_relaxed); }
  return n;
}

}  // namespace dpi

extern "C" {
const char* dpi_last_error(void) { return dpi::t_error; }
int dpi_version(void) { return 100; }
long long dpi_launch_count(void) { return dpi::g_launches.load(); }
}
