// Library-level state: version string, thread-local error text, cached SM count.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>
#include <atomic>

namespace npp {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof g_err, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
void set_error_str(const char* what) {
  strncpy(g_err, what, sizeof g_err - 1);
  g_err[sizeof g_err - 1] = 0;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NPP_PDL");
    v = (e && *e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace npp

extern "C" {
const char* npp_version(void) { return "npp_b200 0.1 (sm_100a)"; }
const char* npp_last_error(void) { return npp::g_err; }
int npp_sm_count(void) { return npp::sm_count(); }
long long npp_launch_count(void) { return npp::g_launches.load(); }
}
