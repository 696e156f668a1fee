// capi.cu -- ABI version and error strings of libfbbev_b200.so
#include <atomic>

#include "common.cuh"

namespace fbbev {
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace fbbev

FBBEV_API long long fbbev_debug_launch_count(void) {
  return fbbev::g_launches.load(std::memory_order_relaxed);
}

FBBEV_API int fbbev_abi_version(void) { return FBBEV_ABI_VERSION; }

FBBEV_API const char* fbbev_error_string(int code) {
  switch (code) {
    case FBBEV_OK: return "ok";
    case FBBEV_ERR_INVALID_ARGUMENT: return "fbbev: invalid argument";
    case FBBEV_ERR_WORKSPACE_TOO_SMALL: return "fbbev: workspace too small";
    case FBBEV_ERR_UNSUPPORTED: return "fbbev: unsupported shape or size";
    default:
      if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
      return "fbbev: unknown error";
  }
}
