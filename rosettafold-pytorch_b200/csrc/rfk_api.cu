// rfk_api.cu — library-level entry points: error strings, version, launch counter, device probes.
#include <atomic>

#include "rfk_common.cuh"

namespace rfk {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Device probes are cached PER DEVICE (index = cudaGetDevice()): a host process may drive several GPUs.
static std::atomic<int> g_arch[kMaxDevices];  // 0 = unknown, 1 = sm_10x, 2 = anything else
static std::atomic<int> g_sms[kMaxDevices];

int check_arch() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return RFK_ERR_UNSUPPORTED_ARCH;
  const bool cacheable = dev >= 0 && dev < kMaxDevices;
  if (cacheable) {
    const int c = g_arch[dev].load(std::memory_order_relaxed);
    if (c) return c == 1 ? RFK_OK : RFK_ERR_UNSUPPORTED_ARCH;
  }
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return RFK_ERR_UNSUPPORTED_ARCH;
  if (cacheable) g_arch[dev].store(major == 10 ? 1 : 2, std::memory_order_relaxed);
  return major == 10 ? RFK_OK : RFK_ERR_UNSUPPORTED_ARCH;
}

int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const bool cacheable = dev >= 0 && dev < kMaxDevices;
  if (cacheable && (n = g_sms[dev].load(std::memory_order_relaxed)) > 0) return n;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  n = n > 0 ? n : 148;
  if (cacheable) g_sms[dev].store(n, std::memory_order_relaxed);
  return n;
}

}  // namespace rfk

extern "C" const char* rfk_strerror(int code) {
  switch (code) {
    case RFK_OK: return "ok";
    case RFK_ERR_BAD_DIMS: return "bad dimensions";
    case RFK_ERR_MISALIGNED: return "misaligned pointer or stride (bf16 operands need 16-byte alignment)";
    case RFK_ERR_UNSUPPORTED_ARCH: return "unsupported GPU architecture (sm_100a required)";
    case RFK_ERR_BAD_DTYPE: return "bad dtype code";
    case RFK_ERR_NULL_POINTER: return "null pointer";
    case RFK_ERR_TMA_ENCODE: return "cuTensorMapEncodeTiled failed";
    case RFK_ERR_WORKSPACE: return "workspace too small";
    case RFK_ERR_UNSUPPORTED: return "unsupported argument combination";
    default: break;
  }
  if (code >= RFK_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(code - RFK_ERR_CUDA_BASE));
  return "unknown rfk error";
}

extern "C" int rfk_version(void) { return 1; }
extern "C" uint64_t rfk_launch_count(void) { return rfk::g_launches.load(); }
