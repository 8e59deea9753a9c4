// Error string, version and device query of the C ABI (include/sga_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace sga {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

// Upper bound on the CTAs (= SMs) the persistent one-CTA-per-SM PointNet kernels (forward, Gram statistics, backward)
// occupy; 0 = all.  A step that runs the graph branch concurrently on a second stream leaves it a few SMs this way
// (serving.CapturedInference, the training forward / backward of MultiModalEncoder).
static int g_persistent_cta_cap = 0;
void set_persistent_cta_cap(int n) { g_persistent_cta_cap = n > 0 ? n : 0; }
int persistent_ctas() {
  const int sms = sm_count();
  return (g_persistent_cta_cap > 0 && g_persistent_cta_cap < sms) ? g_persistent_cta_cap : sms;
}

}  // namespace sga

extern "C" {

const char* sga_last_error(void) { return sga::g_err; }

int sga_version(void) { return 100; }

int sga_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  SGA_CUDA(cudaGetDevice(&dev));
  int n = 0, ma = 0, mi = 0;
  SGA_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  SGA_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
  SGA_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  return SGA_OK;
}

}  // extern "C"
