// Library-level entry points: ABI version, thread-local error string, device check.
#include "common.cuh"
#include <string.h>

namespace vqa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return VQA_ECUDA;
  }
  return VQA_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace vqa

extern "C" int vqa_abi_version(void) { return VQA_ABI_VERSION; }
extern "C" const char* vqa_last_error(void) { return vqa::g_err; }

extern "C" int vqa_device_check(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    vqa::set_error("vqa_device_check: no CUDA device (this library has no CPU fallback)");
    return VQA_ENODEVICE;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    vqa::set_error("vqa_device_check: device %d has compute capability %d.x; libvqacore is built for sm_100a only", dev,
                   major);
    return VQA_ENODEVICE;
  }
  return VQA_OK;
}

// sizeof() of every parameter struct, so a binding in another language can verify its layout
// (tests/test_abi.py checks the ctypes mirror in _lib.py against these).
extern "C" size_t vqa_sizeof(const char* name) {
#define VQA_SZ(T) if (strcmp(name, #T) == 0) return sizeof(T)
  VQA_SZ(vqa_dropout);
  VQA_SZ(vqa_linear_fwd_params);
  VQA_SZ(vqa_linear_bwd_params);
  VQA_SZ(vqa_mutan_fwd_params);
  VQA_SZ(vqa_mutan_bwd_params);
  VQA_SZ(vqa_region_softmax_pool_fwd_params);
  VQA_SZ(vqa_region_softmax_pool_bwd_params);
  VQA_SZ(vqa_cor_compound_fwd_params);
  VQA_SZ(vqa_cor_compound_bwd_params);
  VQA_SZ(vqa_oda_pair_attn_fwd_params);
  VQA_SZ(vqa_oda_pair_attn_bwd_params);
  VQA_SZ(vqa_kld_logsoftmax_params);
  VQA_SZ(vqa_model_fwd_params);
  VQA_SZ(vqa_model_bwd_params);
#undef VQA_SZ
  return 0;
}
