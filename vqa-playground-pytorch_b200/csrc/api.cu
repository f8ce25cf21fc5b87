// Library-level entry points: ABI version, thread-local error string, device check.
#include "common.cuh"
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

namespace vqa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
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

// ---- optional per-op timing of the whole-model plans (CUDA events on the caller's stream) -----
struct ProfRec { std::string name; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};

ProfScope::ProfScope(void* stream, const char* name) : idx(-1), st(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, (cudaStream_t)stream);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  idx = (int)g_prof.size() - 1;
}
bool prof_active() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
ProfScope::~ProfScope() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[idx].b, (cudaStream_t)st);
}

}  // namespace vqa

extern "C" unsigned long long vqa_launch_count(void) { return vqa::g_launches.load(); }

extern "C" int vqa_profile_begin(void) {
  std::lock_guard<std::mutex> lk(vqa::g_prof_mu);
  for (auto& r : vqa::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  vqa::g_prof.clear();
  vqa::g_prof_on.store(1);
  return VQA_OK;
}

// Writes "name=total_ms/count;..." (aggregated by name, first-seen order) into buf. Synchronises the events.
extern "C" int vqa_profile_end(char* buf, size_t cap) {
  using namespace vqa;
  g_prof_on.store(0);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::vector<std::string> names;
  std::vector<double> tot;
  std::vector<int> cnt;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) {
      cudaGetLastError();
      ms = 0.f;
    }
    size_t i = 0;
    for (; i < names.size(); ++i) if (names[i] == r.name) break;
    if (i == names.size()) { names.push_back(r.name); tot.push_back(0); cnt.push_back(0); }
    tot[i] += ms; cnt[i] += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::string out;
  char tmp[256];
  for (size_t i = 0; i < names.size(); ++i) {
    snprintf(tmp, sizeof(tmp), "%s=%.6f/%d;", names[i].c_str(), tot[i], cnt[i]);
    out += tmp;
  }
  if (buf && cap) { strncpy(buf, out.c_str(), cap - 1); buf[cap - 1] = 0; }
  return out.size() < cap ? VQA_OK : VQA_EINVAL;
}

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
  VQA_SZ(vqa_pack_segment);
  VQA_SZ(vqa_bits_segment);
  VQA_SZ(vqa_param_segment);
  VQA_SZ(vqa_clip_adam_params);
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
  VQA_SZ(vqa_peer_allreduce_params);
  VQA_SZ(vqa_gru_gate_fwd_params);
  VQA_SZ(vqa_gru_gate_bwd_params);
  VQA_SZ(vqa_model_fwd_params);
  VQA_SZ(vqa_model_bwd_params);
#undef VQA_SZ
  return 0;
}
