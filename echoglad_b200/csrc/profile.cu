// Optional in-library profiler: when enabled, every launch helper brackets its kernels with CUDA events
// on the caller's stream, so bench.py can report the live per-kernel device time of the timed region
// (roofline.achieved) without a profiler attached.  Disabled (the default) it costs one branch.
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace eg {

std::atomic<long long> g_launches{0};

int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    if (n > kNumSMs) n = kNumSMs;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

bool first_use_on_current_device(std::atomic<unsigned long long>& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;  // unknown device: always set up
  const unsigned long long bit = 1ull << dev;
  return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}
static std::atomic<int> g_profile_on{0};

struct Span {
  cudaEvent_t beg, end;
};
static std::mutex g_mu;
static std::map<std::string, std::vector<Span>> g_spans;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;  // the scope then records nothing (see ProfileScope)
  return e;
}

ProfileScope::ProfileScope(const char* name, cudaStream_t s) : name_(name), stream_(s), on_(g_profile_on.load() != 0) {
  if (!on_) return;
  std::lock_guard<std::mutex> lk(g_mu);
  beg_ = get_event();
  end_ = get_event();
  if (!beg_ || !end_) {  // event creation failed: give back what we got and profile nothing for this scope
    if (beg_) g_pool.push_back(reinterpret_cast<cudaEvent_t>(beg_));
    if (end_) g_pool.push_back(reinterpret_cast<cudaEvent_t>(end_));
    on_ = false;
    return;
  }
  cudaEventRecord(reinterpret_cast<cudaEvent_t>(beg_), stream_);
}

ProfileScope::~ProfileScope() {
  if (!on_) return;
  cudaEventRecord(reinterpret_cast<cudaEvent_t>(end_), stream_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_spans[name_].push_back(Span{reinterpret_cast<cudaEvent_t>(beg_), reinterpret_cast<cudaEvent_t>(end_)});
}

}  // namespace eg

using namespace eg;

extern "C" {

int eg_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_spans)
    for (auto& sp : kv.second) {
      g_pool.push_back(sp.beg);
      g_pool.push_back(sp.end);
    }
  g_spans.clear();
  g_profile_on.store(on ? 1 : 0);
  return EG_OK;
}

int eg_profile_read(const char* name, double* total_ms, int64_t* launches) {
  EG_CHECK_ARG(name && total_ms && launches, "eg_profile_read: NULL argument");
  std::lock_guard<std::mutex> lk(g_mu);
  *total_ms = 0.0;
  *launches = 0;
  auto it = g_spans.find(name);
  if (it == g_spans.end()) return EG_OK;
  for (auto& sp : it->second) {
    EG_CUDA(cudaEventSynchronize(sp.end));
    float ms = 0.f;
    EG_CUDA(cudaEventElapsedTime(&ms, sp.beg, sp.end));
    *total_ms += ms;
    *launches += 1;
  }
  return EG_OK;
}

int eg_profile_names(char* buf, size_t n) {
  EG_CHECK_ARG(buf && n > 0, "eg_profile_names: bad buffer");
  std::lock_guard<std::mutex> lk(g_mu);
  std::string s;
  for (auto& kv : g_spans) {
    if (!s.empty()) s += ",";
    s += kv.first;
  }
  snprintf(buf, n, "%s", s.c_str());
  return EG_OK;
}

int64_t eg_launch_count(void) { return (int64_t)g_launches.load(); }

}  // extern "C"
