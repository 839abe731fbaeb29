// Shared helpers for the echoglad_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/echoglad_b200.h"

#include <atomic>

namespace eg {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// brackets the kernels of one launch helper with CUDA events when eg_profile_enable(1) is active
class ProfileScope {
 public:
  ProfileScope(const char* name, cudaStream_t s);
  ~ProfileScope();
 private:
  const char* name_;
  cudaStream_t stream_;
  bool on_;
  void* beg_ = nullptr;
  void* end_ = nullptr;
};

#define EG_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      eg::set_error(__VA_ARGS__);               \
      return EG_ERR_INVALID;                    \
    }                                           \
  } while (0)

#define EG_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      eg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return EG_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

// after every kernel launch: counts the launch (eg_launch_count) and surfaces launch errors
#define EG_LAUNCH_CHECK()                          \
  do {                                             \
    eg::g_launches.fetch_add(1, std::memory_order_relaxed); \
    EG_CUDA(cudaGetLastError());                   \
  } while (0)

constexpr int kNumSMs = 148;            // B200: UPPER BOUND used to size workspaces / partial-slot tables at compile time
constexpr int kMaxParts = 4 * kNumSMs;  // upper bound on per-CTA partial slots of any reduction
// SM count of the CURRENT device (cudaDevAttrMultiProcessorCount, cached per device), capped at kNumSMs so that
// every grid sized from it fits the workspace layout: the launch helpers size persistent grids with this.
int num_sms();
// true exactly once per (mask, current device): guards per-device one-time setup such as
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), which is a per-device attribute.  Thread-safe.
bool first_use_on_current_device(std::atomic<unsigned long long>& mask);
// workspace: [kMaxParts][2][128] doubles for column statistics + [kNumSMs][128*128] floats for the
// weight-gradient partials + slack
constexpr size_t kStatsBytes = (size_t)kMaxParts * 2 * 128 * sizeof(double);
constexpr size_t kWgradBytes = (size_t)kNumSMs * 128 * 128 * sizeof(float);
constexpr size_t kWorkspaceBytes = kStatsBytes + kWgradBytes + 65536;

// Per-tile gather plan of the fused tensor-core kernel (built on the host in graph.cu, device arrays).
// The unique source rows a tile reads most (own rows, lattice halo, parents) are STAGED in shared memory
// once per K chunk; a tile row then sums up to 7 staged neighbours, up to 4 FAR neighbours read straight
// from global (the 2x2 children of an aux node) and its self loop (staged, always last).  Rows that do
// not fit this shape (hubs of use_connection_nodes, 'grid-diagonal' aux rows) are CSR rows: the plan
// weights are zero and the row is summed from the device CSR instead.
constexpr int kPlanSrc = 216;   // staged source rows per tile (216 x 128 B = 27 KB per K chunk)
constexpr int kPlanStaged = 7;  // staged non-self edges per row
constexpr int kPlanFar = 4;     // far edges per row
struct alignas(16) PlanRow {    // 80 bytes
  uint8_t slot[8];     // [0..6] staged neighbours in summation order, [7] the row itself (self loop)
  int32_t csr_beg;     // CSR rows: first CSR entry, else 0
  int32_t csr_deg;     // CSR rows: entries incl. the self loop, else 0
  float w[8];          // deg^-1/2[u] deg^-1/2[v] of slot[k]; 0 = unused (slot 0 is always a valid row)
  int32_t far_node[4]; // frame-local node ids read from global; unused = node 0 with weight 0
  float far_w[4];
};
static_assert(sizeof(PlanRow) == 80, "PlanRow layout");
struct TilePlan {
  const int4* hdr;      // [tiles] {staged source rows, max staged non-self edges, far edges (0 / 4), has CSR rows}
  const int32_t* src;   // [tiles][kPlanSrc] staged node ids, ascending (-1 = unused)
  const PlanRow* rows;  // [tiles][128]
};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- counter-based dropout RNG: one 64-bit mix per group of 4 consecutive elements -----------------
// keep(element e) <=> 16-bit lane of mix(seed, e/4) >= p * 65536.  Stateless, so backward recomputes it.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  float t = p * 65536.0f;
  return t <= 0.f ? 0u : (t >= 65536.f ? 65536u : (uint32_t)t);
}
// returns 4 keep-flags (bit i = element 4*group+i kept)
__host__ __device__ __forceinline__ uint32_t drop_keep4(uint64_t seed, uint64_t group, uint32_t thr) {
  uint64_t r = mix64(seed ^ (group * 0x9e3779b97f4a7c15ULL));
  uint32_t k = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) k |= (((uint32_t)(r >> (16 * i)) & 0xffffu) >= thr ? 1u : 0u) << i;
  return k;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (evict-first) 128-bit store for outputs that are not re-read by this kernel
__device__ __forceinline__ void st4_stream(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

}  // namespace eg
