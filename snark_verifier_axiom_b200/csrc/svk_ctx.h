// Internal context shared by the libsvk translation units.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/svk.h"
#include "coop_pairing.cuh"
#include "poseidon.cuh"

struct DkDevice {
  G2Line* d_lines_g2 = nullptr;       // SVK_N_LINES
  G2Line* d_lines_neg_sg2 = nullptr;  // SVK_N_LINES
  G2LineX* d_linesx_g2 = nullptr;       // the same lines with their xi-images (coop_pairing.cuh)
  G2LineX* d_linesx_neg_sg2 = nullptr;
  G1Affine g1;                        // svk.g (Montgomery)
  svk_g1 g1_canon;
};

struct KernelStat {
  uint64_t count = 0;
  double ms = 0;
};
struct PendingEvent {
  const char* name;
  cudaEvent_t e0, e1;
};

struct svk_ctx {
  bool profile = false;  // svk_profile_enable: CUDA events around every kernel launch
  cudaEvent_t profile_ref = nullptr;  // device-clock reference of svk_profile_timeline
  std::vector<PendingEvent> pending;
  std::map<std::string, KernelStat> stats;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t done = nullptr;  // blocking-sync event: host-buffer calls sleep instead of spinning (many contexts, few cores)
  std::string err;
  uint64_t launches = 0;
  int sm_count = 0;
  size_t msm_latency_threads_max = 150000;  // per-proof MSM: one thread per TERM while proofs x terms stays under this
  size_t fold_lanes_groups_max = 2048;  // wide fold levels (Straus path): one thread per MEMBER while (groups x sides) stays under this
  size_t fold_dbl_threads_max = 32768;  // fold levels with at most this many (accumulator, side) pairs precompute the doublings beside the sponge
  size_t tape_coop_max = 32768;   // launches with at most this many proofs (sponges) run the warp-cooperative Poseidon (poseidon_coop.cuh)
  size_t decide_coop_max = 512;  // decide calls with at most this many accumulators run one accumulator per BLOCK (k_decide_coop)
  PairingConsts* d_pairing_consts = nullptr;
  PoseidonConsts* d_poseidon = nullptr;
  PoseidonConsts h_poseidon;
  std::vector<DkDevice> dks;
  // scratch buffers (grown on demand, reused across calls)
  void* scratch[24] = {nullptr};
  size_t scratch_sz[24] = {0};
  std::vector<struct ProtocolDevice*> protocols;
  // calls on one context are serialised inside the library (entry points nest: verify_batch -> verify_multi -> ...)
  std::recursive_mutex mu;
  // proof-sharded jobs (csrc/sharded.cu): this rank's NCCL communicator (ncclComm_t), created by svk_nccl_init or attached
  void* nccl_comm = nullptr;
  bool nccl_owned = false;
  int world = 1, rank = 0;
};
// NVTX range per C-ABI call (header-only NVTX 3: a no-op unless a profiler injects itself), so nsys / ncu timelines show the
// reference-level operation (verify batch, fold, decide, msm) above the kernels it launched.
struct SvkNvtxRange {
  explicit SvkNvtxRange(const char* name) { nvtxRangePushA(name); }
  ~SvkNvtxRange() { nvtxRangePop(); }
};
#define SVK_LOCK(ctx)                                              \
  std::lock_guard<std::recursive_mutex> svk_lock_((ctx)->mu);      \
  SvkNvtxRange svk_nvtx_(__func__)

inline int svk_fail(svk_ctx* ctx, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return -1;
}

#define SVK_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return svk_fail(ctx, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

inline int svk_scratch(svk_ctx* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_sz[slot] < bytes) {
    if (ctx->scratch[slot]) cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratch_sz[slot] = 0;
    size_t want = bytes + bytes / 4 + 256;
    SVK_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], want));
    ctx->scratch_sz[slot] = want;
  }
  *out = ctx->scratch[slot];
  return 0;
}

// Kernel launch wrapper: counts the launch and, when profiling is on, brackets it with CUDA events
// on the launching stream (resolved in svk_profile_report).
#define SVK_LAUNCH(ctx, name, ...)                                   \
  do {                                                               \
    PendingEvent pe_{name, nullptr, nullptr};                        \
    if ((ctx)->profile) {                                            \
      cudaEventCreate(&pe_.e0);                                      \
      cudaEventCreate(&pe_.e1);                                      \
      cudaEventRecord(pe_.e0, (ctx)->stream);                        \
    }                                                                \
    __VA_ARGS__;                                                     \
    (ctx)->launches++;                                               \
    if ((ctx)->profile) {                                            \
      cudaEventRecord(pe_.e1, (ctx)->stream);                        \
      (ctx)->pending.push_back(pe_);                                 \
    }                                                                \
  } while (0)

// Wait for everything enqueued on the context stream without burning a core.
inline int svk_wait(svk_ctx* ctx) {
  if (!ctx->done) SVK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->done, cudaEventBlockingSync | cudaEventDisableTiming));
  SVK_CUDA(ctx, cudaEventRecord(ctx->done, ctx->stream));
  SVK_CUDA(ctx, cudaEventSynchronize(ctx->done));
  return 0;
}
