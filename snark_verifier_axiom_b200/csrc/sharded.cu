// Proof-sharded verification behind the C ABI (SURVEY 8e / 8b "one ctx per rank + an ncclComm_t"): every rank verifies and
// folds its own shard, the per-rank batch records (256 B each) are all-gathered with ncclAllGather on the context stream,
// every rank folds batch b over the ranks (`KzgAs`, flat: snark-verifier/src/pcs/kzg/accumulation.rs:29-62) and runs the single
// pairing (decider.rs:60-68).  EC addition is not an ncclRedOp, so "reduce" = all-gather + fold.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded, or the system one), so libsvk.so itself
// has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>

#include <mutex>

#include "svk_ctx.h"
#include "svk_protocol.h"

int svk_fold_launch_seg(svk_ctx* ctx, size_t n_seg, size_t n, const uint8_t* d_accs, size_t group_size, uint8_t* d_out, size_t out_stride);
int svk_decide_launch_strided(svk_ctx* ctx, int dk, size_t n, const void* d_accs, size_t acc_stride, void* d_ok, size_t ok_stride);

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
const int kNcclUint8 = 1;  // ncclDataType_t::ncclUint8

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
  });
  return api;
}
}  // namespace

// [rank][batch] records (256 B) -> accumulators [batch][rank] (128 B)
__global__ void k_gather_accs(u32 world, u32 nb, const uint4* gathered, uint4* accs) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;  // one uint4 (16 B) per thread: 8 per accumulator
  if (t >= world * nb * 8) return;
  u32 q = t & 7, ra = t >> 3, r = ra % world, b = ra / world;
  accs[(size_t)(b * world + r) * 8 + q] = gathered[(size_t)(r * nb + b) * 16 + q];
}

// final record of batch b: ok = every rank's local ok (all proofs read and verified succinctly, local fold found only curve
// points) && cross-rank fold status == 0 && the single pairing accepted
__global__ void k_sharded_verdict(u32 world, u32 nb, const uint8_t* gathered, uint8_t* final_records) {
  u32 b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  bool ok = true;
  for (u32 r = 0; r < world; r++) ok = ok && gathered[(size_t)(r * nb + b) * 256 + 165];
  uint8_t* rec = final_records + (size_t)b * 256;
  int32_t fold_status = *reinterpret_cast<const int32_t*>(rec + 160);
  rec[165] = (ok && fold_status == 0 && rec[164]) ? 1 : 0;
}

extern "C" {

int svk_plonk_fold_multi_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                             const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                             void* d_out_status, void* d_out_records);
int svk_plonk_verify_multi_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                               const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                               void* d_out_status, void* d_out_records);

int svk_nccl_unique_id(uint8_t* out_id) {
  NcclApi& a = nccl();
  if (!a.ok) return -1;
  ncclUniqueId id;
  if (a.GetUniqueId(&id) != 0) return -1;
  memcpy(out_id, id.internal, 128);
  return 0;
}

// Collective over the `world` ranks of the job (one context per rank, each on its own GPU): creates this context's communicator.
int svk_nccl_init(svk_ctx* ctx, int world, int rank, const uint8_t* id_bytes) {
  SVK_LOCK(ctx);
  NcclApi& a = nccl();
  if (!a.ok) return svk_fail(ctx, "NCCL is not available (dlopen of libnccl.so.2 failed)");
  if (world < 1 || rank < 0 || rank >= world) return svk_fail(ctx, "bad world / rank");
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = a.CommInitRank(&comm, world, id, rank);
  if (r != 0) return svk_fail(ctx, "ncclCommInitRank: %s", a.GetErrorString(r));
  if (ctx->nccl_comm && ctx->nccl_owned) a.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = comm;
  ctx->nccl_owned = true;
  ctx->world = world;
  ctx->rank = rank;
  return 0;
}

// The same with a communicator the host already owns (`ncclComm_t` as void*); it is not destroyed with the context.
int svk_nccl_attach(svk_ctx* ctx, void* comm, int world, int rank) {
  SVK_LOCK(ctx);
  if (!nccl().ok) return svk_fail(ctx, "NCCL is not available (dlopen of libnccl.so.2 failed)");
  if (ctx->nccl_comm && ctx->nccl_owned) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = comm;
  ctx->nccl_owned = false;
  ctx->world = world;
  ctx->rank = rank;
  return 0;
}

void svk_nccl_release(svk_ctx* ctx) {
  if (ctx->nccl_comm && ctx->nccl_owned && nccl().ok) nccl().CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

// This rank's shard: n_batches batches of batch_size proofs.  d_out_records: n_batches x 256 B (local), d_gather: world x
// n_batches x 256 B, d_final_records: n_batches x 256 B { global accumulator ; r ; fold_status ; decide_ok ; ok }.
// Everything is enqueued on the context stream; no host synchronisation.
int svk_plonk_verify_sharded_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                                 const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                                 void* d_out_status, void* d_out_records, void* d_gather, void* d_final_records) {
  SVK_LOCK(ctx);
  if (!ctx->nccl_comm || ctx->world <= 1) {  // a job of one rank: the plain call; the final records are the local ones
    if (svk_plonk_verify_multi_dev(ctx, proto, n_batches, batch_size, d_instances, n_instances, d_proofs, proof_stride, d_proof_lens, group_size,
                                   d_out_accs, d_out_status, d_out_records))
      return -1;
    if (d_final_records && d_final_records != d_out_records)
      SVK_CUDA(ctx, cudaMemcpyAsync(d_final_records, d_out_records, n_batches * 256, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  if (svk_plonk_fold_multi_dev(ctx, proto, n_batches, batch_size, d_instances, n_instances, d_proofs, proof_stride, d_proof_lens, group_size,
                               d_out_accs, d_out_status, d_out_records))
    return -1;
  NcclApi& a = nccl();
  ncclResult_t r = a.AllGather(d_out_records, d_gather, n_batches * 256, kNcclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
  if (r != 0) return svk_fail(ctx, "ncclAllGather: %s", a.GetErrorString(r));
  u32 world = (u32)ctx->world, nb = (u32)n_batches;
  uint8_t* d_accs;
  if (svk_scratch(ctx, 23, (size_t)world * nb * 128 + 256, (void**)&d_accs)) return -1;
  u32 total = world * nb * 8;
  SVK_LAUNCH(ctx, "k_gather_accs", k_gather_accs<<<(total + 127) / 128, 128, 0, ctx->stream>>>(world, nb, (const uint4*)d_gather, (uint4*)d_accs));
  if (svk_fold_launch_seg(ctx, nb, world, d_accs, 0, (uint8_t*)d_final_records, 256)) return -1;
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  if (svk_decide_launch_strided(ctx, ctx->protocols[proto]->dk, nb, d_final_records, 256, (uint8_t*)d_final_records + 164, 256)) return -1;
  SVK_LAUNCH(ctx, "k_sharded_verdict",
             k_sharded_verdict<<<(nb + 63) / 64, 64, 0, ctx->stream>>>(world, nb, (const uint8_t*)d_gather, (uint8_t*)d_final_records));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
