// libsvk C-ABI entry points (include/svk.h): context, deciding key, decide.
#include <cstring>

#include "pairing_host.h"
#include "svk_ctx.h"

int svk_decide_launch(svk_ctx* ctx, int dk, size_t n, const void* d_accs, void* d_ok);
int svk_modmul_peak_launch(svk_ctx* ctx, int iters, double* out_rate, double* out_ms);

static thread_local std::string g_create_err;

extern "C" {

int svk_create(int device, svk_ctx** out) {
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libsvk has no CPU fallback)";
    return -1;
  }
  if (device < 0 || device >= n) { g_create_err = "device index out of range"; return -1; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    g_create_err = "libsvk is built for sm_100a only";
    return -1;
  }
  if (cudaSetDevice(device) != cudaSuccess) { g_create_err = "cudaSetDevice failed"; return -1; }
  svk_ctx* ctx = new svk_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; g_create_err = "stream create failed"; return -1; }
  ctx->own_stream = true;
  PairingConsts k = svk_host::make_pairing_consts();
  if (cudaMalloc(&ctx->d_pairing_consts, sizeof k) != cudaSuccess ||
      cudaMemcpy(ctx->d_pairing_consts, &k, sizeof k, cudaMemcpyHostToDevice) != cudaSuccess) {
    delete ctx; g_create_err = "pairing constants upload failed"; return -1;
  }
  *out = ctx;
  return 0;
}

void svk_destroy(svk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& k : ctx->dks) { cudaFree(k.d_lines_g2); cudaFree(k.d_lines_neg_sg2); }
  for (int i = 0; i < 8; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  cudaFree(ctx->d_pairing_consts);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* svk_last_error(svk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int svk_set_stream(svk_ctx* ctx, void* s) {
  if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
  ctx->stream = (cudaStream_t)s;
  return 0;
}

int svk_sync(svk_ctx* ctx) { SVK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return 0; }

uint64_t svk_launch_count(svk_ctx* ctx) { return ctx->launches; }

static bool load_fq_canon(Fq& out, const svk_fe& fe) {
  fe_load_le(out.v, fe.b);
  if (!Fq::is_canonical(out.v)) return false;
  out = out.to_mont();
  return true;
}

int svk_dk_load(svk_ctx* ctx, const svk_deciding_key* dk) {
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  G2Affine g2, sg2;
  bool ok = load_fq_canon(g2.x.c0, dk->g2.x_c0) && load_fq_canon(g2.x.c1, dk->g2.x_c1) && load_fq_canon(g2.y.c0, dk->g2.y_c0) &&
            load_fq_canon(g2.y.c1, dk->g2.y_c1) && load_fq_canon(sg2.x.c0, dk->s_g2.x_c0) && load_fq_canon(sg2.x.c1, dk->s_g2.x_c1) &&
            load_fq_canon(sg2.y.c0, dk->s_g2.y_c0) && load_fq_canon(sg2.y.c1, dk->s_g2.y_c1);
  if (!ok) return svk_fail(ctx, "deciding key: non-canonical G2 coordinate");
  if (!svk_host::g2_on_curve(g2) || !svk_host::g2_on_curve(sg2)) return svk_fail(ctx, "deciding key: G2 point not on the twist");
  DkDevice d;
  d.g1_canon = dk->g1;
  bool id = true;
  for (int i = 0; i < 32; i++) id = id && dk->g1.x.b[i] == 0 && dk->g1.y.b[i] == 0;
  if (id) return svk_fail(ctx, "deciding key: g1 is the identity");
  if (!load_fq_canon(d.g1.x, dk->g1.x) || !load_fq_canon(d.g1.y, dk->g1.y) || !g1_on_curve(d.g1))
    return svk_fail(ctx, "deciding key: g1 not on curve");
  PairingConsts k = svk_host::make_pairing_consts();
  G2Affine neg_sg2 = {sg2.x, sg2.y.neg()};  // `-dk.s_g2` (decider.rs:64)
  std::vector<G2Line> t1 = svk_host::make_line_table(g2, k), t2 = svk_host::make_line_table(neg_sg2, k);
  size_t bytes = sizeof(G2Line) * SVK_N_LINES;
  SVK_CUDA(ctx, cudaMalloc(&d.d_lines_g2, bytes));
  SVK_CUDA(ctx, cudaMalloc(&d.d_lines_neg_sg2, bytes));
  SVK_CUDA(ctx, cudaMemcpy(d.d_lines_g2, t1.data(), bytes, cudaMemcpyHostToDevice));
  SVK_CUDA(ctx, cudaMemcpy(d.d_lines_neg_sg2, t2.data(), bytes, cudaMemcpyHostToDevice));
  ctx->dks.push_back(d);
  return (int)ctx->dks.size() - 1;
}

int svk_kzg_decide_batch_dev(svk_ctx* ctx, int dk, size_t n, const void* d_accs, void* d_out_ok) {
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_decide_launch(ctx, dk, n, d_accs, d_out_ok);
}

int svk_kzg_decide_batch(svk_ctx* ctx, int dk, size_t n, const svk_acc* accs, uint8_t* out_ok) {
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  void *d_in, *d_out;
  if (svk_scratch(ctx, 0, n * sizeof(svk_acc), &d_in) || svk_scratch(ctx, 1, n, &d_out)) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(d_in, accs, n * sizeof(svk_acc), cudaMemcpyHostToDevice, ctx->stream));
  if (svk_decide_launch(ctx, dk, n, d_in, d_out)) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out_ok, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
  SVK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int svk_bench_modmul_peak(svk_ctx* ctx, int iters, double* out_modmul_per_s, double* out_ms) {
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_modmul_peak_launch(ctx, iters, out_modmul_per_s, out_ms);
}

}  // extern "C"
