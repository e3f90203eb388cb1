// libsvk C-ABI entry points (include/svk.h): context, deciding key, decide.
#include <cstring>

#include "compiler.h"
#include "pairing_host.h"
#include "poseidon_host.h"
#include "svk_ctx.h"
#include "svk_protocol.h"

int svk_decide_launch(svk_ctx* ctx, int dk, size_t n, const void* d_accs, void* d_ok);
int svk_modmul_peak_launch(svk_ctx* ctx, int iters, double* out_rate, double* out_ms);
int svk_msm_launch(svk_ctx* ctx, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, uint8_t* d_out, int* d_status);
int svk_g1_mul_batch_launch(svk_ctx* ctx, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, size_t n_points, uint8_t* d_out);
int svk_msm_curve_launch(svk_ctx* ctx, int curve, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, uint8_t* d_out, int* d_status);
int svk_ipa_decide_launch(svk_ctx* ctx, int curve, u32 k, const uint8_t* d_g, size_t n, const uint8_t* d_xi, const uint8_t* d_u,
                          int32_t* d_out_status, int32_t* d_invalid);
int svk_fixed_tables_launch(svk_ctx* ctx, ProtocolDevice* pd);
int svk_blind_points_launch(svk_ctx* ctx, const uint8_t* d_blind, uint8_t* d_slot);
int svk_poseidon_squeeze_launch(svk_ctx* ctx, size_t n, u32 n_in, const uint8_t* d_inputs, uint8_t* d_out, int coop, int* d_bad);
int svk_fold_launch(svk_ctx* ctx, size_t n, const uint8_t* d_accs, size_t group_size, uint8_t* d_out_acc, u32* d_out_r, int32_t* d_status);
int svk_fold_launch_seg(svk_ctx* ctx, size_t n_seg, size_t n, const uint8_t* d_accs, size_t group_size, uint8_t* d_out, size_t out_stride);
int svk_decide_launch_strided(svk_ctx* ctx, int dk, size_t n, const void* d_accs, size_t acc_stride, void* d_ok, size_t ok_stride);
int svk_batch_verdict_launch(svk_ctx* ctx, size_t n_seg, size_t batch, const int32_t* d_status, uint8_t* d_records, size_t rec_stride);
int svk_succinct_verify_launch(svk_ctx* ctx, ProtocolDevice* pd, size_t n, const uint8_t* d_instances, u32 n_instances_given,
                               const uint8_t* d_proofs, size_t proof_stride, const u32* d_proof_lens, uint8_t* d_out_acc,
                               u32* d_out_challenges, int32_t* d_out_status);

extern "C" void svk_nccl_release(svk_ctx* ctx);
static thread_local std::string g_create_err;

// Fixed-base window tables are identical for every context that compiles the same verifying key on the same
// device (bench.py keeps 32 contexts in flight): share one device copy so the 4.7 MB table stays L2-resident.
#include <mutex>
struct SharedTable { G1Affine* d = nullptr; int refs = 0; bool ready = false; };
static std::mutex g_table_mu;
static std::map<std::string, SharedTable> g_tables;

template <class T>
static int upload(svk_ctx* ctx, T** d, const std::vector<T>& v) {
  *d = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  SVK_CUDA(ctx, cudaMalloc(d, bytes));
  if (!v.empty()) SVK_CUDA(ctx, cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// Schedule of the per-proof MSM (verify.cu: k_msm_var + k_msm_sum).  Variable-base terms become k_msm_var items (Straus,
// straus.cuh); in k_msm_sum every one of the SVK_MSM_LANES lanes of a (proof, side) gets an EQUAL number of fixed-base table
// windows (one mixed addition each; vk commitments and g), and the partial sums / scalar == 1 bases are dealt round-robin.
// Returns the algorithmic Fq mults per proof for this side (SURVEY 8d counting).
// `var_lane_base`: index of this side's first k_msm_var lane; `var_lanes`: how many lanes this side may use.
static size_t schedule_msm(const std::vector<MsmTermDev>& terms, std::vector<std::vector<MsmWork>>& var_lanes_items, u32 var_lane_base,
                           u32 var_lanes, std::vector<MsmWork>& work, std::vector<u32>& lane_off, std::vector<FixedSlot>& fixed_sched,
                           u32& fixed_per, u32 fixed_bits, int L) {
  const size_t FW = 256 / fixed_bits;  // table windows per fixed base
  std::vector<std::vector<MsmWork>> lanes(L);
  size_t total = 0;
  int rr = 0;
  u32 n_var = 0;
  std::vector<MsmTermDev> fixed_terms;
  for (auto& t : terms) {
    if (t.slot < 0) { lanes[rr++ % L].push_back({2, t.fixed, t.base, -1, 0, 0}); total += 11; }
    else if (t.fixed) fixed_terms.push_back(t);
    else {
      var_lanes_items[var_lane_base + (n_var % var_lanes)].push_back({0, 0, t.base, t.slot, 0, 0});
      n_var++;
      total += 133 + 50 * 11 + 16 * 7;
    }
  }
  u32 used = std::min(n_var, var_lanes);
  for (u32 l = 0; l < used; l++) {  // this side's k_msm_var partial sums
    lanes[rr++ % L].push_back({3, 0, (int32_t)(var_lane_base + l), -1, 0, 0});
    total += 255 * 7 + 380 + 16;
  }
  // fixed-base table additions: a flat list of (term, window) pairs dealt evenly, `per` per lane, padded with no-ops
  size_t fixed_windows = fixed_terms.size() * FW;
  size_t per = (fixed_windows + L - 1) / L;
  fixed_per = (u32)per;
  fixed_sched.assign(per * L, FixedSlot{-1, 0, 0});
  for (size_t i = 0; i < fixed_windows; i++) {
    const MsmTermDev& ft = fixed_terms[i / FW];
    size_t lane = i / per, j = i % per;
    fixed_sched[lane * per + j] = FixedSlot{ft.base, (int32_t)(i % FW), ft.slot};
    total += 11;
  }
  work.clear(); lane_off.assign(L + 1, 0);
  for (int l = 0; l < L; l++) { lane_off[l] = (u32)work.size(); work.insert(work.end(), lanes[l].begin(), lanes[l].end()); }
  lane_off[L] = (u32)work.size();
  return total + (L - 1) * 16 + 385;  // + lane-tree additions (none for one lane) + to_affine
}

// Frees everything a ProtocolDevice owns (svk_destroy and every error path of protocol_upload).
static void protocol_release(ProtocolDevice* p) {
  if (!p) return;
  cudaFree(p->d_ops); cudaFree(p->d_aux); cudaFree(p->d_consts); cudaFree(p->d_sched); cudaFree(p->d_lhs); cudaFree(p->d_rhs);
  cudaFree(p->d_fixed); cudaFree(p->d_old_idx);
  for (auto& sc : p->sched) {
    cudaFree(sc.d_var_items); cudaFree(sc.d_var_lane_off); cudaFree(sc.d_work_lhs); cudaFree(sc.d_work_rhs);
    cudaFree(sc.d_lane_off_lhs); cudaFree(sc.d_lane_off_rhs); cudaFree(sc.d_fixed_lhs); cudaFree(sc.d_fixed_rhs);
  }
  if (p->d_fixed_tables) {
    std::lock_guard<std::mutex> lk(g_table_mu);
    auto it = g_tables.find(p->table_key);
    if (it != g_tables.end() && --it->second.refs == 0) { cudaFree(it->second.d); g_tables.erase(it); }
  }
  delete p;
}

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
  if (const char* e = getenv("SVK_DECIDE_COOP_MAX")) ctx->decide_coop_max = (size_t)atoll(e);
  if (const char* e = getenv("SVK_TAPE_COOP_MAX")) ctx->tape_coop_max = (size_t)atoll(e);
  if (const char* e = getenv("SVK_FOLD_DBL_THREADS_MAX")) ctx->fold_dbl_threads_max = (size_t)atoll(e);
  if (const char* e = getenv("SVK_FOLD_LANES_GROUPS_MAX")) ctx->fold_lanes_groups_max = (size_t)atoll(e);
  if (const char* e = getenv("SVK_MSM_LATENCY_THREADS_MAX")) ctx->msm_latency_threads_max = (size_t)atoll(e);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; g_create_err = "stream create failed"; return -1; }
  ctx->own_stream = true;
  PairingConsts k = svk_host::make_pairing_consts();
  if (cudaMalloc(&ctx->d_pairing_consts, sizeof k) != cudaSuccess ||
      cudaMemcpy(ctx->d_pairing_consts, &k, sizeof k, cudaMemcpyHostToDevice) != cudaSuccess) {
    delete ctx; g_create_err = "pairing constants upload failed"; return -1;
  }
  if (!svk_host::make_poseidon_consts(ctx->h_poseidon) || cudaMalloc(&ctx->d_poseidon, sizeof(PoseidonConsts)) != cudaSuccess ||
      cudaMemcpy(ctx->d_poseidon, &ctx->h_poseidon, sizeof(PoseidonConsts), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete ctx; g_create_err = "poseidon constants upload failed"; return -1;
  }
  *out = ctx;
  return 0;
}

void svk_destroy(svk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& k : ctx->dks) { cudaFree(k.d_lines_g2); cudaFree(k.d_lines_neg_sg2); cudaFree(k.d_linesx_g2); cudaFree(k.d_linesx_neg_sg2); }
  for (int i = 0; i < 24; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  cudaFree(ctx->d_pairing_consts);
  cudaFree(ctx->d_poseidon);
  for (auto* p : ctx->protocols) protocol_release(p);
  svk_nccl_release(ctx);
  if (ctx->done) cudaEventDestroy(ctx->done);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* svk_last_error(svk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int svk_set_stream(svk_ctx* ctx, void* s) {
  SVK_LOCK(ctx);
  if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
  ctx->stream = (cudaStream_t)s;
  return 0;
}

int svk_sync(svk_ctx* ctx) { SVK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return 0; }

uint64_t svk_launch_count(svk_ctx* ctx) { return ctx->launches; }

int svk_profile_enable(svk_ctx* ctx, int on) {
  SVK_LOCK(ctx);
  ctx->profile = on != 0;
  if (on) {  // reference point of svk_profile_timeline: an event on the (idle) stream, i.e. "now" on the device clock
    if (!ctx->profile_ref) SVK_CUDA(ctx, cudaEventCreate(&ctx->profile_ref));
    SVK_CUDA(ctx, cudaEventRecord(ctx->profile_ref, ctx->stream));
  }
  return 0;
}

// JSON list [["kernel", start_ms, end_ms], ...] of every launch since svk_profile_enable(ctx, 1), times relative to that call on the
// device clock (comparable between contexts enabled back to back); resets the statistics like svk_profile_report.
int svk_profile_timeline(svk_ctx* ctx, char* buf, size_t buf_len) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::string out = "[";
  bool first = true;
  for (auto& pe : ctx->pending) {
    float t0 = 0, t1 = 0;
    if (ctx->profile_ref && cudaEventElapsedTime(&t0, ctx->profile_ref, pe.e0) == cudaSuccess &&
        cudaEventElapsedTime(&t1, ctx->profile_ref, pe.e1) == cudaSuccess) {
      char tmp[160];
      snprintf(tmp, sizeof tmp, "%s[\"%s\", %.4f, %.4f]", first ? "" : ", ", pe.name, t0, t1);
      out += tmp;
      first = false;
    }
    cudaEventDestroy(pe.e0);
    cudaEventDestroy(pe.e1);
  }
  ctx->pending.clear();
  out += "]";
  if (out.size() + 1 > buf_len) return svk_fail(ctx, "timeline buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return (int)out.size();
}

// JSON: {"kernel": {"count": c, "ms": total}, ...}; resets the statistics.  Returns bytes written or -1.
int svk_profile_report(svk_ctx* ctx, char* buf, size_t buf_len) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto& pe : ctx->pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, pe.e0, pe.e1) == cudaSuccess) {
      auto& st = ctx->stats[pe.name];
      st.count++;
      st.ms += ms;
    }
    cudaEventDestroy(pe.e0);
    cudaEventDestroy(pe.e1);
  }
  ctx->pending.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : ctx->stats) {
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s\"%s\": {\"count\": %llu, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
             (unsigned long long)kv.second.count, kv.second.ms);
    out += tmp;
    first = false;
  }
  out += "}";
  ctx->stats.clear();
  if (out.size() + 1 > buf_len) return svk_fail(ctx, "profile buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return (int)out.size();
}

static bool load_fq_canon(Fq& out, const svk_fe& fe) {
  fe_load_le(out.v, fe.b);
  if (!Fq::is_canonical(out.v)) return false;
  out = out.to_mont();
  return true;
}

int svk_dk_load(svk_ctx* ctx, const svk_deciding_key* dk) {
  SVK_LOCK(ctx);
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
  std::vector<G2LineX> x1, x2;
  for (auto& l : t1) x1.push_back({l.neg_lam, l.c3, l.neg_lam.mul_xi(), l.c3.mul_xi()});
  for (auto& l : t2) x2.push_back({l.neg_lam, l.c3, l.neg_lam.mul_xi(), l.c3.mul_xi()});
  SVK_CUDA(ctx, cudaMalloc(&d.d_linesx_g2, sizeof(G2LineX) * SVK_N_LINES));
  SVK_CUDA(ctx, cudaMalloc(&d.d_linesx_neg_sg2, sizeof(G2LineX) * SVK_N_LINES));
  SVK_CUDA(ctx, cudaMemcpy(d.d_linesx_g2, x1.data(), sizeof(G2LineX) * SVK_N_LINES, cudaMemcpyHostToDevice));
  SVK_CUDA(ctx, cudaMemcpy(d.d_linesx_neg_sg2, x2.data(), sizeof(G2LineX) * SVK_N_LINES, cudaMemcpyHostToDevice));
  ctx->dks.push_back(d);
  return (int)ctx->dks.size() - 1;
}

int svk_kzg_decide_batch_dev(svk_ctx* ctx, int dk, size_t n, const void* d_accs, void* d_out_ok) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_decide_launch(ctx, dk, n, d_accs, d_out_ok);
}

int svk_kzg_decide_batch(svk_ctx* ctx, int dk, size_t n, const svk_acc* accs, uint8_t* out_ok) {
  SVK_LOCK(ctx);
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

// ---- PlonkProtocol ingestion --------------------------------------------------------------------
int svk_protocol_compile_ex(svk_ctx* ctx, const uint8_t* blob, size_t len, int mos, int transcript_kind, int dk);
int svk_protocol_compile(svk_ctx* ctx, const uint8_t* blob, size_t len, int mos, int dk) {
  SVK_LOCK(ctx);
  return svk_protocol_compile_ex(ctx, blob, len, mos, SVK_TRANSCRIPT_POSEIDON, dk);
}

static int protocol_upload(svk_ctx* ctx, svk_host::CompiledProtocol& cp, int mos, int transcript_kind, int dk, int force_bits = 0);

int svk_protocol_compile_ex(svk_ctx* ctx, const uint8_t* blob, size_t len, int mos, int transcript_kind, int dk) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (dk < 0 || dk >= (int)ctx->dks.size()) return svk_fail(ctx, "bad deciding-key id %d", dk);
  svk_host::CompiledProtocol cp;
  try {
    cp = svk_host::compile_protocol(blob, len, mos, transcript_kind);
  } catch (svk_host::CompileError& e) {
    return svk_fail(ctx, "protocol compile: %s", e.what());
  }
  return protocol_upload(ctx, cp, mos, transcript_kind, dk);
}

int svk_protocol_compile_bincode(svk_ctx* ctx, const uint8_t* bytes, size_t len, int fe_encoding, int mos, int transcript_kind, int dk,
                                 size_t* consumed, int* fe_used) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (dk < 0 || dk >= (int)ctx->dks.size()) return svk_fail(ctx, "bad deciding-key id %d", dk);
  svk_host::CompiledProtocol cp;
  try {
    cp = svk_host::compile_protocol_bincode(bytes, len, fe_encoding, mos, transcript_kind, consumed, fe_used);
  } catch (svk_host::CompileError& e) {
    return svk_fail(ctx, "protocol compile (bincode): %s", e.what());
  }
  return protocol_upload(ctx, cp, mos, transcript_kind, dk);
}

static int protocol_upload(svk_ctx* ctx, svk_host::CompiledProtocol& cp, int mos, int transcript_kind, int dk, int force_bits) {
  ProtocolDevice* pd = new ProtocolDevice();
  pd->mos = mos;
  pd->transcript_kind = transcript_kind;
  pd->dk = dk;
  pd->verify_valid = cp.verify_valid;
  pd->invalid_reason = cp.invalid_reason;
  pd->n_ops = (u32)cp.ops.size();
  pd->read_ops_end = cp.read_ops_end;
  pd->n_regs = cp.n_regs;
  pd->n_instances = cp.n_instances;
  pd->n_challenges = cp.n_challenges;
  pd->n_scalar_slots = cp.n_scalar_slots;
  pd->proof_len = cp.proof_len;
  pd->n_perm = cp.n_perm;
  pd->n_fr_mul = cp.n_fr_mul;
  pd->num_instance = cp.num_instance;
  pd->n_old = cp.n_old;
  pd->acc_limbs = cp.acc_limbs;
  pd->acc_bits = cp.acc_bits;
  if (pd->n_old && upload(ctx, &pd->d_old_idx, cp.old_acc_idx)) { protocol_release(pd); return -1; }
  pd->n_pre = (u32)cp.preprocessed.size();
  for (auto& p : cp.points) pd->points.push_back({p.byte_offset, (u32)(p.val_x < 0 ? 0 : p.val_x), (u32)(p.val_y < 0 ? 0 : p.val_y)});
  std::vector<G1Affine> fixed;
  for (auto& g : cp.preprocessed) {
    G1Affine a = G1Affine::identity();
    bool id = true;
    for (int i = 0; i < 32; i++) id = id && g.x.b[i] == 0 && g.y.b[i] == 0;
    if (!id) {
      if (!load_fq_canon(a.x, g.x) || !load_fq_canon(a.y, g.y) || !g1_on_curve(a)) {
        protocol_release(pd);
        return svk_fail(ctx, "protocol: preprocessed commitment is not a canonical on-curve point");
      }
    }
    fixed.push_back(a);
  }
  fixed.push_back(ctx->dks[dk].g1);
  auto conv = [&](const std::vector<svk_host::MsmTerm>& in) {
    std::vector<MsmTermDev> out;
    for (auto& t : in) {
      MsmTermDev d;
      if (t.base == SVK_BASE_G) { d.fixed = 1; d.base = (int32_t)pd->n_pre; }
      else if (t.base < (int)pd->n_pre) { d.fixed = 1; d.base = t.base; }
      else { d.fixed = 0; d.base = t.base - (int)pd->n_pre; }
      d.slot = t.slot;
      out.push_back(d);
    }
    return out;
  };
  std::vector<MsmTermDev> lhs = conv(cp.lhs), rhs = conv(cp.rhs);
  pd->n_lhs = (u32)lhs.size();
  pd->n_rhs = (u32)rhs.size();
  pd->h_lhs = lhs;
  pd->h_rhs = rhs;
  // k_msm_var lanes: `var_lanes` threads per proof for the lhs terms, plus lanes for the rhs side when it has scaled variable
  // bases of its own (GWC: rhs = sum u^i W_i; SHPLONK's rhs is W' itself).  A thread carries at most SVK_VAR_TERMS_MAX (16) terms.
  u32 lhs_var = 0, rhs_var = 0;
  for (auto& t : lhs) lhs_var += (t.slot >= 0 && !t.fixed) ? 1 : 0;
  for (auto& t : rhs) rhs_var += (t.slot >= 0 && !t.fixed) ? 1 : 0;
  // 16-bit windows halve the table additions of k_msm_sum; their tables (67 MB per base) are used while they stay under ~3 GB
  pd->fixed_bits = fixed.size() <= 48 ? SVK_FIXED_BITS_LARGE : SVK_FIXED_BITS_SMALL;
  if (const char* e = getenv("SVK_FIXED_BITS")) pd->fixed_bits = atoi(e) == 16 ? 16 : 8;
  if (force_bits) pd->fixed_bits = (u32)force_bits;
  auto fail = [&](int rc) { protocol_release(pd); return rc; };
  for (int which = 0; which < 2; which++) {
    MsmSched& sc = pd->sched[which];
    // [0]: measured on B200 (profiles/r1_notes.md) with the signed-window Straus core: 1 lane = 1.37 M proofs/s, 2 lanes = 1.33 M
    // (one more 255-doubling chain per proof).  [1]: one lane per term.
    u32 want = which == 0 ? 1 : 16;
    if (which == 0)
      if (const char* e = getenv("SVK_VAR_LANES")) want = (u32)std::max(1, std::min(16, atoi(e)));
    u32 ll = std::max<u32>(std::min<u32>(want, std::max<u32>(lhs_var, 1)), (lhs_var + 15) / 16);
    u32 rl = rhs_var ? std::max<u32>(std::min<u32>(want, rhs_var), (rhs_var + 15) / 16) : 0;
    sc.var_lanes = ll;
    std::vector<std::vector<MsmWork>> vlanes(ll + rl);
    std::vector<MsmWork> wl, wr, var_items;
    std::vector<u32> ol, orr;
    std::vector<FixedSlot> fl, fr;
    sc.msm_lanes = which == 0 ? 1 : SVK_MSM_LANES_LATENCY;
    if (which == 0)
      if (const char* e = getenv("SVK_MSM_LANES")) { int v = atoi(e); sc.msm_lanes = (v == 2 || v == 4) ? (u32)v : 1; }
    sc.msm_work_modmul = schedule_msm(lhs, vlanes, 0, ll, wl, ol, fl, sc.fixed_per_lhs, pd->fixed_bits, (int)sc.msm_lanes) +
                         schedule_msm(rhs, vlanes, ll, std::max<u32>(rl, 1), wr, orr, fr, sc.fixed_per_rhs, pd->fixed_bits, (int)sc.msm_lanes);
    if (upload(ctx, &sc.d_fixed_lhs, fl) || upload(ctx, &sc.d_fixed_rhs, fr)) return fail(-1);
    std::vector<u32> vloff;
    for (auto& l : vlanes) {
      vloff.push_back((u32)var_items.size());
      var_items.insert(var_items.end(), l.begin(), l.end());
      sc.var_terms_per_thread = std::max<u32>(sc.var_terms_per_thread, (u32)l.size());
    }
    vloff.push_back((u32)var_items.size());
    pd->n_var = (u32)var_items.size();
    sc.var_lanes_total = (u32)vloff.size() - 1;
    if (sc.var_terms_per_thread > 16) { protocol_release(pd); return svk_fail(ctx, "too many variable-base terms per thread"); }
    if (upload(ctx, &sc.d_var_lane_off, vloff) || upload(ctx, &sc.d_var_items, var_items) || upload(ctx, &sc.d_work_lhs, wl) ||
        upload(ctx, &sc.d_lane_off_lhs, ol) || upload(ctx, &sc.d_work_rhs, wr) || upload(ctx, &sc.d_lane_off_rhs, orr))
      return fail(-1);
  }
  if (upload(ctx, &pd->d_ops, cp.ops) || upload(ctx, &pd->d_aux, cp.aux) || upload(ctx, &pd->d_consts, cp.consts) ||
      upload(ctx, &pd->d_sched, pd->points) || upload(ctx, &pd->d_lhs, lhs) || upload(ctx, &pd->d_rhs, rhs) || upload(ctx, &pd->d_fixed, fixed))
    return fail(-1);
  // Fixed-base tables, shared by every context of the process that compiles the same key on the same device.  The entry is
  // created, filled and marked ready under the mutex: a second context either fills it or finds it complete.
  pd->table_key = std::to_string(ctx->device) + ":" + std::to_string(pd->fixed_bits) + ":" + std::string((const char*)fixed.data(), fixed.size() * sizeof(G1Affine));
  {
    std::lock_guard<std::mutex> lk(g_table_mu);
    SharedTable& st = g_tables[pd->table_key];
    if (!st.d) {
      size_t bytes = fixed.size() * (256 / pd->fixed_bits) * ((size_t)1 << pd->fixed_bits) * sizeof(G1Affine);
      if (cudaMalloc(&st.d, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        g_tables.erase(pd->table_key);
        bool retry = pd->fixed_bits == SVK_FIXED_BITS_LARGE;
        protocol_release(pd);
        if (retry) goto retry_small;  // the 16-bit tables do not fit next to the caller's allocations: 8-bit windows (128x smaller)
        return svk_fail(ctx, "fixed-base table allocation of %zu bytes failed (SVK_FIXED_BITS=8 selects the 128x smaller tables)", bytes);
      }
      st.ready = false;
    }
    st.refs++;
    pd->d_fixed_tables = st.d;
    if (!st.ready) {
      if (svk_fixed_tables_launch(ctx, pd)) {  // synchronises the stream
        std::string err = ctx->err;
        st.refs--;
        pd->d_fixed_tables = nullptr;
        if (st.refs == 0) { cudaFree(st.d); g_tables.erase(pd->table_key); }
        protocol_release(pd);
        ctx->err = err;
        return -1;
      }
      st.ready = true;
    }
  }
  ctx->protocols.push_back(pd);
  return (int)ctx->protocols.size() - 1;
retry_small:
  return protocol_upload(ctx, cp, mos, transcript_kind, dk, SVK_FIXED_BITS_SMALL);
}

int svk_protocol_info(svk_ctx* ctx, int proto, uint32_t* out) {
  SVK_LOCK(ctx);
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  ProtocolDevice* pd = ctx->protocols[proto];
  out[0] = pd->proof_len; out[1] = pd->n_instances; out[2] = pd->n_challenges; out[3] = pd->n_regs; out[4] = pd->n_ops;
  out[5] = (u32)pd->n_perm; out[6] = pd->verify_valid ? 1 : 0; out[7] = (u32)pd->n_fr_mul; out[8] = pd->n_lhs; out[9] = pd->n_rhs;
  out[10] = (u32)pd->points.size(); out[11] = pd->n_scalar_slots; out[12] = (u32)pd->sched[0].msm_work_modmul; out[13] = (u32)(pd->n_var * (133 + 50 * 11 + 16 * 7) + pd->sched[0].var_lanes_total * (255 * 7 + 380));  /* straus.cuh: table 8 dbl + 7 madd + normalisation, ~50 mixed additions; 255 doublings + one inversion per lane */ out[14] = pd->n_var; out[15] = pd->sched[0].var_lanes_total;
  out[16] = pd->n_old; out[17] = pd->acc_limbs; out[18] = pd->acc_bits; out[19] = (u32)pd->transcript_kind;
  return 0;
}

// The per-column instance check of `PlonkProof::read` (proof.rs:66-69): `instances[i].len() == protocol.num_instance[i]` for
// every column.  The batch entry points carry the FLAT count only; a host binding calls this once per snark shape and
// reports SVK_INVALID_INSTANCES for a snark whose columns differ ([[a], [b, c]] against num_instance [2, 1] has the right total).
int svk_plonk_instance_shape_ok(svk_ctx* ctx, int proto, uint32_t n_cols, const uint32_t* col_lens) {
  SVK_LOCK(ctx);
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  const std::vector<u32>& want = ctx->protocols[proto]->num_instance;
  if (want.size() != n_cols) return 0;
  for (size_t i = 0; i < want.size(); i++)
    if (want[i] != col_lens[i]) return 0;
  return 1;
}

// ---- the unevaluated `Msm` of the final accumulator (util/msm.rs:20-24): terms and per-proof scalars -------------
int svk_protocol_msm_terms(svk_ctx* ctx, int proto, int side, int32_t* out, size_t max_terms) {
  SVK_LOCK(ctx);
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  const std::vector<MsmTermDev>& t = side ? ctx->protocols[proto]->h_rhs : ctx->protocols[proto]->h_lhs;
  if (t.size() > max_terms) return svk_fail(ctx, "terms buffer too small");
  for (size_t i = 0; i < t.size(); i++) { out[3 * i] = t[i].fixed; out[3 * i + 1] = t[i].base; out[3 * i + 2] = t[i].slot; }
  return (int)t.size();
}

int svk_plonk_msm_scalars_batch(svk_ctx* ctx, int proto, size_t n, const svk_fe* instances, uint32_t n_instances, const uint8_t* proofs,
                                size_t proof_stride, const uint32_t* proof_lens, svk_fe* out_scalars, svk_fe* out_challenges,
                                int32_t* out_status) {
  SVK_LOCK(ctx);
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  ProtocolDevice* pd = ctx->protocols[proto];
  std::vector<svk_acc> accs(n * pd->accs_per_proof());
  if (svk_plonk_succinct_verify_batch(ctx, proto, n, instances, n_instances, proofs, proof_stride, proof_lens, accs.data(), out_challenges,
                                      out_status))
    return -1;
  // scratch slot 5 still holds the scalars of that call: [slot][item] x 32 B  ->  [item][slot]
  size_t ns = pd->n_scalar_slots;
  std::vector<uint8_t> tmp(ns * n * 32);
  SVK_CUDA(ctx, cudaMemcpy(tmp.data(), ctx->scratch[5], tmp.size(), cudaMemcpyDeviceToHost));
  for (size_t s = 0; s < ns; s++)
    for (size_t i = 0; i < n; i++) memcpy(out_scalars[i * ns + s].b, &tmp[(s * n + i) * 32], 32);
  return 0;
}

int svk_plonk_succinct_verify_batch_dev(svk_ctx* ctx, int proto, size_t n, const void* d_instances, uint32_t n_instances,
                                        const void* d_proofs, size_t proof_stride, const void* d_proof_lens, void* d_out_acc,
                                        void* d_out_challenges, void* d_out_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  return svk_succinct_verify_launch(ctx, ctx->protocols[proto], n, (const uint8_t*)d_instances, n_instances, (const uint8_t*)d_proofs,
                                    proof_stride, (const u32*)d_proof_lens, (uint8_t*)d_out_acc, (u32*)d_out_challenges,
                                    (int32_t*)d_out_status);
}

int svk_plonk_succinct_verify_batch(svk_ctx* ctx, int proto, size_t n, const svk_fe* instances, uint32_t n_instances,
                                    const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_lens, svk_acc* out_acc,
                                    svk_fe* out_challenges, int32_t* out_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  if (n == 0) return 0;
  ProtocolDevice* pd = ctx->protocols[proto];
  cudaStream_t s = ctx->stream;
  size_t inst_bytes = n * (size_t)n_instances * 32, proof_bytes = n * proof_stride, chal_bytes = n * (size_t)pd->n_challenges * 32;
  uint8_t* d_io;
  size_t off_inst = 0, off_proofs = (inst_bytes + 255) / 256 * 256, off_lens = off_proofs + (proof_bytes + 255) / 256 * 256,
         off_acc = off_lens + (n * 4 + 255) / 256 * 256, off_chal = off_acc + n * 128 * pd->accs_per_proof(), off_status = off_chal + (chal_bytes + 255) / 256 * 256,
         total = off_status + n * 4;
  if (svk_scratch(ctx, 0, total, (void**)&d_io)) return -1;
  if (inst_bytes) SVK_CUDA(ctx, cudaMemcpyAsync(d_io + off_inst, instances, inst_bytes, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(d_io + off_proofs, proofs, proof_bytes, cudaMemcpyHostToDevice, s));
  if (proof_lens) SVK_CUDA(ctx, cudaMemcpyAsync(d_io + off_lens, proof_lens, n * 4, cudaMemcpyHostToDevice, s));
  if (svk_succinct_verify_launch(ctx, pd, n, d_io + off_inst, n_instances, d_io + off_proofs, proof_stride,
                                 proof_lens ? (const u32*)(d_io + off_lens) : nullptr, d_io + off_acc, (u32*)(d_io + off_chal),
                                 (int32_t*)(d_io + off_status)))
    return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out_acc, d_io + off_acc, n * 128 * pd->accs_per_proof(), cudaMemcpyDeviceToHost, s));
  if (out_challenges && chal_bytes) SVK_CUDA(ctx, cudaMemcpyAsync(out_challenges, d_io + off_chal, chal_bytes, cudaMemcpyDeviceToHost, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(out_status, d_io + off_status, n * 4, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  return 0;
}

// ---- KzgAs fold ------------------------------------------------------------------------------------
int svk_kzg_as_fold_dev(svk_ctx* ctx, size_t n, const void* d_accs, size_t group_size, void* d_out_acc, void* d_out_r, void* d_out_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_fold_launch(ctx, n, (const uint8_t*)d_accs, group_size, (uint8_t*)d_out_acc, (u32*)d_out_r, (int32_t*)d_out_status);
}

int svk_kzg_as_fold(svk_ctx* ctx, size_t n, const svk_acc* accs, size_t group_size, svk_acc* out_acc, svk_fe* out_r, int32_t* out_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n == 0) return svk_fail(ctx, "fold of zero accumulators");
  uint8_t* d;
  if (svk_scratch(ctx, 0, n * 128 + 256, (void**)&d)) return -1;
  uint8_t* d_out = d + n * 128;  // 128 acc + 32 r + 4 status
  cudaStream_t s = ctx->stream;
  SVK_CUDA(ctx, cudaMemcpyAsync(d, accs, n * 128, cudaMemcpyHostToDevice, s));
  if (svk_fold_launch(ctx, n, d, group_size, d_out, (u32*)(d_out + 128), (int32_t*)(d_out + 160))) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out_acc, d_out, 128, cudaMemcpyDeviceToHost, s));
  if (out_r) SVK_CUDA(ctx, cudaMemcpyAsync(out_r, d_out + 128, 32, cudaMemcpyDeviceToHost, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(out_status, d_out + 160, 4, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  return 0;
}

// `KzgAs::{read_proof, verify}` with `KzgAsVerifyingKey(true)` (zk; pcs/kzg/accumulation.rs:29-62, 113-136): `as_proof` = the 64 bytes
// the prover wrote (two compressed blind points).  out_acc = sum_i r^i acc_i + r^n blind, r squeezed after absorbing all n + 1 pairs.
int svk_kzg_as_fold_zk(svk_ctx* ctx, size_t n, const svk_acc* accs, const uint8_t* as_proof, size_t as_proof_len, svk_acc* out_acc, svk_fe* out_r,
                       int32_t* out_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n == 0) return svk_fail(ctx, "fold of zero accumulators");
  if (as_proof_len < 64) {  // `read_ec_point` hits the end of the stream
    *out_status = SVK_TRANSCRIPT | (SVK_T_EOF << 8);
    memset(out_acc, 0, sizeof *out_acc);
    if (out_r) memset(out_r, 0, sizeof *out_r);
    return 0;
  }
  uint8_t* d;
  if (svk_scratch(ctx, 0, (n + 1) * 128 + 512, (void**)&d)) return -1;
  uint8_t* d_out = d + (n + 1) * 128;  // 128 acc + 32 r + 4 status
  uint8_t* d_blind = d_out + 256;
  cudaStream_t s = ctx->stream;
  SVK_CUDA(ctx, cudaMemcpyAsync(d, accs, n * 128, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(d_blind, as_proof, 64, cudaMemcpyHostToDevice, s));
  if (svk_blind_points_launch(ctx, d_blind, d + n * 128)) return -1;
  if (svk_fold_launch(ctx, n + 1, d, 0, d_out, (u32*)(d_out + 128), (int32_t*)(d_out + 160))) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out_acc, d_out, 128, cudaMemcpyDeviceToHost, s));
  if (out_r) SVK_CUDA(ctx, cudaMemcpyAsync(out_r, d_out + 128, 32, cudaMemcpyDeviceToHost, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(out_status, d_out + 160, 4, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  return 0;
}

int svk_kzg_as_fold_multi_dev(svk_ctx* ctx, size_t n_seg, size_t n, const void* d_accs, size_t group_size, void* d_out_records) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_fold_launch_seg(ctx, n_seg, n, (const uint8_t*)d_accs, group_size, (uint8_t*)d_out_records, 256);
}

int svk_kzg_decide_records_dev(svk_ctx* ctx, int dk, size_t n_records, void* d_records) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_decide_launch_strided(ctx, dk, n_records, d_records, 256, (uint8_t*)d_records + 164, 256);
}

// ---- PlonkVerifier::verify over batches: succinct verify each proof, fold every batch, ONE pairing per batch ----
// n_batches batches of batch_size proofs, back to back.  d_out_records: n_batches x 256 B
//   { svk_acc folded ; svk_fe r ; int32 fold_status ; uint8 decide_ok ; uint8 ok }
int svk_plonk_verify_multi_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                               const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                               void* d_out_status, void* d_out_records) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  if (n_batches == 0 || batch_size == 0) return svk_fail(ctx, "empty batch");
  ProtocolDevice* pd = ctx->protocols[proto];
  size_t n = n_batches * batch_size;
  uint8_t* rec = (uint8_t*)d_out_records;
  if (svk_succinct_verify_launch(ctx, pd, n, (const uint8_t*)d_instances, n_instances, (const uint8_t*)d_proofs, proof_stride,
                                 (const u32*)d_proof_lens, (uint8_t*)d_out_accs, nullptr, (int32_t*)d_out_status))
    return -1;
  // a proof contributes its new accumulator and the old ones it carries, in that order (verifier/plonk.rs:86-91)
  if (svk_fold_launch_seg(ctx, n_batches, batch_size * pd->accs_per_proof(), (const uint8_t*)d_out_accs, group_size, rec, 256)) return -1;
  if (svk_decide_launch_strided(ctx, pd->dk, n_batches, rec, 256, rec + 164, 256)) return -1;
  return svk_batch_verdict_launch(ctx, n_batches, batch_size, (const int32_t*)d_out_status, rec, 256);
}

// The same without the pairing: what a RANK of a proof-sharded job runs -- its per-batch folded accumulators are
// all-gathered, folded across ranks and decided ONCE (snark_verifier_axiom_b200/distributed.py).  Records carry decide_ok = 1
// ("not decided here"), so `ok` = every proof read and verified succinctly and the fold found only curve points.
int svk_plonk_fold_multi_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                             const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                             void* d_out_status, void* d_out_records) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  if (n_batches == 0 || batch_size == 0) return svk_fail(ctx, "empty batch");
  ProtocolDevice* pd = ctx->protocols[proto];
  size_t n = n_batches * batch_size;
  uint8_t* rec = (uint8_t*)d_out_records;
  if (svk_succinct_verify_launch(ctx, pd, n, (const uint8_t*)d_instances, n_instances, (const uint8_t*)d_proofs, proof_stride,
                                 (const u32*)d_proof_lens, (uint8_t*)d_out_accs, nullptr, (int32_t*)d_out_status))
    return -1;
  if (svk_fold_launch_seg(ctx, n_batches, batch_size * pd->accs_per_proof(), (const uint8_t*)d_out_accs, group_size, rec, 256)) return -1;
  SVK_CUDA(ctx, cudaMemset2DAsync(rec + 164, 256, 1, 1, n_batches, ctx->stream));
  return svk_batch_verdict_launch(ctx, n_batches, batch_size, (const int32_t*)d_out_status, rec, 256);
}

int svk_plonk_verify_batch_dev(svk_ctx* ctx, int proto, size_t n, const void* d_instances, uint32_t n_instances, const void* d_proofs,
                               size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs, void* d_out_status,
                               void* d_out_folded) {
  SVK_LOCK(ctx);
  return svk_plonk_verify_multi_dev(ctx, proto, 1, n, d_instances, n_instances, d_proofs, proof_stride, d_proof_lens, group_size, d_out_accs,
                                    d_out_status, d_out_folded);
}

int svk_plonk_verify_multi(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const svk_fe* instances, uint32_t n_instances,
                           const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_lens, size_t group_size, int locate_failures,
                           int32_t* out_status, uint8_t* out_records) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (proto < 0 || proto >= (int)ctx->protocols.size()) return svk_fail(ctx, "bad protocol id %d", proto);
  if (n_batches == 0 || batch_size == 0) return svk_fail(ctx, "empty batch");
  size_t n = n_batches * batch_size;
  size_t apk = ctx->protocols[proto]->accs_per_proof();
  cudaStream_t s = ctx->stream;
  size_t inst_bytes = n * (size_t)n_instances * 32, proof_bytes = n * proof_stride;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t off_proofs = al(inst_bytes), off_lens = off_proofs + al(proof_bytes), off_accs = off_lens + al(n * 4), off_status = off_accs + n * apk * 128,
         off_rec = off_status + al(n * 4), off_dec = off_rec + n_batches * 256, total = off_dec + al(n * apk);
  uint8_t* d;
  if (svk_scratch(ctx, 0, total, (void**)&d)) return -1;
  if (inst_bytes) SVK_CUDA(ctx, cudaMemcpyAsync(d, instances, inst_bytes, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(d + off_proofs, proofs, proof_bytes, cudaMemcpyHostToDevice, s));
  if (proof_lens) SVK_CUDA(ctx, cudaMemcpyAsync(d + off_lens, proof_lens, n * 4, cudaMemcpyHostToDevice, s));
  if (svk_plonk_verify_multi_dev(ctx, proto, n_batches, batch_size, d, n_instances, d + off_proofs, proof_stride,
                                 proof_lens ? d + off_lens : nullptr, group_size, d + off_accs, d + off_status, d + off_rec))
    return -1;
  // results go straight into the caller's buffers (pinned buffers keep these copies asynchronous)
  SVK_CUDA(ctx, cudaMemcpyAsync(out_status, d + off_status, n * 4, cudaMemcpyDeviceToHost, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(out_records, d + off_rec, n_batches * 256, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  if (locate_failures) {
    // a batch whose proofs all read fine but whose folded pairing failed: decide each accumulator to name the culprits
    for (size_t b = 0; b < n_batches; b++) {
      if (out_records[b * 256 + 165]) continue;
      bool all_ok = true;
      for (size_t i = 0; i < batch_size; i++) all_ok = all_ok && out_status[b * batch_size + i] == 0;
      if (!all_ok) continue;
      std::vector<uint8_t> oks(batch_size * apk);  // `decide_all` over [new, old...] of every proof (verifier/plonk.rs:131-134)
      if (svk_decide_launch(ctx, ctx->protocols[proto]->dk, batch_size * apk, d + off_accs + b * batch_size * apk * 128, d + off_dec)) return -1;
      SVK_CUDA(ctx, cudaMemcpyAsync(oks.data(), d + off_dec, batch_size * apk, cudaMemcpyDeviceToHost, s));
      if (svk_wait(ctx)) return -1;
      for (size_t i = 0; i < batch_size * apk; i++)
        if (!oks[i]) out_status[b * batch_size + i / apk] = SVK_ASSERTION_FAILURE;
    }
  }
  return 0;
}

int svk_plonk_verify_batch(svk_ctx* ctx, int proto, size_t n, const svk_fe* instances, uint32_t n_instances, const uint8_t* proofs,
                           size_t proof_stride, const uint32_t* proof_lens, size_t group_size, int locate_failures, int32_t* out_status,
                           svk_acc* out_folded, uint8_t* out_ok) {
  SVK_LOCK(ctx);
  uint8_t rec[256];
  if (svk_plonk_verify_multi(ctx, proto, 1, n, instances, n_instances, proofs, proof_stride, proof_lens, group_size, locate_failures,
                             out_status, rec))
    return -1;
  if (out_folded) memcpy(out_folded, rec, 128);
  *out_ok = rec[165];
  return 0;
}

// ---- MSM ------------------------------------------------------------------------------------------
int svk_msm_g1_dev(svk_ctx* ctx, size_t n, const void* d_scalars, const void* d_points, void* d_out, void* d_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_msm_launch(ctx, n, (const uint8_t*)d_scalars, (const uint8_t*)d_points, (uint8_t*)d_out, (int*)d_status);
}

int svk_msm_g1(svk_ctx* ctx, size_t n, const svk_fe* scalars, const svk_g1* points, svk_g1* out, int32_t* out_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint8_t* d;
  size_t off_p = (n * 32 + 255) / 256 * 256, off_o = off_p + (n * 64 + 255) / 256 * 256;
  if (svk_scratch(ctx, 0, off_o + 256, (void**)&d)) return -1;
  if (n) {
    SVK_CUDA(ctx, cudaMemcpyAsync(d, scalars, n * 32, cudaMemcpyHostToDevice, s));
    SVK_CUDA(ctx, cudaMemcpyAsync(d + off_p, points, n * 64, cudaMemcpyHostToDevice, s));
  }
  if (svk_msm_launch(ctx, n, d, d + off_p, d + off_o, (int*)(d + off_o + 64))) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out, d + off_o, 64, cudaMemcpyDeviceToHost, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(out_status, d + off_o + 64, 4, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  return 0;
}

// ---- MSM over a chosen curve + the IPA decider -------------------------------------------
int svk_msm_curve_dev(svk_ctx* ctx, int curve, size_t n, const void* d_scalars, const void* d_points, void* d_out, void* d_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_msm_curve_launch(ctx, curve, n, (const uint8_t*)d_scalars, (const uint8_t*)d_points, (uint8_t*)d_out, (int*)d_status);
}

int svk_msm_curve(svk_ctx* ctx, int curve, size_t n, const svk_fe* scalars, const svk_g1* points, svk_g1* out, int32_t* out_status) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint8_t* d;
  size_t off_p = (n * 32 + 255) / 256 * 256, off_o = off_p + (n * 64 + 255) / 256 * 256;
  if (svk_scratch(ctx, 0, off_o + 256, (void**)&d)) return -1;
  if (n) {
    SVK_CUDA(ctx, cudaMemcpyAsync(d, scalars, n * 32, cudaMemcpyHostToDevice, s));
    SVK_CUDA(ctx, cudaMemcpyAsync(d + off_p, points, n * 64, cudaMemcpyHostToDevice, s));
  }
  if (svk_msm_curve_launch(ctx, curve, n, d, d + off_p, d + off_o, (int*)(d + off_o + 64))) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out, d + off_o, 64, cudaMemcpyDeviceToHost, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(out_status, d + off_o + 64, 4, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  return 0;
}

int svk_ipa_decide_batch_dev(svk_ctx* ctx, int curve, uint32_t k, const void* d_g, size_t n, const void* d_xi, const void* d_u,
                             void* d_out_status, void* d_invalid) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_ipa_decide_launch(ctx, curve, k, (const uint8_t*)d_g, n, (const uint8_t*)d_xi, (const uint8_t*)d_u, (int32_t*)d_out_status,
                               (int32_t*)d_invalid);
}

int svk_ipa_decide_batch(svk_ctx* ctx, int curve, uint32_t k, const svk_g1* g, size_t n, const svk_fe* xi, const svk_g1* u,
                         int32_t* out_status, int32_t* out_invalid) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (k == 0 || k > 26) return svk_fail(ctx, "ipa_decide: k out of range");
  if (n == 0) return svk_fail(ctx, "ipa_decide: no accumulators (the reference asserts !accumulators.is_empty(), pcs/ipa/decider.rs:61)");
  cudaStream_t s = ctx->stream;
  size_t m = (size_t)1 << k;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t off_xi = al(m * 64), off_u = off_xi + al(n * (size_t)k * 32), off_st = off_u + al(n * 64), off_inv = off_st + al(n * 4);
  uint8_t* d;
  if (svk_scratch(ctx, 0, off_inv + 256, (void**)&d)) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(d, g, m * 64, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(d + off_xi, xi, n * (size_t)k * 32, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(d + off_u, u, n * 64, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemsetAsync(d + off_inv, 0, 4, s));
  if (svk_ipa_decide_launch(ctx, curve, k, d, n, d + off_xi, d + off_u, (int32_t*)(d + off_st), (int32_t*)(d + off_inv))) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out_status, d + off_st, n * 4, cudaMemcpyDeviceToHost, s));
  int32_t inv = 0;
  SVK_CUDA(ctx, cudaMemcpyAsync(&inv, d + off_inv, 4, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  if (out_invalid) *out_invalid = inv;
  return 0;
}

int svk_g1_mul_batch_dev(svk_ctx* ctx, size_t n, const void* d_scalars, const void* d_points, size_t n_points, void* d_out) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_g1_mul_batch_launch(ctx, n, (const uint8_t*)d_scalars, (const uint8_t*)d_points, n_points, (uint8_t*)d_out);
}

int svk_g1_mul_batch(svk_ctx* ctx, size_t n, const svk_fe* scalars, const svk_g1* points, size_t n_points, svk_g1* out) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  cudaStream_t s = ctx->stream;
  uint8_t* d;
  size_t off_p = (n * 32 + 255) / 256 * 256, off_o = off_p + (n_points * 64 + 255) / 256 * 256;
  if (svk_scratch(ctx, 0, off_o + n * 64, (void**)&d)) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(d, scalars, n * 32, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(d + off_p, points, n_points * 64, cudaMemcpyHostToDevice, s));
  if (svk_g1_mul_batch_launch(ctx, n, d, d + off_p, n_points, d + off_o)) return -1;
  SVK_CUDA(ctx, cudaMemcpyAsync(out, d + off_o, n * 64, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  return 0;
}

// Test hook: out[i] = Poseidon::new().update(inputs[i][0..n_inputs]).squeeze()  (util/hash/poseidon.rs:448-467,
// T = 3, RATE = 2, R_F = 8, R_P = 57).  schedule 0: one thread per sponge, 1: the warp-cooperative permutation.
// Returns -1 when an input is not a canonical Fr.
int svk_poseidon_squeeze(svk_ctx* ctx, size_t n, const svk_fe* inputs, uint32_t n_inputs, int schedule, svk_fe* out) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  cudaStream_t s = ctx->stream;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t in_bytes = n * (size_t)n_inputs * 32, off_out = al(in_bytes + 32), off_bad = off_out + al(n * 32);
  uint8_t* d;
  if (svk_scratch(ctx, 0, off_bad + 256, (void**)&d)) return -1;
  if (in_bytes) SVK_CUDA(ctx, cudaMemcpyAsync(d, inputs, in_bytes, cudaMemcpyHostToDevice, s));
  SVK_CUDA(ctx, cudaMemsetAsync(d + off_bad, 0, 4, s));
  if (svk_poseidon_squeeze_launch(ctx, n, n_inputs, d, d + off_out, schedule != 0, (int*)(d + off_bad))) return -1;
  int bad = 0;
  SVK_CUDA(ctx, cudaMemcpyAsync(out, d + off_out, n * 32, cudaMemcpyDeviceToHost, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(&bad, d + off_bad, 4, cudaMemcpyDeviceToHost, s));
  if (svk_wait(ctx)) return -1;
  if (bad) return svk_fail(ctx, "poseidon_squeeze: input is not a canonical Fr");
  return 0;
}

int svk_bench_modmul_peak(svk_ctx* ctx, int iters, double* out_modmul_per_s, double* out_ms) {
  SVK_LOCK(ctx);
  SVK_CUDA(ctx, cudaSetDevice(ctx->device));
  return svk_modmul_peak_launch(ctx, iters, out_modmul_per_s, out_ms);
}

}  // extern "C"
