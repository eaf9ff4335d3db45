// abi.cu -- extern "C" entry points of libegt_b200.so (see include/egt_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "kernels.h"
#include "fused.h"
#include "wide.h"
#include "node_blas.h"

namespace egt {

static thread_local char g_err[512] = "";
static thread_local int g_last_path = 0;
static int g_force_staged = 0;
    // egt_debug_force_staged(): route every shape through the staged kernels

void set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int n = snprintf(g_err, sizeof(g_err), "egt_b200[%d]: ", code);
  vsnprintf(g_err + n, sizeof(g_err) - n, fmt, ap);
  va_end(ap);
}

// ---- launch counter / per-kernel event profiler -------------------------------------------
struct ProfSlot { const char *name; double ms; long count; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static long g_launches = 0;
static std::vector<ProfSlot> g_slots;
struct Pending { int slot; cudaEvent_t e0, e1; };
static std::vector<Pending> g_pending;

LaunchScope::LaunchScope(const char *name, cudaStream_t stream) : slot(-1), st(stream), e1(nullptr) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ++g_launches;
  if (!g_prof_on) return;
  for (size_t i = 0; i < g_slots.size(); ++i)
    if (!strcmp(g_slots[i].name, name)) slot = (int)i;
  if (slot < 0) { g_slots.push_back(ProfSlot{name, 0.0, 0}); slot = (int)g_slots.size() - 1; }
  cudaEvent_t e0;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  g_pending.push_back(Pending{slot, e0, e1});
}
LaunchScope::~LaunchScope() {
  if (e1) cudaEventRecord(e1, st);
}

bool pdl_enabled() {
  // measured on B200 (C0): overlapping the successor's set-up with the predecessor's tail made the step 7 %
  // SLOWER (0.246 vs 0.229 ms) -- the early CTAs compete for the SMs the tail still needs -- so it is opt-in
  static const bool on = getenv("EGT_PDL") && atoi(getenv("EGT_PDL")) != 0;
  return on;
}

// ---- a side stream for the small kernels that depend on the weights only (wide_prep) or feed nothing downstream
// (wide_bwd_finalize): they run next to the node kernels of the same call instead of between them.  Fork / join with
// events, so a call captured into a CUDA graph records parallel branches.  Created on the first call of a device that
// is not being captured (stream creation is not allowed inside a global-mode capture); EGT_SIDE_STREAM=0 turns it off.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr, aux = nullptr; bool ok = false; };
static SideStream g_side[16];
static SideStream *side_stream(cudaStream_t st) {
  static const bool off = getenv("EGT_SIDE_STREAM") && atoi(getenv("EGT_SIDE_STREAM")) == 0;
  if (off) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream &x = g_side[dev];
  if (!x.ok) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;
    if (cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&x.aux, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    x.ok = true;
  }
  return &x;
}
// side waits for everything issued on st so far; returns the stream to launch the branch on (st itself when off)
static cudaStream_t side_fork(SideStream *x, cudaStream_t st) {
  if (!x) return st;
  if (cudaEventRecord(x->fork, st) != cudaSuccess || cudaStreamWaitEvent(x->s, x->fork, 0) != cudaSuccess) return st;
  return x->s;
}
// mark: remember what the branch has been given so far; wait: st waits for the marked work only (the branch may already
// have been given more, which keeps running next to st); join = mark + wait
static int side_mark(SideStream *x, cudaStream_t st, cudaStream_t branch) {
  if (!x || branch == st) return EGT_OK;
  EGT_CHECK_CUDA(cudaEventRecord(x->join, branch));
  return EGT_OK;
}
static int side_wait(SideStream *x, cudaStream_t st, cudaStream_t branch) {
  if (!x || branch == st) return EGT_OK;
  EGT_CHECK_CUDA(cudaStreamWaitEvent(st, x->join, 0));
  return EGT_OK;
}
static int side_join(SideStream *x, cudaStream_t st, cudaStream_t branch) {
  int rc = side_mark(x, st, branch);
  return rc ? rc : side_wait(x, st, branch);
}

static size_t esize(int dtype) { return dtype == EGT_F32 ? 4 : 2; }
static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static int check_attn_cfg(const egt_attn_cfg_t *c) {
  EGT_REQUIRE(c != nullptr, EGT_E_ARG, "cfg is NULL");
  EGT_REQUIRE(c->B > 0 && c->N > 0 && c->h > 0 && c->dk > 0, EGT_E_SHAPE, "B,N,h,dk must be positive (got %d,%d,%d,%d)",
              c->B, c->N, c->h, c->dk);
  EGT_REQUIRE(c->dk <= 16, EGT_E_SHAPE, "dk=%d > 16 is not supported", c->dk);
  EGT_REQUIRE(c->dtype == EGT_F32 || c->dtype == EGT_BF16, EGT_E_DTYPE, "dtype must be EGT_F32 or EGT_BF16");
  // egt_layers.py:20-24
  EGT_REQUIRE(!(c->scale_degree && !c->gate_input), EGT_E_ARG, "scale_degree requires gate_input");
  EGT_REQUIRE(c->scaler_type == EGT_SCALER_LOG || c->scaler_type == EGT_SCALER_LINEAR, EGT_E_ARG,
              "scaler_type must be log or linear");
  EGT_REQUIRE(c->attn_mask >= EGT_MASK_NONE && c->attn_mask <= EGT_MASK_ADJ_U8, EGT_E_ARG, "bad attn_mask kind");
  EGT_REQUIRE(c->random_mask_prob >= 0.f && c->random_mask_prob < 1.f, EGT_E_ARG, "random_mask_prob out of range");
  EGT_REQUIRE(c->attn_dropout >= 0.f && c->attn_dropout < 1.f, EGT_E_ARG, "attn_dropout out of range");
  EGT_REQUIRE(c->num_virtual_nodes >= 0 && c->num_virtual_nodes <= c->N, EGT_E_ARG, "num_virtual_nodes out of range");
  return EGT_OK;
}

static AttnParams make_attn_params(const egt_attn_cfg_t *c) {
  AttnParams P;
  memset(&P, 0, sizeof(P));
  P.B = c->B; P.N = c->N; P.h = c->h; P.dk = c->dk;
  P.scale = 1.0f / sqrtf((float)c->dk);
  P.has_clip = c->has_clip; P.clip_lo = c->clip_lo; P.clip_hi = c->clip_hi;
  P.scale_degree = c->scale_degree; P.scaler_type = c->scaler_type;
  P.num_virtual_nodes = c->num_virtual_nodes;
  P.attn_mask = c->attn_mask;
  P.rand_mask = c->training && c->random_mask_prob > 0.f;
  P.dropout = c->training && c->attn_dropout > 0.f;
  P.random_mask_prob = c->random_mask_prob; P.attn_dropout = c->attn_dropout;
  P.seed = c->seed; P.offset = c->offset; P.offset_dev = c->offset_dev;
  P.dq_scale = 1.0f;
  return P;
}

static int check_device() {
  int dev = 0;
  EGT_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0;
  EGT_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  EGT_REQUIRE(major == 10, EGT_E_ARCH, "libegt_b200 is built for sm_100a only (device has compute capability %d.x)", major);
  return EGT_OK;
}

// workspace carving -------------------------------------------------------------------------
struct BlockWs {
  char *E, *G, *Hhat, *dHext, *dE, *dG;      // [pairs,h] dtype
  char *dS, *As;                              // [pairs,h] dtype (staged backward: row pass -> column pass)
  float *row_ws;                              // [2,B,N,h]
  char *d_v_att, *d_qkv;                      // [R,d], [R,3d] dtype
  float *hn, *dhn;                            // [R,d] f32
  char *prep;                                 // FusedPrep (fused path only)
  char *nblas;                                // node_blas.cu scratch (wide path, d != 64)
  float *d_qkv_f32, *partials;                // fused backward: [R,3d] f32, [ctas,FPART] f32
  char *d_qkv_bf;                             // fused backward, N <= 128 (one row tile per graph): [R,3d] bf16 instead
  size_t total;
};
static BlockWs carve(const egt_block_cfg_t *c, int backward, void *base, bool has_de_out = true) {
  const egt_attn_cfg_t &a = c->attn;
  size_t es = esize(a.dtype);
  size_t pairs = (size_t)a.B * a.N * a.N, R = (size_t)a.B * a.N, d = (size_t)a.h * a.dk;
  size_t off = 0;
  char *b = (char *)base;
  BlockWs w;
  memset(&w, 0, sizeof(w));
  auto take = [&](size_t bytes) { char *p = b ? b + off : nullptr; off += align_up(bytes); return p; };
  bool edge = c->edge_channel_type != EGT_EDGE_NONE;
  bool residual = c->edge_channel_type >= EGT_EDGE_RESIDUAL;
  const bool wide = wide_supported(c, a.dtype) && !g_force_staged;
  const bool fused = (fused_supported(c, a.dtype) && !g_force_staged) || wide;
  if (fused) w.prep = take(wide ? sizeof(WidePrep) : sizeof(FusedPrep));
  if (wide && !(d == 64 && a.h == 8) && node_blas_supported((int)d)) w.nblas = take(node_blas_workspace_bytes((int)R, (int)d));
  if (fused && !backward) { w.total = off; return w; }
  if (fused && backward && has_de_out && (!wide || wide_bwd_supported(c))) {      // fused backward: nothing of shape [pairs,h] is materialised
    w.d_v_att = take(R * d * es);
    w.d_qkv_f32 = (float *)take(R * 3 * d * sizeof(float));
    if (!wide && a.N <= 128 && fused_bwd_key_splits(a.B, a.N) == 1) w.d_qkv_bf = take(R * 3 * d * 2);
    w.partials = (float *)take(wide ? wide_bwd_partials_floats(c) * sizeof(float)
                                    : (size_t)a.B * ((a.N + 127) / 128) * fused_bwd_key_splits(a.B, a.N) * FPART * sizeof(float));
    w.hn = (float *)take(R * d * sizeof(float));
    w.dhn = (float *)take(R * d * sizeof(float));
    w.total = off;
    return w;
  }
  if (edge) {
    w.E = take(pairs * a.h * es);
    if (c->gate_attention) w.G = take(pairs * a.h * es);
  }
  if (residual) w.Hhat = take(pairs * a.h * es);
  if (backward) {
    if (residual) w.dHext = take(pairs * a.h * es);
    if (edge) {
      w.dE = take(pairs * a.h * es);
      if (c->gate_attention) w.dG = take(pairs * a.h * es);
    }
    w.row_ws = (float *)take(2 * R * a.h * sizeof(float));
    w.dS = take(pairs * a.h * es);
    w.As = take(pairs * a.h * es);
    w.d_v_att = take(R * d * es);
    w.d_qkv = take(R * 3 * d * es);
    w.hn = (float *)take(R * d * sizeof(float));
    w.dhn = (float *)take(R * d * sizeof(float));
  }
  w.total = off;
  return w;
}

static int check_block_cfg(const egt_block_cfg_t *c) {
  EGT_REQUIRE(c != nullptr, EGT_E_ARG, "cfg is NULL");
  EGT_REQUIRE(c->edge_channel_type >= EGT_EDGE_NONE && c->edge_channel_type <= EGT_EDGE_CONSTRAINED, EGT_E_ARG,
              "bad edge_channel_type");
  // graph_xformer_model_base.py:46-47
  EGT_REQUIRE(!(c->attn.scale_degree && !c->gate_attention), EGT_E_ARG, "scale_degree only works with gate_attention");
  EGT_REQUIRE(!(c->attn.scale_degree && c->edge_channel_type == EGT_EDGE_NONE), EGT_E_ARG,
              "scale_degree needs gates, i.e. an edge channel");
  if (c->edge_channel_type != EGT_EDGE_NONE) {
    EGT_REQUIRE(c->d_e > 0, EGT_E_SHAPE, "d_e must be positive");
    EGT_REQUIRE(c->attn.h <= 16, EGT_E_SHAPE, "block path supports h <= 16 (got %d)", c->attn.h);
  }
  return EGT_OK;
}

static egt_attn_cfg_t derived_attn_cfg(const egt_block_cfg_t *c) {
  egt_attn_cfg_t a = c->attn;
  a.edge_input = c->edge_channel_type != EGT_EDGE_NONE;                       // graph_xformer_model_base.py:120
  a.gate_input = a.edge_input && c->gate_attention;                          // :121
  a.attn_mask = c->edge_channel_type == EGT_EDGE_CONSTRAINED ? EGT_MASK_ADJ_U8 : EGT_MASK_NONE;   // :122
  return a;
}

static EdgeParams make_edge_params(const egt_block_cfg_t *c, const egt_block_weights_t *w) {
  EdgeParams p;
  memset(&p, 0, sizeof(p));
  p.pairs = (size_t)c->attn.B * c->attn.N * c->attn.N;
  p.d_e = c->d_e; p.h = c->attn.h;
  p.has_ln = c->edge_channel_type >= EGT_EDGE_RESIDUAL;
  p.ln_eps = c->ln_eps;
  p.gated = c->gate_attention;
  p.act = c->edge_act; p.act_alpha = c->edge_act_alpha;
  p.ln_g = w->norm_edge_gamma; p.ln_b = w->norm_edge_beta;
  p.w_e = w->dense_edge_b_kernel; p.b_e = w->dense_edge_b_bias;
  p.w_g = w->attention_gates_kernel; p.b_g = w->attention_gates_bias;
  p.w_r = w->dense_edge_r_kernel; p.b_r = w->dense_edge_r_bias;
  return p;
}

}  // namespace egt

using namespace egt;

extern "C" {

int egt_abi_version(void) { return EGT_ABI_VERSION; }
const char *egt_last_error(void) { return g_err; }
int egt_last_path(void) { return g_last_path; }

float egt_rng_uniform_host(uint64_t seed, uint64_t offset, uint32_t stream_id, uint64_t idx) {
  return rng_uniform(seed, offset, stream_id, idx);
}

int egt_debug_force_staged(int on) { g_force_staged = on != 0; return EGT_OK; }

long egt_launch_count(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  return g_launches;
}

int egt_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  if (on) { g_slots.clear(); }
  return EGT_OK;
}

// Synchronises the device, folds pending event pairs into per-kernel totals and writes up to `max`
// entries: names (64 bytes each, NUL padded), total milliseconds, launch counts.  Returns #entries.
int egt_profile_read(char *names_host, double *ms_host, long *counts_host, int max) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto &p : g_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) { g_slots[p.slot].ms += ms; g_slots[p.slot].count++; }
    cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
  }
  g_pending.clear();
  int n = 0;
  for (auto &sl : g_slots) {
    if (n >= max) break;
    memset(names_host + 64 * n, 0, 64);
    strncpy(names_host + 64 * n, sl.name, 63);
    ms_host[n] = sl.ms; counts_host[n] = sl.count;
    ++n;
  }
  return n;
}

int64_t egt_block_param_layout(const egt_block_cfg_t *c, int64_t *off) {
  if (!c) return -1;
  int64_t d = (int64_t)c->attn.h * c->attn.dk, de = c->d_e, h = c->attn.h;
  bool edge = c->edge_channel_type != EGT_EDGE_NONE;
  bool residual = c->edge_channel_type >= EGT_EDGE_RESIDUAL;
  bool gated = edge && c->gate_attention;
  int64_t sizes[14] = {d, d, d * 3 * d, 3 * d, d * d, d,
                       residual ? de : 0, residual ? de : 0,
                       gated ? de * h : 0, gated ? h : 0,
                       edge ? de * h : 0, edge ? h : 0,
                       residual ? h * de : 0, residual ? de : 0};
  int64_t tot = 0;
  for (int i = 0; i < 14; ++i) {
    if (off) off[i] = sizes[i] ? tot : -1;
    tot += sizes[i];
  }
  return tot;
}

int egt_attn_fwd(const egt_attn_cfg_t *cfg, const void *qkv, const void *E, const void *G, const void *M,
                 const uint8_t *mask, void *v_att, void *h_hat, void *a_tild, float *lse, float *deg,
                 void *stream) {
  int rc = check_attn_cfg(cfg);
  if (rc) return rc;
  if ((rc = check_device())) return rc;
  EGT_REQUIRE(qkv && v_att && lse && deg, EGT_E_ARG, "qkv, v_att, lse, deg must be non-NULL");
  EGT_REQUIRE(!cfg->edge_input || E, EGT_E_ARG, "edge_input set but E is NULL");
  EGT_REQUIRE(!cfg->gate_input || G, EGT_E_ARG, "gate_input set but G is NULL");
  EGT_REQUIRE(cfg->attn_mask == EGT_MASK_NONE || M, EGT_E_ARG, "attn_mask set but M is NULL");
  AttnParams P = make_attn_params(cfg);
  P.qkv = qkv; P.E = cfg->edge_input ? E : nullptr; P.G = cfg->gate_input ? G : nullptr;
  P.M = M; P.mask = mask; P.v_att = v_att; P.h_hat = h_hat; P.a_tild = a_tild; P.lse = lse; P.deg = deg;
  return attn_staged_fwd(P, cfg->dtype, (cudaStream_t)stream);
}

int egt_attn_bwd(const egt_attn_cfg_t *cfg, const void *qkv, const void *E, const void *G, const void *M,
                 const uint8_t *mask, const float *lse, const float *deg, const void *d_v_att,
                 const void *d_h_hat, void *d_qkv, void *dE, void *dG, float *row_ws, void *stream) {
  int rc = check_attn_cfg(cfg);
  if (rc) return rc;
  if ((rc = check_device())) return rc;
  EGT_REQUIRE(qkv && lse && deg && d_v_att && d_qkv && row_ws, EGT_E_ARG,
              "qkv, lse, deg, d_v_att, d_qkv, row_ws must be non-NULL");
  AttnParams P = make_attn_params(cfg);
  P.qkv = qkv; P.E = cfg->edge_input ? E : nullptr; P.G = cfg->gate_input ? G : nullptr;
  P.M = M; P.mask = mask; P.lse = (float *)lse; P.deg = (float *)deg;
  P.d_v_att = d_v_att; P.d_h_hat = d_h_hat; P.d_qkv = d_qkv;
  P.dE = cfg->edge_input ? dE : nullptr; P.dG = cfg->gate_input ? dG : nullptr;
  P.row_ws = row_ws;
  return attn_staged_bwd(P, cfg->dtype, (cudaStream_t)stream);
}

size_t egt_block_workspace_bytes(const egt_block_cfg_t *cfg, int32_t backward) {
  if (!cfg) return 0;
  size_t a = carve(cfg, backward, nullptr, true).total, b = carve(cfg, backward, nullptr, false).total;
  return (a > b ? a : b) + 256;
}

int egt_block_fwd(const egt_block_cfg_t *cfg, const egt_block_weights_t *w, const egt_block_fwd_io_t *io,
                  void *stream) {
  int rc = check_block_cfg(cfg);
  if (rc) return rc;
  egt_attn_cfg_t a = derived_attn_cfg(cfg);
  if ((rc = check_attn_cfg(&a))) return rc;
  if ((rc = check_device())) return rc;
  EGT_REQUIRE(w && io, EGT_E_ARG, "weights / io is NULL");
  EGT_REQUIRE(io->h && io->h_out && io->qkv && io->v_att && io->lse && io->deg, EGT_E_ARG,
              "h, h_out, qkv, v_att, lse, deg must be non-NULL");
  const bool edge = cfg->edge_channel_type != EGT_EDGE_NONE;
  const bool residual = cfg->edge_channel_type >= EGT_EDGE_RESIDUAL;
  EGT_REQUIRE(!edge || io->e, EGT_E_ARG, "e is NULL");
  EGT_REQUIRE(!residual || io->e_out, EGT_E_ARG, "e_out is NULL");
  EGT_REQUIRE(cfg->edge_channel_type != EGT_EDGE_CONSTRAINED || io->adj, EGT_E_ARG, "constrained needs adj");
  cudaStream_t st = (cudaStream_t)stream;
  const int R = a.B * a.N, d = a.h * a.dk;

  // node side: LN_h + QKV (graph_xformer_model_base.py:108-114)
  const bool fused = fused_supported(cfg, a.dtype) && !g_force_staged;
  size_t need = egt_block_workspace_bytes(cfg, 0);
  EGT_REQUIRE(io->workspace_bytes >= need && (io->workspace || need <= 256), EGT_E_ARG,
              "workspace too small: %zu < %zu", io->workspace_bytes, need);
  BlockWs ws = carve(cfg, 0, io->workspace);
  const bool wide = !fused && wide_supported(cfg, a.dtype) && !g_force_staged;
  if (wide) {    // width-generic fused path (wide_fwd.cu); Q is stored pre-scaled by dk^-0.5
    const bool node_tc = d == 64 && a.h == 8;          // the tcgen05 node kernels (node_tc.cu) serve model width 64
    SideStream *side = side_stream(st);
    // the folded weights depend on the weights only: their kernel runs next to the node kernels, not in front of them
    cudaStream_t sb = side_fork(side, st);
    if ((rc = wide_prep_launch(cfg, w, (WidePrep *)ws.prep, sb))) return rc;
    if (node_tc) {
      if ((rc = node_qkv_launch(io->h, w->norm_mha_gamma, w->norm_mha_beta, cfg->ln_eps, w->dense_qkv_kernel,
                                w->dense_qkv_bias, 1.0f / sqrtf((float)a.dk), io->qkv, R, w, a.clip_lo, a.clip_hi, nullptr, st))) return rc;
    } else if (ws.nblas) {
      if ((rc = node_blas_qkv(io->h, w, cfg->ln_eps, 1.0f / sqrtf((float)a.dk), io->qkv, R, d, ws.nblas, st))) return rc;
    } else {
      LinearArgs lq;
      memset(&lq, 0, sizeof(lq));
      lq.x = io->h; lq.W = w->dense_qkv_kernel; lq.bias = w->dense_qkv_bias; lq.out = io->qkv;
      lq.ln_gamma = w->norm_mha_gamma; lq.ln_beta = w->norm_mha_beta; lq.ln_eps = cfg->ln_eps;
      lq.scale = 1.0f / sqrtf((float)a.dk); lq.scale_cols = d;
      lq.R = R; lq.din = d; lq.dout = 3 * d;
      if ((rc = linear_launch(lq, a.dtype, st))) return rc;
    }
    if ((rc = side_join(side, st, sb))) return rc;
    g_last_path = 1;
    WideFwdArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.B = a.B; fa.N = a.N; fa.mask = io->mask; fa.prep = (const WidePrep *)ws.prep;
    fa.v_att = (__nv_bfloat16 *)io->v_att; fa.lse = io->lse; fa.deg = io->deg;
    fa.clip_lo = a.clip_lo; fa.clip_hi = a.clip_hi; fa.ln_eps = cfg->ln_eps;
    fa.scale_degree = a.scale_degree; fa.scaler_type = a.scaler_type; fa.num_virtual_nodes = a.num_virtual_nodes;
    fa.rand_mask = a.training && a.random_mask_prob > 0.f;
    fa.rand_thr = (uint32_t)ceilf(a.random_mask_prob * 65536.0f - 0.5f);
    fa.seed = a.seed; fa.offset = a.offset; fa.offset_dev = a.offset_dev;
    if ((rc = wide_fwd_launch(cfg, fa, io->e, io->e_out, io->qkv, st))) return rc;
    if (node_tc) return node_out_launch(io->v_att, io->h, w->dense_mha_kernel, w->dense_mha_bias, io->h_out, R, st);
    if (ws.nblas) return node_blas_out(io->v_att, io->h, w, io->h_out, R, d, ws.nblas, st);
    LinearArgs lo;   // h' = V_att W_O + b_O + h  (graph_xformer_model_base.py:136-140)
    memset(&lo, 0, sizeof(lo));
    lo.x = io->v_att; lo.W = w->dense_mha_kernel; lo.bias = w->dense_mha_bias; lo.res = io->h; lo.out = io->h_out;
    lo.R = R; lo.din = d; lo.dout = d;
    return linear_launch(lo, a.dtype, st);
  }
  if (fused) {   // Q is stored pre-scaled by dk^-0.5; the kernel's extra CTA writes the derived weights
    if ((rc = node_qkv_launch(io->h, w->norm_mha_gamma, w->norm_mha_beta, cfg->ln_eps, w->dense_qkv_kernel,
                              w->dense_qkv_bias, 1.0f / sqrtf((float)a.dk), io->qkv, R, w, a.clip_lo, a.clip_hi,
                              (FusedPrep *)ws.prep, st))) return rc;
  } else {
    LinearArgs lq;
    memset(&lq, 0, sizeof(lq));
    lq.x = io->h; lq.W = w->dense_qkv_kernel; lq.bias = w->dense_qkv_bias; lq.out = io->qkv;
    lq.ln_gamma = w->norm_mha_gamma; lq.ln_beta = w->norm_mha_beta; lq.ln_eps = cfg->ln_eps;
    lq.R = R; lq.din = d; lq.dout = 3 * d;
    if ((rc = linear_launch(lq, a.dtype, st))) return rc;
  }

  g_last_path = fused ? 1 : 0;

  if (fused) {
    FusedFwdArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.B = a.B; fa.N = a.N; fa.mask = io->mask; fa.prep = (const FusedPrep *)ws.prep;
    fa.v_att = (__nv_bfloat16 *)io->v_att; fa.lse = io->lse; fa.deg = io->deg;
    fa.clip_lo = a.clip_lo; fa.clip_hi = a.clip_hi; fa.ln_eps = cfg->ln_eps;
    fa.scale_degree = a.scale_degree; fa.scaler_type = a.scaler_type; fa.num_virtual_nodes = a.num_virtual_nodes;
    fa.rand_mask = a.training && a.random_mask_prob > 0.f;
    fa.rand_thr = (uint32_t)ceilf(a.random_mask_prob * 65536.0f - 0.5f);
    fa.seed = a.seed; fa.offset = a.offset; fa.offset_dev = a.offset_dev;
    // the output projection + residual runs in the same kernel (V_att goes from registers to the tensor core)
    fa.w_o = w->dense_mha_kernel; fa.b_o = w->dense_mha_bias;
    fa.h = (const __nv_bfloat16 *)io->h; fa.h_out = (__nv_bfloat16 *)io->h_out;
    return fused_fwd_launch(fa, io->e, io->e_out, io->qkv, st);
  }

  EdgeParams ep = make_edge_params(cfg, w);
  if (edge) {
    ep.e = io->e; ep.E = ws.E; ep.G = ws.G;
    if ((rc = edge_proj_fwd_launch(ep, a.dtype, st))) return rc;
  }
  AttnParams P = make_attn_params(&a);
  P.qkv = io->qkv; P.E = ws.E; P.G = a.gate_input ? ws.G : nullptr; P.M = io->adj; P.mask = io->mask;
  P.v_att = io->v_att; P.h_hat = ws.Hhat; P.lse = io->lse; P.deg = io->deg;
  if ((rc = attn_staged_fwd(P, a.dtype, st))) return rc;
  if (residual) {
    ep.h_hat = ws.Hhat; ep.e_out = io->e_out;
    if ((rc = edge_out_fwd_launch(ep, a.dtype, st))) return rc;
  }
  // h' = V_att W_O + b_O + h  (graph_xformer_model_base.py:136-140)
  LinearArgs lo;
  memset(&lo, 0, sizeof(lo));
  lo.x = io->v_att; lo.W = w->dense_mha_kernel; lo.bias = w->dense_mha_bias; lo.res = io->h; lo.out = io->h_out;
  lo.R = R; lo.din = d; lo.dout = d;
  return linear_launch(lo, a.dtype, st);
}

int egt_block_bwd(const egt_block_cfg_t *cfg, const egt_block_weights_t *w, const egt_block_grads_t *g,
                  const egt_block_bwd_io_t *io, void *stream) {
  int rc = check_block_cfg(cfg);
  if (rc) return rc;
  egt_attn_cfg_t a = derived_attn_cfg(cfg);
  if ((rc = check_attn_cfg(&a))) return rc;
  if ((rc = check_device())) return rc;
  EGT_REQUIRE(w && g && io, EGT_E_ARG, "weights / grads / io is NULL");
  EGT_REQUIRE(io->h && io->qkv && io->v_att && io->lse && io->deg && io->dh_out && io->dh, EGT_E_ARG,
              "h, qkv, v_att, lse, deg, dh_out, dh must be non-NULL");
  const bool edge = cfg->edge_channel_type != EGT_EDGE_NONE;
  const bool residual = cfg->edge_channel_type >= EGT_EDGE_RESIDUAL;
  EGT_REQUIRE(!edge || (io->e && io->de), EGT_E_ARG, "e / de is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const int R = a.B * a.N, d = a.h * a.dk;
  size_t need = egt_block_workspace_bytes(cfg, 1);
  EGT_REQUIRE(io->workspace && io->workspace_bytes >= need, EGT_E_ARG, "workspace too small: %zu < %zu",
              io->workspace_bytes, need);
  const bool have_de = io->de_out != nullptr;
  BlockWs ws = carve(cfg, 1, io->workspace, have_de);
  const bool wide = wide_supported(cfg, a.dtype) && !g_force_staged;
  const bool fused = (fused_supported(cfg, a.dtype) && !g_force_staged) || wide;   // the forward saved a pre-scaled Q
  const bool fused_bwd = fused && !wide && have_de;
  const bool wide_bwd = wide && have_de && wide_bwd_supported(cfg);
  g_last_path = (fused_bwd || wide_bwd) ? 1 : 0;

  if (wide_bwd) {   // width-generic fused backward (wide_bwd.cu); node side on node_tc.cu (d = 64) or the staged kernels
    const bool node_tc = d == 64 && a.h == 8;
    SideStream *side = side_stream(st);
    cudaStream_t sb = side_fork(side, st);             // folded weights next to the first node kernel
    if ((rc = wide_prep_launch(cfg, w, (WidePrep *)ws.prep, sb))) return rc;
    if ((rc = side_mark(side, st, sb))) return rc;     // what wide_bwd has to wait for
    if (node_tc) {
      const bool idle_sms = a.B * ((a.N + 127) / 128) + 20 <= 148;      // the fused kernel's grid leaves SMs for the side launch
      if ((rc = node_bwd1_launch(io->dh_out, io->v_att, w->dense_mha_kernel, ws.d_v_att, g->dense_mha_kernel,
                                 g->dense_mha_bias, R, w, a.clip_lo, a.clip_hi, nullptr, st, idle_sms ? sb : st))) return rc;
    } else if (ws.nblas) {
      if ((rc = node_blas_bwd1(io->dh_out, io->v_att, w, g, ws.d_v_att, R, d, ws.nblas, st, sb))) return rc;   // dW_O, db_O: side
    } else {
      LinearArgs l1;  // dV_att = dh' W_O^T ; dW_O += V_att^T dh' ; db_O += colsum(dh')
      memset(&l1, 0, sizeof(l1));
      l1.x = io->dh_out; l1.W = w->dense_mha_kernel; l1.trans = 1; l1.out = ws.d_v_att; l1.R = R; l1.din = d; l1.dout = d;
      if ((rc = linear_launch(l1, a.dtype, st))) return rc;
      XtyArgs x1;
      memset(&x1, 0, sizeof(x1));
      x1.X = io->v_att; x1.Y = io->dh_out; x1.dW = g->dense_mha_kernel; x1.db = g->dense_mha_bias; x1.R = R; x1.dx = d; x1.dy = d;
      if ((rc = xty_launch(x1, a.dtype, st))) return rc;
    }
    if ((rc = side_wait(side, st, sb))) return rc;
    const int tiles = (a.N + 127) / 128;
    if (tiles > 1 || wide_bwd_splits(cfg) > 1) EGT_CHECK_CUDA(cudaMemsetAsync(ws.d_qkv_f32, 0, (size_t)R * 3 * d * sizeof(float), st));
    WideBwdArgs fb;
    memset(&fb, 0, sizeof(fb));
    fb.B = a.B; fb.N = a.N; fb.mask = io->mask; fb.prep = (const WidePrep *)ws.prep;
    fb.v_att = (const __nv_bfloat16 *)io->v_att; fb.d_v_att = (const __nv_bfloat16 *)ws.d_v_att;
    fb.lse = io->lse; fb.deg = io->deg; fb.d_qkv = ws.d_qkv_f32; fb.partials = ws.partials;
    fb.clip_lo = a.clip_lo; fb.clip_hi = a.clip_hi; fb.dq_scale = 1.0f / sqrtf((float)a.dk); fb.ln_eps = cfg->ln_eps;
    fb.scale_degree = a.scale_degree; fb.scaler_type = a.scaler_type; fb.num_virtual_nodes = a.num_virtual_nodes;
    fb.rand_mask = a.training && a.random_mask_prob > 0.f;
    fb.rand_thr = (uint32_t)ceilf(a.random_mask_prob * 65536.0f - 0.5f);
    fb.seed = a.seed; fb.offset = a.offset; fb.offset_dev = a.offset_dev;
    if ((rc = wide_bwd_launch(cfg, fb, io->e, io->de_out, io->de, io->qkv, st))) return rc;
    // the fold of the edge-side weight gradients feeds nothing downstream: next to the last node kernels (it writes the
    // edge weights' gradients, they write the node weights')
    sb = side_fork(side, st);
    if ((rc = wide_bwd_finalize_launch(cfg, ws.partials, w, g, (const WidePrep *)ws.prep, sb))) return rc;
    if (node_tc) {
      if ((rc = node_bwd2_launch(io->h, io->dh_out, ws.d_qkv_f32, nullptr, w->norm_mha_gamma, w->norm_mha_beta, cfg->ln_eps,
                                 w->dense_qkv_kernel, io->dh, g->dense_qkv_kernel, g->dense_qkv_bias, g->norm_mha_gamma,
                                 g->norm_mha_beta, R, nullptr, 0, w, g, st, st))) return rc;   // the side stream is busy with the fold
      return side_join(side, st, sb);
    }
    // node side: hn = LN(h); dW_qkv += hn^T dqkv; dhn = dqkv W_qkv^T; dh = LN_bwd(dhn) + dh'
    if (ws.nblas) {
      if ((rc = node_blas_bwd2(io->h, ws.d_qkv_f32, w, g, cfg->ln_eps, ws.dhn, R, d, ws.nblas, st, sb, side ? side->aux : nullptr))) return rc;
      LnBwdArgs lb;
      memset(&lb, 0, sizeof(lb));
      lb.x = io->h; lb.dy = ws.dhn; lb.dres = io->dh_out; lb.gamma = w->norm_mha_gamma; lb.eps = cfg->ln_eps;
      lb.dx = io->dh; lb.dgamma = g->norm_mha_gamma; lb.dbeta = g->norm_mha_beta; lb.R = R; lb.D = d;
      if ((rc = ln_bwd_launch(lb, a.dtype, st))) return rc;
      return side_join(side, st, sb);
    }
    if ((rc = side_join(side, st, sb))) return rc;     // staged node kernels below: keep them behind the fold
    LinearArgs l2;
    memset(&l2, 0, sizeof(l2));
    l2.x = io->h; l2.ln_gamma = w->norm_mha_gamma; l2.ln_beta = w->norm_mha_beta; l2.ln_eps = cfg->ln_eps;
    l2.xn_out = ws.hn; l2.R = R; l2.din = d; l2.dout = 0;
    if ((rc = linear_launch(l2, a.dtype, st))) return rc;
    XtyArgs x2;
    memset(&x2, 0, sizeof(x2));
    x2.X = ws.hn; x2.x_f32 = 1; x2.Y = ws.d_qkv_f32; x2.y_f32 = 1; x2.dW = g->dense_qkv_kernel; x2.db = g->dense_qkv_bias;
    x2.R = R; x2.dx = d; x2.dy = 3 * d;
    if ((rc = xty_launch(x2, a.dtype, st))) return rc;
    LinearArgs l3;
    memset(&l3, 0, sizeof(l3));
    l3.x = ws.d_qkv_f32; l3.x_f32 = 1; l3.W = w->dense_qkv_kernel; l3.trans = 1; l3.out = ws.dhn; l3.out_f32 = 1;
    l3.R = R; l3.din = 3 * d; l3.dout = d;
    if ((rc = linear_launch(l3, a.dtype, st))) return rc;
    LnBwdArgs lb;
    memset(&lb, 0, sizeof(lb));
    lb.x = io->h; lb.dy = ws.dhn; lb.dres = io->dh_out; lb.gamma = w->norm_mha_gamma; lb.eps = cfg->ln_eps;
    lb.dx = io->dh; lb.dgamma = g->norm_mha_gamma; lb.dbeta = g->norm_mha_beta; lb.R = R; lb.D = d;
    return ln_bwd_launch(lb, a.dtype, st);
  }

  if (fused_bwd) {
    // dW_O / db_O feed nothing in the backward pass: a second small launch on the side stream, next to the fused backward
    // kernel (whose 128-CTA grid leaves SMs idle at the headline shape); joined before the call returns
    // -- only when that grid really leaves 20 SMs idle: otherwise the second launch competes with it
    const int bwd_ctas = a.B * ((a.N + 127) / 128) * fused_bwd_key_splits(a.B, a.N);
    SideStream *side = bwd_ctas + 20 <= 148 ? side_stream(st) : nullptr;
    cudaStream_t sb = side_fork(side, st);
    if ((rc = node_bwd1_launch(io->dh_out, io->v_att, w->dense_mha_kernel, ws.d_v_att, g->dense_mha_kernel,
                               g->dense_mha_bias, R, w, a.clip_lo, a.clip_hi, (FusedPrep *)ws.prep, st, sb))) return rc;
    // N x N part in one kernel (fused_bwd.cu): de, dQ|dK|dV and the edge-side weight-gradient partial sums
    const int tiles = (a.N + 127) / 128, ksplit = fused_bwd_key_splits(a.B, a.N);
    if (tiles > 1 || ksplit > 1) EGT_CHECK_CUDA(cudaMemsetAsync(ws.d_qkv_f32, 0, (size_t)R * 3 * d * sizeof(float), st));
    FusedBwdArgs fb;
    memset(&fb, 0, sizeof(fb));
    fb.B = a.B; fb.N = a.N; fb.mask = io->mask; fb.prep = (const FusedPrep *)ws.prep;
    fb.v_att = (const __nv_bfloat16 *)io->v_att; fb.d_v_att = (const __nv_bfloat16 *)ws.d_v_att;
    fb.lse = io->lse; fb.deg = io->deg; fb.d_qkv = ws.d_qkv_f32; fb.partials = ws.partials;
    // one row tile per graph: every dQ / dK / dV element has exactly one writer, so the kernel writes bf16 -- what the
    // node kernel's tensor-core products consume -- and node_bwd2 loads its tiles by TMA instead of converting float32
    fb.d_qkv_bf = (__nv_bfloat16 *)ws.d_qkv_bf;
    fb.clip_lo = a.clip_lo; fb.clip_hi = a.clip_hi; fb.dq_scale = 1.0f / sqrtf((float)a.dk); fb.ln_eps = cfg->ln_eps;
    fb.scale_degree = a.scale_degree; fb.scaler_type = a.scaler_type; fb.num_virtual_nodes = a.num_virtual_nodes;
    fb.rand_mask = a.training && a.random_mask_prob > 0.f;
    fb.rand_thr = (uint32_t)ceilf(a.random_mask_prob * 65536.0f - 0.5f);
    fb.seed = a.seed; fb.offset = a.offset; fb.offset_dev = a.offset_dev;
    if ((rc = fused_bwd_launch(fb, io->e, io->de_out, io->de, io->qkv, st))) return rc;
    // (splitting this kernel into its dh half and its weight-gradient half on two streams was measured: the dh half alone
    // takes as long as the whole kernel -- 26 us, the float32 -> bf16 staging of dqkv -- and the step got 2 % slower)
    if ((rc = node_bwd2_launch(io->h, io->dh_out, ws.d_qkv_f32, ws.d_qkv_bf, w->norm_mha_gamma, w->norm_mha_beta, cfg->ln_eps,
                               w->dense_qkv_kernel, io->dh, g->dense_qkv_kernel, g->dense_qkv_bias, g->norm_mha_gamma,
                               g->norm_mha_beta, R, ws.partials, a.B * tiles * ksplit, w, g, st, st))) return rc;
    return side_join(side, st, sb);
  }

  // dV_att = dh' W_O^T ; dW_O += V_att^T dh' ; db_O += colsum(dh')
  LinearArgs l1;
  memset(&l1, 0, sizeof(l1));
  l1.x = io->dh_out; l1.W = w->dense_mha_kernel; l1.trans = 1; l1.out = ws.d_v_att; l1.R = R; l1.din = d; l1.dout = d;
  if ((rc = linear_launch(l1, a.dtype, st))) return rc;
  XtyArgs x1;
  memset(&x1, 0, sizeof(x1));
  x1.X = io->v_att; x1.Y = io->dh_out; x1.dW = g->dense_mha_kernel; x1.db = g->dense_mha_bias; x1.R = R; x1.dx = d; x1.dy = d;
  if ((rc = xty_launch(x1, a.dtype, st))) return rc;

  EdgeParams ep = make_edge_params(cfg, w);
  ep.g_ln_g = g->norm_edge_gamma; ep.g_ln_b = g->norm_edge_beta;
  ep.g_w_e = g->dense_edge_b_kernel; ep.g_b_e = g->dense_edge_b_bias;
  ep.g_w_g = g->attention_gates_kernel; ep.g_b_g = g->attention_gates_bias;
  ep.e = io->e;
  if (edge) {   // recompute E, G
    ep.E = ws.E; ep.G = ws.G;
    if ((rc = edge_proj_fwd_launch(ep, a.dtype, st))) return rc;
  }
  const bool have_de_out = residual && io->de_out;
  if (have_de_out) {   // dH_ext = de' W_r^T
    EdgeParams e1 = ep;
    e1.de_out = io->de_out; e1.d_h_ext = ws.dHext; e1.g_w_r = nullptr; e1.g_b_r = nullptr;
    if ((rc = edge_out_bwd_launch(e1, a.dtype, st))) return rc;
  }
  AttnParams P = make_attn_params(&a);
  P.qkv = io->qkv; P.E = ws.E; P.G = a.gate_input ? ws.G : nullptr; P.M = io->adj; P.mask = io->mask;
  P.lse = (float *)io->lse; P.deg = (float *)io->deg;
  P.d_v_att = ws.d_v_att; P.d_h_hat = have_de_out ? ws.dHext : nullptr; P.d_qkv = ws.d_qkv;
  P.dE = ws.dE; P.dG = a.gate_input ? ws.dG : nullptr; P.row_ws = ws.row_ws;
  P.h_hat = have_de_out ? ws.Hhat : nullptr;     // row pass re-materialises H_hat for dW_r
  P.v_att = const_cast<void *>(io->v_att);       // saved forward output: D = sum_dd dV_att * V_att (attn_fast.cu)
  P.dS_ws = ws.dS; P.As_ws = ws.As;
  if (fused) { P.dq_scale = P.scale; P.scale = 1.0f; }   // the fused forward saved a pre-scaled Q
  if ((rc = attn_staged_bwd(P, a.dtype, st))) return rc;
  if (have_de_out) {   // dW_r += H_hat^T de' ; db_r += colsum(de')
    EdgeParams e2 = ep;
    e2.de_out = io->de_out; e2.h_hat = ws.Hhat; e2.d_h_ext = nullptr;
    e2.g_w_r = g->dense_edge_r_kernel; e2.g_b_r = g->dense_edge_r_bias;
    if ((rc = edge_out_bwd_launch(e2, a.dtype, st))) return rc;
  }
  if (edge) {
    ep.dE = ws.dE; ep.dG = ws.dG; ep.de = io->de; ep.de_out = io->de_out;
    if ((rc = edge_proj_bwd_launch(ep, a.dtype, st))) return rc;
  }
  // node side: hn = LN(h); dW_qkv += hn^T dqkv; dhn = dqkv W_qkv^T; dh = LN_bwd(dhn) + dh'
  LinearArgs l2;
  memset(&l2, 0, sizeof(l2));
  l2.x = io->h; l2.ln_gamma = w->norm_mha_gamma; l2.ln_beta = w->norm_mha_beta; l2.ln_eps = cfg->ln_eps;
  l2.xn_out = ws.hn; l2.R = R; l2.din = d; l2.dout = 0;
  if ((rc = linear_launch(l2, a.dtype, st))) return rc;
  XtyArgs x2;
  memset(&x2, 0, sizeof(x2));
  x2.X = ws.hn; x2.x_f32 = 1; x2.Y = ws.d_qkv; x2.dW = g->dense_qkv_kernel; x2.db = g->dense_qkv_bias;
  x2.R = R; x2.dx = d; x2.dy = 3 * d;
  if ((rc = xty_launch(x2, a.dtype, st))) return rc;
  LinearArgs l3;
  memset(&l3, 0, sizeof(l3));
  l3.x = ws.d_qkv; l3.W = w->dense_qkv_kernel; l3.trans = 1; l3.out = ws.dhn; l3.out_f32 = 1;
  l3.R = R; l3.din = 3 * d; l3.dout = d;
  if ((rc = linear_launch(l3, a.dtype, st))) return rc;
  LnBwdArgs lb;
  memset(&lb, 0, sizeof(lb));
  lb.x = io->h; lb.dy = ws.dhn; lb.dres = io->dh_out; lb.gamma = w->norm_mha_gamma; lb.eps = cfg->ln_eps;
  lb.dx = io->dh; lb.dgamma = g->norm_mha_gamma; lb.dbeta = g->norm_mha_beta; lb.R = R; lb.D = d;
  return ln_bwd_launch(lb, a.dtype, st);
}

}  // extern "C"
