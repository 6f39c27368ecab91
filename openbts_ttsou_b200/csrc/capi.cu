// capi.cu -- the extern "C" surface declared in include/btsdsp.h: context, table construction,
// single-vector calls (host pointers), batched calls (device pointers), host-buffer pipelines.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/btsdsp.h"
#include "kernels.cuh"
#include "tables_host.h"

using namespace btsdsp;

namespace {

std::string g_create_error;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct btsdsp_ctx {
  int device = 0, sps = 1;
  DevTables *T = nullptr;       // device
  DevTables *hT = nullptr;      // host mirror (pinned)
  cudaStream_t st = nullptr, st_in = nullptr, st_out = nullptr, st_side = nullptr;
  long long rx_seg = 0;         // btsdsp_rx_stream_dev: chunks per segment of the overlapped pipeline (0 = one launch each)
  int rx_res_ctas = 0;          // ... and the resampler's CTA cap while it shares the GPU with the demod kernels
  int rx_light = 1;             // ... and whether those launches use the resampler's light (one-tile) configuration
  long long host_seg = 4000;    // layer-3 pipelines: chunks per copy/compute segment (4000 = 27.6 MB of complex64 samples)
  std::string err;
  std::atomic<long long> launches{0};
  DevBuf buf[16];               // grow-only device scratch, by role
  DevBuf pin[4];                // grow-only pinned staging
  std::vector<cudaEvent_t> events;
  // scratch of the layer-2 calls, one set per caller stream: calls on one stream are ordered by the stream, calls on
  // different streams may overlap on the device and must not share the EqParams records / correlation rows
  struct StreamScratch { cudaStream_t st; DevBuf eqp, scratch, res; };
  std::vector<StreamScratch *> per_stream;
  std::mutex mu;                // guards per_stream / err / launches when layer-2 calls come from several host threads
  std::atomic<int> graphs_alive{0};   // btsdsp_graph objects not yet destroyed (they pin the scratch buffers' addresses)
  bool copy_only = false;       // btsdsp_set_copy_only: the host pipelines move their bytes but launch nothing
  bool timing = false;          // btsdsp_set_timing: bracket the kernels of the receive path with events
  cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
  int tev_used = 0;
};

namespace {

enum { B_A = 0, B_B, B_C, B_D, B_SCRATCH, B_RAW, B_RES, B_FLAG, B_AMP, B_TOA, B_SOFT, B_TSC, B_EQP };

int fail(btsdsp_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
  char msg[512];
  if (e != cudaSuccess) snprintf(msg, sizeof msg, "%s: %s", what, cudaGetErrorString(e));
  else snprintf(msg, sizeof msg, "%s", what);
  if (c) { std::lock_guard<std::mutex> lk(c->mu); c->err = msg; } else g_create_error = msg;
  return code;
}

#define CK(call)                                                          \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) return fail(ctx, BTSDSP_ECUDA, #call, e__);   \
  } while (0)

#define ARG(cond)                                                                 \
  do {                                                                            \
    if (!(cond)) return fail(ctx, BTSDSP_EINVAL, "invalid argument: " #cond);     \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int grow(btsdsp_ctx *ctx, DevBuf &b, size_t bytes, bool pinned = false) {
  if (bytes <= b.cap) return BTSDSP_OK;
  // a captured graph holds the addresses of the scratch buffers its calls used: they must not move while it lives
  if (b.p && ctx->graphs_alive.load() > 0)
    return fail(ctx, BTSDSP_EINVAL, "a scratch buffer would have to grow while a captured graph refers to it: destroy the graph, "
                                    "or issue the largest call once before capturing");
  size_t want = bytes + bytes / 8 + 256;
  if (b.p) { if (pinned) cudaFreeHost(b.p); else cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  cudaError_t e = pinned ? cudaMallocHost(&b.p, want) : cudaMalloc(&b.p, want);
  if (e != cudaSuccess) { b.p = nullptr; return fail(ctx, BTSDSP_ENOMEM, pinned ? "cudaMallocHost" : "cudaMalloc", e); }
  b.cap = want;
  return BTSDSP_OK;
}
#define GROW(slot, bytes)                                         \
  do {                                                            \
    int r__ = grow(ctx, ctx->buf[slot], (bytes));                 \
    if (r__ != BTSDSP_OK) return r__;                             \
  } while (0)

template <class Tp> Tp *dbuf(btsdsp_ctx *ctx, int slot) { return (Tp *)ctx->buf[slot].p; }

// the scratch set of a caller stream (created on first use; a handful of streams per context is the expected use)
btsdsp_ctx::StreamScratch *stream_scratch(btsdsp_ctx *ctx, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (auto *s : ctx->per_stream) if (s->st == st) return s;
  auto *s = new btsdsp_ctx::StreamScratch();
  s->st = st;
  ctx->per_stream.push_back(s);
  return s;
}
#define GROWBUF(b, bytes)                                         \
  do {                                                            \
    int r__ = grow(ctx, (b), (bytes));                            \
    if (r__ != BTSDSP_OK) return r__;                             \
  } while (0)

int check_launch(btsdsp_ctx *ctx, const char *what, int nlaunch = 1) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, BTSDSP_ECUDA, what, e);
  ctx->launches += nlaunch;
  return BTSDSP_OK;
}
#define LAUNCHED(what, n)                              \
  do {                                                 \
    int r__ = check_launch(ctx, what, n);              \
    if (r__ != BTSDSP_OK) return r__;                  \
  } while (0)

// Everything the reference computes at start-up.  Host: the libm-in-double parts (trig tables :207-212,
// pulse :411-430) and the filter scaling; device: everything that is float arithmetic over the tables
// (rotation tables, sinc grid, midamble / RACH autocorrelation peaks) -- through the product's own kernels.
int build_tables(btsdsp_ctx *ctx) {
  DevTables *h = ctx->hT;
  const int sps = ctx->sps;
  host_fill_tables(h, sps);

  CK(cudaMemcpyAsync(ctx->T, h, sizeof(DevTables), cudaMemcpyHostToDevice, ctx->st));
  launch_init_tables(ctx->T, ctx->st);
  LAUNCHED("init tables", 2);

  // generateMidamble :779-828 and generateRACHSequence :830-857 on the device
  GROW(B_A, 64 * kMaxSps * sizeof(cf));
  GROW(B_B, 64 * kMaxSps * sizeof(cf));
  GROW(B_C, 64 * kMaxSps * sizeof(cf));
  GROW(B_D, 256);
  cf *dMid = dbuf<cf>(ctx, B_A), *dFull = dbuf<cf>(ctx, B_B), *dAc = dbuf<cf>(ctx, B_C);
  uint8_t *dBits = dbuf<uint8_t>(ctx, B_D);
  cf *dPeak = (cf *)(dBits + 64);
  float *dIdx = (float *)(dBits + 128);
  for (int t = 0; t < 8; t++) {
    uint8_t bits[26];
    for (int i = 0; i < 26; i++) bits[i] = kTSC[t][i] == '1';
    CK(cudaMemcpyAsync(dBits, bits, 26, cudaMemcpyHostToDevice, ctx->st));
    const int nmid = 16 * sps, nfull = 26 * sps;
    launch_modulate_impulse(ctx->T, dBits + 5, 16, dMid, ctx->st);                                  // :794-797
    launch_modulate(ctx->T, dBits, 26, 1, 0, nullptr, 0, dFull, nfull, ctx->st);                    // :798-801
    launch_scale_vector(dMid, nmid, 0, mk(-1.0F, 0.0F), ctx->st);                                   // :811
    launch_scale_vector(dFull, nfull, 0, mk(0.0F, 1.0F), ctx->st);                                  // :812
    launch_convolve(dFull, nfull, 0, dMid, nmid, 0, dAc, (nmid % 2) ? nmid / 2 : nmid / 2 - 1, nfull, 1, ctx->st);  // :814
    launch_peak_detect(ctx->T, dAc, nfull, dPeak, dIdx, nullptr, ctx->st);                          // :821
    LAUNCHED("midamble init", 6);
    cf gain; float toa;
    CK(cudaMemcpyAsync(&gain, dPeak, sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&toa, dIdx, sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(ctx->T->mid_seq[t], dMid, nmid * sizeof(cf), cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    toa = toa - (float)(5 * sps);                                                                   // :822
    CK(cudaMemcpy(&ctx->T->mid_toa[t], &toa, sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&ctx->T->mid_gain[t], &gain, sizeof(cf), cudaMemcpyHostToDevice));
  }
  {
    uint8_t bits[41];
    for (int i = 0; i < 41; i++) bits[i] = kRACH[i] == '1';
    CK(cudaMemcpyAsync(dBits, bits, 41, cudaMemcpyHostToDevice, ctx->st));
    const int n = 41 * sps;
    launch_modulate(ctx->T, dBits, 41, 1, 0, nullptr, 0, dFull, n, ctx->st);                        // :837-840
    launch_convolve(dFull, n, 0, dFull, n, 0, dAc, (n % 2) ? n / 2 : n / 2 - 1, n, 1, ctx->st);      // :844
    launch_peak_detect(ctx->T, dAc, n, dPeak, dIdx, nullptr, ctx->st);                              // :851
    LAUNCHED("rach init", 3);
    cf gain; float toa;
    CK(cudaMemcpyAsync(&gain, dPeak, sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&toa, dIdx, sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(ctx->T->rach_seq, dFull, n * sizeof(cf), cudaMemcpyDeviceToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    CK(cudaMemcpy(&ctx->T->rach_toa, &toa, sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(&ctx->T->rach_gain, &gain, sizeof(cf), cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(h, ctx->T, sizeof(DevTables), cudaMemcpyDeviceToHost));
  upload_resampler_taps(h);
  if (sps == 1) upload_rach_taps(h);   // the tuned access-burst kernels are sps == 1 only; __constant__ data is per device
  CK(cudaGetLastError());
  return BTSDSP_OK;
}

BurstSrc make_src(const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens, long long first, int sps) {
  BurstSrc s;
  s.base = (const cf *)bursts; s.pitch = pitch; s.lens = lens; s.first = first; s.sps = sps;
  return s;
}

int conv_sizes(int la, int lb, int span, int *start, int *outsz) {   // sigProcLib.cpp:278-304
  switch (span) {
    case BTSDSP_FULL_SPAN:    *start = 0;  *outsz = la + lb - 1; break;
    case BTSDSP_OVERLAP_ONLY: *start = la; *outsz = abs(la - lb) + 1; break;
    case BTSDSP_START_ONLY:   *start = 0;  *outsz = la; break;
    case BTSDSP_WITH_TAIL:    *start = lb; *outsz = la; break;
    case BTSDSP_NO_DELAY:     *start = (lb % 2) ? lb / 2 : lb / 2 - 1; *outsz = la; break;
    default: return -1;
  }
  return 0;
}

int conv_or_corr(btsdsp_ctx *ctx, int corr, const btsdsp_cf32 *a, int la, int a_real, const btsdsp_cf32 *b, int lb,
                 int b_real, btsdsp_cf32 *c, int cap, int span) {
  ARG(ctx && a && b && la > 0 && lb > 0);
  int start, outsz;
  if (conv_sizes(la, lb, span, &start, &outsz)) return fail(ctx, BTSDSP_EINVAL, "unknown span type");
  if (cap < outsz || !c) return outsz;
  DeviceGuard g(ctx->device);
  GROW(B_A, la * sizeof(cf)); GROW(B_B, lb * sizeof(cf)); GROW(B_C, outsz * sizeof(cf));
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), a, la * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_B), b, lb * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_convolve(dbuf<cf>(ctx, B_A), la, a_real, dbuf<cf>(ctx, B_B), lb, b_real, dbuf<cf>(ctx, B_C), start, outsz, corr,
                  ctx->st);
  LAUNCHED("convolve", 1);
  CK(cudaMemcpyAsync(c, dbuf<cf>(ctx, B_C), outsz * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return outsz;
}

}  // namespace

extern "C" {

int btsdsp_version(void) { return 100; }

int btsdsp_create(btsdsp_ctx **out, int device, int sps) {
  btsdsp_ctx *ctx = nullptr;
  if (!out) return fail(nullptr, BTSDSP_EINVAL, "ctx pointer is null");
  *out = nullptr;
  if (!(sps == 1 || sps == 2 || sps == 4)) return fail(nullptr, BTSDSP_EINVAL, "sps must be 1, 2 or 4");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, BTSDSP_ECUDA, "no CUDA device: this library has no CPU fallback", e);
  if (device < 0 || device >= ndev) return fail(nullptr, BTSDSP_EINVAL, "device index out of range");
  DeviceGuard g(device);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, BTSDSP_ECUDA, "cudaGetDeviceProperties", e);
  if (prop.major < 10) return fail(nullptr, BTSDSP_ECUDA, "device is not sm_100 (B200) class; kernels are built for sm_100a only");
  ctx = new btsdsp_ctx;
  ctx->device = device;
  ctx->sps = sps;
  int rc = BTSDSP_OK;
  auto setup = [&]() -> int {
    CK(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->st_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->st_out, cudaStreamNonBlocking));
    {
      int lo = 0, hi = 0;                                      // the resampler side stream outranks the demod kernels
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&ctx->st_side, cudaStreamNonBlocking, hi));
    }
    if (const char *e = getenv("BTSDSP_RX_SEG")) ctx->rx_seg = atoll(e);
    if (const char *e = getenv("BTSDSP_RX_RES_CTAS")) ctx->rx_res_ctas = atoi(e);
    if (const char *e = getenv("BTSDSP_RX_LIGHT")) ctx->rx_light = atoi(e);
    if (const char *e = getenv("BTSDSP_HOST_SEG")) { const long long v = atoll(e); if (v >= 250) ctx->host_seg = v; }
    CK(cudaMalloc(&ctx->T, sizeof(DevTables)));
    CK(cudaMallocHost(&ctx->hT, sizeof(DevTables)));
    int ce = configure_kernels();
    if (!ce) ce = configure_resamplers();
    if (ce) return fail(ctx, BTSDSP_ECUDA, "cudaFuncSetAttribute", (cudaError_t)ce);
    return build_tables(ctx);
  };
  rc = setup();
  if (rc != BTSDSP_OK) {
    g_create_error = ctx->err;
    btsdsp_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return BTSDSP_OK;
}

int btsdsp_destroy(btsdsp_ctx *ctx) {
  if (!ctx) return BTSDSP_OK;
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  for (auto &b : ctx->buf) if (b.p) cudaFree(b.p);
  for (auto *ss : ctx->per_stream) { if (ss->eqp.p) cudaFree(ss->eqp.p); if (ss->scratch.p) cudaFree(ss->scratch.p); if (ss->res.p) cudaFree(ss->res.p); delete ss; }
  for (auto &b : ctx->pin) if (b.p) cudaFreeHost(b.p);
  for (auto ev : ctx->events) cudaEventDestroy(ev);
  for (auto ev : ctx->tev) if (ev) cudaEventDestroy(ev);
  if (ctx->T) cudaFree(ctx->T);
  if (ctx->hT) cudaFreeHost(ctx->hT);
  if (ctx->st) cudaStreamDestroy(ctx->st);
  if (ctx->st_in) cudaStreamDestroy(ctx->st_in);
  if (ctx->st_out) cudaStreamDestroy(ctx->st_out);
  if (ctx->st_side) cudaStreamDestroy(ctx->st_side);
  delete ctx;
  return BTSDSP_OK;
}

const char *btsdsp_last_error(const btsdsp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int btsdsp_device(const btsdsp_ctx *ctx) { return ctx ? ctx->device : -1; }
int btsdsp_sps(const btsdsp_ctx *ctx) { return ctx ? ctx->sps : -1; }
long long btsdsp_launch_count(const btsdsp_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }
/* optional per-kernel timing of btsdsp_demod_normal_dev: events around k_detect_design and k_equalize_fast */
int btsdsp_set_timing(btsdsp_ctx *ctx, int enable) {
  ARG(ctx);
  DeviceGuard g(ctx->device);
  if (enable && !ctx->tev[0])
    for (int i = 0; i < 4; i++) CK(cudaEventCreate(&ctx->tev[i]));
  ctx->timing = enable != 0;
  ctx->tev_used = 0;
  return BTSDSP_OK;
}
int btsdsp_get_timing(btsdsp_ctx *ctx, float *detect_ms, float *equalize_ms) {
  ARG(ctx && detect_ms && equalize_ms);
  if (!ctx->timing || ctx->tev_used < 3) return fail(ctx, BTSDSP_EINVAL, "no timed btsdsp_demod_normal_dev call recorded");
  DeviceGuard g(ctx->device);
  CK(cudaEventSynchronize(ctx->tev[2]));
  CK(cudaEventElapsedTime(detect_ms, ctx->tev[0], ctx->tev[1]));
  CK(cudaEventElapsedTime(equalize_ms, ctx->tev[1], ctx->tev[2]));
  return BTSDSP_OK;
}
/* measurement aid: the copy roofline of the host-buffer pipelines (same buffers, segments, streams and events, no kernels) */
int btsdsp_set_copy_only(btsdsp_ctx *ctx, int enable) {
  ARG(ctx);
  ctx->copy_only = enable != 0;
  return BTSDSP_OK;
}
/* ---- CUDA graphs for small-batch callers: record a sequence of layer-2 calls once, replay it with one launch ---- */
struct btsdsp_graph {
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;       // kernels in the captured sequence
  long long mark = 0;           // launch counter at capture begin
};
int btsdsp_graph_begin(btsdsp_ctx *ctx, void *stream, btsdsp_graph **out) {
  ARG(ctx && out && stream);    // the legacy default stream cannot be captured
  DeviceGuard g(ctx->device);
  auto *gr = new btsdsp_graph();
  gr->mark = ctx->launches.load();
  cudaError_t e = cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) { delete gr; return fail(ctx, BTSDSP_ECUDA, "cudaStreamBeginCapture", e); }
  ctx->graphs_alive++;
  *out = gr;
  return BTSDSP_OK;
}
int btsdsp_graph_end(btsdsp_ctx *ctx, void *stream, btsdsp_graph *gr) {
  ARG(ctx && gr && !gr->exec);
  DeviceGuard g(ctx->device);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &graph);
  if (e != cudaSuccess || !graph) return fail(ctx, BTSDSP_ECUDA, "cudaStreamEndCapture", e);
  gr->launches = ctx->launches.load() - gr->mark;
  ctx->launches = gr->mark;     // nothing has run yet: the kernels are counted when the graph is launched
  e = cudaGraphInstantiate(&gr->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { gr->exec = nullptr; return fail(ctx, BTSDSP_ECUDA, "cudaGraphInstantiate", e); }
  return BTSDSP_OK;
}
int btsdsp_graph_launch(btsdsp_ctx *ctx, btsdsp_graph *gr, void *stream) {
  ARG(ctx && gr && gr->exec);
  DeviceGuard g(ctx->device);
  CK(cudaGraphLaunch(gr->exec, (cudaStream_t)stream));
  ctx->launches += gr->launches;
  return BTSDSP_OK;
}
int btsdsp_graph_destroy(btsdsp_ctx *ctx, btsdsp_graph *gr) {
  ARG(ctx && gr);
  DeviceGuard g(ctx->device);
  if (gr->exec) cudaGraphExecDestroy(gr->exec);
  delete gr;
  ctx->graphs_alive--;
  return BTSDSP_OK;
}

int btsdsp_synchronize(btsdsp_ctx *ctx) {
  ARG(ctx);
  DeviceGuard g(ctx->device);
  CK(cudaDeviceSynchronize());
  return BTSDSP_OK;
}

int btsdsp_get_table(btsdsp_ctx *ctx, int id, int idx, float *dst, int cap) {
  ARG(ctx);
  const DevTables *h = ctx->hT;
  const int sps = ctx->sps;
  const void *src = nullptr;
  int n = 0;
  float meta[3];
  switch (id) {
    case BTSDSP_T_COS: src = h->cosT; n = kTrig + 1; break;
    case BTSDSP_T_SIN: src = h->sinT; n = kTrig + 1; break;
    case BTSDSP_T_ROT: src = h->rot; n = 2 * 157 * sps; break;
    case BTSDSP_T_REVROT: src = h->revrot; n = 2 * 157 * sps; break;
    case BTSDSP_T_PULSE: src = h->pulse; n = 2 * h->pulse_len; break;
    case BTSDSP_T_MID_SEQ: ARG(idx >= 0 && idx < 8); src = h->mid_seq[idx]; n = 2 * 16 * sps; break;
    case BTSDSP_T_MID_META:
      ARG(idx >= 0 && idx < 8);
      meta[0] = h->mid_toa[idx]; meta[1] = h->mid_gain[idx].x; meta[2] = h->mid_gain[idx].y; src = meta; n = 3; break;
    case BTSDSP_T_RACH_SEQ: src = h->rach_seq; n = 2 * 41 * sps; break;
    case BTSDSP_T_RACH_META:
      meta[0] = h->rach_toa; meta[1] = h->rach_gain.x; meta[2] = h->rach_gain.y; src = meta; n = 3; break;
    case BTSDSP_T_LPF_RX: src = h->lpf_rx; n = kRxTaps; break;
    case BTSDSP_T_LPF_TX: src = h->lpf_tx; n = kTxTaps; break;
    default: return fail(ctx, BTSDSP_EINVAL, "unknown table id");
  }
  if (dst && cap >= n) memcpy(dst, src, n * sizeof(float));
  return n;
}

void *btsdsp_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void btsdsp_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- layer 1 ---------------------------------------------------------------------------------------
int btsdsp_convolve(btsdsp_ctx *ctx, const btsdsp_cf32 *a, int la, int a_real, const btsdsp_cf32 *b, int lb, int b_real,
                    btsdsp_cf32 *c, int cap, int span) {
  return conv_or_corr(ctx, 0, a, la, a_real, b, lb, b_real, c, cap, span);
}
int btsdsp_correlate(btsdsp_ctx *ctx, const btsdsp_cf32 *a, int la, int a_real, const btsdsp_cf32 *b, int lb, int b_real,
                     btsdsp_cf32 *c, int cap, int span) {
  return conv_or_corr(ctx, 1, a, la, a_real, b, lb, b_real, c, cap, span);
}

int btsdsp_scale_vector(btsdsp_ctx *ctx, btsdsp_cf32 *v, int n, int real_only, btsdsp_cf32 scale) {
  ARG(ctx && v && n > 0);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf));
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), v, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_scale_vector(dbuf<cf>(ctx, B_A), n, real_only, mk(scale.re, scale.im), ctx->st);
  LAUNCHED("scale_vector", 1);
  CK(cudaMemcpyAsync(v, dbuf<cf>(ctx, B_A), n * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

int btsdsp_delay_vector(btsdsp_ctx *ctx, btsdsp_cf32 *v, int n, float delay) {
  ARG(ctx && v && n > 0);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_B, n * sizeof(cf));
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), v, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_delay_vector(ctx->T, dbuf<cf>(ctx, B_A), n, delay, dbuf<cf>(ctx, B_B), ctx->st);
  LAUNCHED("delay_vector", 2);
  CK(cudaMemcpyAsync(v, dbuf<cf>(ctx, B_A), n * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

int btsdsp_peak_detect(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, btsdsp_cf32 *peak, float *peak_index,
                       float *avg_power) {
  ARG(ctx && v && n > 1 && peak);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_D, 256);
  cf *dp = dbuf<cf>(ctx, B_D);
  float *di = (float *)(dp + 1), *da = di + 1;
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), v, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_peak_detect(ctx->T, dbuf<cf>(ctx, B_A), n, dp, di, da, ctx->st);
  LAUNCHED("peak_detect", 1);
  float res[4];
  CK(cudaMemcpyAsync(res, dp, 16, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  peak->re = res[0]; peak->im = res[1];
  if (peak_index) *peak_index = res[2];
  if (avg_power) *avg_power = res[3];
  return BTSDSP_OK;
}

int btsdsp_interpolate_point(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, float ix, btsdsp_cf32 *out) {
  ARG(ctx && v && n > 0 && out);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_D, 256);
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), v, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_interp_point(ctx->T, dbuf<cf>(ctx, B_A), n, ix, dbuf<cf>(ctx, B_D), ctx->st);
  LAUNCHED("interpolate_point", 1);
  CK(cudaMemcpyAsync(out, dbuf<cf>(ctx, B_D), sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

int btsdsp_energy_detect(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, unsigned window, float threshold,
                         float *avg_power, int *detected) {
  ARG(ctx && v && n > 0 && detected);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_D, 256);
  float *da = dbuf<float>(ctx, B_D);
  int *df = (int *)(da + 1);
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), v, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_energy_detect(dbuf<cf>(ctx, B_A), n, window, threshold, da, df, ctx->st);
  LAUNCHED("energy_detect", 1);
  float res[2];
  CK(cudaMemcpyAsync(res, da, 8, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  if (avg_power) *avg_power = res[0];
  memcpy(detected, &res[1], 4);
  return BTSDSP_OK;
}

int btsdsp_modulate_burst(btsdsp_ctx *ctx, const uint8_t *bits, int nbits, int guard, btsdsp_cf32 *out, int cap) {
  ARG(ctx && bits && nbits > 0 && guard >= 0);
  const int n = ctx->sps * (nbits + guard);
  ARG(n <= 157 * ctx->sps);            // the rotation table is 157*sps long (sigProcLib.cpp:215)
  if (cap < n || !out) return n;
  DeviceGuard g(ctx->device);
  GROW(B_D, nbits + 64); GROW(B_A, n * sizeof(cf));
  CK(cudaMemcpyAsync(dbuf<uint8_t>(ctx, B_D), bits, nbits, cudaMemcpyHostToDevice, ctx->st));
  launch_modulate(ctx->T, dbuf<uint8_t>(ctx, B_D), nbits, 1, guard, nullptr, 0, dbuf<cf>(ctx, B_A), n, ctx->st);
  LAUNCHED("modulate", 1);
  CK(cudaMemcpyAsync(out, dbuf<cf>(ctx, B_A), n * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return n;
}

int btsdsp_analyze_traffic_burst(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, unsigned tsc, float threshold,
                                 btsdsp_cf32 *amplitude, float *toa, int request_channel, btsdsp_cf32 *chan,
                                 float *chan_offset, int *detected) {
  ARG(ctx && burst && amplitude && toa && detected && tsc < 8);
  const int sps = ctx->sps;
  ARG(n >= 92 * sps);                   // the correlation window is samples [56*sps, 92*sps) (:951)
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_D, 1024); GROW(B_SCRATCH, scratch_per_burst(sps) * sizeof(cf));
  uint8_t *d = dbuf<uint8_t>(ctx, B_D);
  struct { int32_t len; int32_t flag; cf amp; float toa; float off; cf chan[6 * kMaxSps]; uint8_t tsc; } hres;
  int32_t *dLen = (int32_t *)d, *dFlag = dLen + 1;
  cf *dAmp = (cf *)(d + 8);
  float *dToa = (float *)(d + 16), *dOff = dToa + 1;
  cf *dChan = (cf *)(d + 24);
  uint8_t *dTsc = d + 24 + sizeof(cf) * 6 * kMaxSps;
  hres.len = n; hres.tsc = (uint8_t)tsc;
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), burst, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(dLen, &hres.len, 4, cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(dTsc, &hres.tsc, 1, cudaMemcpyHostToDevice, ctx->st));
  NormalOut o = {dFlag, dAmp, dToa, dChan, dOff, nullptr, nullptr, nullptr, 0};
  launch_analyze(ctx->T, make_src((const btsdsp_cf32 *)dbuf<cf>(ctx, B_A), n, dLen, 0, sps), dTsc, 1, threshold,
                 request_channel, o, dbuf<cf>(ctx, B_SCRATCH), n > kBurstRows, ctx->st);
  LAUNCHED("analyze", 1);
  CK(cudaMemcpyAsync(&hres, d, 24 + sizeof(cf) * 6 * kMaxSps, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  *detected = hres.flag;
  amplitude->re = hres.amp.x; amplitude->im = hres.amp.y;
  *toa = hres.toa;
  if (request_channel && hres.flag) {
    if (chan) memcpy(chan, hres.chan, sizeof(cf) * 6 * sps);
    if (chan_offset) *chan_offset = hres.off;
  }
  return BTSDSP_OK;
}

int btsdsp_detect_rach_burst(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, float threshold, btsdsp_cf32 *amplitude,
                             float *toa, int *detected) {
  ARG(ctx && burst && amplitude && toa && detected && n > 1);
  const int sps = ctx->sps;
  ARG(n <= 157 * sps);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_D, 1024); GROW(B_SCRATCH, scratch_per_burst(sps) * sizeof(cf));
  uint8_t *d = dbuf<uint8_t>(ctx, B_D);
  int32_t *dLen = (int32_t *)d, *dFlag = dLen + 1;
  cf *dAmp = (cf *)(d + 8);
  float *dToa = (float *)(d + 16);
  struct { int32_t len, flag; cf amp; float toa; } hres;
  hres.len = n;
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), burst, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(dLen, &hres.len, 4, cudaMemcpyHostToDevice, ctx->st));
  NormalOut o = {dFlag, dAmp, dToa, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  launch_rach(ctx->T, make_src((const btsdsp_cf32 *)dbuf<cf>(ctx, B_A), n, dLen, 0, sps), 1, threshold, 0, o,
              dbuf<cf>(ctx, B_SCRATCH), 0, ctx->st);
  LAUNCHED("detect_rach", 1);
  CK(cudaMemcpyAsync(&hres, d, 20, cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  *detected = hres.flag;
  amplitude->re = hres.amp.x; amplitude->im = hres.amp.y;
  *toa = hres.toa;
  return BTSDSP_OK;
}

int btsdsp_design_dfe(btsdsp_ctx *ctx, const btsdsp_cf32 *chan, int nchan, float snr, int nf, btsdsp_cf32 *w,
                      btsdsp_cf32 *b) {
  ARG(ctx && chan && w && b);
  ARG(nf >= 1 && nf <= kDfeMax && nchan >= 1 && nchan <= nf);   // the reference overruns G1 when nchan > Nf
  DeviceGuard g(ctx->device);
  GROW(B_D, 1024);
  cf *d = dbuf<cf>(ctx, B_D);
  cf *dW = d + 32, *dB = d + 64;
  CK(cudaMemcpyAsync(d, chan, nchan * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_design_dfe_generic(d, nchan, snr, nf, dW, dB, ctx->st);
  LAUNCHED("design_dfe", 1);
  CK(cudaMemcpyAsync(w, dW, nf * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  if (nchan > 1) CK(cudaMemcpyAsync(b, dB, (nchan - 1) * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

int btsdsp_equalize_burst(btsdsp_ctx *ctx, btsdsp_cf32 *burst, int n, float toa, const btsdsp_cf32 *w, int nw,
                          const btsdsp_cf32 *b, int nb, float *soft) {
  ARG(ctx && burst && w && b && soft);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "equalizeBurst assumes symbol-rate sampling (sps == 1)");
  ARG(n > 0 && n <= 157 && nw >= 1 && nw <= kDfeMax && nb >= 0 && nb <= kDfeMax);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_B, (n + kDfeMax) * sizeof(cf)); GROW(B_C, n * sizeof(float)); GROW(B_D, 1024);
  cf *d = dbuf<cf>(ctx, B_D);
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), burst, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(d, w, nw * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  if (nb) CK(cudaMemcpyAsync(d + 32, b, nb * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_equalize_generic(ctx->T, dbuf<cf>(ctx, B_A), n, toa, d, nw, d + 32, nb, dbuf<cf>(ctx, B_B),
                          dbuf<float>(ctx, B_C), ctx->st);
  LAUNCHED("equalize", 1);
  CK(cudaMemcpyAsync(burst, dbuf<cf>(ctx, B_A), n * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaMemcpyAsync(soft, dbuf<float>(ctx, B_C), n * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

int btsdsp_demodulate_burst(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, btsdsp_cf32 channel, float toa,
                            float *soft) {
  ARG(ctx && burst && soft && n > 0);
  const int sps = ctx->sps;
  ARG(n <= 157 * sps);
  DeviceGuard g(ctx->device);
  const int ns = (sps > 1) ? n / sps : n;
  GROW(B_A, n * sizeof(cf)); GROW(B_C, (n + 8) * sizeof(float)); GROW(B_D, 1024);
  GROW(B_SCRATCH, scratch_per_burst(sps) * sizeof(cf));
  uint8_t *d = dbuf<uint8_t>(ctx, B_D);
  struct { int32_t len; int32_t pad; cf amp; float toa; } h;
  h.len = n; h.pad = 0; h.amp = mk(channel.re, channel.im); h.toa = toa;
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), burst, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(d, &h, sizeof h, cudaMemcpyHostToDevice, ctx->st));
  launch_demodulate(ctx->T, make_src((const btsdsp_cf32 *)dbuf<cf>(ctx, B_A), n, (const int32_t *)d, 0, sps), 1,
                    (const cf *)(d + 8), (const float *)(d + 16), dbuf<float>(ctx, B_C), ns, dbuf<cf>(ctx, B_SCRATCH),
                    ctx->st);
  LAUNCHED("demodulate", 1);
  CK(cudaMemcpyAsync(soft, dbuf<float>(ctx, B_C), ns * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return ns;
}

int btsdsp_polyphase_resample(btsdsp_ctx *ctx, const btsdsp_cf32 *x, int n, int P, int Q, int lpf, btsdsp_cf32 *out,
                              int cap) {
  ARG(ctx && x && n > 0 && P > 0 && Q > 0 && (lpf == 0 || lpf == 1));
  const int outn = (int)ceil(n * (float)P / (float)Q);     // :1171
  if (cap < outn || !out) return outn;
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_B, outn * sizeof(cf));
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), x, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_resample_generic(dbuf<cf>(ctx, B_A), n, P, Q, lpf ? ctx->T->lpf_tx : ctx->T->lpf_rx, lpf ? kTxTaps : kRxTaps,
                          dbuf<cf>(ctx, B_B), outn, ctx->st);
  LAUNCHED("polyphase_resample", 1);
  CK(cudaMemcpyAsync(out, dbuf<cf>(ctx, B_B), outn * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return outn;
}

/* polyphaseResampleVector :1157 with the caller's own filter (ntaps complex taps; lpf_real: only their real parts are
 * used, the realOnly branch :1194-1200) */
int btsdsp_polyphase_resample_taps(btsdsp_ctx *ctx, const btsdsp_cf32 *x, int n, int P, int Q, const btsdsp_cf32 *lpf,
                                   int ntaps, int lpf_real, btsdsp_cf32 *out, int cap) {
  ARG(ctx && x && lpf && n > 0 && P > 0 && Q > 0 && ntaps > 0);
  const int outn = (int)ceil(n * (float)P / (float)Q);     // :1171
  if (cap < outn || !out) return outn;
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_B, outn * sizeof(cf)); GROW(B_C, ntaps * sizeof(cf));
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), x, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_C), lpf, ntaps * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  launch_resample_taps(dbuf<cf>(ctx, B_A), n, P, Q, dbuf<cf>(ctx, B_C), ntaps, lpf_real, dbuf<cf>(ctx, B_B), outn, ctx->st);
  LAUNCHED("polyphase_resample_taps", 1);
  CK(cudaMemcpyAsync(out, dbuf<cf>(ctx, B_B), outn * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return outn;
}

/* createLPF :1102-1150: filter_len == 651 gives the 651-tap prototype, anything else the 961-entry vector whose first
 * filter_len entries are the 961-tap prototype's (the rest stay zero); scaled by float(gain_dc / sum) with the sum of the
 * copied taps in double.  filter_len > 961 overruns both the vector and the table in the reference: EINVAL here. */
int btsdsp_create_lpf(btsdsp_ctx *ctx, int filter_len, float gain_dc, float *taps, int cap) {
  ARG(ctx && filter_len > 0 && filter_len <= kRxTaps);
  const int n = (filter_len == kTxTaps) ? kTxTaps : kRxTaps;
  if (!taps || cap < n) return n;
  memset(taps, 0, (size_t)n * sizeof(float));
  create_lpf(filter_len == kTxTaps ? LPF651_BITS : LPF961_BITS, filter_len, gain_dc, taps);
  return n;
}

/* addVector :746 (x += y over the shorter length), offsetVector :760, conjugateVector :733, vectorSlicer :507, in place
 * on x, GMSKRotate / GMSKReverseRotate :232-264 (n <= 157*sps); vectorNorm2 :146 returns sum |x|^2 in *result (x unchanged) */
int btsdsp_vector_op(btsdsp_ctx *ctx, int op, btsdsp_cf32 *x, int n, int real_only, const btsdsp_cf32 *y, int ny,
                     btsdsp_cf32 scalar, float *result) {
  ARG(ctx && x && n > 0 && op >= BTSDSP_VOP_ADD && op <= BTSDSP_VOP_REVROTATE);
  ARG((op != BTSDSP_VOP_ROTATE && op != BTSDSP_VOP_REVROTATE) || n <= 157 * ctx->sps);   /* the rotation tables' length, :215 */
  ARG(op != BTSDSP_VOP_ADD || (y && ny > 0));
  ARG(op != BTSDSP_VOP_NORM2 || result);
  DeviceGuard g(ctx->device);
  GROW(B_A, n * sizeof(cf)); GROW(B_D, 256);
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), x, n * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  if (op == BTSDSP_VOP_ADD) {
    GROW(B_B, ny * sizeof(cf));
    CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_B), y, ny * sizeof(cf), cudaMemcpyHostToDevice, ctx->st));
  }
  launch_vector_op(ctx->T, op, dbuf<cf>(ctx, B_A), n, real_only, dbuf<cf>(ctx, B_B), ny, mk(scalar.re, scalar.im), dbuf<float>(ctx, B_D),
                   ctx->st);
  LAUNCHED("vector_op", 1);
  if (op == BTSDSP_VOP_NORM2) CK(cudaMemcpyAsync(result, dbuf<float>(ctx, B_D), sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
  else CK(cudaMemcpyAsync(x, dbuf<cf>(ctx, B_A), n * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

/* the RX datagram's RSSI byte source, (int) floor(20.0*log10(9450.0/|amp|)) (Transceiver.cpp:400), for n magnitudes */
int btsdsp_trx_rssi(btsdsp_ctx *ctx, const float *abs_amp, int n, int32_t *rssi) {
  ARG(ctx && abs_amp && rssi && n > 0);
  DeviceGuard g(ctx->device);
  GROW(B_A, (size_t)n * sizeof(float)); GROW(B_B, (size_t)n * sizeof(int32_t));
  CK(cudaMemcpyAsync(dbuf<float>(ctx, B_A), abs_amp, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx->st));
  launch_rssi(ctx->T, dbuf<float>(ctx, B_A), n, dbuf<int>(ctx, B_B), ctx->st);
  LAUNCHED("trx_rssi", 1);
  CK(cudaMemcpyAsync(rssi, dbuf<int>(ctx, B_B), (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

// ---- layer 2 ---------------------------------------------------------------------------------------
int btsdsp_modulate_dev(btsdsp_ctx *ctx, const uint8_t *bits, int nbits, long long n, int guard, long long first,
                        btsdsp_cf32 *out, long long pitch, void *stream) {
  ARG(ctx && bits && out && nbits > 0 && n >= 0 && pitch >= 0);
  ARG(nbits + (guard < 0 ? 9 : guard) <= 157);
  if (pitch == 0) ARG(nbits == 148 && guard < 0);
  DeviceGuard g(ctx->device);
  launch_modulate(ctx->T, bits, nbits, n, guard, nullptr, first, (cf *)out, pitch, (cudaStream_t)stream);
  LAUNCHED("modulate", n > 0);
  return BTSDSP_OK;
}

int btsdsp_resample_rx_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, int has_history, long long nchunks,
                           btsdsp_cf32 *out, void *stream) {
  ARG(ctx && raw && out && nchunks >= 0);
  DeviceGuard g(ctx->device);
  launch_resample_rx(ctx->T, (const cf *)raw, has_history, nchunks, (cf *)out, (cudaStream_t)stream);
  LAUNCHED("resample_rx", nchunks > 0);
  return BTSDSP_OK;
}

int btsdsp_resample_rx_i16_dev(btsdsp_ctx *ctx, const int16_t *iq, int swap_iq, int has_history, long long nchunks,
                               btsdsp_cf32 *out, void *stream) {
  ARG(ctx && iq && out && nchunks >= 0);
  DeviceGuard g(ctx->device);
  const int r = launch_resample_rx_i16(iq, swap_iq, has_history, nchunks, (cf *)out, (cudaStream_t)stream);
  if (r < 0) return fail(ctx, BTSDSP_EINVAL, "int16 ingest needs 16-byte aligned pointers");
  LAUNCHED("resample_rx_i16", r);
  return BTSDSP_OK;
}

int btsdsp_demod_normal_u8_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                               long long first, const uint8_t *tsc, long long n, float detect_thr, float gate_thr,
                               float snr_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, uint8_t *soft_u8,
                               int soft_pitch_bytes, void *stream) {
  ARG(ctx && bursts && tsc && soft_u8 && n >= 0 && pitch >= 0 && soft_pitch_bytes >= 148 && soft_pitch_bytes % 4 == 0);
  ARG((reinterpret_cast<uintptr_t>(soft_u8) & 3) == 0);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the DFE path assumes symbol-rate sampling (sps == 1)");
  DeviceGuard g(ctx->device);
  NormalOut o = {flag, (cf *)amp, toa, nullptr, nullptr, nullptr, nullptr, nullptr, soft_pitch_bytes};
  o.soft_u8 = soft_u8;
  auto *ss = stream_scratch(ctx, (cudaStream_t)stream);
  GROWBUF(ss->eqp, demod_scratch_bytes(n));
  const int nl = launch_demod_normal(ctx->T, make_src(bursts, pitch, lens, first, 1), tsc, n, detect_thr, gate_thr,
                                     snr_thr, o, ss->eqp.p, (cudaStream_t)stream);
  LAUNCHED("demod_normal_u8", nl);
  return BTSDSP_OK;
}

int btsdsp_resample_tx_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *in, int has_history, long long nchunks, int16_t *out,
                           void *stream) {
  ARG(ctx && in && out && nchunks >= 0);
  ARG((reinterpret_cast<uintptr_t>(out) & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0);   // {I,Q} pairs are stored as 4-byte units
  DeviceGuard g(ctx->device);
  launch_resample_tx(ctx->T, (const cf *)in, has_history, nchunks, out, (cudaStream_t)stream);
  LAUNCHED("resample_tx", nchunks > 0);
  return BTSDSP_OK;
}

int btsdsp_demod_normal_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                            long long first, const uint8_t *tsc, long long n, float detect_thr, float gate_thr,
                            float snr_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch,
                            btsdsp_cf32 *chan, float *chan_off, btsdsp_cf32 *w, btsdsp_cf32 *b, void *stream) {
  ARG(ctx && bursts && tsc && n >= 0 && pitch >= 0);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the DFE path assumes symbol-rate sampling (sps == 1)");
  ARG(!soft || soft_pitch >= 148);
  DeviceGuard g(ctx->device);
  NormalOut o = {flag, (cf *)amp, toa, (cf *)chan, chan_off, (cf *)w, (cf *)b, soft, soft_pitch};
  cudaStream_t st = (cudaStream_t)stream;
  auto *ss = stream_scratch(ctx, st);
  GROWBUF(ss->eqp, demod_scratch_bytes(n));
  if (ctx->timing) CK(cudaEventRecord(ctx->tev[0], st));
  const int nl = launch_demod_normal(ctx->T, make_src(bursts, pitch, lens, first, 1), tsc, n, detect_thr, gate_thr,
                                     snr_thr, o, ss->eqp.p, st, ctx->timing ? ctx->tev[1] : nullptr);
  if (ctx->timing) { CK(cudaEventRecord(ctx->tev[2], st)); ctx->tev_used = 3; }
  LAUNCHED("demod_normal", nl);
  return BTSDSP_OK;
}

int btsdsp_analyze_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                       long long first, const uint8_t *tsc, long long n, float detect_thr, int request_channel,
                       int32_t *flag, btsdsp_cf32 *amp, float *toa, btsdsp_cf32 *chan, float *chan_off, void *stream) {
  ARG(ctx && bursts && tsc && n >= 0 && pitch >= 0);
  DeviceGuard g(ctx->device);
  auto *ss = stream_scratch(ctx, (cudaStream_t)stream);
  if (ctx->sps != 1) GROWBUF(ss->scratch, (size_t)n * scratch_per_burst(ctx->sps) * sizeof(cf));
  NormalOut o = {flag, (cf *)amp, toa, (cf *)chan, chan_off, nullptr, nullptr, nullptr, 0};
  launch_analyze(ctx->T, make_src(bursts, pitch, lens, first, ctx->sps), tsc, n, detect_thr, request_channel, o,
                 (cf *)ss->scratch.p, 0, (cudaStream_t)stream);
  LAUNCHED("analyze", n > 0);
  return BTSDSP_OK;
}

int btsdsp_rach_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens, long long first,
                    long long n, float detect_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft,
                    int soft_pitch, void *stream) {
  ARG(ctx && bursts && n >= 0 && pitch >= 0);
  ARG(!soft || soft_pitch >= 157);
  DeviceGuard g(ctx->device);
  auto *ss = stream_scratch(ctx, (cudaStream_t)stream);
  if (ctx->sps != 1) GROWBUF(ss->scratch, (size_t)n * scratch_per_burst(ctx->sps) * sizeof(cf));
  NormalOut o = {flag, (cf *)amp, toa, nullptr, nullptr, nullptr, nullptr, soft, soft_pitch};
  GROWBUF(ss->eqp, rach_scratch_bytes(n));
  const int nl = launch_rach(ctx->T, make_src(bursts, pitch, lens, first, ctx->sps), n, detect_thr, soft != nullptr, o,
                             (cf *)ss->scratch.p, 0, (cudaStream_t)stream, ss->eqp.p);
  LAUNCHED("rach", nl);
  return BTSDSP_OK;
}

int btsdsp_design_dfe_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *chan, const float *snr, long long n, btsdsp_cf32 *w,
                          btsdsp_cf32 *b, void *stream) {
  ARG(ctx && chan && snr && w && b && n >= 0);
  DeviceGuard g(ctx->device);
  launch_design_dfe((const cf *)chan, snr, n, (cf *)w, (cf *)b, (cudaStream_t)stream);
  LAUNCHED("design_dfe", n > 0);
  return BTSDSP_OK;
}

int btsdsp_equalize_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                        long long first, long long n, const float *toa, const btsdsp_cf32 *w, const btsdsp_cf32 *b,
                        float *soft, int soft_pitch, btsdsp_cf32 *burst_out, long long out_pitch, void *stream) {
  ARG(ctx && bursts && toa && w && b && soft && n >= 0 && pitch >= 0 && soft_pitch >= 148);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "equalizeBurst assumes symbol-rate sampling (sps == 1)");
  DeviceGuard g(ctx->device);
  launch_equalize(ctx->T, make_src(bursts, pitch, lens, first, 1), n, toa, (const cf *)w, (const cf *)b, soft,
                  soft_pitch, (cf *)burst_out, out_pitch, (cudaStream_t)stream);
  LAUNCHED("equalize", n > 0);
  return BTSDSP_OK;
}

int btsdsp_demodulate_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                          long long first, long long n, const btsdsp_cf32 *amp, const float *toa, float *soft,
                          int soft_pitch, void *stream) {
  ARG(ctx && bursts && amp && toa && soft && n >= 0 && pitch >= 0 && soft_pitch >= 157);
  DeviceGuard g(ctx->device);
  auto *ss = stream_scratch(ctx, (cudaStream_t)stream);
  GROWBUF(ss->scratch, (size_t)n * scratch_per_burst(ctx->sps) * sizeof(cf));
  launch_demodulate(ctx->T, make_src(bursts, pitch, lens, first, ctx->sps), n, (const cf *)amp, toa, soft, soft_pitch,
                    (cf *)ss->scratch.p, (cudaStream_t)stream);
  LAUNCHED("demodulate", n > 0);
  return BTSDSP_OK;
}

// ---- layer 3 ---------------------------------------------------------------------------------------
static int rx_stream_dev_impl(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, int has_history, long long nchunks, const uint8_t *tsc,
                              long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                              btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch, void *stream) {
  ARG(ctx && raw && tsc && nchunks > 0 && nbursts >= 0);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the RX stream path runs at sps == 1");
  ARG(((nbursts + 3) / 4) * 625 <= nchunks * 585);
  ARG(!soft || soft_pitch >= 148);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  auto *ss = stream_scratch(ctx, st);
  GROWBUF(ss->res, (size_t)nchunks * 585 * sizeof(cf));
  GROWBUF(ss->eqp, demod_scratch_bytes(nbursts + 1));
  cf *dRes = (cf *)ss->res.p;
  const long long seg = ctx->rx_seg;
  if (seg <= 0 || nchunks <= seg) {
    // one resampler launch, then the two demod kernels, all on the caller's stream
    launch_resample_rx(ctx->T, (const cf *)raw, has_history, nchunks, dRes, st);
    NormalOut o = {flag, (cf *)amp, toa, nullptr, nullptr, nullptr, nullptr, soft, soft_pitch};
    const int nl = launch_demod_normal(ctx->T, make_src((const btsdsp_cf32 *)dRes, 0, nullptr, 0, 1), tsc, nbursts, detect_thr,
                                       gate_thr, snr_thr, o, ss->eqp.p, st);
    LAUNCHED("rx_stream", 1 + nl);
    return BTSDSP_OK;
  }
  // Segmented: the HBM-bound resampler of segment s+1 runs on a side stream, on a capped number of SMs, while the
  // FP32-bound demod kernels of segment s run on the caller's stream (complementary resources).
  // The whole raw stream is resident, so a segment's 192-sample history is simply the samples before it.
  const long long nseg = (nchunks + seg - 1) / seg;
  while ((long long)ctx->events.size() < nseg + 2) {
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->events.push_back(ev);
  }
  cudaStream_t side = ctx->st_side;
  CK(cudaEventRecord(ctx->events[nseg], st));                   // the inputs are ready where the caller's stream stands
  CK(cudaStreamWaitEvent(side, ctx->events[nseg], 0));
  long long done_bursts = 0;
  int nl = 0;
  // resampler launches run one segment AHEAD of the demod kernels (R(0), R(1), D(0), R(2), D(1), ...): R(s+1) is queued
  // before D(s), so its CTAs (one per SM, the light configuration) are resident when D(s)'s CTAs fill the rest of the SMs
  auto resample_seg = [&](long long s) {
    const long long c0 = s * seg, c1 = (c0 + seg < nchunks) ? c0 + seg : nchunks;
    launch_resample_rx(ctx->T, (const cf *)raw + c0 * 864, has_history || c0 > 0, c1 - c0, dRes + c0 * 585, side,
                       s == 0 ? 0 : ctx->rx_res_ctas, s == 0 ? 0 : ctx->rx_light);
    nl++;
    cudaEventRecord(ctx->events[s], side);
  };
  resample_seg(0);
  for (long long s = 0; s < nseg; s++) {
    const long long c0 = s * seg, c1 = (c0 + seg < nchunks) ? c0 + seg : nchunks;
    if (s + 1 < nseg) resample_seg(s + 1);
    CK(cudaStreamWaitEvent(st, ctx->events[s], 0));
    long long avail = (c1 * 585 / 625) * 4;                     // bursts wholly inside the samples resampled so far
    if (avail > nbursts || c1 == nchunks) avail = nbursts;
    const long long nb = avail - done_bursts;
    if (nb > 0) {
      NormalOut o = {flag ? flag + done_bursts : nullptr, amp ? (cf *)amp + done_bursts : nullptr,
                     toa ? toa + done_bursts : nullptr, nullptr, nullptr, nullptr, nullptr,
                     soft ? soft + done_bursts * (long long)soft_pitch : nullptr, soft_pitch};
      nl += launch_demod_normal(ctx->T, make_src((const btsdsp_cf32 *)dRes, 0, nullptr, done_bursts, 1), tsc + done_bursts, nb,
                                detect_thr, gate_thr, snr_thr, o, (unsigned char *)ss->eqp.p + demod_scratch_bytes(done_bursts), st);
    }
    done_bursts = avail;
  }
  LAUNCHED("rx_stream (segmented)", nl);
  return BTSDSP_OK;
}

int btsdsp_rx_stream_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, long long nchunks, const uint8_t *tsc,
                         long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                         btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch, void *stream) {
  return rx_stream_dev_impl(ctx, raw, 0, nchunks, tsc, nbursts, detect_thr, gate_thr, snr_thr, flag, amp, toa, soft, soft_pitch, stream);
}
/* a later piece of a running stream: raw[-192..-1] hold the previous samples (RadioInterface::pullBuffer keeps them in
 * rcvHistory, radioInterface.cpp:238-259); the piece must begin on a 117-frame boundary of the stream */
int btsdsp_rx_stream_cont_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, int has_history, long long nchunks, const uint8_t *tsc,
                              long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                              btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch, void *stream) {
  return rx_stream_dev_impl(ctx, raw, has_history != 0, nchunks, tsc, nbursts, detect_thr, gate_thr, snr_thr, flag, amp, toa, soft,
                            soft_pitch, stream);
}

// Host-buffer receive pipeline: the stream is cut into segments of whole frame groups; segment s is
// copied H2D on st_in while segment s-1 is resampled + demodulated on st and segment s-2's results
// return D2H on st_out.  The whole raw / resampled stream stays resident on the device so the
// resampler's 192-sample history and bursts straddling segment borders need no special casing.
// `raw` holds complex64 samples (i16 == 0) or int16 {I,Q} pairs (i16 != 0); `soft` receives floats (u8 == 0, pitch
// in floats) or the datagram bytes (u8 != 0, pitch in bytes).
static int rx_stream_host_impl(btsdsp_ctx *ctx, const void *raw_, int i16, int swap_iq, long long nchunks,
                               const uint8_t *tsc, long long nbursts, float detect_thr, float gate_thr, float snr_thr,
                               int32_t *flag, btsdsp_cf32 *amp, float *toa, void *soft_, int u8, int soft_pitch) {
  ARG(ctx && raw_ && tsc && nchunks > 0 && nbursts >= 0);
  const size_t in_sz = i16 ? 4 : 8, soft_sz = u8 ? 1 : 4;
  const unsigned char *raw = (const unsigned char *)raw_;
  unsigned char *soft = (unsigned char *)soft_;
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the RX stream path runs at sps == 1");
  ARG(((nbursts + 3) / 4) * 625 <= nchunks * 585);
  ARG(!soft || (soft_pitch >= 148 && (!u8 || soft_pitch % 4 == 0)));
  DeviceGuard g(ctx->device);
  GROW(B_RAW, (size_t)nchunks * 864 * in_sz);
  GROW(B_RES, (size_t)nchunks * 585 * sizeof(cf));
  GROW(B_TSC, (size_t)nbursts + 16);
  GROW(B_FLAG, (size_t)(nbursts + 1) * sizeof(int32_t));
  GROW(B_AMP, (size_t)(nbursts + 1) * sizeof(cf));
  GROW(B_TOA, (size_t)(nbursts + 1) * sizeof(float));
  if (soft) GROW(B_SOFT, (size_t)(nbursts + 1) * soft_pitch * soft_sz);
  GROW(B_EQP, demod_scratch_bytes(nbursts + 1));
  unsigned char *dRaw = dbuf<unsigned char>(ctx, B_RAW);
  cf *dRes = dbuf<cf>(ctx, B_RES);
  uint8_t *dTsc = dbuf<uint8_t>(ctx, B_TSC);
  int32_t *dFlag = dbuf<int32_t>(ctx, B_FLAG);
  cf *dAmp = dbuf<cf>(ctx, B_AMP);
  float *dToa = dbuf<float>(ctx, B_TOA);
  unsigned char *dSoft = soft ? dbuf<unsigned char>(ctx, B_SOFT) : nullptr;

  const long long seg = ctx->host_seg;              // chunks per segment (default 4000: 27.6 MB of complex64 samples)
  const long long nseg = (nchunks + seg - 1) / seg;
  while ((long long)ctx->events.size() < 2 * nseg + 2) {
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->events.push_back(ev);
  }
  CK(cudaMemcpyAsync(dTsc, tsc, (size_t)nbursts, cudaMemcpyHostToDevice, ctx->st_in));
  long long done_bursts = 0;
  for (long long s = 0; s < nseg; s++) {
    const long long c0 = s * seg, c1 = (c0 + seg < nchunks) ? c0 + seg : nchunks;
    cudaEvent_t evIn = ctx->events[2 * s], evK = ctx->events[2 * s + 1];
    CK(cudaMemcpyAsync(dRaw + c0 * 864 * in_sz, raw + c0 * 864 * in_sz, (size_t)(c1 - c0) * 864 * in_sz,
                       cudaMemcpyHostToDevice, ctx->st_in));
    CK(cudaEventRecord(evIn, ctx->st_in));
    CK(cudaStreamWaitEvent(ctx->st, evIn, 0));
    if (ctx->copy_only) {}
    else if (i16) launch_resample_rx_i16((const int16_t *)(dRaw + c0 * 864 * in_sz), swap_iq, c0 > 0, c1 - c0, dRes + c0 * 585, ctx->st);
    else launch_resample_rx(ctx->T, (const cf *)(dRaw + c0 * 864 * in_sz), c0 > 0, c1 - c0, dRes + c0 * 585, ctx->st);
    // bursts wholly inside the samples resampled so far: groups of 4 slots = 625 samples
    long long avail = (c1 * 585 / 625) * 4;
    if (avail > nbursts || c1 == nchunks) avail = nbursts;
    const long long nb = avail - done_bursts;
    int nl = ctx->copy_only ? 0 : 1;
    if (nb > 0 && !ctx->copy_only) {
      NormalOut o = {dFlag + done_bursts, dAmp + done_bursts, dToa + done_bursts, nullptr, nullptr, nullptr, nullptr,
                     nullptr, soft_pitch};
      if (dSoft && u8) o.soft_u8 = dSoft + done_bursts * soft_pitch;
      else if (dSoft) o.soft = (float *)(dSoft + done_bursts * soft_pitch * soft_sz);
      nl = 1 + launch_demod_normal(ctx->T, make_src((const btsdsp_cf32 *)dRes, 0, nullptr, done_bursts, 1),
                                   dTsc + done_bursts, nb, detect_thr, gate_thr, snr_thr, o,
                                   dbuf<unsigned char>(ctx, B_EQP) + demod_scratch_bytes(done_bursts), ctx->st);
    }
    LAUNCHED("rx_stream_host", nl);
    CK(cudaEventRecord(evK, ctx->st));
    if (nb > 0) {
      CK(cudaStreamWaitEvent(ctx->st_out, evK, 0));
      if (flag) CK(cudaMemcpyAsync(flag + done_bursts, dFlag + done_bursts, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->st_out));
      if (amp) CK(cudaMemcpyAsync(amp + done_bursts, dAmp + done_bursts, nb * sizeof(cf), cudaMemcpyDeviceToHost, ctx->st_out));
      if (toa) CK(cudaMemcpyAsync(toa + done_bursts, dToa + done_bursts, nb * sizeof(float), cudaMemcpyDeviceToHost, ctx->st_out));
      if (soft) CK(cudaMemcpyAsync(soft + done_bursts * soft_pitch * soft_sz, dSoft + done_bursts * soft_pitch * soft_sz,
                                   (size_t)nb * soft_pitch * soft_sz, cudaMemcpyDeviceToHost, ctx->st_out));
    }
    done_bursts = avail;
  }
  CK(cudaStreamSynchronize(ctx->st_out));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

int btsdsp_rx_stream_host(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, long long nchunks, const uint8_t *tsc,
                          long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                          btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch) {
  return rx_stream_host_impl(ctx, raw, 0, 0, nchunks, tsc, nbursts, detect_thr, gate_thr, snr_thr, flag, amp, toa, soft, 0,
                             soft_pitch);
}

int btsdsp_rx_stream_wire_host(btsdsp_ctx *ctx, const int16_t *iq, int swap_iq, long long nchunks, const uint8_t *tsc,
                               long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                               btsdsp_cf32 *amp, float *toa, uint8_t *soft_u8) {
  return rx_stream_host_impl(ctx, iq, 1, swap_iq, nchunks, tsc, nbursts, detect_thr, gate_thr, snr_thr, flag, amp, toa,
                             soft_u8, 1, 148);
}

int btsdsp_tx_stream_dev(btsdsp_ctx *ctx, const uint8_t *bits148, long long n, int16_t *out, void *stream) {
  ARG(ctx && bits148 && out && n > 0);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the TX stream path runs at sps == 1");
  ARG(n % 4 == 0 && (n / 4 * 625) % 585 == 0);
  DeviceGuard g(ctx->device);
  // one fused kernel: the modulated stream never leaves shared memory (resample.cu: k_tx_fused)
  launch_tx_fused(ctx->T, bits148, nullptr, n, out, (cudaStream_t)stream);
  LAUNCHED("tx_stream", 1);
  return BTSDSP_OK;
}

int btsdsp_tx_streams_dev(btsdsp_ctx *ctx, const uint8_t *bits148, long long n, int nstreams, int16_t *out, void *stream) {
  ARG(ctx && bits148 && out && n > 0 && nstreams > 0);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the TX stream path runs at sps == 1");
  ARG(n % 4 == 0 && (n / 4 * 625) % 585 == 0);
  DeviceGuard g(ctx->device);
  launch_tx_fused(ctx->T, bits148, nullptr, n, out, (cudaStream_t)stream, nstreams);
  LAUNCHED("tx_streams", 1);
  return BTSDSP_OK;
}

int btsdsp_tx_stream_host(btsdsp_ctx *ctx, const uint8_t *bits148, long long n, int16_t *out) {
  ARG(ctx && bits148 && out && n > 0);
  ARG(n % 4 == 0 && (n / 4 * 625) % 585 == 0);
  DeviceGuard g(ctx->device);
  const long long nchunks = n / 4 * 625 / 585;
  GROW(B_TSC, (size_t)n * 148);
  GROW(B_RAW, (size_t)nchunks * 864 * 2 * sizeof(int16_t));
  CK(cudaMemcpyAsync(dbuf<uint8_t>(ctx, B_TSC), bits148, (size_t)n * 148, cudaMemcpyHostToDevice, ctx->st));
  int r = btsdsp_tx_stream_dev(ctx, dbuf<uint8_t>(ctx, B_TSC), n, dbuf<int16_t>(ctx, B_RAW), ctx->st);
  if (r != BTSDSP_OK) return r;
  CK(cudaMemcpyAsync(out, dbuf<int16_t>(ctx, B_RAW), (size_t)nchunks * 864 * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost,
                     ctx->st));
  CK(cudaStreamSynchronize(ctx->st));
  return BTSDSP_OK;
}

int btsdsp_demod_normal_host(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                             const uint8_t *tsc, long long n, float detect_thr, float gate_thr, float snr_thr,
                             int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch,
                             btsdsp_cf32 *chan, float *chan_off, btsdsp_cf32 *w, btsdsp_cf32 *b) {
  ARG(ctx && bursts && tsc && n > 0 && pitch >= 157);
  ARG(!soft || soft_pitch >= 148);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the DFE path assumes symbol-rate sampling (sps == 1)");
  DeviceGuard g(ctx->device);
  // one flat device arena: bursts | lens | tsc | flag | amp | toa | off | chan | w | b | soft
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t o_b = take((size_t)n * pitch * sizeof(cf)), o_len = take((size_t)n * 4), o_tsc = take((size_t)n),
               o_flag = take((size_t)n * 4), o_amp = take((size_t)n * 8), o_toa = take((size_t)n * 4),
               o_off = take((size_t)n * 4), o_chan = take((size_t)n * 48), o_w = take((size_t)n * 56),
               o_fb = take((size_t)n * 40), o_soft = take((size_t)n * soft_pitch * 4);
  GROW(B_RAW, total);
  uint8_t *d = dbuf<uint8_t>(ctx, B_RAW);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(d + o_b, bursts, (size_t)n * pitch * sizeof(cf), cudaMemcpyHostToDevice, st));
  if (lens) CK(cudaMemcpyAsync(d + o_len, lens, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_tsc, tsc, (size_t)n, cudaMemcpyHostToDevice, st));
  int r = btsdsp_demod_normal_dev(ctx, (const btsdsp_cf32 *)(d + o_b), pitch, lens ? (const int32_t *)(d + o_len) : nullptr,
                                  0, d + o_tsc, n, detect_thr, gate_thr, snr_thr, (int32_t *)(d + o_flag),
                                  (btsdsp_cf32 *)(d + o_amp), (float *)(d + o_toa), soft ? (float *)(d + o_soft) : nullptr,
                                  soft_pitch, chan ? (btsdsp_cf32 *)(d + o_chan) : nullptr,
                                  chan_off ? (float *)(d + o_off) : nullptr, w ? (btsdsp_cf32 *)(d + o_w) : nullptr,
                                  b ? (btsdsp_cf32 *)(d + o_fb) : nullptr, st);
  if (r != BTSDSP_OK) return r;
  if (flag) CK(cudaMemcpyAsync(flag, d + o_flag, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (amp) CK(cudaMemcpyAsync(amp, d + o_amp, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (toa) CK(cudaMemcpyAsync(toa, d + o_toa, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (chan_off) CK(cudaMemcpyAsync(chan_off, d + o_off, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (chan) CK(cudaMemcpyAsync(chan, d + o_chan, (size_t)n * 48, cudaMemcpyDeviceToHost, st));
  if (w) CK(cudaMemcpyAsync(w, d + o_w, (size_t)n * 56, cudaMemcpyDeviceToHost, st));
  if (b) CK(cudaMemcpyAsync(b, d + o_fb, (size_t)n * 40, cudaMemcpyDeviceToHost, st));
  if (soft) CK(cudaMemcpyAsync(soft, d + o_soft, (size_t)n * soft_pitch * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

int btsdsp_rach_host(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens, long long n,
                     float detect_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch) {
  ARG(ctx && bursts && n > 0 && pitch >= 157 * ctx->sps);
  ARG(!soft || soft_pitch >= 157);
  DeviceGuard g(ctx->device);
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t o_b = take((size_t)n * pitch * sizeof(cf)), o_len = take((size_t)n * 4), o_flag = take((size_t)n * 4),
               o_amp = take((size_t)n * 8), o_toa = take((size_t)n * 4), o_soft = take((size_t)n * soft_pitch * 4);
  GROW(B_RAW, total);
  uint8_t *d = dbuf<uint8_t>(ctx, B_RAW);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(d + o_b, bursts, (size_t)n * pitch * sizeof(cf), cudaMemcpyHostToDevice, st));
  if (lens) CK(cudaMemcpyAsync(d + o_len, lens, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  int r = btsdsp_rach_dev(ctx, (const btsdsp_cf32 *)(d + o_b), pitch, lens ? (const int32_t *)(d + o_len) : nullptr, 0, n,
                          detect_thr, (int32_t *)(d + o_flag), (btsdsp_cf32 *)(d + o_amp), (float *)(d + o_toa),
                          soft ? (float *)(d + o_soft) : nullptr, soft_pitch, st);
  if (r != BTSDSP_OK) return r;
  if (flag) CK(cudaMemcpyAsync(flag, d + o_flag, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (amp) CK(cudaMemcpyAsync(amp, d + o_amp, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (toa) CK(cudaMemcpyAsync(toa, d + o_toa, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (soft) CK(cudaMemcpyAsync(soft, d + o_soft, (size_t)n * soft_pitch * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

/* ---- caller policy: Transceiver::pullRadioVector + driveReceiveFIFO over batches (trx_policy.cuh) ---- */
struct btsdsp_trx {
  int narfcn = 0;
  TrxState *d_state = nullptr;
  std::vector<uint8_t> tsc;             // [narfcn]
  std::vector<uint8_t> chan_type;       // [narfcn][8]
  DevBuf meta, scratch, io;             // device: per-burst kind/tsc/rach maps; pass scratch; host-API staging
  DevBuf pin;                           // pinned host staging of the maps
  DevBuf radio, res;                    // radio chain: per-ARFCN int16 staging (history + chunks), resampled streams
  bool have_history = false;            // the radio chain has seen earlier chunks (their last 192 samples are staged)
  cudaEvent_t meta_done = nullptr;      // the previous call's map upload has been consumed
  cudaStream_t side = nullptr;          // access-burst kernels run here, next to the normal-burst kernels
  cudaEvent_t fork_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  TrxVariant var;                       // main transceiver, or the second one (btsdsp_trx_set_variant_52m)
  // The slot map of a pull (per burst: correlator type, TSC, access-burst numbering) depends on FN only through FN mod 102
  // (expectedCorrType looks at FN % 51 and FN % 2; the hyperframe is a multiple of 102), so a receiver that pulls batch
  // after batch sees the same few maps again and again: they are kept on the device, keyed by (fn0 % 102, nframes).
  struct SlotMap { void *dev = nullptr; long long nr = 0; };
  std::unordered_map<long long, SlotMap> maps;
  size_t map_bytes = 0;
};
static void trx_drop_maps(btsdsp_trx *t) {
  for (auto &kv : t->maps) if (kv.second.dev) cudaFree(kv.second.dev);
  t->maps.clear();
  t->map_bytes = 0;
}

int btsdsp_trx_create(btsdsp_ctx *ctx, int narfcn, const uint8_t *tsc, const uint8_t *chan_type, int start_fn,
                      btsdsp_trx **out) {
  ARG(ctx && out && narfcn > 0 && tsc && chan_type && start_fn >= 0 && start_fn < kHyperframe);
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the receive policy assumes symbol-rate sampling (sps == 1)");
  for (int a = 0; a < narfcn; a++) {
    ARG(tsc[a] < 8);
    for (int tn = 0; tn < 8; tn++) ARG(chan_type[a * 8 + tn] <= CT_LOOPBACK);
  }
  DeviceGuard g(ctx->device);
  btsdsp_trx *t = new btsdsp_trx;
  t->narfcn = narfcn;
  t->tsc.assign(tsc, tsc + narfcn);
  t->chan_type.assign(chan_type, chan_type + (size_t)narfcn * 8);
  std::vector<TrxState> h(narfcn);
  memset(h.data(), 0, h.size() * sizeof(TrxState));
  for (int a = 0; a < narfcn; a++) {                       // Transceiver::Transceiver, Transceiver.cpp:72-89
    h[a].thr = 250.0;
    h[a].prev_false_fn = start_fn;
    h[a].tsc = tsc[a];
    for (int tn = 0; tn < 8; tn++) { h[a].chan_type[tn] = chan_type[a * 8 + tn]; h[a].est_fn[tn] = start_fn; }
  }
  cudaError_t e = cudaMalloc(&t->d_state, h.size() * sizeof(TrxState));
  if (e == cudaSuccess) e = cudaMemcpy(t->d_state, h.data(), h.size() * sizeof(TrxState), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->meta_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking);
  for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&t->fork_ev[k], cudaEventDisableTiming);
  if (e != cudaSuccess) { if (t->d_state) cudaFree(t->d_state); delete t; return fail(ctx, BTSDSP_ECUDA, "trx_create", e); }
  *out = t;
  return BTSDSP_OK;
}

int btsdsp_trx_destroy(btsdsp_ctx *ctx, btsdsp_trx *t) {
  ARG(ctx && t);
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  trx_drop_maps(t);
  if (t->d_state) cudaFree(t->d_state);
  if (t->meta.p) cudaFree(t->meta.p);
  if (t->scratch.p) cudaFree(t->scratch.p);
  if (t->io.p) cudaFree(t->io.p);
  if (t->radio.p) cudaFree(t->radio.p);
  if (t->res.p) cudaFree(t->res.p);
  if (t->pin.p) cudaFreeHost(t->pin.p);
  if (t->meta_done) cudaEventDestroy(t->meta_done);
  if (t->side) cudaStreamDestroy(t->side);
  for (int k = 0; k < 4; k++) if (t->fork_ev[k]) cudaEventDestroy(t->fork_ev[k]);
  delete t;
  return BTSDSP_OK;
}

int btsdsp_trx_state_bytes(void) { return (int)sizeof(TrxState); }

int btsdsp_trx_get_state(btsdsp_ctx *ctx, btsdsp_trx *t, int arfcn, void *dst, int cap) {
  ARG(ctx && t && dst && arfcn >= 0 && arfcn < t->narfcn && cap >= (int)sizeof(TrxState));
  DeviceGuard g(ctx->device);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dst, t->d_state + arfcn, sizeof(TrxState), cudaMemcpyDeviceToHost));
  return BTSDSP_OK;
}

int btsdsp_trx_set_slot(btsdsp_ctx *ctx, btsdsp_trx *t, int arfcn, int tn, int chan_type) {   /* SETSLOT, Transceiver.cpp:549 */
  ARG(ctx && t && arfcn >= 0 && arfcn < t->narfcn && tn >= 0 && tn < 8 && chan_type >= 0 && chan_type <= CT_LOOPBACK);
  DeviceGuard g(ctx->device);
  CK(cudaDeviceSynchronize());
  trx_drop_maps(t);                                           // the cached slot maps were built from the old combination
  t->chan_type[(size_t)arfcn * 8 + tn] = (uint8_t)chan_type;
  CK(cudaMemcpy(&t->d_state[arfcn].chan_type[tn], &chan_type, sizeof(int), cudaMemcpyHostToDevice));
  return BTSDSP_OK;
}

/* the second transceiver variant's receive policy (Transceiver52M/Transceiver.cpp:268-404); max_expected_delay is its
 * mMaxExpectedDelay (the SETMAXDLY control command) */
int btsdsp_trx_set_variant_52m(btsdsp_ctx *ctx, btsdsp_trx *t, int enable, int max_expected_delay) {
  ARG(ctx && t && max_expected_delay >= 0 && max_expected_delay <= 60);
  t->var.v52m = enable != 0;
  t->var.max_toa = (unsigned)max_expected_delay;
  t->var.need_dfe = !t->var.v52m || max_expected_delay > 1;                                 // :272
  return BTSDSP_OK;
}

static int trx_pull_impl(btsdsp_ctx *ctx, btsdsp_trx *t, const btsdsp_cf32 *bursts, long long pitch, long long stream_pitch,
                         int nframes, int fn0, int32_t *valid, uint8_t *dgram, int dgram_pitch, void *stream) {
  ARG(ctx && t && bursts && valid && dgram && nframes > 0 && fn0 >= 0);
  ARG(pitch >= 157 || (pitch == 0 && stream_pitch >= (long long)nframes * 1250));
  fn0 %= kHyperframe;
  ARG(dgram_pitch >= 160 && dgram_pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(dgram) & 7) == 0);
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int A = t->narfcn;
  const long long n = (long long)nframes * A * 8;
  ARG(n < (1LL << 31));
  // the slot map is a pure function of (channel combination, FN): evaluate it here, upload one byte per burst
  auto grow_buf = [&](DevBuf &b, size_t bytes, bool pinned) { return grow(ctx, b, bytes, pinned); };
  const size_t o_kind = 0, o_tsc = (size_t)((n + 255) & ~255LL), o_slot = 2 * o_tsc, o_idx = o_slot + (size_t)n * 4;
  const size_t meta_bytes = o_idx + (size_t)n * 4 + 256;
  const long long key = (long long)(fn0 % 102) * (1LL << 32) + nframes;
  auto hit = t->maps.find(key);
  uint8_t *dm = nullptr;
  long long nr = 0;
  int r = BTSDSP_OK;
  if (hit != t->maps.end()) {
    dm = (uint8_t *)hit->second.dev;
    nr = hit->second.nr;
    CK(cudaStreamWaitEvent(st, t->meta_done, 0));          // the latest map upload (possibly issued on another stream) has landed
  } else {
    CK(cudaEventSynchronize(t->meta_done));                // the previous upload has left the pinned staging buffer:
    r = grow_buf(t->pin, meta_bytes, true);                // only now may it be reused -- or freed by a grow
    if (r != BTSDSP_OK) return r;
    uint8_t *hp = (uint8_t *)t->pin.p;
    uint8_t *kind = hp + o_kind, *tscb = hp + o_tsc;
    int *slot = (int *)(hp + o_slot), *idx = (int *)(hp + o_idx);
    for (int f = 0; f < nframes; f++) {
      const int fn = (int)(((long long)fn0 + f) % kHyperframe);
      for (int a = 0; a < A; a++) {
        const long long i0 = ((long long)f * A + a) * 8;
        for (int tn = 0; tn < 8; tn++) {
          const int c = expected_corr_type(t->chan_type[(size_t)a * 8 + tn], fn);
          kind[i0 + tn] = (uint8_t)c;
          tscb[i0 + tn] = t->tsc[a];
          if (c == CORR_RACH) { slot[i0 + tn] = (int)nr; idx[nr++] = (int)(i0 + tn); } else slot[i0 + tn] = -1;
        }
      }
    }
    if (t->map_bytes + meta_bytes > (size_t)512 << 20) trx_drop_maps(t);     // a caller that never repeats a phase: start over
    void *dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, meta_bytes);
    if (e != cudaSuccess) return fail(ctx, BTSDSP_ENOMEM, "cudaMalloc (slot map)", e);
    CK(cudaMemcpyAsync(dev, hp, meta_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(t->meta_done, st));
    btsdsp_trx::SlotMap m;
    m.dev = dev; m.nr = nr;
    t->maps[key] = m;
    t->map_bytes += meta_bytes;
    dm = (uint8_t *)dev;
  }
  r = grow_buf(t->scratch, trx_scratch_bytes(n, nr, A, t->var.v52m && !t->var.need_dfe), false);
  if (r != BTSDSP_OK) return r;
  const int nl = launch_trx_pull(ctx->T, t->d_state, A, nframes, fn0, (const cf *)bursts, pitch, stream_pitch, dm + o_kind,
                                 dm + o_tsc, (const int *)(dm + o_idx), (const int *)(dm + o_slot), nr, t->scratch.p, valid,
                                 dgram, dgram_pitch, st, t->side, t->fork_ev, t->var);
  LAUNCHED("trx_pull", nl);
  return BTSDSP_OK;
}

int btsdsp_trx_pull_dev(btsdsp_ctx *ctx, btsdsp_trx *t, const btsdsp_cf32 *bursts, long long pitch, int nframes, int fn0,
                        int32_t *valid, uint8_t *dgram, int dgram_pitch, void *stream) {
  ARG(pitch >= 157);
  return trx_pull_impl(ctx, t, bursts, pitch, 0, nframes, fn0, valid, dgram, dgram_pitch, stream);
}

int btsdsp_trx_pull_streams_dev(btsdsp_ctx *ctx, btsdsp_trx *t, const btsdsp_cf32 *streams, long long stream_pitch, int nframes,
                                int fn0, int32_t *valid, uint8_t *dgram, int dgram_pitch, void *stream) {
  return trx_pull_impl(ctx, t, streams, 0, stream_pitch, nframes, fn0, valid, dgram, dgram_pitch, stream);
}

/* radio samples in, datagrams out: RadioInterface::pullBuffer (int16 -> float, 65/96 resample with the running 192-sample
 * history, radioInterface.cpp:197-273) + slot cutting (:370-394) + pullRadioVector + driveReceiveFIFO, for narfcn radios */
int btsdsp_trx_radio_host(btsdsp_ctx *ctx, btsdsp_trx *t, const int16_t *iq, long long iq_pitch, long long nchunks, int swap_iq,
                          int fn0, int32_t *valid, uint8_t *dgram, int dgram_pitch) {
  ARG(ctx && t && iq && valid && dgram && nchunks > 0 && nchunks % 250 == 0 && iq_pitch >= nchunks * 864 && dgram_pitch >= 158);
  DeviceGuard g(ctx->device);
  const int A = t->narfcn;
  const int nframes = (int)(nchunks / 250 * 117);
  const long long n = (long long)nframes * A * 8;
  const long long raw_pitch = 192 + nchunks * 864;                 // int16 pairs per ARFCN in the staging buffer
  const long long res_pitch = nchunks * 585;                       // resampled samples per ARFCN (a multiple of 2)
  int r = grow(ctx, t->radio, (size_t)A * raw_pitch * 4, false);
  if (r != BTSDSP_OK) return r;
  r = grow(ctx, t->res, (size_t)A * res_pitch * sizeof(cf), false);
  if (r != BTSDSP_OK) return r;
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t o_v = take((size_t)n * 4), o_d = take((size_t)n * 160);
  r = grow(ctx, t->io, total, false);
  if (r != BTSDSP_OK) return r;
  cudaStream_t st = ctx->st;
  int16_t *draw = (int16_t *)t->radio.p;
  cf *dres = (cf *)t->res.p;
  uint8_t *dio = (uint8_t *)t->io.p;
  // all radios' new samples behind their 192-sample histories, ONE resampler launch over all streams, then the last
  // 192 raw samples of every stream become the next call's history
  CK(cudaMemcpy2DAsync(draw + 192 * 2, (size_t)raw_pitch * 4, iq, (size_t)iq_pitch * 4, (size_t)nchunks * 864 * 4, (size_t)A,
                       cudaMemcpyHostToDevice, st));
  if (launch_resample_rx_i16_multi(draw + 192 * 2, raw_pitch, A, swap_iq, t->have_history ? 1 : 0, nchunks, dres, res_pitch, st) < 0)
    return fail(ctx, BTSDSP_EINVAL, "trx_radio: resampler rejected the staging pointers");
  CK(cudaMemcpy2DAsync(draw, (size_t)raw_pitch * 4, draw + (size_t)nchunks * 864 * 2, (size_t)raw_pitch * 4, 192 * 4, (size_t)A,
                       cudaMemcpyDeviceToDevice, st));
  LAUNCHED("trx_radio resample", 1);
  t->have_history = true;
  r = trx_pull_impl(ctx, t, (const btsdsp_cf32 *)dres, 0, res_pitch, nframes, fn0, (int32_t *)(dio + o_v), dio + o_d, 160, st);
  if (r != BTSDSP_OK) return r;
  CK(cudaMemcpyAsync(valid, dio + o_v, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpy2DAsync(dgram, dgram_pitch, dio + o_d, 160, 158, (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

int btsdsp_trx_pull_host(btsdsp_ctx *ctx, btsdsp_trx *t, const btsdsp_cf32 *bursts, long long pitch, int nframes, int fn0,
                         int32_t *valid, uint8_t *dgram, int dgram_pitch) {
  ARG(ctx && t && bursts && valid && dgram && nframes > 0 && pitch >= 157 && dgram_pitch >= 158);
  DeviceGuard g(ctx->device);
  const long long n = (long long)nframes * t->narfcn * 8;
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t o_b = take((size_t)n * pitch * sizeof(cf)), o_v = take((size_t)n * 4), o_d = take((size_t)n * 160);
  int r = grow(ctx, t->io, total, false);
  if (r != BTSDSP_OK) return r;
  uint8_t *d = (uint8_t *)t->io.p;
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(d + o_b, bursts, (size_t)n * pitch * sizeof(cf), cudaMemcpyHostToDevice, st));
  r = btsdsp_trx_pull_dev(ctx, t, (const btsdsp_cf32 *)(d + o_b), pitch, nframes, fn0, (int32_t *)(d + o_v), d + o_d, 160, st);
  if (r != BTSDSP_OK) return r;
  CK(cudaMemcpyAsync(valid, d + o_v, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpy2DAsync(dgram, dgram_pitch, d + o_d, 160, 158, (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

/* ---- TX datagrams: Transceiver::driveTransmitPriorityQueue (:582-632) + addRadioVector (:100-114) + pushBuffer ---- */
static int tx_datagrams_impl(btsdsp_ctx *ctx, const uint8_t *dgram, long long n, int dgram_pitch, int fn0, int nframes,
                             const uint8_t *filler, int16_t *out, long long *placed, bool v52m) {
  ARG(ctx && (dgram || n == 0) && out && n >= 0 && dgram_pitch >= 154 && fn0 >= 0 && nframes > 0 && (v52m || nframes % 117 == 0));
  if (ctx->sps != 1) return fail(ctx, BTSDSP_EUNSUPPORTED, "the TX stream path runs at sps == 1");
  DeviceGuard g(ctx->device);
  const long long nslots = (long long)nframes * 8, nchunks = nslots / 4 * 625 / 585;
  std::vector<uint8_t> bits((size_t)nslots * 148, 0);
  // the second variant scales at modulate time: fillers by 13500 (Transceiver52M/Transceiver.cpp:74), bursts by
  // 13500 * pow(10, -RSSI/10) (:111); its radio then only casts to short (no resampler, no further gain)
  std::vector<float> scale((size_t)nslots, filler ? (v52m ? 13500.0F : 1.0F) : 0.0F);
  if (filler) for (long long sl = 0; sl < nslots; sl++) memcpy(&bits[(size_t)sl * 148], filler, 148);
  long long ok = 0;
  for (long long i = 0; i < n; i++) {
    const uint8_t *d = dgram + i * (long long)dgram_pitch;
    const int tn = (int)(signed char)d[0];
    long long fn = 0;
    for (int k = 0; k < 4; k++) fn = (fn << 8) | d[1 + k];                                  // :593-595
    const int rssi = (int)(signed char)d[5];                                                // `(int) buffer[5]`, a char
    if (tn < 0 || tn > 7) continue;
    const long long f = fn_delta((int)(fn % kHyperframe), fn0 % kHyperframe);
    if (f < 0 || f >= nframes) continue;
    const long long sl = f * 8 + tn;
    memcpy(&bits[(size_t)sl * 148], d + 6, 148);                                            // :620-623
    scale[(size_t)sl] = v52m ? (float)(13500.0 * pow(10, -rssi / 10)) : (float)pow(10, -rssi / 10);   // :108, integer division
    ok++;
  }
  if (placed) *placed = ok;
  if (v52m) {
    const long long nsamp = nslots / 4 * 625;
    GROW(B_TSC, (size_t)nslots * 148);
    GROW(B_TOA, (size_t)nslots * 4);
    GROW(B_RES, (size_t)nsamp * sizeof(cf));
    GROW(B_RAW, (size_t)nsamp * 2 * sizeof(int16_t));
    cudaStream_t st = ctx->st;
    CK(cudaMemcpyAsync(dbuf<uint8_t>(ctx, B_TSC), bits.data(), bits.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dbuf<float>(ctx, B_TOA), scale.data(), scale.size() * 4, cudaMemcpyHostToDevice, st));
    launch_modulate(ctx->T, dbuf<uint8_t>(ctx, B_TSC), 148, nslots, -1, nullptr, 0, dbuf<cf>(ctx, B_RES), 0, st, dbuf<float>(ctx, B_TOA));
    launch_usrpify(dbuf<cf>(ctx, B_RES), nsamp, dbuf<int16_t>(ctx, B_RAW), st);
    LAUNCHED("tx_datagrams_52m", 2);
    CK(cudaMemcpyAsync(out, dbuf<int16_t>(ctx, B_RAW), (size_t)nsamp * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BTSDSP_OK;
  }
  GROW(B_TSC, (size_t)nslots * 148);
  GROW(B_TOA, (size_t)nslots * 4);
  GROW(B_RAW, (size_t)nchunks * 864 * 2 * sizeof(int16_t));
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(dbuf<uint8_t>(ctx, B_TSC), bits.data(), bits.size(), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dbuf<float>(ctx, B_TOA), scale.data(), scale.size() * 4, cudaMemcpyHostToDevice, st));
  launch_tx_fused(ctx->T, dbuf<uint8_t>(ctx, B_TSC), dbuf<float>(ctx, B_TOA), nslots, dbuf<int16_t>(ctx, B_RAW), st);
  LAUNCHED("tx_datagrams", 1);
  CK(cudaMemcpyAsync(out, dbuf<int16_t>(ctx, B_RAW), (size_t)nchunks * 864 * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));                             // also keeps `bits` / `scale` alive until the copies are done
  return BTSDSP_OK;
}

int btsdsp_tx_datagrams_host(btsdsp_ctx *ctx, const uint8_t *dgram, long long n, int dgram_pitch, int fn0, int nframes,
                             const uint8_t *filler, int16_t *out, long long *placed) {
  return tx_datagrams_impl(ctx, dgram, n, dgram_pitch, fn0, nframes, filler, out, placed, false);
}
int btsdsp_tx_datagrams_52m_host(btsdsp_ctx *ctx, const uint8_t *dgram, long long n, int dgram_pitch, int fn0, int nframes,
                                 const uint8_t *filler, int16_t *out, long long *placed) {
  return tx_datagrams_impl(ctx, dgram, n, dgram_pitch, fn0, nframes, filler, out, placed, true);
}

/* ---- L1 FEC after the path: XCCH deinterleave + Viterbi + Fire-code check (fec.cuh) ---- */
int btsdsp_xcch_decode_dev(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nframes, uint8_t *u,
                           int32_t *ok, void *stream) {
  ARG(ctx && soft_u8 && u && ok && nframes >= 0 && burst_pitch >= 148);
  DeviceGuard g(ctx->device);
  const int nl = launch_xcch_decode(soft_u8, burst_pitch, nframes, u, ok, (cudaStream_t)stream);
  LAUNCHED("xcch_decode", nl);
  return BTSDSP_OK;
}

int btsdsp_xcch_decode_host(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nframes, uint8_t *u,
                            int32_t *ok) {
  ARG(ctx && soft_u8 && u && ok && nframes > 0 && burst_pitch >= 148);
  DeviceGuard g(ctx->device);
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t o_s = take((size_t)nframes * 4 * burst_pitch), o_u = take((size_t)nframes * 228), o_k = take((size_t)nframes * 4);
  GROW(B_RAW, total);
  uint8_t *d = dbuf<uint8_t>(ctx, B_RAW);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(d + o_s, soft_u8, (size_t)nframes * 4 * burst_pitch, cudaMemcpyHostToDevice, st));
  int r = btsdsp_xcch_decode_dev(ctx, d + o_s, burst_pitch, nframes, d + o_u, (int32_t *)(d + o_k), st);
  if (r != BTSDSP_OK) return r;
  CK(cudaMemcpyAsync(u, d + o_u, (size_t)nframes * 228, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ok, d + o_k, (size_t)nframes * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

int btsdsp_tch_decode_dev(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nblocks, uint8_t *d, int32_t *good,
                          int32_t *stolen, uint8_t *facch_u, int32_t *facch_ok, void *stream) {
  ARG(ctx && soft_u8 && nblocks >= 0 && burst_pitch >= 148);
  DeviceGuard g(ctx->device);
  const int nl = launch_tch_decode(soft_u8, burst_pitch, nblocks, d, good, stolen, facch_u, facch_ok, (cudaStream_t)stream);
  LAUNCHED("tch_decode", nl);
  return BTSDSP_OK;
}

int btsdsp_tch_decode_host(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nblocks, uint8_t *d, int32_t *good,
                           int32_t *stolen, uint8_t *facch_u, int32_t *facch_ok) {
  ARG(ctx && soft_u8 && nblocks > 0 && burst_pitch >= 148);
  DeviceGuard g(ctx->device);
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t nbytes = (size_t)(4 * nblocks + 4) * burst_pitch;
  const size_t o_s = take(nbytes), o_d = take((size_t)nblocks * 260), o_g = take((size_t)nblocks * 4), o_t = take((size_t)nblocks * 4),
               o_u = take((size_t)nblocks * 228), o_k = take((size_t)nblocks * 4);
  GROW(B_RAW, total);
  uint8_t *b = dbuf<uint8_t>(ctx, B_RAW);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(b + o_s, soft_u8, nbytes, cudaMemcpyHostToDevice, st));
  int r = btsdsp_tch_decode_dev(ctx, b + o_s, burst_pitch, nblocks, b + o_d, (int32_t *)(b + o_g), (int32_t *)(b + o_t), b + o_u,
                                (int32_t *)(b + o_k), st);
  if (r != BTSDSP_OK) return r;
  if (d) CK(cudaMemcpyAsync(d, b + o_d, (size_t)nblocks * 260, cudaMemcpyDeviceToHost, st));
  if (good) CK(cudaMemcpyAsync(good, b + o_g, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
  if (stolen) CK(cudaMemcpyAsync(stolen, b + o_t, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
  if (facch_u) CK(cudaMemcpyAsync(facch_u, b + o_u, (size_t)nblocks * 228, cudaMemcpyDeviceToHost, st));
  if (facch_ok) CK(cudaMemcpyAsync(facch_ok, b + o_k, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

/* ---- L1 encoders on the transmit side (fec.cuh): L2 frames / speech frames -> the 148-bit normal bursts modulateBurst takes ---- */
static unsigned tsc_word_of(int tsc) {               /* midamble bit i (GSM 05.02 5.2.3) in bit 25 - i */
  unsigned w = 0;
  if (tsc >= 0 && tsc < 8)
    for (int i = 0; i < 26; i++) w |= (unsigned)(kTSC[tsc][i] == '1') << (25 - i);
  return w;
}
int btsdsp_xcch_encode_dev(btsdsp_ctx *ctx, const uint8_t *frames, long long nframes, int lsb8msb, int tsc, uint8_t *bursts,
                           int burst_pitch, void *stream) {
  ARG(ctx && frames && bursts && nframes >= 0 && burst_pitch >= 148 && tsc >= -1 && tsc < 8);
  DeviceGuard g(ctx->device);
  const int nl = launch_xcch_encode(frames, nframes, lsb8msb != 0, tsc_word_of(tsc), tsc >= 0, bursts, burst_pitch, (cudaStream_t)stream);
  LAUNCHED("xcch_encode", nl);
  return BTSDSP_OK;
}
int btsdsp_xcch_encode_host(btsdsp_ctx *ctx, const uint8_t *frames, long long nframes, int lsb8msb, int tsc, uint8_t *bursts,
                            int burst_pitch) {
  ARG(ctx && frames && bursts && nframes > 0 && burst_pitch >= 148 && tsc >= -1 && tsc < 8);
  DeviceGuard g(ctx->device);
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t nout = (size_t)nframes * 4 * burst_pitch;
  const size_t o_f = take((size_t)nframes * 184), o_b = take(nout);
  GROW(B_RAW, total);
  uint8_t *d = dbuf<uint8_t>(ctx, B_RAW);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(d + o_f, frames, (size_t)nframes * 184, cudaMemcpyHostToDevice, st));
  if (burst_pitch > 148) CK(cudaMemsetAsync(d + o_b, 0, nout, st));          /* bytes between the bursts: defined */
  int r = btsdsp_xcch_encode_dev(ctx, d + o_f, nframes, lsb8msb, tsc, d + o_b, burst_pitch, st);
  if (r != BTSDSP_OK) return r;
  CK(cudaMemcpyAsync(bursts, d + o_b, nout, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}
int btsdsp_tch_encode_dev(btsdsp_ctx *ctx, const uint8_t *d260, const uint8_t *f184, const uint8_t *steal, long long nblocks, int lsb8msb,
                          int tsc, const uint8_t *carry, uint8_t *bursts, int burst_pitch, void *stream) {
  ARG(ctx && d260 && f184 && steal && bursts && nblocks >= 0 && burst_pitch >= 148 && tsc >= -1 && tsc < 8);
  DeviceGuard g(ctx->device);
  const int nl = launch_tch_encode(d260, f184, steal, nblocks, lsb8msb != 0, tsc_word_of(tsc), tsc >= 0, carry, bursts, burst_pitch,
                                   (cudaStream_t)stream);
  LAUNCHED("tch_encode", nl);
  return BTSDSP_OK;
}
int btsdsp_tch_encode_host(btsdsp_ctx *ctx, const uint8_t *d260, const uint8_t *f184, const uint8_t *steal, long long nblocks, int lsb8msb,
                           int tsc, const uint8_t *carry, uint8_t *bursts, int burst_pitch) {
  ARG(ctx && d260 && f184 && steal && bursts && nblocks > 0 && burst_pitch >= 148 && tsc >= -1 && tsc < 8);
  DeviceGuard g(ctx->device);
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t nout = (size_t)(4 * nblocks + 4) * burst_pitch, ncar = (size_t)4 * burst_pitch;
  const size_t o_d = take((size_t)nblocks * 260), o_f = take((size_t)nblocks * 184), o_s = take((size_t)nblocks), o_c = take(ncar),
               o_b = take(nout);
  GROW(B_RAW, total);
  uint8_t *d = dbuf<uint8_t>(ctx, B_RAW);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(d + o_d, d260, (size_t)nblocks * 260, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_f, f184, (size_t)nblocks * 184, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_s, steal, (size_t)nblocks, cudaMemcpyHostToDevice, st));
  if (carry) CK(cudaMemcpyAsync(d + o_c, carry, ncar, cudaMemcpyHostToDevice, st));
  if (burst_pitch > 148) CK(cudaMemsetAsync(d + o_b, 0, nout, st));
  int r = btsdsp_tch_encode_dev(ctx, d + o_d, d + o_f, d + o_s, nblocks, lsb8msb, tsc, carry ? d + o_c : nullptr, d + o_b, burst_pitch, st);
  if (r != BTSDSP_OK) return r;
  CK(cudaMemcpyAsync(bursts, d + o_b, nout, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

int btsdsp_rach_decode_dev(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long n, uint8_t *u, int32_t *fields,
                           void *stream) {
  ARG(ctx && soft_u8 && fields && n >= 0 && burst_pitch >= 148);
  DeviceGuard g(ctx->device);
  const int nl = launch_rach_decode(soft_u8, burst_pitch, n, u, fields, (cudaStream_t)stream);
  LAUNCHED("rach_decode", nl);
  return BTSDSP_OK;
}

int btsdsp_rach_decode_host(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long n, uint8_t *u, int32_t *fields) {
  ARG(ctx && soft_u8 && fields && n > 0 && burst_pitch >= 148);
  DeviceGuard g(ctx->device);
  size_t total = 0;
  auto take = [&total](size_t bytes) { size_t o = (total + 255) & ~(size_t)255; total = o + bytes; return o; };
  const size_t o_s = take((size_t)n * burst_pitch), o_u = take((size_t)n * 18), o_k = take((size_t)n * 4);
  GROW(B_RAW, total);
  uint8_t *d = dbuf<uint8_t>(ctx, B_RAW);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(d + o_s, soft_u8, (size_t)n * burst_pitch, cudaMemcpyHostToDevice, st));
  int r = btsdsp_rach_decode_dev(ctx, d + o_s, burst_pitch, n, d + o_u, (int32_t *)(d + o_k), st);
  if (r != BTSDSP_OK) return r;
  if (u) CK(cudaMemcpyAsync(u, d + o_u, (size_t)n * 18, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(fields, d + o_k, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BTSDSP_OK;
}

/* ---- the reference's second transceiver variant (Transceiver52M/sigProcLib.cpp), the functions that differ ---- */
int btsdsp_analyze_52m_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens, long long first,
                           const uint8_t *tsc, long long n, float detect_thr, unsigned max_toa, int request_channel,
                           int32_t *flag, btsdsp_cf32 *amp, float *toa, btsdsp_cf32 *chan, float *chan_off, void *stream) {
  ARG(ctx && bursts && tsc && n >= 0 && pitch >= 0 && max_toa <= 60);
  DeviceGuard g(ctx->device);
  auto *ss = stream_scratch(ctx, (cudaStream_t)stream);
  GROWBUF(ss->scratch, (size_t)(n > 0 ? n : 1) * analyze_52m_scratch_stride(max_toa, ctx->sps) * sizeof(cf));
  NormalOut o = {flag, (cf *)amp, toa, (cf *)chan, chan_off, nullptr, nullptr, nullptr, 0};
  const int nl = launch_analyze_52m(ctx->T, make_src(bursts, pitch, lens, first, ctx->sps), tsc, n, detect_thr, max_toa,
                                    request_channel, o, (cf *)ss->scratch.p, (cudaStream_t)stream);
  LAUNCHED("analyze_52m", nl);
  return BTSDSP_OK;
}

int btsdsp_analyze_traffic_burst_52m(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, unsigned tsc, float threshold,
                                     unsigned max_toa, int request_channel, int *detected, btsdsp_cf32 *amp, float *toa,
                                     btsdsp_cf32 *chan, float *chan_off) {
  ARG(ctx && burst && detected && amp && toa && tsc < 8 && n >= 148 * ctx->sps && n <= 157 * kMaxSps && max_toa <= 60);
  ARG(!request_channel || (chan && chan_off));
  DeviceGuard g(ctx->device);
  const int sps = ctx->sps;
  GROW(B_A, (size_t)n * sizeof(cf));
  GROW(B_D, 1024);
  cudaStream_t st = ctx->st;
  uint8_t *d = dbuf<uint8_t>(ctx, B_D);
  int32_t *dflag = (int32_t *)d; cf *damp = (cf *)(d + 16); float *dtoa = (float *)(d + 32), *doff = (float *)(d + 48);
  cf *dchan = (cf *)(d + 64); uint8_t *dtsc = d + 512; int32_t *dlen = (int32_t *)(d + 528);
  const uint8_t t8 = (uint8_t)tsc; const int32_t l32 = n;
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), burst, (size_t)n * sizeof(cf), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dtsc, &t8, 1, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dlen, &l32, 4, cudaMemcpyHostToDevice, st));
  int r = btsdsp_analyze_52m_dev(ctx, (const btsdsp_cf32 *)dbuf<cf>(ctx, B_A), n, dlen, 0, dtsc, 1, threshold, max_toa,
                                 request_channel, dflag, (btsdsp_cf32 *)damp, dtoa, (btsdsp_cf32 *)dchan, doff, st);
  if (r != BTSDSP_OK) return r;
  int32_t f = 0;
  CK(cudaMemcpyAsync(&f, dflag, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(amp, damp, sizeof(cf), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(toa, dtoa, 4, cudaMemcpyDeviceToHost, st));
  if (request_channel) {
    CK(cudaMemcpyAsync(chan, dchan, (size_t)6 * sps * sizeof(cf), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(chan_off, doff, 4, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  *detected = f;
  return BTSDSP_OK;
}

int btsdsp_energy_detect_52m(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, unsigned window, float threshold, float *avg_pwr,
                             int *above) {
  ARG(ctx && v && above && n > 0 && window > 0 && 4LL * ((long long)(window < (unsigned)n ? window : n) - 1) < n);
  DeviceGuard g(ctx->device);
  GROW(B_A, (size_t)n * sizeof(cf));
  GROW(B_D, 64);
  cudaStream_t st = ctx->st;
  CK(cudaMemcpyAsync(dbuf<cf>(ctx, B_A), v, (size_t)n * sizeof(cf), cudaMemcpyHostToDevice, st));
  float *davg = dbuf<float>(ctx, B_D);
  int *dflag = (int *)(davg + 4);
  launch_energy_detect_52m(dbuf<cf>(ctx, B_A), n, window, threshold, davg, dflag, st);
  LAUNCHED("energy_detect_52m", 1);
  int f = 0; float a = 0.0F;
  CK(cudaMemcpyAsync(&f, dflag, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&a, davg, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *above = f;
  if (avg_pwr) *avg_pwr = a;
  return BTSDSP_OK;
}

int btsdsp_resample_rx_i16_streams_dev(btsdsp_ctx *ctx, const int16_t *iq, long long iq_pitch, int nstreams, int swap_iq,
                                       int has_history, long long nchunks, btsdsp_cf32 *out, long long out_pitch, void *stream) {
  ARG(ctx && iq && out && nstreams > 0 && nchunks >= 0 && iq_pitch >= nchunks * 864 && out_pitch >= nchunks * 585);
  DeviceGuard g(ctx->device);
  if (launch_resample_rx_i16_multi(iq, iq_pitch, nstreams, swap_iq, has_history, nchunks, (cf *)out, out_pitch, (cudaStream_t)stream) < 0)
    return fail(ctx, BTSDSP_EINVAL, "resample_rx_i16_streams: pointers must be 16-byte aligned, pitches multiples of 4 (iq) / 2 (out) samples");
  LAUNCHED("resample_rx_i16_streams", nchunks > 0);
  return BTSDSP_OK;
}

}  // extern "C"
