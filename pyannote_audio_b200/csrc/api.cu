// extern "C" surface of libb200diar.so (declared in include/b200diar.h): context, weight ingestion
// (BN folding, layout transforms, fp16 conversion on the host), and the forward entry points.
#include "../../include/b200diar.h"
#include "audio.cuh"
#include "cluster.cuh"
#include "common.cuh"
#include "emb.cuh"
#include "post.cuh"
#include "seg.cuh"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

using namespace b200;

struct b200_ctx {
  int device = 0;
  int num_sms = 148;
  int seg_gemm_impl = 1;   // 1 = split-fp16 tcgen05 GEMMs for the LSTM input projections / linear layers, 0 = fp32 SIMT
  int seg_conv_impl = 1;   // 1 = SincNet Conv1d(k=5) layers on tcgen05 (split fp16), 0 = fp32 CUDA-core kernel
  int seg_rec_impl = 1;    // 1 = LSTM recurrence on the tensor cores (needs seg_gemm_impl = 1), 0 = fp32 SIMT cluster kernel
  int conv_fuse = 1;       // 1 = layer1 BasicBlocks as one fused kernel (conv_block32_kernel) when conv_impl == 8
  int conv_ghost = 0;      // 1 = TMEM rings with ghost blocks (no seam-split MMAs) in conv_tc4 / conv_block32
  int conv_fold = 1;       // 1 = conv_tc3 loads one pixel box per (kh, channel block), kw taps are descriptor shifts
  int conv_scfold = 1;     // 1 = layer2.0: the 1x1 stride-2 shortcut rides in the padded weight rows of the stride-2 conv
  int conv_impl = 8;   // channels-as-M conv for C_out >= 128, strip-streaming conv for the narrow stride-1 3x3, per-tap conv otherwise
  int seg_max_batch = 4736;     // chunks per segmentation sub-batch (37 LSTM tiles of 128 sequences x 2 directions = 74 clusters)
  // chunks per embedding sub-batch.  296 = 2 x 148: the persistent conv kernels stride their items over 148 CTAs and
  // every layer's item count is a multiple of the sub-batch (8 / 4 strips, 20 / 10 pixel tiles per segment), so all
  // CTAs get the same number of items (256 left the last wave 30-92 % full: 516 -> 511 ms per bench step)
  int emb_max_batch = 296;
  int fbank_share = 1;          // 1 = overlapping hop-aligned chunks share their fbank frames (emb.cuh: FbankRun)
  int64_t launches = 0;
  SegWeights seg;
  EmbWeights emb;
  std::vector<void*> owned_seg, owned_emb;   // device allocations holding the weights of each network
  std::vector<void*>* owned = &owned_seg;    // where upload() records allocations (set by the load entry points)
  void* ws = nullptr;
  size_t ws_cap = 0;
  long long* d_off = nullptr;
  int* d_valid = nullptr;
  int* d_frame0 = nullptr;         // first fbank row of every chunk inside its sub-batch
  b200::FbankRun* d_runs = nullptr; // fbank runs of all sub-batches, back to back
  int meta_cap = 0;
  // optional CUDA-event timers around the dominant kernels (bench.py's live roofline measurement)
  int profile = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> trunk_events, seg_events;
  std::vector<cudaEvent_t> event_pool;
  int64_t trunk_segments = 0, seg_chunks = 0;
  // polyphase resampling tables, one per (orig, new) ratio seen (device copies, freed with the ctx)
  struct ResampleTable { int orig, nw, width; float* dev; };
  std::vector<ResampleTable> resample_tables;
};

namespace {

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

template <typename T>
int upload(b200_ctx* ctx, const std::vector<T>& h, T** out) {
  void* p = nullptr;
  B200_CUDA_OK(cudaMalloc(&p, h.size() * sizeof(T) + 16));
  B200_CUDA_OK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  ctx->owned->push_back(p);
  *out = reinterpret_cast<T*>(p);
  return B200_OK;
}

cudaEvent_t take_event(b200_ctx* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct ScopedTimer {
  b200_ctx* ctx;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* sink;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  ScopedTimer(b200_ctx* c, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* s, cudaStream_t stream)
      : ctx(c), sink(s), st(stream) {
    if (ctx->profile) { a = take_event(ctx); b = take_event(ctx); cudaEventRecord(a, st); }
  }
  ~ScopedTimer() {
    if (ctx->profile) { cudaEventRecord(b, st); sink->push_back({a, b}); }
  }
};

// a reload replaces the previous upload of the same network: free it (after the device drained) instead of leaking
void release_weights(b200_ctx* ctx, std::vector<void*>* list) {
  if (!list->empty()) cudaDeviceSynchronize();
  for (void* p : *list) cudaFree(p);
  list->clear();
  ctx->owned = list;
}

int ensure_ws(b200_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_cap) return B200_OK;
  if (ctx->ws) { cudaDeviceSynchronize(); cudaFree(ctx->ws); ctx->ws = nullptr; ctx->ws_cap = 0; }
  B200_CUDA_OK(cudaMalloc(&ctx->ws, bytes));
  ctx->ws_cap = bytes;
  return B200_OK;
}

int ensure_meta(b200_ctx* ctx, int n) {
  if (n <= ctx->meta_cap) return B200_OK;
  if (ctx->d_off) {
    cudaDeviceSynchronize();
    cudaFree(ctx->d_off); cudaFree(ctx->d_valid); cudaFree(ctx->d_frame0); cudaFree(ctx->d_runs);
  }
  const int cap = n + 1024;
  B200_CUDA_OK(cudaMalloc((void**)&ctx->d_off, sizeof(long long) * cap));
  B200_CUDA_OK(cudaMalloc((void**)&ctx->d_valid, sizeof(int) * cap));
  B200_CUDA_OK(cudaMalloc((void**)&ctx->d_frame0, sizeof(int) * cap));
  B200_CUDA_OK(cudaMalloc((void**)&ctx->d_runs, sizeof(b200::FbankRun) * cap));
  ctx->meta_cap = cap;
  return B200_OK;
}

int push_meta(b200_ctx* ctx, const int64_t* off, const int32_t* valid, int n, cudaStream_t st) {
  int rc = ensure_meta(ctx, n);
  if (rc) return rc;
  for (int i = 0; i < n; ++i)
    B200_CHECK(valid[i] >= 0 && valid[i] <= kChunk && off[i] >= 0, B200_ERR_INVALID,
               "chunk %d: offset %lld / valid %d out of range", i, (long long)off[i], (int)valid[i]);
  B200_CUDA_OK(cudaMemcpyAsync(ctx->d_off, off, sizeof(long long) * n, cudaMemcpyHostToDevice, st));
  B200_CUDA_OK(cudaMemcpyAsync(ctx->d_valid, valid, sizeof(int) * n, cudaMemcpyHostToDevice, st));
  return B200_OK;
}

// fbank plan of a list of chunks processed in sub-batches of nbmax (emb.cuh: FbankRun): per sub-batch the runs
// [run_base[s], run_base[s + 1]) and the number of fbank rows; frame0 / runs go to the device with the chunk table.
// share = 0 (and every short or unaligned chunk): one private run per chunk, row0 = b * 998.
struct FbankPlan {
  std::vector<int> run_base, nrows;
};
void plan_fbank(const int64_t* off, const int32_t* valid, int n, int nbmax, bool share,
                std::vector<b200::FbankRun>* runs_out, std::vector<int>* frame0_out, FbankPlan* plan) {
  constexpr int kHop = 160;
  std::vector<b200::FbankRun>& runs = *runs_out;
  std::vector<int>& frame0 = *frame0_out;
  runs.clear();
  runs.reserve((size_t)n);
  frame0.assign((size_t)n, 0);
  plan->run_base.clear();
  plan->nrows.clear();
  for (int c0 = 0; c0 < n; c0 += nbmax) {
    const int nb = (n - c0) < nbmax ? (n - c0) : nbmax;
    plan->run_base.push_back((int)runs.size());
    int rows = 0, run_rows = 0;
    bool open = false;                                       // the last run is made of full chunks and may be extended
    for (int b = 0; b < nb; ++b) {
      const long long o = off[c0 + b];
      const bool full = share && valid[c0 + b] == kChunk;
      if (full && open) {
        const long long d = o - runs.back().src;
        if (d >= 0 && d % kHop == 0 && d / kHop <= run_rows) {
          const int f0 = (int)(d / kHop);
          frame0[c0 + b] = runs.back().row0 + f0;
          if (f0 + kFbankFrames > run_rows) { rows += f0 + kFbankFrames - run_rows; run_rows = f0 + kFbankFrames; }
          continue;
        }
      }
      runs.push_back(b200::FbankRun{o, rows, full ? INT_MAX : (int)valid[c0 + b]});
      frame0[c0 + b] = rows;
      rows += kFbankFrames;
      run_rows = kFbankFrames;
      open = full;
    }
    plan->nrows.push_back(rows);
  }
  plan->run_base.push_back((int)runs.size());
}

int push_fbank_plan(b200_ctx* ctx, const int64_t* off, const int32_t* valid, int n, int nbmax, bool share,
                    FbankPlan* plan, cudaStream_t st) {
  std::vector<b200::FbankRun> runs;
  std::vector<int> frame0;
  plan_fbank(off, valid, n, nbmax, share, &runs, &frame0, plan);
  B200_CUDA_OK(cudaMemcpyAsync(ctx->d_frame0, frame0.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  B200_CUDA_OK(cudaMemcpyAsync(ctx->d_runs, runs.data(), sizeof(b200::FbankRun) * runs.size(), cudaMemcpyHostToDevice, st));
  return B200_OK;
}

// ---- kaldi mel bank / window / twiddles (torchaudio.compliance.kaldi.get_mel_banks, _feature_window_function) ----
double mel_scale(double f) { return 1127.0 * std::log(1.0 + f / 700.0); }

int build_fbank_constants(b200_ctx* ctx) {
  EmbWeights& E = ctx->emb;
  std::vector<float> window(400);
  for (int i = 0; i < 400; ++i) window[i] = (float)(0.54 - 0.46 * std::cos(2.0 * M_PI * i / 399.0));
  std::vector<float> tw(512);
  for (int k = 0; k < 256; ++k) {
    tw[2 * k] = (float)std::cos(2.0 * M_PI * k / 512.0);
    tw[2 * k + 1] = (float)(-std::sin(2.0 * M_PI * k / 512.0));
  }
  const int nb = 80, nfft = 256;
  const double low = 20.0, high = 8000.0, bw = 16000.0 / 512.0;
  const double ml = mel_scale(low), mh = mel_scale(high), delta = (mh - ml) / (nb + 1);
  std::vector<float> w;
  std::vector<int> st(nb), ln(nb), off(nb);
  for (int b = 0; b < nb; ++b) {
    const double left = ml + b * delta, center = ml + (b + 1.0) * delta, right = ml + (b + 2.0) * delta;
    int first = -1, last = -1;
    std::vector<float> row(nfft, 0.f);
    for (int k = 0; k < nfft; ++k) {
      const double mel = mel_scale(bw * k);
      const double up = (mel - left) / (center - left), down = (right - mel) / (right - center);
      const double v = std::fmax(0.0, std::fmin(up, down));
      row[k] = (float)v;
      if (v > 0) { if (first < 0) first = k; last = k; }
    }
    if (first < 0) { first = 0; last = -1; }
    st[b] = first; ln[b] = last - first + 1; off[b] = (int)w.size();
    for (int k = first; k <= last; ++k) w.push_back(row[k]);
  }
  int rc;
  if ((rc = upload(ctx, window, &E.window))) return rc;
  if ((rc = upload(ctx, tw, &E.twiddle))) return rc;
  if ((rc = upload(ctx, w, &E.mel_w))) return rc;
  if ((rc = upload(ctx, st, &E.mel_start))) return rc;
  if ((rc = upload(ctx, ln, &E.mel_len))) return rc;
  if ((rc = upload(ctx, off, &E.mel_off))) return rc;
  return B200_OK;
}

// fold eval-mode BatchNorm2d into a conv: w' = w * g/sqrt(v+eps),  b' = beta - mean * g/sqrt(v+eps)
int make_conv(b200_ctx* ctx, const b200_conv_bn& src, int cin, int cout, int k, int stride, ConvLayer* L) {
  B200_CHECK(src.conv_weight && src.bn_weight && src.bn_bias && src.bn_mean && src.bn_var, B200_ERR_INVALID,
             "missing conv/bn tensor (cin=%d cout=%d)", cin, cout);
  L->C_in = cin; L->C_out = cout; L->ksize = k; L->stride = stride;
  std::vector<__half> w((size_t)k * k * cout * cin);
  std::vector<float> bias(cout);
  for (int co = 0; co < cout; ++co) {
    const float s = src.bn_weight[co] / std::sqrt(src.bn_var[co] + 1e-5f);
    bias[co] = src.bn_bias[co] - src.bn_mean[co] * s;
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < k * k; ++t)
        w[((size_t)t * cout + co) * cin + ci] = __float2half(src.conv_weight[((size_t)co * cin + ci) * k * k + t] * s);
  }
  int rc;
  if ((rc = upload(ctx, w, &L->w))) return rc;
  if ((rc = upload(ctx, bias, &L->bias))) return rc;
  if (k == 3 && stride == 1 && cin == cout && (cin == 32 || cin == 64)) {   // conv_tc4_kernel: [kw][(kh, c_out)][c_in]
    const int C = cin;
    std::vector<__half> w4((size_t)9 * C * C);
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        for (int co = 0; co < C; ++co)
          for (int ci = 0; ci < C; ++ci)
            w4[(((size_t)kw * 3 + kh) * C + co) * C + ci] = w[((size_t)(kh * 3 + kw) * cout + co) * cin + ci];
    if ((rc = upload(ctx, w4, &L->w4))) return rc;
  }
  if ((k == 3 && stride == 1 && cout >= 128) || stride == 2) {
    // copy for the channels-as-M kernel (rows padded to 128): wide stride-1 convs, the stride-2 convs and the
    // 1x1 stride-2 shortcuts
    const int rows = (cout + 127) / 128 * 128;
    std::vector<__half> w3((size_t)k * k * rows * cin, __float2half(0.f));
    for (int t = 0; t < k * k; ++t)
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci) w3[((size_t)t * rows + co) * cin + ci] = w[((size_t)t * cout + co) * cin + ci];
    if ((rc = upload(ctx, w3, &L->w3))) return rc;
  }
  return B200_OK;
}

}  // namespace

extern "C" {

const char* b200_last_error(void) { return b200::last_error(); }
int b200_version(void) { return 100; }

int b200_ctx_create(b200_ctx** out, int device) {
  B200_CHECK(out != nullptr, B200_ERR_INVALID, "ctx pointer is NULL");
  int ndev = 0;
  B200_CUDA_OK(cudaGetDeviceCount(&ndev));
  B200_CHECK(device >= 0 && device < ndev, B200_ERR_INVALID, "device %d not available (%d devices)", device, ndev);
  cudaDeviceProp prop;
  B200_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  B200_CHECK(prop.major == 10, B200_ERR_STATE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
             device, prop.major, prop.minor);
  b200_ctx* c = new b200_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if (const char* e = std::getenv("B200_CONV_IMPL")) c->conv_impl = std::atoi(e);
  if (const char* e = std::getenv("B200_EMB_MAX_BATCH")) c->emb_max_batch = std::atoi(e) > 0 ? std::atoi(e) : c->emb_max_batch;
  if (const char* e = std::getenv("B200_SEG_MAX_BATCH")) c->seg_max_batch = std::atoi(e) > 0 ? std::atoi(e) : c->seg_max_batch;
  *out = c;
  return B200_OK;
}

int b200_ctx_destroy(b200_ctx* ctx) {
  if (!ctx) return B200_OK;
  DeviceGuard g(ctx->device);
  cudaDeviceSynchronize();
  for (void* p : ctx->owned_seg) cudaFree(p);
  for (void* p : ctx->owned_emb) cudaFree(p);
  for (auto& t : ctx->resample_tables) cudaFree(t.dev);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->d_off) cudaFree(ctx->d_off);
  if (ctx->d_valid) cudaFree(ctx->d_valid);
  if (ctx->d_frame0) cudaFree(ctx->d_frame0);
  if (ctx->d_runs) cudaFree(ctx->d_runs);
  delete ctx;
  return B200_OK;
}

int b200_ctx_set_option(b200_ctx* ctx, const char* key, int64_t value) {
  B200_CHECK(ctx && key, B200_ERR_INVALID, "NULL ctx/key");
  std::string k(key);
  if (k == "conv_impl") ctx->conv_impl = (int)value;
  else if (k == "seg_max_batch") ctx->seg_max_batch = (int)value;
  else if (k == "emb_max_batch") ctx->emb_max_batch = (int)value;
  else if (k == "profile") ctx->profile = (int)value;
  else if (k == "seg_gemm_impl") ctx->seg_gemm_impl = (int)value;
  else if (k == "seg_rec_impl") ctx->seg_rec_impl = (int)value;
  else if (k == "seg_conv_impl") ctx->seg_conv_impl = (int)value;
  else if (k == "conv_fuse") ctx->conv_fuse = (int)value;
  else if (k == "conv_ghost") ctx->conv_ghost = (int)value;
  else if (k == "conv_fold") ctx->conv_fold = (int)value;
  else if (k == "conv_scfold") ctx->conv_scfold = (int)value;
  else if (k == "fbank_share") ctx->fbank_share = (int)value;
  else B200_CHECK(false, B200_ERR_INVALID, "unknown option '%s'", key);
  B200_CHECK(ctx->seg_max_batch >= 1 && ctx->emb_max_batch >= 1 && ctx->conv_impl >= 0 && ctx->conv_impl <= 8,
             B200_ERR_INVALID, "option '%s' value %lld out of range", key, (long long)value);
  return B200_OK;
}

int64_t b200_ctx_launch_count(const b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200_ctx_timer(b200_ctx* ctx, const char* name, double* total_ms, int64_t* units) {
  B200_CHECK(ctx && name && total_ms && units, B200_ERR_INVALID, "bad arguments");
  DeviceGuard g(ctx->device);
  std::string k(name);
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* v = nullptr;
  int64_t* u = nullptr;
  if (k == "trunk") { v = &ctx->trunk_events; u = &ctx->trunk_segments; }
  else if (k == "seg") { v = &ctx->seg_events; u = &ctx->seg_chunks; }
  else B200_CHECK(false, B200_ERR_INVALID, "unknown timer '%s'", name);
  B200_CUDA_OK(cudaDeviceSynchronize());
  double ms = 0.0;
  for (auto& pr : *v) {
    float t = 0.f;
    B200_CUDA_OK(cudaEventElapsedTime(&t, pr.first, pr.second));
    ms += t;
    ctx->event_pool.push_back(pr.first);
    ctx->event_pool.push_back(pr.second);
  }
  v->clear();
  *total_ms = ms;
  *units = *u;
  *u = 0;
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------------
static int64_t gcd64(int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; }

int64_t b200_audio_num_frames(int64_t frames_in, int32_t sr_in, int32_t sr_out) {
  if (frames_in <= 0 || sr_in <= 0 || sr_out <= 0) return 0;
  if (sr_in == sr_out) return frames_in;
  const int64_t g = gcd64(sr_in, sr_out), orig = sr_in / g, nw = sr_out / g;
  return (nw * frames_in + orig - 1) / orig;                // ceil(new * length / orig)
}

int b200_audio_ingest(b200_ctx* ctx, const void* pcm, int32_t format, int32_t channels, int64_t frames_in,
                      int32_t sr_in, int32_t sr_out, int32_t channel, float* out, int64_t out_capacity, void* stream) {
  B200_CHECK(ctx && pcm && out && channels >= 1 && frames_in >= 0 && sr_in > 0 && sr_out > 0, B200_ERR_INVALID,
             "bad arguments");
  B200_CHECK(format == B200_PCM_S16_INTERLEAVED || format == B200_PCM_F32_PLANAR, B200_ERR_INVALID,
             "unknown PCM format %d", (int)format);
  B200_CHECK(channel < channels, B200_ERR_INVALID, "channel %d of a %d-channel file", (int)channel, (int)channels);
  const int64_t frames_out = b200_audio_num_frames(frames_in, sr_in, sr_out);
  B200_CHECK(frames_out <= out_capacity, B200_ERR_INVALID, "output holds %lld samples, %lld needed",
             (long long)out_capacity, (long long)frames_out);
  if (frames_out == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  const int64_t gg = gcd64(sr_in, sr_out);
  const int orig = (int)(sr_in / gg), nw = (int)(sr_out / gg);
  const b200_ctx::ResampleTable* tab = nullptr;
  for (auto& t : ctx->resample_tables)
    if (t.orig == orig && t.nw == nw) tab = &t;
  if (!tab) {
    b200_ctx::ResampleTable t{orig, nw, 0, nullptr};
    std::vector<float> host;
    if (orig == nw) { host.assign(1, 1.0f); t.width = 0; }   // same rate: y[i] = 1.0 * x[i] (exact)
    else resample_table(orig, nw, &t.width, &host);
    B200_CUDA_OK(cudaMalloc((void**)&t.dev, host.size() * sizeof(float)));
    B200_CUDA_OK(cudaMemcpy(t.dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->resample_tables.push_back(t);
    tab = &ctx->resample_tables.back();
  }
  ctx->launches += 1;
  return audio_ingest(pcm, format, channels, frames_in, channel, tab->dev, tab->orig, tab->nw, tab->width, out,
                      frames_out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------
int b200_seg_load(b200_ctx* ctx, const b200_seg_weights* w) {
  B200_CHECK(ctx && w, B200_ERR_INVALID, "NULL ctx/weights");
  DeviceGuard g(ctx->device);
  SegWeights& S = ctx->seg;
  B200_CHECK(w->lstm_layers >= 1 && w->lstm_layers <= 4, B200_ERR_INVALID, "lstm_layers=%d unsupported",
             (int)w->lstm_layers);
  S.loaded = false;
  release_weights(ctx, &ctx->owned_seg);
  S.lstm_layers = w->lstm_layers;
  S.wav_w = w->wav_norm_weight;
  S.wav_b = w->wav_norm_bias;
  int rc;
  {  // half filter bank [126][80]; the kernel relies on the (anti)symmetry of ParamSincFB filters
    B200_CHECK(w->sinc_filters, B200_ERR_INVALID, "sinc_filters is NULL");
    std::vector<float> f(126 * 80);
    for (int ch = 0; ch < 80; ++ch) {
      const float* r = w->sinc_filters + ch * 251;
      const float sign = ch < 40 ? 1.f : -1.f;
      float mx = 0.f, err = 0.f;
      for (int k = 0; k < 125; ++k) {
        mx = std::fmax(mx, std::fabs(r[k]));
        err = std::fmax(err, std::fabs(r[k] - sign * r[250 - k]));
        f[k * 80 + ch] = r[k];
      }
      if (ch >= 40) err = std::fmax(err, std::fabs(r[125]));
      B200_CHECK(err <= 1e-6f * (mx + 1e-30f) + 1e-12f, B200_ERR_INVALID,
                 "sinc filter %d is not (anti)symmetric (err %g): not a ParamSincFB bank", ch, (double)err);
      f[125 * 80 + ch] = ch < 40 ? r[125] : 0.f;
    }
    if ((rc = upload(ctx, f, &S.sinc_f))) return rc;
    {   // tensor-core layout of the full bank: [k-step][128 rows][16 taps] as fp16 (hi, lo), taps 251..255 zero
      std::vector<__half> hi((size_t)16 * 128 * 16, __float2half(0.f)), lo(hi.size(), __float2half(0.f));
      for (int ch = 0; ch < 80; ++ch)
        for (int k = 0; k < 251; ++k) {
          const float v = w->sinc_filters[ch * 251 + k];
          const size_t o = ((size_t)(k / 16) * 128 + ch) * 16 + (k % 16);
          hi[o] = __float2half(v);
          lo[o] = __float2half(v - __half2float(hi[o]));
        }
      if ((rc = upload(ctx, hi, &S.sinc_tc_hi))) return rc;
      if ((rc = upload(ctx, lo, &S.sinc_tc_lo))) return rc;
    }
  }
  const int nch[3] = {80, 60, 60};
  for (int i = 0; i < 3; ++i) {
    B200_CHECK(w->norm_weight[i] && w->norm_bias[i], B200_ERR_INVALID, "norm1d.%d missing", i);
    std::vector<float> gmm(w->norm_weight[i], w->norm_weight[i] + nch[i]), bt(w->norm_bias[i], w->norm_bias[i] + nch[i]);
    if ((rc = upload(ctx, gmm, &S.in_gamma[i]))) return rc;
    if ((rc = upload(ctx, bt, &S.in_beta[i]))) return rc;
  }
  const int cin[2] = {80, 60};
  for (int i = 0; i < 2; ++i) {
    B200_CHECK(w->conv_weight[i] && w->conv_bias[i], B200_ERR_INVALID, "conv1d.%d missing", i + 1);
    std::vector<float> wc((size_t)cin[i] * 5 * 60), bc(w->conv_bias[i], w->conv_bias[i] + 60);
    for (int co = 0; co < 60; ++co)
      for (int ci = 0; ci < cin[i]; ++ci)
        for (int k = 0; k < 5; ++k) wc[((size_t)ci * 5 + k) * 60 + co] = w->conv_weight[i][((size_t)co * cin[i] + ci) * 5 + k];
    if ((rc = upload(ctx, wc, &S.conv_w[i]))) return rc;
    if ((rc = upload(ctx, bc, &S.conv_b[i]))) return rc;
    {   // tensor-core layout: [channel block of 16][tap][128 rows = c_out (60 real)][16 c_in] as fp16 (hi, lo)
      const int ncb = (cin[i] + 15) / 16;
      std::vector<__half> hi((size_t)ncb * 5 * 128 * 16, __float2half(0.f)), lo(hi.size(), __float2half(0.f));
      for (int co = 0; co < 60; ++co)
        for (int ci = 0; ci < cin[i]; ++ci)
          for (int k = 0; k < 5; ++k) {
            const float v = w->conv_weight[i][((size_t)co * cin[i] + ci) * 5 + k];
            const size_t o = ((((size_t)(ci / 16) * 5 + k) * 128) + co) * 16 + (ci % 16);
            hi[o] = __float2half(v);
            lo[o] = __float2half(v - __half2float(hi[o]));
          }
      if ((rc = upload(ctx, hi, &S.conv_tc_hi[i]))) return rc;
      if ((rc = upload(ctx, lo, &S.conv_tc_lo[i]))) return rc;
    }
  }
  for (int l = 0; l < S.lstm_layers; ++l) {
    const int I = l == 0 ? 60 : 256, Kp = l == 0 ? 64 : 256;
    S.k_in[l] = Kp;
    std::vector<float> wih((size_t)1024 * Kp, 0.f), bg(1024), whh((size_t)2 * 2 * 128 * 256);
    for (int d = 0; d < 2; ++d) {
      const float *Wi = w->lstm_w_ih[l * 2 + d], *Wh = w->lstm_w_hh[l * 2 + d];
      const float *bi = w->lstm_b_ih[l * 2 + d], *bh = w->lstm_b_hh[l * 2 + d];
      B200_CHECK(Wi && Wh && bi && bh, B200_ERR_INVALID, "lstm layer %d dir %d missing", l, d);
      for (int u = 0; u < 128; ++u)
        for (int gt = 0; gt < 4; ++gt) {
          const int n = d * 512 + u * 4 + gt, src = gt * 128 + u;
          for (int k = 0; k < I; ++k) wih[(size_t)n * Kp + k] = Wi[(size_t)src * I + k];
          bg[n] = bi[src] + bh[src];
        }
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 128; ++k)
          for (int p = 0; p < 2; ++p)
            for (int tx = 0; tx < 32; ++tx)
              for (int gt = 0; gt < 4; ++gt) {
                const int unit = 64 * r + 2 * tx + p;
                whh[(((size_t)(d * 2 + r) * 128 + k) * 256) + p * 128 + tx * 4 + gt] = Wh[(size_t)(gt * 128 + unit) * 128 + k];
              }
    }
    {
      std::vector<__half> hi(wih.size()), lo(wih.size());
      for (size_t i = 0; i < wih.size(); ++i) {
        hi[i] = __float2half(wih[i]);
        lo[i] = __float2half(wih[i] - __half2float(hi[i]));
      }
      if ((rc = upload(ctx, hi, &S.w_ih_hi[l]))) return rc;
      if ((rc = upload(ctx, lo, &S.w_ih_lo[l]))) return rc;
    }
    if ((rc = upload(ctx, wih, &S.w_ih[l]))) return rc;
    if ((rc = upload(ctx, bg, &S.b_g[l]))) return rc;
    if ((rc = upload(ctx, whh, &S.w_hh[l]))) return rc;
    {   // tensor-core recurrence: rows (dir, rank, unit_local, gate), k contiguous (permuted, below), as fp16 (hi, lo)
      std::vector<__half> hi((size_t)1024 * 128), lo((size_t)1024 * 128);
      for (int d = 0; d < 2; ++d) {
        const float* Wh = w->lstm_w_hh[l * 2 + d];
        for (int r = 0; r < 2; ++r)
          for (int ul = 0; ul < 64; ++ul)
            for (int gt = 0; gt < 4; ++gt)
              for (int k = 0; k < 128; ++k) {
                const float v = Wh[(size_t)(gt * 128 + 64 * r + ul) * 128 + k];
                // K order inside a k-block of 64 units: (chunk, half, unit-in-chunk) for unit = half*32 + chunk*8 + u,
                // so that the units the epilogue warps finish together form one 16-wide k-step (seg_lstm_tc.cu)
                const int kl = k & 63, kp = (k & 64) | (((kl & 31) >> 3) << 4) | ((kl >> 5) << 3) | (kl & 7);
                const size_t o = ((size_t)((d * 2 + r) * 256 + ul * 4 + gt)) * 128 + kp;
                hi[o] = __float2half(v);
                lo[o] = __float2half(v - __half2float(hi[o]));
              }
      }
      if ((rc = upload(ctx, hi, &S.w_hh_hi[l]))) return rc;
      if ((rc = upload(ctx, lo, &S.w_hh_lo[l]))) return rc;
    }
  }
  const int lin_in[2] = {256, 128};
  for (int i = 0; i < 2; ++i) {
    B200_CHECK(w->linear_weight[i] && w->linear_bias[i], B200_ERR_INVALID, "linear.%d missing", i);
    std::vector<float> lw(w->linear_weight[i], w->linear_weight[i] + 128 * lin_in[i]), lb(w->linear_bias[i], w->linear_bias[i] + 128);
    {
      std::vector<__half> hi(lw.size()), lo(lw.size());
      for (size_t j = 0; j < lw.size(); ++j) {
        hi[j] = __float2half(lw[j]);
        lo[j] = __float2half(lw[j] - __half2float(hi[j]));
      }
      if ((rc = upload(ctx, hi, &S.lin_w_hi[i]))) return rc;
      if ((rc = upload(ctx, lo, &S.lin_w_lo[i]))) return rc;
    }
    if ((rc = upload(ctx, lw, &S.lin_w[i]))) return rc;
    if ((rc = upload(ctx, lb, &S.lin_b[i]))) return rc;
  }
  B200_CHECK(w->classifier_weight && w->classifier_bias, B200_ERR_INVALID, "classifier missing");
  std::vector<float> cw(w->classifier_weight, w->classifier_weight + 7 * 128), cb(w->classifier_bias, w->classifier_bias + 7);
  if ((rc = upload(ctx, cw, &S.cls_w))) return rc;
  if ((rc = upload(ctx, cb, &S.cls_b))) return rc;
  S.loaded = true;
  return B200_OK;
}

int b200_emb_load(b200_ctx* ctx, const b200_emb_weights* w) {
  B200_CHECK(ctx && w, B200_ERR_INVALID, "NULL ctx/weights");
  DeviceGuard g(ctx->device);
  EmbWeights& E = ctx->emb;
  int rc;
  E.loaded = false;
  release_weights(ctx, &ctx->owned_emb);
  if ((rc = build_fbank_constants(ctx))) return rc;
  {
    const b200_conv_bn& s = w->stem;
    B200_CHECK(s.conv_weight && s.bn_weight && s.bn_bias && s.bn_mean && s.bn_var, B200_ERR_INVALID, "stem missing");
    std::vector<float> cw(32 * 9), cb(32);
    for (int c = 0; c < 32; ++c) {
      const float sc = s.bn_weight[c] / std::sqrt(s.bn_var[c] + 1e-5f);
      cb[c] = s.bn_bias[c] - s.bn_mean[c] * sc;
      for (int k = 0; k < 9; ++k) cw[c * 9 + k] = s.conv_weight[c * 9 + k] * sc;
    }
    if ((rc = upload(ctx, cw, &E.conv1_w))) return rc;
    if ((rc = upload(ctx, cb, &E.conv1_b))) return rc;
  }
  E.blocks.clear();
  const int planes[4] = {32, 64, 128, 256}, nblk[4] = {3, 4, 6, 3}, strides[4] = {1, 2, 2, 2};
  int in_planes = 32, bi = 0;
  for (int l = 0; l < 4; ++l)
    for (int i = 0; i < nblk[l]; ++i, ++bi) {
      BlockWeights B;
      const int s = i == 0 ? strides[l] : 1;
      if ((rc = make_conv(ctx, w->block_conv1[bi], in_planes, planes[l], 3, s, &B.conv1))) return rc;
      if ((rc = make_conv(ctx, w->block_conv2[bi], planes[l], planes[l], 3, 1, &B.conv2))) return rc;
      B.has_shortcut = (s != 1 || in_planes != planes[l]);
      if (B.has_shortcut) {
        if ((rc = make_conv(ctx, w->block_shortcut[bi], in_planes, planes[l], 1, s, &B.shortcut))) return rc;
        if (planes[l] == 64 && s == 2) {
          // the 64 zero-padded rows of conv1's 128-row channels-as-M weight tiles carry the 1x1 shortcut (centre tap)
          const int cin = in_planes, cout = 64;
          const b200_conv_bn &c1 = w->block_conv1[bi], &sc = w->block_shortcut[bi];
          std::vector<__half> w3s((size_t)9 * 128 * cin, __float2half(0.f));
          std::vector<float> bias_s(128);
          for (int co = 0; co < cout; ++co) {
            const float s1 = c1.bn_weight[co] / std::sqrt(c1.bn_var[co] + 1e-5f);
            const float s2 = sc.bn_weight[co] / std::sqrt(sc.bn_var[co] + 1e-5f);
            bias_s[co] = c1.bn_bias[co] - c1.bn_mean[co] * s1;
            bias_s[64 + co] = sc.bn_bias[co] - sc.bn_mean[co] * s2;
            for (int ci = 0; ci < cin; ++ci) {
              for (int t = 0; t < 9; ++t)
                w3s[((size_t)t * 128 + co) * cin + ci] = __float2half(c1.conv_weight[((size_t)co * cin + ci) * 9 + t] * s1);
              w3s[((size_t)4 * 128 + 64 + co) * cin + ci] = __float2half(sc.conv_weight[(size_t)co * cin + ci] * s2);
            }
          }
          if ((rc = upload(ctx, w3s, &B.conv1.w3s))) return rc;
          if ((rc = upload(ctx, bias_s, &B.conv1.bias_s))) return rc;
        }
      }
      in_planes = planes[l];
      E.blocks.push_back(B);
    }
  B200_CHECK(w->seg1_weight && w->seg1_bias, B200_ERR_INVALID, "seg_1 missing");
  std::vector<float> sw(w->seg1_weight, w->seg1_weight + (size_t)256 * 5120), sb(w->seg1_bias, w->seg1_bias + 256);
  {
    std::vector<__half> hi(sw.size()), lo(sw.size());
    for (size_t i = 0; i < sw.size(); ++i) {
      hi[i] = __float2half(sw[i]);
      lo[i] = __float2half(sw[i] - __half2float(hi[i]));
    }
    if ((rc = upload(ctx, hi, &E.seg1_w_hi))) return rc;
    if ((rc = upload(ctx, lo, &E.seg1_w_lo))) return rc;
  }
  if ((rc = upload(ctx, sw, &E.seg1_w))) return rc;
  if ((rc = upload(ctx, sb, &E.seg1_b))) return rc;
  E.loaded = true;
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------------
static int seg_run(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid, int n,
                   uint8_t* classes, float* logp, float* sinc_out, cudaStream_t st) {
  B200_CHECK(ctx && ctx->seg.loaded, B200_ERR_STATE, "segmentation weights not loaded");
  B200_CHECK(wav && chunk_off && chunk_valid && n >= 0, B200_ERR_INVALID, "bad arguments");
  if (n == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  const int nbmax = n < ctx->seg_max_batch ? n : ctx->seg_max_batch;
  const size_t x0_bytes = align_up((size_t)nbmax * kFrames * 64 * sizeof(float), 1024);
  const size_t sinc_b = sincnet_workspace_bytes(nbmax), lstm_b = lstm_workspace_bytes(nbmax);
  const size_t big = sinc_b > lstm_b ? sinc_b : lstm_b;    // the two phases reuse the same region
  int rc = ensure_ws(ctx, x0_bytes + big + 4096);
  if (rc) return rc;
  if ((rc = push_meta(ctx, chunk_off, chunk_valid, n, st))) return rc;
  float* x0 = reinterpret_cast<float*>(ctx->ws);
  void* region = reinterpret_cast<char*>(ctx->ws) + x0_bytes;
  for (int c0 = 0; c0 < n; c0 += nbmax) {
    const int nb = (n - c0) < nbmax ? (n - c0) : nbmax;
    ScopedTimer timer(ctx, &ctx->seg_events, st);
    if (ctx->profile) ctx->seg_chunks += nb;
    float* x0_dst = sinc_out ? sinc_out + (size_t)c0 * kFrames * 64 : x0;
    if ((rc = sincnet_forward(ctx->seg, wav, ctx->d_off + c0, ctx->d_valid + c0, nb, region, x0_dst, ctx->seg_conv_impl,
                              ctx->num_sms, st)))
      return rc;
    ctx->launches += 8;
    if (sinc_out) continue;
    if ((rc = lstm_head_forward(ctx->seg, x0, nb, region, classes + (size_t)c0 * kFrames,
                                logp ? logp + (size_t)c0 * kFrames * kClasses : nullptr, ctx->num_sms,
                                ctx->seg_gemm_impl, ctx->seg_rec_impl, st)))
      return rc;
    ctx->launches += 2 * ctx->seg.lstm_layers + 3;
  }
  return B200_OK;
}

int b200_seg_forward(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                     int32_t num_chunks, uint8_t* classes, float* logp, void* stream) {
  if (num_chunks == 0) return B200_OK;
  B200_CHECK(classes != nullptr, B200_ERR_INVALID, "classes is NULL");
  return seg_run(ctx, wav, chunk_off, chunk_valid, num_chunks, classes, logp, nullptr, (cudaStream_t)stream);
}

__global__ void strip_pad_kernel(const float* __restrict__ x64, float* __restrict__ out, size_t rows) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 60) return;
  out[idx] = x64[(idx / 60) * 64 + idx % 60];
}

int b200_sincnet_forward(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                         int32_t num_chunks, float* out, void* stream) {
  B200_CHECK(out != nullptr, B200_ERR_INVALID, "out is NULL");
  if (num_chunks == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  float* tmp = nullptr;
  const size_t rows = (size_t)num_chunks * kFrames;
  B200_CUDA_OK(cudaMalloc((void**)&tmp, rows * 64 * sizeof(float)));
  int rc = seg_run(ctx, wav, chunk_off, chunk_valid, num_chunks, nullptr, nullptr, tmp, st);
  if (rc == B200_OK) {
    strip_pad_kernel<<<(unsigned)((rows * 60 + 255) / 256), 256, 0, st>>>(tmp, out, rows);
    ctx->launches += 1;
    cudaStreamSynchronize(st);
  }
  cudaFree(tmp);
  return rc;
}

int b200_powerset_to_multilabel(b200_ctx* ctx, const uint8_t* classes, int64_t n, uint8_t* multilabel, void* stream) {
  B200_CHECK(ctx && classes && multilabel && n >= 0, B200_ERR_INVALID, "bad arguments");
  if (n == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return powerset_to_multilabel(classes, n, multilabel, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------
struct EmbWs {
  float *fbank, *fmean, *stats;
  __half *A, *Bf, *Cf;
};
static size_t carve_emb(int NB, void* base, EmbWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 1024);
    void* p = base ? (char*)base + off : nullptr;
    off += bytes;
    return p;
  };
  EmbWs t;
  const size_t act = (size_t)NB * kMel * kFbankFrames * 32 * sizeof(__half);   // largest activation (layer1)
  t.fbank = (float*)take((size_t)NB * kFbankFrames * kMel * sizeof(float));
  t.fmean = (float*)take((size_t)NB * kMel * sizeof(float));
  t.stats = (float*)take((size_t)NB * kSpeakers * 2 * kStatsDim * sizeof(float));
  t.A = (__half*)take(act);
  t.Bf = (__half*)take(act);
  t.Cf = (__half*)take(act);
  if (w) *w = t;
  return align_up(off, 1024);
}

static int conv_flags(const b200_ctx* ctx) {
  return (ctx->conv_ghost ? kConvGhost : 0) | (ctx->conv_fold ? kConvFold : 0);
}

// one BasicBlock on nb segments: A -> (Bf, Cf) -> A, in place on the residual
static int block_run(b200_ctx* ctx, const BlockWeights& B, __half* A, __half* Bf, __half* Cf, int nb, int H, int Wd,
                     cudaStream_t st) {
  const int s = B.conv1.stride;
  const int impl1 = (ctx->conv_impl == 2) ? (s == 1 ? 1 : 0) : ctx->conv_impl;
  const int impl_s1 = ctx->conv_impl == 2 ? 1 : ctx->conv_impl;
  const int Ho = (H + 2 - 3) / s + 1, Wo = (Wd + 2 - 3) / s + 1;
  int rc;
  if (ctx->conv_impl == 8 && ctx->conv_scfold && B.has_shortcut && B.conv1.w3s) {
    // layer2.0: stride-2 conv1 and the 1x1 shortcut in one launch (the shortcut rides in the padded weight rows)
    if ((rc = conv_s2_shortcut_forward(B.conv1, A, Bf, Cf, nb, H, Wd, ctx->num_sms, st))) return rc;
    if ((rc = conv_forward(B.conv2, Bf, Cf, A, nb, Ho, Wo, 1, impl_s1, ctx->num_sms, st, conv_flags(ctx)))) return rc;
    ctx->launches += 2;
    return B200_OK;
  }
  if ((rc = conv_forward(B.conv1, A, nullptr, Bf, nb, H, Wd, 1, impl1, ctx->num_sms, st, conv_flags(ctx)))) return rc;
  const __half* res = A;
  if (B.has_shortcut) {
    if ((rc = conv_forward(B.shortcut, A, nullptr, Cf, nb, H, Wd, 0, impl1, ctx->num_sms, st, conv_flags(ctx)))) return rc;
    res = Cf;
    ctx->launches += 1;
  }
  if ((rc = conv_forward(B.conv2, Bf, res, A, nb, Ho, Wo, 1, impl_s1, ctx->num_sms, st, conv_flags(ctx)))) return rc;
  ctx->launches += 2;
  return B200_OK;
}

// conv1 + 16 BasicBlocks; returns the buffer holding the result (NHWC fp16 [nb][10][125][256])
static int trunk_run(b200_ctx* ctx, const EmbWs& w, const int* frame0, int nb, cudaStream_t st,
                     const __half** result) {
  const EmbWeights& E = ctx->emb;
  int rc;
  int H = kMel, Wd = kFbankFrames;
  // (running stem + layer1 in L2-sized groups of segments was measured 12-35 % slower than whole sub-batches:
  //  small grids lose more to tails and launch gaps than the L2 hits return)
  __half* cur = w.A;            // current activation; the other two buffers are scratch
  __half* s1 = w.Bf;
  __half* s2 = w.Cf;
  if ((rc = conv1_forward(w.fbank, w.fmean, frame0, E.conv1_w, E.conv1_b, cur, nb, st))) return rc;
  ctx->launches += 1;
  for (const BlockWeights& B : E.blocks) {
    const int s = B.conv1.stride;
    if (ctx->conv_impl == 8 && ctx->conv_fuse && !B.has_shortcut && s == 1 && B.conv1.C_in == 32 && B.conv1.w4 &&
        B.conv2.w4) {
      // layer1: the whole block in one kernel, intermediate activation kept in shared memory
      if ((rc = conv_block32_forward(B.conv1, B.conv2, cur, s1, nb, H, Wd, ctx->num_sms, st, ctx->conv_ghost))) return rc;
      ctx->launches += 1;
      __half* t = cur; cur = s1; s1 = t;
      continue;
    }
    if ((rc = block_run(ctx, B, cur, s1, s2, nb, H, Wd, st))) return rc;
    H = (H + 2 - 3) / s + 1; Wd = (Wd + 2 - 3) / s + 1;
  }
  B200_CHECK(H == 10 && Wd == kEmbT, B200_ERR_STATE, "unexpected trunk output %dx%d", H, Wd);
  *result = cur;
  return B200_OK;
}

static int emb_forward_impl(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                            int32_t num_chunks, const uint8_t* masks, float* emb, float* const* emb_peers,
                            int32_t n_peers, void* stream);

int b200_emb_forward(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                     int32_t num_chunks, const uint8_t* masks, float* emb, void* stream) {
  return emb_forward_impl(ctx, wav, chunk_off, chunk_valid, num_chunks, masks, emb, nullptr, 0, stream);
}

int b200_emb_forward_push(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                          int32_t num_chunks, const uint8_t* masks, float* emb, float* const* emb_peers,
                          int32_t n_peers, void* stream) {
  B200_CHECK(n_peers >= 0 && n_peers <= 7 && (n_peers == 0 || emb_peers), B200_ERR_INVALID, "bad peer list");
  return emb_forward_impl(ctx, wav, chunk_off, chunk_valid, num_chunks, masks, emb, emb_peers, n_peers, stream);
}

int b200_push(b200_ctx* ctx, const void* src, int64_t bytes, void* const* dsts, int32_t n_dsts, void* stream) {
  B200_CHECK(ctx && src && (n_dsts == 0 || dsts) && bytes >= 0 && n_dsts >= 0, B200_ERR_INVALID, "bad arguments");
  if (bytes == 0 || n_dsts == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return push_bytes(src, bytes, dsts, n_dsts, (cudaStream_t)stream);
}

static int emb_forward_impl(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                            int32_t num_chunks, const uint8_t* masks, float* emb, float* const* emb_peers,
                            int32_t n_peers, void* stream) {
  B200_CHECK(ctx && ctx->emb.loaded, B200_ERR_STATE, "embedding weights not loaded");
  B200_CHECK(wav && chunk_off && chunk_valid && masks && emb && num_chunks >= 0, B200_ERR_INVALID, "bad arguments");
  if (num_chunks == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int nbmax = num_chunks < ctx->emb_max_batch ? num_chunks : ctx->emb_max_batch;
  // pooled statistics of ALL chunks as fp16 (hi, lo) pairs -> one tensor-core GEMM for the Linear 5120 -> 256
  const size_t rows = (size_t)num_chunks * kSpeakers;
  const size_t split_bytes = align_up(rows * 2 * kStatsDim * sizeof(__half), 1024);
  const size_t sub_bytes = carve_emb(nbmax, nullptr, nullptr);
  int rc = ensure_ws(ctx, sub_bytes + 2 * split_bytes + 4096);
  if (rc) return rc;
  if ((rc = push_meta(ctx, chunk_off, chunk_valid, num_chunks, st))) return rc;
  FbankPlan plan;
  if ((rc = push_fbank_plan(ctx, chunk_off, chunk_valid, num_chunks, nbmax, ctx->fbank_share != 0, &plan, st))) return rc;
  EmbWs w;
  carve_emb(nbmax, ctx->ws, &w);
  __half* st_hi = reinterpret_cast<__half*>(reinterpret_cast<char*>(ctx->ws) + sub_bytes);
  __half* st_lo = reinterpret_cast<__half*>(reinterpret_cast<char*>(ctx->ws) + sub_bytes + split_bytes);
  const __half* feat = nullptr;
  for (int c0 = 0; c0 < num_chunks; c0 += nbmax) {
    const int nb = (num_chunks - c0) < nbmax ? (num_chunks - c0) : nbmax;
    const int sb = c0 / nbmax;
    if ((rc = fbank_forward(ctx->emb, wav, ctx->d_runs + plan.run_base[sb], plan.run_base[sb + 1] - plan.run_base[sb],
                            plan.nrows[sb], ctx->d_frame0 + c0, nb, w.fbank, w.fmean, st)))
      return rc;
    ctx->launches += 2;
    {
      ScopedTimer timer(ctx, &ctx->trunk_events, st);
      if ((rc = trunk_run(ctx, w, ctx->d_frame0 + c0, nb, st, &feat))) return rc;
    }
    if (ctx->profile) ctx->trunk_segments += nb;
    const size_t o = (size_t)c0 * kSpeakers * 2 * kStatsDim;
    if ((rc = stats_pool_forward(feat, masks + (size_t)c0 * kSpeakers * kFrames, nullptr, st_hi + o, st_lo + o, nb, st)))
      return rc;
    ctx->launches += 1;
  }
  // the Linear 5120 -> 256 of ALL chunks; with peers its epilogue also pushes every tile to the other GPUs (fused
  // all-gather of the embeddings over NVLink)
  rc = gemm_tc_split(st_hi, st_lo, 2 * kStatsDim, ctx->emb.seg1_w_hi, ctx->emb.seg1_w_lo, 2 * kStatsDim, emb, kEmbDim,
                     nullptr, nullptr, 0, ctx->emb.seg1_b, (int)rows, kEmbDim, 2 * kStatsDim, 0, ctx->num_sms, st,
                     emb_peers, n_peers);
  ctx->launches += 1;
  return rc;
}

int b200_emb_fbank(b200_ctx* ctx, const float* wav, const int64_t* chunk_off, const int32_t* chunk_valid,
                   int32_t num_chunks, float* fbank, void* stream) {
  B200_CHECK(ctx && ctx->emb.loaded, B200_ERR_STATE, "embedding weights not loaded");
  B200_CHECK(wav && chunk_off && chunk_valid && fbank && num_chunks >= 0, B200_ERR_INVALID, "bad arguments");
  if (num_chunks == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_ws(ctx, (size_t)num_chunks * kMel * sizeof(float) + 4096);
  if (rc) return rc;
  if ((rc = push_meta(ctx, chunk_off, chunk_valid, num_chunks, st))) return rc;
  FbankPlan plan;                                            // output layout [B][998][80]: one private run per chunk
  if ((rc = push_fbank_plan(ctx, chunk_off, chunk_valid, num_chunks, num_chunks, false, &plan, st))) return rc;
  float* fmean = reinterpret_cast<float*>(ctx->ws);
  if ((rc = fbank_forward(ctx->emb, wav, ctx->d_runs, num_chunks, plan.nrows[0], ctx->d_frame0, num_chunks, fbank, fmean,
                          st)))
    return rc;
  ctx->launches += 3;
  return fbank_center(fbank, fmean, num_chunks, st);
}

int64_t b200_emb_fbank_plan(const int64_t* chunk_off, const int32_t* chunk_valid, int32_t num_chunks, int32_t sub_batch,
                            int32_t share, int32_t* frame0, int32_t* rows_per_sub_batch) {
  if (!chunk_off || !chunk_valid || num_chunks < 0 || sub_batch < 1) return -1;
  for (int i = 0; i < num_chunks; ++i)
    if (chunk_valid[i] < 0 || chunk_valid[i] > kChunk || chunk_off[i] < 0) return -1;
  std::vector<b200::FbankRun> runs;
  std::vector<int> f0;
  FbankPlan plan;
  plan_fbank(chunk_off, chunk_valid, num_chunks, sub_batch, share != 0, &runs, &f0, &plan);
  if (frame0) std::copy(f0.begin(), f0.end(), frame0);
  if (rows_per_sub_batch) std::copy(plan.nrows.begin(), plan.nrows.end(), rows_per_sub_batch);
  return (int64_t)runs.size();
}

int b200_emb_trunk(b200_ctx* ctx, const float* fbank, int32_t num_chunks, float* frames, void* stream) {
  B200_CHECK(ctx && ctx->emb.loaded, B200_ERR_STATE, "embedding weights not loaded");
  B200_CHECK(fbank && frames && num_chunks >= 0, B200_ERR_INVALID, "bad arguments");
  if (num_chunks == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int nbmax = num_chunks < ctx->emb_max_batch ? num_chunks : ctx->emb_max_batch;
  int rc = ensure_ws(ctx, carve_emb(nbmax, nullptr, nullptr) + 4096);
  if (rc) return rc;
  EmbWs w;
  carve_emb(nbmax, ctx->ws, &w);
  for (int c0 = 0; c0 < num_chunks; c0 += nbmax) {
    const int nb = (num_chunks - c0) < nbmax ? (num_chunks - c0) : nbmax;
    B200_CUDA_OK(cudaMemcpyAsync(w.fbank, fbank + (size_t)c0 * kFbankFrames * kMel,
                                 (size_t)nb * kFbankFrames * kMel * sizeof(float), cudaMemcpyDeviceToDevice, st));
    B200_CUDA_OK(cudaMemsetAsync(w.fmean, 0, (size_t)nb * kMel * sizeof(float), st));
    const __half* feat = nullptr;
    if ((rc = trunk_run(ctx, w, nullptr, nb, st, &feat))) return rc;
    if ((rc = frames_to_nchw(feat, frames + (size_t)c0 * 256 * 10 * kEmbT, nb, st))) return rc;
    ctx->launches += 1;
  }
  return B200_OK;
}

int b200_stats_pool(b200_ctx* ctx, const float* seq, const float* weights, float* out, int32_t B, int32_t F, int32_t T,
                    int32_t S, int32_t Tw, void* stream) {
  B200_CHECK(ctx && seq && out && B >= 0 && F > 0 && T > 0 && S > 0, B200_ERR_INVALID, "bad arguments");
  if (B == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return stats_pool_generic(seq, weights, out, B, F, T, S, weights ? Tw : T, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------
int b200_speaker_count(b200_ctx* ctx, const uint8_t* seg, const int32_t* start_frame, int32_t num_chunks,
                       int32_t num_frames, uint8_t* count, void* stream) {
  B200_CHECK(ctx && seg && start_frame && count && num_chunks > 0 && num_frames > 0, B200_ERR_INVALID, "bad arguments");
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return speaker_count(seg, start_frame, num_chunks, num_frames, count, (cudaStream_t)stream);
}

int b200_reconstruct(b200_ctx* ctx, const uint8_t* seg, const int8_t* hard_clusters, const int32_t* start_frame,
                     int32_t num_chunks, int32_t num_frames, const uint8_t* count, int32_t num_clusters_out,
                     uint8_t* discrete, void* stream) {
  B200_CHECK(ctx && seg && hard_clusters && start_frame && count && discrete && num_chunks > 0 && num_frames > 0,
             B200_ERR_INVALID, "bad arguments");
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return reconstruct(seg, (const signed char*)hard_clusters, start_frame, num_chunks, num_frames, num_clusters_out,
                     count, discrete, (cudaStream_t)stream);
}

int b200_aggregate(b200_ctx* ctx, const float* scores, const int32_t* start_frame, int32_t num_chunks,
                   int32_t num_frames, int32_t num_classes, const double* hamming, const double* warm_up,
                   int32_t skip_average, float missing, float epsilon, float* out, void* stream) {
  B200_CHECK(ctx && scores && start_frame && out && num_chunks > 0 && num_frames > 0 && num_classes > 0,
             B200_ERR_INVALID, "bad arguments");
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return aggregate_scores(scores, start_frame, num_chunks, num_frames, num_classes, hamming, warm_up, skip_average,
                          missing, epsilon, out, (cudaStream_t)stream);
}

int b200_powerset_speech(b200_ctx* ctx, const uint8_t* classes, int64_t n, float* speech, void* stream) {
  B200_CHECK(ctx && classes && speech && n >= 0, B200_ERR_INVALID, "bad arguments");
  if (n == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return powerset_speech(classes, n, speech, (cudaStream_t)stream);
}

int b200_frame_transitions(b200_ctx* ctx, const uint8_t* discrete, int32_t num_frames, int32_t num_clusters,
                           int32_t cap, int32_t* buf, void* stream) {
  B200_CHECK(ctx && discrete && buf && num_frames > 0 && num_clusters > 0 && cap > 0, B200_ERR_INVALID,
             "bad arguments");
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return frame_transitions(discrete, num_frames, num_clusters, cap, buf, (cudaStream_t)stream);
}

int b200_clean_frames(b200_ctx* ctx, const uint8_t* seg, int32_t num_chunks, int32_t* clean, uint8_t* active,
                      void* stream) {
  B200_CHECK(ctx && seg && clean && active && num_chunks >= 0, B200_ERR_INVALID, "bad arguments");
  if (num_chunks == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return clean_frames(seg, num_chunks, clean, active, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------
int b200_linkage_centroid_batched(b200_ctx* ctx, const double* x, const int32_t* row_offsets, int32_t num_problems,
                                  int32_t dim, int32_t normalize, double* Z, void* stream) {
  B200_CHECK(ctx && x && row_offsets && Z && num_problems >= 1 && dim >= 1, B200_ERR_INVALID, "bad arguments");
  for (int f = 0; f < num_problems; ++f)
    B200_CHECK(row_offsets[f + 1] >= row_offsets[f] && row_offsets[f + 1] - row_offsets[f] <= 32768, B200_ERR_INVALID,
               "linkage: problem %d has %d observations (row offsets must be non-decreasing, at most 32768 per problem "
               "= about 3 h of audio at a 1 s step: cluster longer recordings in windows)", f,
               (int)(row_offsets[f + 1] - row_offsets[f]));
  DeviceGuard g(ctx->device);
  int rc = ensure_ws(ctx, linkage_workspace_bytes_batched(row_offsets, num_problems, dim));
  if (rc) return rc;
  ctx->launches += 2 + num_problems;
  return linkage_centroid_batched(x, row_offsets, num_problems, dim, normalize, Z, ctx->ws, (cudaStream_t)stream);
}

int b200_linkage_centroid(b200_ctx* ctx, const double* x, int32_t n, int32_t dim, int32_t normalize, double* Z,
                          void* stream) {
  B200_CHECK(n >= 2, B200_ERR_INVALID, "linkage needs at least 2 observations");
  const int32_t offs[2] = {0, n};
  return b200_linkage_centroid_batched(ctx, x, offs, 1, dim, normalize, Z, stream);
}

int b200_fcluster_distance(const double* Z, int32_t n, double t, int32_t* labels) {
  B200_CHECK(Z && labels && n >= 1, B200_ERR_INVALID, "bad arguments");
  return fcluster_distance(Z, n, t, labels);
}

int b200_plda_transform(b200_ctx* ctx, const double* x, int32_t n, int32_t Din, int32_t Dout, int32_t L,
                        const double* mean1, const double* mean2, const double* lda, const double* mu,
                        const double* trT, double* fea, void* stream) {
  B200_CHECK(ctx && x && mean1 && mean2 && lda && mu && trT && fea && n >= 0 && Din >= 1 && Dout >= 1 && L >= 1 &&
                 L <= Dout && Din + Dout <= 4096, B200_ERR_INVALID, "bad arguments");
  if (n == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return plda_transform(x, n, Din, Dout, L, mean1, mean2, lda, mu, trT, fea, (cudaStream_t)stream);
}

int b200_weighted_centroids(b200_ctx* ctx, const double* q, int32_t n, int32_t S, const int32_t* kept, int32_t K,
                            const double* train, int32_t dim, double* centroids, void* stream) {
  B200_CHECK(ctx && q && kept && train && centroids && n >= 1 && S >= 1 && K >= 0 && dim >= 1, B200_ERR_INVALID,
             "bad arguments");
  if (K == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return weighted_centroids(q, n, S, kept, K, train, dim, centroids, (cudaStream_t)stream);
}

int b200_cdist_cosine(b200_ctx* ctx, const double* a, int32_t m, const double* b, int32_t k, int32_t dim, double* d,
                      void* stream) {
  B200_CHECK(ctx && a && b && d && m >= 0 && k >= 1 && dim >= 1, B200_ERR_INVALID, "bad arguments");
  if (m == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return cdist_cosine(a, m, b, k, dim, d, (cudaStream_t)stream);
}

int b200_vbx_batched(b200_ctx* ctx, const double* fea, const double* phi, const int32_t* n, const int32_t* S,
                     int32_t num_problems, int32_t D, double Fa, double Fb, int32_t max_iters, double epsilon,
                     double* gamma, double* pi, int32_t* iters, void* stream) {
  B200_CHECK(ctx && fea && phi && n && S && gamma && pi && num_problems >= 1 && D >= 1 && max_iters >= 1,
             B200_ERR_INVALID, "bad arguments");
  for (int f = 0; f < num_problems; ++f) B200_CHECK(n[f] >= 0 && S[f] >= 0, B200_ERR_INVALID, "negative problem size");
  DeviceGuard g(ctx->device);
  int rc = ensure_ws(ctx, vbx_workspace_bytes_batched(n, S, num_problems, D));
  if (rc) return rc;
  ctx->launches += 1;
  return vbx_run_batched(fea, phi, n, S, num_problems, D, Fa, Fb, max_iters, epsilon, gamma, pi, iters, ctx->ws,
                         (cudaStream_t)stream);
}

int b200_vbx(b200_ctx* ctx, const double* fea, const double* phi, int32_t n, int32_t D, int32_t S, double Fa,
             double Fb, int32_t max_iters, double epsilon, double* gamma, double* pi, int32_t* iters, void* stream) {
  B200_CHECK(n >= 1 && S >= 1, B200_ERR_INVALID, "bad arguments");
  return b200_vbx_batched(ctx, fea, phi, &n, &S, 1, D, Fa, Fb, max_iters, epsilon, gamma, pi, iters, stream);
}

int b200_assign(b200_ctx* ctx, const double* soft, int32_t num_chunks, int32_t num_clusters, int32_t constrained,
                int8_t* hard, void* stream) {
  B200_CHECK(ctx && soft && hard && num_chunks >= 0 && num_clusters >= 1, B200_ERR_INVALID, "bad arguments");
  if (num_chunks == 0) return B200_OK;
  DeviceGuard g(ctx->device);
  ctx->launches += 1;
  return assign_clusters(soft, num_chunks, num_clusters, constrained, (signed char*)hard, (cudaStream_t)stream);
}

}  // extern "C"
