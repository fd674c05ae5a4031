// Embedding path: kaldi-compatible log-mel fbank and masked statistics pooling.
//
// fbank follows torchaudio.compliance.kaldi.fbank as called by the reference
//   (/root/reference/src/pyannote/audio/models/embedding/wespeaker/__init__.py:113-139):
//   x*32768 -> frames of 400 @ hop 160 (snip_edges) -> remove DC -> pre-emphasis 0.97 (replicate pad) ->
//   Hamming -> zero-pad to 512 -> |rFFT|^2 -> 80 triangular mel bins (20 Hz..Nyquist) -> log(max(., eps)).
// The global-mean centring over frames (:137-139) is produced as a separate [B][80] vector that the first
// conv subtracts on load.
//
// stats pooling follows models/blocks/pooling.py:30-61,76-130 via resnet.py:61-66 (TSTP): nearest
// interpolation of the 589-frame mask onto the 125 trunk frames, weighted mean and weighted unbiased std.
#include "common.cuh"
#include "emb.cuh"

namespace b200 {

constexpr int kFrameLen = 400;
constexpr int kFrameHop = 160;
constexpr int kFft = 512;
constexpr float kEps = 1.1920928955078125e-07f;

__device__ __forceinline__ int bitrev9(int x) { return __brev((unsigned)x) >> 23; }

// smem index with one pad word per 16 elements: the register-blocked FFT passes below read 16 consecutive or
// 16-strided elements per lane, both conflict-free with this padding
__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }
constexpr int kFftPad = kFft + kFft / 16;   // 544

// 4 radix-2 DIT stages on 16 values held in registers; tw(s, j) returns the twiddle of the butterfly whose upper
// input is local element j in local stage s
template <typename TW>
__device__ __forceinline__ void fft16_stages(float (&xr)[16], float (&xi)[16], TW tw) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int half = 1 << s;
#pragma unroll
    for (int bf = 0; bf < 8; ++bf) {
      const int pos = bf & (half - 1);
      const int j0 = ((bf >> s) << (s + 1)) + pos, j1 = j0 + half;
      float wr, wi;
      tw(s, pos, wr, wi);
      const float tr = wr * xr[j1] - wi * xi[j1];
      const float ti = wr * xi[j1] + wi * xr[j1];
      const float ur = xr[j0], ui = xi[j0];
      xr[j0] = ur + tr; xi[j0] = ui + ti;
      xr[j1] = ur - tr; xi[j1] = ui - ti;
    }
  }
}

// Round 2: one warp computes TWO frames.  A 512-point real FFT is a 256-point complex FFT of z[n] = x[2n] + i x[2n+1]
// followed by   X[k] = E[k] + W_512^k O[k],  E = (Z[k] + conj Z[256-k]) / 2,  O = (Z[k] - conj Z[256-k]) / 2i,
// and stages 0-7 of the register-blocked 512-point DIT schedule below already ARE two independent 256-point FFTs
// on the two halves of the array (stage 8 was the only one that mixed them): frame A lives in elements [0, 256),
// frame B in [256, 512), half a warp each.  Same arithmetic per butterfly, less than half of it per frame
// (the padded imaginary half of the old complex transform was all zeros): 663 -> see profiles/ us per 256 segments.
__global__ void __launch_bounds__(256) fbank_kernel(const float* __restrict__ wav,
                                                    const FbankRun* __restrict__ runs, int nruns, int nrows,
                                                    const float* __restrict__ window,
                                                    const float* __restrict__ twiddle, const float* __restrict__ mel_w,
                                                    const int* __restrict__ mel_start, const int* __restrict__ mel_len,
                                                    const int* __restrict__ mel_off, float* __restrict__ out) {
  __shared__ float s_re[8][kFftPad];
  __shared__ float s_im[8][kFftPad];
  __shared__ float s_tw[256][2];
  __shared__ long long s_src[16];                           // first sample of each of this block's 16 frame rows
  __shared__ int s_lim[16];                                 // samples of the run still valid from there
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_tw[0][0])[i] = twiddle[i];
  if (threadIdx.x < 16) {
    // row -> run: last run whose first row is <= row (runs are sorted by row0, rows of a run are hop-spaced frames)
    int row = blockIdx.x * 16 + threadIdx.x;
    if (row >= nrows) row = nrows - 1;
    int lo = 0, hi = nruns - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (runs[mid].row0 <= row) lo = mid; else hi = mid - 1;
    }
    const FbankRun r = runs[lo];
    const int local = row - r.row0;
    s_src[threadIdx.x] = r.src + (long long)local * kFrameHop;
    const long long left = (long long)r.limit - (long long)local * kFrameHop;
    s_lim[threadIdx.x] = left < 0 ? 0 : (left > kFrameLen ? kFrameLen : (int)left);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = lane & 15, f = lane >> 4;                  // half-warp f owns frame row 2 * pair + f
  const int pair = blockIdx.x * 8 + warp;
  if (2 * pair >= nrows) return;                           // warp-uniform
  const int row = 2 * pair + f;                            // row == nrows (odd tail): computed on a clamped source, not stored
  float* re = s_re[warp];
  float* im = s_im[warp];
  const float* xw = wav + s_src[2 * warp + f];
  const int valid = s_lim[2 * warp + f];
  const int h0 = f << 8;                                   // this frame's half of the arrays

  // load + scale, frame mean (sample i = l + 16 j)
  float x[25];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 25; ++j) {
    const int g = l + 16 * j;                              // sample of the frame; xw already points at the frame
    const float v = (g < valid) ? xw[g] * 32768.0f : 0.f;
    x[j] = v;
    sum += v;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)kFrameLen;
  // DC removal, pre-emphasis (previous sample from the neighbouring lane; sample 0 replicates itself), window;
  // z[n] = v[2n] + i v[2n+1] scattered to bit-reversed order: even lanes write real parts, odd lanes imaginary parts
  float* dst = (l & 1) ? im : re;
  float carry = 0.f;                                       // lane 15's sample of the previous j (for lane 0)
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int i = l + 16 * j;
    float v = 0.f;
    if (j < 25) {
      const float cur = x[j] - mean;
      float prev = __shfl_up_sync(0xffffffffu, cur, 1, 16);
      if (l == 0) prev = (j == 0) ? cur : carry;
      carry = __shfl_sync(0xffffffffu, cur, 15, 16);
      v = (cur - 0.97f * prev) * window[i];
    }
    dst[fpad(h0 + (int)(__brev((unsigned)(i >> 1)) >> 24))] = v;
  }
  __syncwarp();
  // two 256-point radix-2 DIT FFTs (one per half-warp) as 4 + 4 stages: two register-blocked passes of 16 values
  float xr[16], xi[16];
  {  // stages 0-3: lane owns elements 16 lane .. 16 lane + 15 (lanes 16-31: the second frame's half)
#pragma unroll
    for (int j = 0; j < 16; ++j) { xr[j] = re[fpad(16 * lane + j)]; xi[j] = im[fpad(16 * lane + j)]; }
    fft16_stages(xr, xi, [&](int s, int pos, float& wr, float& wi) {
      const int k = pos * (256 >> s);                      // multiples of 32: compile-time after unrolling
      wr = s_tw[k][0];
      wi = s_tw[k][1];
    });
#pragma unroll
    for (int j = 0; j < 16; ++j) { re[fpad(16 * lane + j)] = xr[j]; im[fpad(16 * lane + j)] = xi[j]; }
  }
  __syncwarp();
  {  // stages 4-7: lane owns elements e0 + 16 j of its frame's half
    const int e0 = l + h0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { xr[j] = re[fpad(e0 + 16 * j)]; xi[j] = im[fpad(e0 + 16 * j)]; }
    fft16_stages(xr, xi, [&](int s, int pos, float& wr, float& wi) {
      // global stage 4 + s, position inside the butterfly group = l + 16 pos
      const int k = (l + 16 * pos) * (16 >> s);
      wr = s_tw[k][0];
      wi = s_tw[k][1];
    });
#pragma unroll
    for (int j = 0; j < 16; ++j) { re[fpad(e0 + 16 * j)] = xr[j]; im[fpad(e0 + 16 * j)] = xi[j]; }
  }
  __syncwarp();
  // real-FFT split + power spectrum of bins 0..255 (the Nyquist bin carries zero mel weight: kaldi pads the bank with
  // a zero column); this lane's bins are k = l + 16 j
  float pw[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int k = l + 16 * j, km = (256 - k) & 255;
    const float zr = re[fpad(h0 + k)], zi = im[fpad(h0 + k)];
    const float mr = re[fpad(h0 + km)], mi = im[fpad(h0 + km)];
    const float er = 0.5f * (zr + mr), ei = 0.5f * (zi - mi);
    const float orr = 0.5f * (zi + mi), oi = -0.5f * (zr - mr);
    const float wr = s_tw[k][0], wi = s_tw[k][1];
    const float xr_ = er + (wr * orr - wi * oi), xi_ = ei + (wr * oi + wi * orr);
    const float a = sqrtf(xr_ * xr_ + xi_ * xi_);            // reference: rfft().abs().pow(2)
    pw[j] = a * a;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j) re[h0 + l + 16 * j] = pw[j];
  __syncwarp();
  for (int m = l; m < kMel; m += 16) {
    const int st = mel_start[m], ln = mel_len[m], off = mel_off[m];
    float acc = 0.f;
    for (int i = 0; i < ln; ++i) acc = fmaf(re[h0 + st + i], mel_w[off + i], acc);
    if (row < nrows) out[(size_t)row * kMel + m] = logf(fmaxf(acc, kEps));
  }
}

__global__ void __launch_bounds__(640) fbank_mean_kernel(const float* __restrict__ fb, const int* __restrict__ frame0,
                                                         float* __restrict__ fmean) {
  // 8 groups of 80 threads each sum an eighth of the frames in fp64, combined in group order
  __shared__ double part[8][kMel];
  const int b = blockIdx.x, m = threadIdx.x % kMel, g = threadIdx.x / kMel;
  const size_t r0 = frame0 ? (size_t)frame0[b] : (size_t)b * kFbankFrames;
  double s = 0.0;
  for (int t = g; t < kFbankFrames; t += 8) s += (double)fb[(r0 + t) * kMel + m];
  part[g][m] = s;
  __syncthreads();
  if (g == 0) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[k][m];
    fmean[b * kMel + m] = (float)(tot / kFbankFrames);
  }
}

__global__ void fbank_center_kernel(float* __restrict__ fb, const float* __restrict__ fmean, size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int m = idx % kMel;
  const int b = idx / ((size_t)kMel * kFbankFrames);
  fb[idx] -= fmean[b * kMel + m];
}

int fbank_center(float* fbank, const float* fmean, int B, cudaStream_t stream) {
  const size_t total = (size_t)B * kFbankFrames * kMel;
  fbank_center_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(fbank, fmean, total);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

__global__ void frames_to_nchw_kernel(const __half* __restrict__ feat, float* __restrict__ out, size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over NCHW output
  if (idx >= total) return;
  const int t = idx % kEmbT;
  const int h = (idx / kEmbT) % 10;
  const int c = (idx / (kEmbT * 10)) % 256;
  const size_t b = idx / ((size_t)kEmbT * 10 * 256);
  out[idx] = __half2float(feat[((b * 10 + h) * kEmbT + t) * 256 + c]);
}

int frames_to_nchw(const __half* feat, float* out, int B, cudaStream_t stream) {
  const size_t total = (size_t)B * 256 * 10 * kEmbT;
  frames_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(feat, out, total);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

int fbank_forward(const EmbWeights& W, const float* wav, const FbankRun* runs, int nruns, int nrows,
                  const int* frame0, int B, float* fbank, float* fmean, cudaStream_t stream) {
  const unsigned grid = (unsigned)ceil_div(nrows, 16);      // 8 warps x 2 frame rows per block
  fbank_kernel<<<grid, 256, 0, stream>>>(wav, runs, nruns, nrows, W.window, W.twiddle, W.mel_w, W.mel_start,
                                         W.mel_len, W.mel_off, fbank);
  B200_CUDA_OK(cudaGetLastError());
  fbank_mean_kernel<<<B, 640, 0, stream>>>(fbank, frame0, fmean);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------
// masked stats pooling on the trunk output, all 3 local speakers from one pass
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_pool_kernel(const __half* __restrict__ feat,
                                                         const unsigned char* __restrict__ masks,
                                                         float* __restrict__ stats, __half* __restrict__ stats_hi,
                                                         __half* __restrict__ stats_lo) {
  // grid (10 freq rows, B); thread = channel c; feat[b][h][t][c]
  __shared__ float s_w[kSpeakers][kEmbT];
  const int h = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
  for (int i = threadIdx.x; i < kSpeakers * kEmbT; i += blockDim.x) {
    const int s = i / kEmbT, t = i % kEmbT;
    // F.interpolate(mode="nearest"): src = floor(dst * in / out)   (pooling.py:116-117)
    const int src = (int)(((long long)t * kFrames) / kEmbT);
    s_w[s][t] = (float)masks[((size_t)b * kSpeakers + s) * kFrames + src];
  }
  __syncthreads();
  const __half* fp = feat + (((size_t)b * 10 + h) * kEmbT) * 256 + c;
  float v1[kSpeakers], v2[kSpeakers], sx[kSpeakers];
#pragma unroll
  for (int s = 0; s < kSpeakers; ++s) { v1[s] = 0.f; v2[s] = 0.f; sx[s] = 0.f; }
  for (int t = 0; t < kEmbT; ++t) {
    const float x = __half2float(fp[(size_t)t * 256]);
#pragma unroll
    for (int s = 0; s < kSpeakers; ++s) {
      const float w = s_w[s][t];
      v1[s] += w;
      v2[s] += w * w;
      sx[s] += x * w;
    }
  }
  float mean[kSpeakers], sd[kSpeakers];
#pragma unroll
  for (int s = 0; s < kSpeakers; ++s) {
    v1[s] += 1e-8f;
    mean[s] = sx[s] / v1[s];
    sd[s] = 0.f;
  }
  for (int t = 0; t < kEmbT; ++t) {
    const float x = __half2float(fp[(size_t)t * 256]);
#pragma unroll
    for (int s = 0; s < kSpeakers; ++s) {
      const float d = x - mean[s];
      sd[s] += d * d * s_w[s][t];
    }
  }
#pragma unroll
  for (int s = 0; s < kSpeakers; ++s) {
    const float var = sd[s] / (v1[s] - v2[s] / v1[s] + 1e-8f);
    const size_t row = ((size_t)b * kSpeakers + s) * (2 * kStatsDim);
    const float sdv = sqrtf(var);
    if (stats) {
      stats[row + c * 10 + h] = mean[s];
      stats[row + kStatsDim + c * 10 + h] = sdv;
    }
    if (stats_hi) {   // (hi, lo) fp16 split consumed by gemm_tc_split (Linear 5120 -> 256)
      const __half mh = __float2half_rn(mean[s]), sh = __float2half_rn(sdv);
      stats_hi[row + c * 10 + h] = mh;
      stats_lo[row + c * 10 + h] = __float2half_rn(mean[s] - __half2float(mh));
      stats_hi[row + kStatsDim + c * 10 + h] = sh;
      stats_lo[row + kStatsDim + c * 10 + h] = __float2half_rn(sdv - __half2float(sh));
    }
  }
}

int stats_pool_forward(const __half* feat, const unsigned char* masks, float* stats, __half* stats_hi,
                       __half* stats_lo, int B, cudaStream_t stream) {
  dim3 grid(10, B);
  stats_pool_kernel<<<grid, 256, 0, stream>>>(feat, masks, stats, stats_hi, stats_lo);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

// generic fp32 version (any F, T, S, Tw) used by the known-answer tests of the reference
__global__ void stats_pool_generic_kernel(const float* __restrict__ seq, const float* __restrict__ w,
                                          float* __restrict__ out, int B, int F, int T, int S, int Tw) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S * F) return;
  const int f = idx % F, s = (idx / F) % S, b = idx / (F * S);
  const float* x = seq + ((size_t)b * F + f) * T;
  float mean, sd;
  if (w == nullptr) {
    float sum = 0.f;
    for (int t = 0; t < T; ++t) sum += x[t];
    mean = sum / T;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += (x[t] - mean) * (x[t] - mean);
    sd = sqrtf(acc / (T - 1));
  } else {
    const float* ww = w + ((size_t)b * S + s) * Tw;
    float v1 = 0.f, v2 = 0.f, sx = 0.f;
    for (int t = 0; t < T; ++t) {
      const float wt = ww[(int)(((long long)t * Tw) / T)];
      v1 += wt; v2 += wt * wt; sx += x[t] * wt;
    }
    v1 += 1e-8f;
    mean = sx / v1;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const float wt = ww[(int)(((long long)t * Tw) / T)];
      acc += (x[t] - mean) * (x[t] - mean) * wt;
    }
    sd = sqrtf(acc / (v1 - v2 / v1 + 1e-8f));
  }
  float* o = out + ((size_t)b * S + s) * 2 * F;
  o[f] = mean;
  o[F + f] = sd;
}

int stats_pool_generic(const float* seq, const float* w, float* out, int B, int F, int T, int S, int Tw,
                       cudaStream_t stream) {
  const int total = B * S * F;
  stats_pool_generic_kernel<<<ceil_div(total, 128), 128, 0, stream>>>(seq, w, out, B, F, T, S, Tw);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
