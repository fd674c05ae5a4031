// Embedding path (kaldi fbank -> ResNet34 trunk -> masked stats pooling -> Linear) declarations.
#pragma once
#include "common.cuh"

namespace b200 {

struct ConvLayer {
  int C_in = 0, C_out = 0, ksize = 3, stride = 1;
  __half* w = nullptr;     // device, [tap][C_out][C_in] fp16, BN scale folded
  __half* w4 = nullptr;    // 32->32 / 64->64 stride-1 layers: [kw][(kh, c_out) = 3C][c_in] for conv_tc4_kernel
  __half* w3 = nullptr;    // same with C_out zero-padded to a multiple of 128 (conv_tc3_kernel's A operand)
  float* bias = nullptr;   // device, [C_out] fp32 (folded BN shift)
  // conv_tc3_kernel only: a stride-2 3x3 conv with C_out = 64 carries the block's 1x1 stride-2 shortcut in the 64
  // otherwise zero-padded rows of its 128-row weight tiles (shortcut weights at the centre tap): one launch, two outputs
  __half* w3s = nullptr;   // [9][128][C_in]: rows 0..63 conv, rows 64..127 shortcut (centre tap only)
  float* bias_s = nullptr; // [128]: conv bias | shortcut bias
};

struct ConvParams {
  int B, H_in, W_in, C_in;
  int H_out, W_out, C_out;
  int taps_h, taps_w, stride, pad;
  int Ck, kblocks, tiles_w, num_tiles, relu;
  const float* bias;
  const __half* residual;
  __half* out;
  uint32_t a_bytes, b_bytes, nstages, idesc, swizzle;
};

struct BlockWeights {
  ConvLayer conv1, conv2, shortcut;
  bool has_shortcut = false;
};

struct EmbWeights {
  bool loaded = false;
  float* conv1_w = nullptr;      // [32][9] folded
  float* conv1_b = nullptr;      // [32]
  std::vector<BlockWeights> blocks;   // 16 BasicBlocks
  float* seg1_w = nullptr;       // [256][5120] fp32 (PyTorch layout)
  float* seg1_b = nullptr;       // [256]
  __half* seg1_w_hi = nullptr;   // fp16 (hi, lo) split of seg1_w for the tensor-core GEMM
  __half* seg1_w_lo = nullptr;
  // fbank constants
  float* window = nullptr;       // [400] hamming
  float* mel_w = nullptr;        // packed non-zero mel weights
  int* mel_start = nullptr;      // [80] first fft bin
  int* mel_len = nullptr;        // [80] number of bins
  int* mel_off = nullptr;        // [80] offset into mel_w
  float* twiddle = nullptr;      // [256][2] cos/sin(-2 pi k / 512)
};

// flags of conv_forward / conv_block32_forward
constexpr int kConvGhost = 1;   // TMEM rings with ghost blocks (conv_tc4 / conv_block32)
constexpr int kConvFold = 2;    // conv_tc3: horizontal taps as descriptor shifts of one pixel box per (kh, channel block)
// impl: 0 = SIMT reference conv, 1 = tcgen05 tensor-core conv
int conv_forward(const ConvLayer& L, const __half* in, const __half* residual, __half* out, int B, int H_in, int W_in,
                 int relu, int impl, int num_sms, cudaStream_t stream, int flags = kConvGhost | kConvFold);
// stride-2 3x3 conv (C_out = 64) + the block's 1x1 stride-2 shortcut in ONE conv_tc3_kernel launch (L.w3s):
// out = relu(conv(in) + b), out_sc = shortcut(in) + b_sc
int conv_s2_shortcut_forward(const ConvLayer& L, const __half* in, __half* out, __half* out_sc, int B, int H_in,
                             int W_in, int num_sms, cudaStream_t stream);
// fused BasicBlock of layer1 (two 32->32 stride-1 convs + identity shortcut), out must not alias in
int conv_block32_forward(const ConvLayer& L1, const ConvLayer& L2, const __half* in, __half* out, int B, int H, int W,
                         int num_sms, cudaStream_t stream, int ghost = 1);
// frame0 (device, [B], may be NULL = b * 998): first fbank row of each segment, see fbank_forward
int conv1_forward(const float* fbank, const float* fmean, const int* frame0, const float* w, const float* bias,
                  __half* out, int B, cudaStream_t stream);

// fbank with SHARED FRAMES.  The sliding chunks of a file overlap by 90 % and a chunk step of 16000 samples is exactly
// 100 frame hops, so frame k of chunk c IS frame k - 100 of chunk c + 1: the same 400 samples through the same
// arithmetic.  The host groups a sub-batch's chunks into runs of hop-aligned, overlapping full chunks; a run's frames
// are computed once into consecutive rows of `fbank` ([nrows][80] fp32) and segment b reads rows
// frame0[b] .. frame0[b] + 997 (conv1_forward, the per-segment mean).  Chunks that are short (valid < 160000: samples
// past `limit` read as zero) or not hop-aligned get a private run, which is also the layout of the public
// b200_emb_fbank ([B][998][80]).  Bit-identical to one private run per chunk; 10x fewer frames on a pipeline batch.
struct FbankRun {
  long long src;   // sample offset of the run's first frame in `wav`
  int row0;        // first row of the run in `fbank`
  int limit;       // valid samples counted from src (INT_MAX for runs of full chunks)
};
int fbank_forward(const EmbWeights& W, const float* wav, const FbankRun* runs, int nruns, int nrows,
                  const int* frame0, int B, float* fbank, float* fmean, cudaStream_t stream);
int fbank_center(float* fbank, const float* fmean, int B, cudaStream_t stream);
// NHWC fp16 [B][10][125][256] -> NCHW fp32 [B][256][10][125]
int frames_to_nchw(const __half* feat, float* out, int B, cudaStream_t stream);

// masked statistics pooling: feat [B][10][125][256] fp16 NHWC, masks [B][3][589] u8 -> stats [B*3][5120] fp32
int stats_pool_forward(const __half* feat, const unsigned char* masks, float* stats, __half* stats_hi,
                       __half* stats_lo, int B, cudaStream_t stream);
// generic weighted pooling used by the known-answer tests: seq [B][F][T] fp32, w [B][S][Tw] fp32 -> [B][S][2F]
int stats_pool_generic(const float* seq, const float* w, float* out, int B, int F, int T, int S, int Tw,
                       cudaStream_t stream);

}  // namespace b200
