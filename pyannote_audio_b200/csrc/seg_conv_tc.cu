// SincNet Conv1d(k=5) + MaxPool1d(3) layers on the tensor cores (split-precision fp16, fp32-level accuracy).
//
// Reference: /root/reference/src/pyannote/audio/models/blocks/sincnet.py:163-184 (conv1d -> pool1d -> norm1d ->
// leaky_relu for layers 2 and 3: Conv1d(80, 60, 5) and Conv1d(60, 60, 5), MaxPool1d(3, stride 3)).
//
// Channels-as-M implicit GEMM, one tile = 240 conv outputs (80 pooled) of one chunk:
//     D[co][t] = sum_{tap, ci} W[co][ci][tap] * X[t + tap][ci]            M = 128 (60 real rows), N = 240, K = 16
//   * B operand: the normalised + leaky-relu'd input, channels-last [b][pos][ci] as fp16 (hi, lo) pairs written by
//     in_apply_split_kernel; one TMA box of 244 positions x 16 channels (boxes are limited to 256 rows) per channel block (SWIZZLE_32B), the five
//     taps are descriptor shifts of one 32-byte pixel row inside that box;
//   * A operand: weight tiles [128][16] per (channel block, tap), streamed with the B tile (hi and lo);
//   * three MMAs per product (lo*hi, hi*lo, hi*hi) into one fp32 TMEM accumulator, 2 accumulator stages;
//   * epilogue: thread = output channel, positions are TMEM columns, so MaxPool1d(3) is a max over three adjacent
//     registers; + bias, store the pooled row, per-(chunk, channel, tile) InstanceNorm partial sums in fp64.
// The fp32 CUDA-core kernel conv5_pool_kernel (seg_sincnet.cu) stays as the A/B reference (seg_conv_impl = 0).
#include "common.cuh"
#include "seg.cuh"
#include "tc_common.cuh"

namespace b200 {

constexpr int kC5Threads = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kC5Pooled = 80;              // pooled outputs per tile (240 conv outputs = N)
constexpr int kC5Rows = 3 * kC5Pooled + 4;  // input positions per tile (TMA box rows, <= 256)
constexpr uint32_t kC5BTile = 8192;        // 244 positions x 32 B (7808) rounded to 1 KB
constexpr uint32_t kC5ATile = 4096;        // 128 rows x 32 B
constexpr uint32_t kC5Stage = 2 * kC5BTile + 10 * kC5ATile;   // B hi | B lo | A hi x5 taps | A lo x5 taps = 57344
constexpr int kC5Stages = 3;

// ---- InstanceNorm affine + leaky_relu + transpose to channels-last fp16 (hi, lo) -------------------------------
template <int C>
__global__ void __launch_bounds__(256) in_apply_split_kernel(const float* __restrict__ P, const float2* __restrict__ affine,
                                                            int Cpad, int L, __half* __restrict__ Xh,
                                                            __half* __restrict__ Xl) {
  __shared__ float tile[80][65];
  const int b = blockIdx.y, l0 = blockIdx.x * 64;
  {
    // thread = (position l, channels cb, cb + 4, ...): all loads of a thread are issued before the first use (the
    // rolled loop kept 1-2 loads in flight per thread and the kernel ran at 40 % of the HBM rate)
    const int l = threadIdx.x & 63, cb = threadIdx.x >> 6;
    constexpr int NJ = C / 4;
    float v[NJ];
    const bool ok = l0 + l < L;
    const float* src = P + ((size_t)b * C + cb) * L + l0 + l;
#pragma unroll
    for (int j = 0; j < NJ; ++j) v[j] = ok ? __ldg(src + (size_t)(4 * j) * L) : 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float2 af = affine[b * C + cb + 4 * j];
      float y = fmaf(v[j], af.x, af.y);
      y = y > 0.f ? y : 0.01f * y;
      tile[cb + 4 * j][l] = ok ? y : 0.f;
    }
  }
  __syncthreads();
  // 8 channels per thread: one 16-byte store each for hi and lo, consecutive threads -> consecutive chunks
  const int groups = Cpad / 8;
  for (int i = threadIdx.x; i < 64 * groups; i += 256) {
    const int l = i / groups, c0 = (i - l * groups) * 8;
    if (l0 + l >= L) continue;
    uint4 uh, ul;
    __half2* ph = reinterpret_cast<__half2*>(&uh);
    __half2* pl = reinterpret_cast<__half2*>(&ul);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = c0 + 2 * e < C ? tile[c0 + 2 * e][l] : 0.f;
      const float v1 = c0 + 2 * e + 1 < C ? tile[c0 + 2 * e + 1][l] : 0.f;
      const __half2 h = __floats2half2_rn(v0, v1);
      const float2 hb = __half22float2(h);
      ph[e] = h;
      pl[e] = __floats2half2_rn(v0 - hb.x, v1 - hb.y);
    }
    const size_t o = ((size_t)b * L + l0 + l) * Cpad + c0;
    *reinterpret_cast<uint4*>(Xh + o) = uh;
    *reinterpret_cast<uint4*>(Xl + o) = ul;
  }
}

struct Conv5TcParams {
  int NB, L_pool, ncb, tiles, num_items, ntiles_part;
  const float* bias;      // [60]
  float* Pout;            // [NB][60][L_pool]
  double2* part;          // [NB][60][ntiles_part]
};

__global__ void __launch_bounds__(kC5Threads, 1)
conv5_tc_kernel(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl, Conv5TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_full = base, bar_empty = base + 32, bar_tfull = base + 64, bar_tempty = base + 80;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + 96);
  float* s_bias = reinterpret_cast<float*>(gbase + 256);
  const uint32_t stage0 = base + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x < 64) s_bias[threadIdx.x] = threadIdx.x < 60 ? p.bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kC5Stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one_sync();
    uint32_t stage = 0, phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int tile = item % p.tiles, b = item / p.tiles;
      for (int cb = 0; cb < p.ncb; ++cb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (leader) {
          const uint32_t st = stage0 + stage * kC5Stage;
          mbar_expect_tx(bar_full + 8 * stage, 2u * (uint32_t)kC5Rows * 32u + 10u * kC5ATile);
          tma_load_3d(&tmXh, bar_full + 8 * stage, st, cb * 16, tile * (3 * kC5Pooled), b);
          tma_load_3d(&tmXl, bar_full + 8 * stage, st + kC5BTile, cb * 16, tile * (3 * kC5Pooled), b);
          for (int tap = 0; tap < 5; ++tap) {
            tma_load_3d(&tmWh, bar_full + 8 * stage, st + 2 * kC5BTile + tap * kC5ATile, 0, 0, cb * 5 + tap);
            tma_load_3d(&tmWl, bar_full + 8 * stage, st + 2 * kC5BTile + (5 + tap) * kC5ATile, 0, 0, cb * 5 + tap);
          }
        }
        __syncwarp();
        if (++stage == kC5Stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one_sync();
    const uint32_t dhi = desc_hi(256u, 6u);                 // 32-byte rows, SWIZZLE_32B, 8-row groups of 256 B
    const uint32_t idesc = (1u << 4) | ((uint32_t)((3 * kC5Pooled) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;
      for (int cb = 0; cb < p.ncb; ++cb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (leader) {
          const uint32_t st = stage0 + stage * kC5Stage;
          const uint32_t bh = desc_lo(st), bl = desc_lo(st + kC5BTile);
#pragma unroll
          for (uint32_t tap = 0; tap < 5; ++tap) {
            const uint32_t ah = desc_lo(st + 2 * kC5BTile + tap * kC5ATile);
            const uint32_t al = desc_lo(st + 2 * kC5BTile + (5 + tap) * kC5ATile);
            // tap = the same box, `tap` positions (32-byte rows) further; small cross terms first
            tc_mma_f16(d_tmem, desc_from(dhi, al), desc_from(dhi, bh + 2 * tap), idesc, (cb | (int)tap) != 0);
            tc_mma_f16(d_tmem, desc_from(dhi, ah), desc_from(dhi, bl + 2 * tap), idesc, 1);
            tc_mma_f16(d_tmem, desc_from(dhi, ah), desc_from(dhi, bh + 2 * tap), idesc, 1);
          }
          tc_commit(bar_empty + 8 * stage);
        }
        __syncwarp();
        if (++stage == kC5Stages) { stage = 0; phase ^= 1; }
      }
      if (leader) tc_commit(bar_tfull + 8 * acc);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    const int q = warp & 3;
    const int co = q * 32 + lane;                            // output channel = TMEM lane (rows >= 60 are padding)
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int tile = item % p.tiles, b = item / p.tiles;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      if (q < 2) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u;
        const float bias = s_bias[co & 63];
        float* s_o = reinterpret_cast<float*>(gbase + 1024 + (size_t)kC5Stages * kC5Stage);   // [64][33]
        const int nrows = q == 0 ? 32 : 28;                 // channels 32..59 in the second warp
        double s = 0.0, ss = 0.0;
        for (int batch = 0; batch < 3; ++batch) {           // 96 + 96 + 48 columns -> 32 + 32 + 16 pooled values
          uint32_t r[96];
          tc_ld32(taddr + batch * 96, r);
          tc_ld32(taddr + batch * 96 + 32, r + 32);          // (last batch: columns 240..255 are never written, unused)
          if (batch < 2) tc_ld32(taddr + batch * 96 + 64, r + 64);
          const int npool = batch < 2 ? 32 : 16;
          const int p0 = tile * kC5Pooled + batch * 32;
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < npool) {
              const float v = fmaxf(fmaxf(__uint_as_float(r[3 * i]), __uint_as_float(r[3 * i + 1])),
                                    __uint_as_float(r[3 * i + 2])) + bias;
              if (co < 60) {
                s_o[co * 33 + i] = v;
                if (p0 + i < p.L_pool) { s4[i & 3] += v; q4[i & 3] = fmaf(v, v, q4[i & 3]); }
              }
            }
          }
          // fp32 partial sums over <= 8 values each, folded into fp64 once per batch (no serial fp64 chain)
          s += (double)((s4[0] + s4[1]) + (s4[2] + s4[3]));
          ss += (double)((q4[0] + q4[1]) + (q4[2] + q4[3]));
          __syncwarp();
          // transposed store: one channel row per instruction, lanes along the positions (coalesced)
          if (lane < npool && p0 + lane < p.L_pool) {
            for (int j = 0; j < nrows; ++j) {
              const int cr = q * 32 + j;
              p.Pout[((size_t)b * 60 + cr) * p.L_pool + p0 + lane] = s_o[cr * 33 + lane];
            }
          }
          __syncwarp();
        }
        if (co < 60) p.part[((size_t)b * 60 + co) * p.ntiles_part + tile] = make_double2(s, ss);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

static int make_map3(CUtensorMap* tm, const __half* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1) {
  PFN_encodeTiled enc = get_encode();
  B200_CHECK(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(conv5) failed: %d", (int)r);
  return B200_OK;
}

int in_apply_split(const float* P, const float2* affine, int NB, int C, int Cpad, int L, __half* Xh, __half* Xl,
                   cudaStream_t stream) {
  B200_CHECK(C == 80 || C == 60, B200_ERR_INVALID, "in_apply_split: %d channels unsupported", C);
  if (C == 80) in_apply_split_kernel<80><<<dim3(ceil_div(L, 64), NB), 256, 0, stream>>>(P, affine, Cpad, L, Xh, Xl);
  else in_apply_split_kernel<60><<<dim3(ceil_div(L, 64), NB), 256, 0, stream>>>(P, affine, Cpad, L, Xh, Xl);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

// Xh/Xl [NB][L_in][Cpad] (Cpad = 16 * ncb), Wh/Wl [ncb * 5][128][16] -> Pout [NB][60][L_pool], part [NB][60][ntiles_part]
int conv5_tc_forward(const __half* Xh, const __half* Xl, const __half* Wh, const __half* Wl, const float* bias, int NB,
                     int L_in, int L_pool, int ncb, float* Pout, double2* part, int ntiles_part, int num_sms,
                     cudaStream_t stream) {
  Conv5TcParams p{};
  p.NB = NB; p.L_pool = L_pool; p.ncb = ncb; p.bias = bias; p.Pout = Pout; p.part = part; p.ntiles_part = ntiles_part;
  p.tiles = ceil_div(L_pool, kC5Pooled);
  B200_CHECK(p.tiles <= ntiles_part, B200_ERR_STATE, "conv5_tc: partial-sum buffer too small");
  p.num_items = NB * p.tiles;
  CUtensorMap tmXh, tmXl, tmWh, tmWl;
  int rc;
  const int Cpad = 16 * ncb;
  if ((rc = make_map3(&tmXh, Xh, Cpad, L_in, NB, 16, kC5Rows))) return rc;
  if ((rc = make_map3(&tmXl, Xl, Cpad, L_in, NB, 16, kC5Rows))) return rc;
  if ((rc = make_map3(&tmWh, Wh, 16, 128, (uint64_t)ncb * 5, 16, 128))) return rc;
  if ((rc = make_map3(&tmWl, Wl, 16, 128, (uint64_t)ncb * 5, 16, 128))) return rc;
  const size_t smem = 1024 + 1024 + (size_t)kC5Stages * kC5Stage + 64 * 33 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(conv5_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = p.num_items < num_sms ? p.num_items : num_sms;
  conv5_tc_kernel<<<grid, kC5Threads, smem, stream>>>(tmXh, tmXl, tmWh, tmWl, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
