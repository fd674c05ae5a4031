// SincNet first layer on the tensor cores: 80 sinc band-pass FIRs (K = 251, stride 10) + abs + MaxPool1d(3).
//
// Reference: /root/reference/src/pyannote/audio/models/blocks/sincnet.py:70-90,163-176 (ParamSincFB 80 x 251, stride
// 10; torch.abs on the first layer; MaxPool1d(3, 3)).  Split-precision fp16 like gemm_tc.cu (fp32-level accuracy).
//
//     D[f][t] = sum_k filt[f][k] * xn[10 t + k]          M = 128 (80 real rows), N = 240 positions, K = 256 (251 real)
//
// The B operand is an im2col of the waveform whose rows start 20 bytes apart, which TMA cannot express (16-byte
// stride granularity), so the CTA builds it: warp 0 -- eight builder warps stage the normalised samples of a tile as fp16 (hi, lo) in shared
// memory (raw values of the next tile prefetched into registers) and (one row per thread) copy 16-sample rows into the swizzled K-major layout tcgen05.mma expects (st.shared +
// fence.proxy.async, the technique of conv_block32_kernel), one k-step (K = 16) per ring stage; the filter tiles of
// that k-step (hi, lo, 8 KB) arrive in the same stage by TMA (8 stages: with the 128 KB bank resident only 4 fitted
// and ncu showed builders and MMA warp waiting on each other, tensor pipe 38 % active).  Epilogue: thread = filter, positions are TMEM columns: abs, max over 3 adjacent
// columns, store, InstanceNorm partial sums in fp64.  48 MMAs per 240 positions.
// The fp32 CUDA-core kernel sinc_pool_kernel (seg_sincnet.cu) stays as the A/B reference (seg_conv_impl = 0).
#include "common.cuh"
#include "seg.cuh"
#include "tc_common.cuh"

namespace b200 {

constexpr int kSTThreads = 448;            // warp 0 samples (+ filter TMA), warp 1 MMA, warps 2-9 builders, 10-13 epilogue
constexpr int kSTPool = 80;                // pooled outputs per tile
constexpr int kSTN = 3 * kSTPool;          // 240 conv outputs per tile = N
constexpr int kSTSamples = 2656;           // 10 * 239 + 256 = 2646 samples per tile, rounded to 83 x 32
constexpr uint32_t kSTXsOff = 1024;                          // [hi | lo][2656] fp16 (10624 B)
constexpr uint32_t kSTEpOff = 1024 + 10624 + 128;            // epilogue staging [80 filters][33] fp32 (10560 B)
constexpr uint32_t kSTBOff = 23552;                          // ring of k-step stages
constexpr uint32_t kSTStage = 24576;                         // B hi 8 KB | B lo 8 KB | A hi 4 KB | A lo 4 KB
constexpr uint32_t kSTStages = 8;                            // deep ring: the build -> MMA -> commit round trip is ~1k cycles
constexpr uint32_t kSTSmem = kSTBOff + kSTStages * kSTStage; // 220160

struct SincTcParams {
  const float* wav;
  const long long* chunk_off;
  const int* chunk_valid;
  const float2* affine;     // per-chunk waveform InstanceNorm affine
  float* P0;                // [NB][80][5325]
  double2* part;            // [NB][80][ntiles_part]
  int NB, tiles, num_items, ntiles_part;
  int early_raw;            // A/B knob B200_SINC_EARLY=1: request the next tile's samples before the k-step loop (round 1)
};

__global__ void __launch_bounds__(kSTThreads, 1)
sinc_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, SincTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t bar_bfull = base + 32, bar_bempty = base + 96, bar_afull = base + 160;
  const uint32_t bar_tfull = base + 224, bar_tempty = base + 240;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + 256);
  __half* xs = reinterpret_cast<__half*>(gbase + kSTXsOff);   // [buf][hl][kSTSamples]
  const uint32_t b_smem = base + kSTBOff;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < kSTStages; ++i) {
      mbar_init(bar_bfull + 8 * i, 8); mbar_init(bar_bempty + 8 * i, 1); mbar_init(bar_afull + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 4); }
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
    // ---- filter tiles of every k-step into the ring stage the builders fill -------------------------------------
    const bool leader = elect_one_sync();
    uint32_t cnt = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      for (int ks = 0; ks < 16; ++ks, ++cnt) {
        const uint32_t stage = cnt & (kSTStages - 1);
        mbar_wait(bar_bempty + 8 * stage, ((cnt / kSTStages) & 1u) ^ 1u);
        if (leader) {
          const uint32_t st = b_smem + stage * kSTStage;
          mbar_expect_tx(bar_afull + 8 * stage, 8192u);
          tma_load_3d(&tmAh, bar_afull + 8 * stage, st + 16384u, 0, 0, ks);
          tma_load_3d(&tmAl, bar_afull + 8 * stage, st + 20480u, 0, 0, ks);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------------------------
    const bool leader = elect_one_sync();
    const uint32_t dhi = desc_hi(256u, 6u);                 // 32-byte rows, SWIZZLE_32B
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kSTN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t cnt = 0, acc = 0, acc_phase = 0;               // cnt = global k-step counter -> ring stage
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;
      for (uint32_t ks = 0; ks < 16; ++ks, ++cnt) {
        const uint32_t stage = cnt & (kSTStages - 1), ph = (cnt / kSTStages) & 1u;
        mbar_wait(bar_afull + 8 * stage, ph);
        mbar_wait(bar_bfull + 8 * stage, ph);
        tc_fence_after();
        if (leader) {
          const uint32_t st = b_smem + stage * kSTStage;
          const uint32_t bh = desc_lo(st), bl = desc_lo(st + 8192u);
          const uint32_t ah = desc_lo(st + 16384u), al = desc_lo(st + 20480u);
          tc_mma_f16(d_tmem, desc_from(dhi, al), desc_from(dhi, bh), idesc, ks != 0u);
          tc_mma_f16(d_tmem, desc_from(dhi, ah), desc_from(dhi, bl), idesc, 1);
          tc_mma_f16(d_tmem, desc_from(dhi, ah), desc_from(dhi, bh), idesc, 1);
          tc_commit(bar_bempty + 8 * stage);
        }
        __syncwarp();
      }
      if (leader) tc_commit(bar_tfull + 8 * acc);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp < 10) {
    // ---- builders: im2col rows (16 samples = 32 bytes, 20 bytes apart in the stream) into the swizzled B tiles -----
    const int tb = (warp - 2) * 32 + lane;                  // 0..255: row tb (rows >= 240 do not exist)
    // The 256 builder threads also stage the tile's samples: raw values of the NEXT tile are prefetched into registers
    // while this tile is built (a single staging warp was latency-bound: 15 us per tile against 3 us of MMAs).
    constexpr int kPer = (kSTSamples + 255) / 256;          // 11 samples per thread
    float rawv[kPer];
    auto load_raw = [&](int it_) {
      const int tile = it_ % p.tiles, b = it_ / p.tiles;
      const float* x = p.wav + p.chunk_off[b];
      const int valid = p.chunk_valid[b];
      const int s0 = tile * (10 * kSTN);
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tb + 256 * j, g = s0 + i;
        rawv[j] = (i < kSTSamples && g < valid) ? __ldg(x + g) : 0.f;
      }
    };
    __half* xh = xs;
    __half* xl = xs + kSTSamples;
    const uint32_t xs_h = base + kSTXsOff;                  // byte address of the hi samples
    uint32_t cnt = 0;
    if ((int)blockIdx.x < p.num_items) load_raw(blockIdx.x);
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int tile = item % p.tiles, b = item / p.tiles;
      const float2 af = p.affine[b];
      const int s0 = tile * (10 * kSTN);
      asm volatile("bar.sync 2, 256;" ::: "memory");        // everyone has finished reading the previous tile's samples
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tb + 256 * j;
        if (i < kSTSamples) {
          const float v = (s0 + i < kChunk) ? fmaf(rawv[j], af.x, af.y) : 0.f;   // same padding rule as sinc_pool_kernel
          const __half h = __float2half_rn(v);
          xh[i] = h;
          xl[i] = __float2half_rn(v - __half2float(h));
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");        // samples complete
      if (p.early_raw && item + (int)gridDim.x < p.num_items) load_raw(item + (int)gridDim.x);
      for (uint32_t ks = 0; ks < 16; ++ks, ++cnt) {
        const uint32_t stage = cnt & (kSTStages - 1);
        mbar_wait(bar_bempty + 8 * stage, ((cnt / kSTStages) & 1u) ^ 1u);
        {
          const int r = tb;
          if (r < kSTN) {
#pragma unroll
            for (uint32_t hl = 0; hl < 2; ++hl) {
              const uint32_t src = xs_h + hl * (kSTSamples * 2u) + (uint32_t)(10 * r + 16 * (int)ks) * 2u;   // 4-byte aligned
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[j]) : "r"(src + 4u * j));
              const uint32_t row = b_smem + stage * kSTStage + hl * 8192u + (uint32_t)r * 32u;
              const uint32_t sw = (row >> 7) & 1u;          // SWIZZLE_32B: 16-byte chunk index ^ address bit 7
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((0u ^ sw) << 4)), "r"(w[0]),
                           "r"(w[1]), "r"(w[2]), "r"(w[3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((1u ^ sw) << 4)), "r"(w[4]),
                           "r"(w[5]), "r"(w[6]), "r"(w[7])
                           : "memory");
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tcgen05.mma reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_bfull + 8 * stage);
      }
      // The next tile's raw samples are requested only now: the proxy fence above is a MEMBAR that waits for every
      // outstanding memory operation of the warp, so with the global loads issued before the k-step loop (round 1) the
      // first fence of every tile sat out their DRAM latency (~1.5 us against 3.2 us of MMAs per tile: the "unexplained"
      // 2x over the MMA floor).  Here the latency falls into the slack of the 8-stage ring instead.
      if (!p.early_raw && item + (int)gridDim.x < p.num_items) load_raw(item + (int)gridDim.x);
    }
  } else {
    // ---- epilogue: abs, MaxPool1d(3), store, InstanceNorm partial sums -------------------------------------------
    const int q = warp & 3;
    const int f = q * 32 + lane;                            // filter = TMEM lane (rows >= 80 are padding)
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int tile = item % p.tiles, b = item / p.tiles;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      if (q < 3) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u;
        float* s_o = reinterpret_cast<float*>(gbase + kSTEpOff);       // [80][33]; a warp only touches its own rows
        const int nrows = q < 2 ? 32 : 16;                           // filters 64..79 in the third warp
        double s = 0.0, ss = 0.0;
        for (int batch = 0; batch < 3; ++batch) {           // 96 + 96 + 48 columns -> 32 + 32 + 16 pooled values
          uint32_t r[96];
          tc_ld32(taddr + batch * 96, r);
          tc_ld32(taddr + batch * 96 + 32, r + 32);          // (last batch: columns 240..255 are never written, unused)
          if (batch < 2) tc_ld32(taddr + batch * 96 + 64, r + 64);
          const int npool = batch < 2 ? 32 : 16;
          const int p0 = tile * kSTPool + batch * 32;
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < npool) {
              const float v = fmaxf(fmaxf(fabsf(__uint_as_float(r[3 * i])), fabsf(__uint_as_float(r[3 * i + 1]))),
                                    fabsf(__uint_as_float(r[3 * i + 2])));
              if (f < 80) {
                s_o[f * 33 + i] = v;
                if (p0 + i < kPool0) { s4[i & 3] += v; q4[i & 3] = fmaf(v, v, q4[i & 3]); }
              }
            }
          }
          // fp32 partial sums over <= 8 non-negative values each, folded into fp64 once per batch (a serial chain of
          // 160 dependent fp64 adds per tile kept the epilogue warps -- the bottleneck of this kernel -- busy)
          s += (double)((s4[0] + s4[1]) + (s4[2] + s4[3]));
          ss += (double)((q4[0] + q4[1]) + (q4[2] + q4[3]));
          __syncwarp();
          // transposed store: one filter row per instruction, lanes along the positions (coalesced 128-byte rows
          // instead of 32 scattered 4-byte stores)
          if (lane < npool && p0 + lane < kPool0) {
            for (int j = 0; j < nrows; ++j) {
              const int fr = q * 32 + j;
              p.P0[((size_t)b * 80 + fr) * kPool0 + p0 + lane] = s_o[fr * 33 + lane];
            }
          }
          __syncwarp();
        }
        if (f < 80) p.part[((size_t)b * 80 + f) * p.ntiles_part + tile] = make_double2(s, ss);
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

static int make_a_map(CUtensorMap* tm, const __half* ptr) {
  PFN_encodeTiled enc = get_encode();
  B200_CHECK(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {16, 128, 16};                       // [k-step][filter row][16 taps]
  cuuint64_t strides[2] = {32, 128 * 32};
  cuuint32_t box[3] = {16, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(sinc) failed: %d", (int)r);
  return B200_OK;
}

// Ah/Al: filter bank [16 k-steps][128 rows (80 real)][16 taps] fp16 (hi, lo); tap k = 16 ks + kk, taps >= 251 zero
int sinc_tc_forward(const float* wav, const long long* chunk_off, const int* chunk_valid, const float2* affine,
                    const __half* Ah, const __half* Al, int NB, float* P0, double2* part, int ntiles_part, int num_sms,
                    cudaStream_t stream) {
  SincTcParams p{};
  p.wav = wav; p.chunk_off = chunk_off; p.chunk_valid = chunk_valid; p.affine = affine; p.P0 = P0; p.part = part;
  p.NB = NB; p.ntiles_part = ntiles_part;
  p.tiles = ceil_div(kPool0, kSTPool);
  B200_CHECK(p.tiles <= ntiles_part, B200_ERR_STATE, "sinc_tc: partial-sum buffer too small");
  p.num_items = NB * p.tiles;
  { const char* e = getenv("B200_SINC_EARLY"); p.early_raw = e ? atoi(e) : 0; }
  CUtensorMap tmAh, tmAl;
  int rc;
  if ((rc = make_a_map(&tmAh, Ah))) return rc;
  if ((rc = make_a_map(&tmAl, Al))) return rc;
  const size_t smem = 1024 + kSTSmem;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(sinc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = p.num_items < num_sms ? p.num_items : num_sms;
  sinc_tc_kernel<<<grid, kSTThreads, smem, stream>>>(tmAh, tmAl, p);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
