// LSTM recurrence on the 5th-gen tensor cores (split-precision fp16, fp32-level accuracy).
//
// Reference: nn.LSTM inside PyanNet (/root/reference/src/pyannote/audio/models/segmentation/PyanNet.py:98-110,226-228);
// per step   gates = Gx[t] + h_{t-1} W_hh^T,   c = f*c + i*g,   h = o*tanh(c)   (gate order i, f, g, o).
//
// A 2-CTA cluster owns a tile of 128 sequences of one direction for all 589 steps.  CTA `rank` owns hidden units
// [64 rank, 64 rank + 64), i.e. 256 gate columns ordered (unit, gate):
//   * B operand: its W_hh slice as fp16 (hi, lo) pairs, 2 x 64 KB, loaded once by TMA and resident in shared memory;
//   * A operand: h_{t-1} of all 128 units for the 128 sequences as fp16 (hi, lo), 2 x 32 KB, K-major SWIZZLE_128B,
//     written every step by the epilogue warps of BOTH CTAs (own half locally, the other half through DSMEM);
//   * D: 128 sequences x 256 gate columns fp32 in TMEM, TWO accumulators (steps alternate);  24 MMAs per step
//     (8 k-steps x {lo*hi, hi*lo, hi*hi});
//   * epilogue (8 warps, thread = sequence x 32 cells, four chunks of 8 cells): tcgen05.ld, + Gx, gates, cell state in
//     registers, h -> (hi, lo) -> both CTAs' A tiles + the layer output;
//   * round 2: the K order of a CTA's 64 units is (chunk, half, unit-in-chunk), so the 16 units that the eight epilogue
//     warps finish with chunk c are exactly k-step c of that CTA's k-block: h_ready is FOUR barriers (one per chunk)
//     and the MMA warp issues the 6 MMAs of k-step c of step t+1 into the other accumulator as soon as chunk c of
//     step t is written by both CTAs -- only the last 6 of the 24 MMAs are exposed (3216 -> ~800 cycles per step);
//   * round 2 gate math (scripts/micro/gate_bench.cu): the MUFU pipe is the bound of the epilogue; the forget gate
//     shares the reciprocal of the input/candidate pair (7 instead of 8 MUFU per cell) and NP of the 5 exponentials
//     run on the FMA pipe (magic-number floor, degree-6 polynomial, exponent add).
// Gx comes from gemm_tc_split in "gx layout" [b/32][t][col/4][b%32][4] so that the 32 lanes of an epilogue warp
// (32 sequences) read 512 contiguous bytes per float4 column group.
//
// Per-step handshakes (all mbarriers live in the consumer's shared memory):
//   acc_full  : tcgen05.commit of this CTA's MMAs                                   -> epilogue warps
//   peer_done : the PEER's MMAs of this step have completed (remote arrive)         -> safe to overwrite its A tile
//   h_ready[c]: 8 local + 8 remote epilogue warps have written chunk c of h_t (release.cluster) -> MMA warp, k-step c of step t+1
#include "common.cuh"
#include "seg.cuh"
#include "tc_common.cuh"
#include <cstdlib>

namespace b200 {

constexpr int kRecThreads = 320;          // warp 0: TMA (weights) + MMA issue, warp 1: TMEM owner, warps 2-9: epilogue
constexpr uint32_t kRecABytes = 65536;    // h (hi | lo) x 2 k-blocks of [128 rows][64 k] fp16
constexpr uint32_t kRecPhaseBytes = 16384; // h bytes landing in one CTA per chunk: 2 CTAs x 256 threads x (hi + lo) x 16 B
constexpr uint32_t kRecWBytes = 131072;   // W (hi | lo) x 2 k-blocks of [256 rows][64 k] fp16
constexpr int kRecDefaultNP = 0;          // exponentials per cell on the FMA pipe (see lstm_cell)

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Spin without acquire semantics (an acquiring try_wait invalidates L1 on every poll); `acquire` adds ONE cluster-scope
// fence after the spin.  The MMA warp (no memory operations in flight) acquires the h tiles written by the peer; the
// epilogue warps only need the control dependency of peer_done (their fence would also wait for the Gx prefetch).
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, bool acquire) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
  if (acquire) asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
// Relaxed spin, then ONE acquiring test of the completed phase: pairs with the cluster-scope release of the st.async
// transactions without the MEMBAR.GPU + ERRBAR of a separate fence (the MMA warp spent 80 % of its time in fences).
__device__ __forceinline__ void mbar_wait_cluster_acq(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t map_to_peer(uint32_t saddr, uint32_t peer) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(peer));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
// Asynchronous 16-byte store into the shared memory of a CTA of the cluster (own or peer) that completes 16 bytes of
// the transaction count of an mbarrier in the SAME CTA when the data has landed: the writer needs no fence and no
// arrive.  (With st.shared + fence.proxy.async + mbarrier.arrive.release per chunk the epilogue warps spent 52 % of
// their samples in MEMBAR / FENCE.VIEW.ASYNC, which wait for every outstanding global load / store of the warp.)
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t raddr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// generic-proxy shared-memory writes (own CTA and peer, through DSMEM) -> async-proxy reads by tcgen05.mma.  The
// unqualified fence.proxy.async also drains global memory (MEMBAR.ALL + ERRBAR: 14 % of all stall samples in ncu).
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 2^(s x) on the FMA pipe for s x in [-126, 40] (the caller clamps x): round(s x) by the magic-number add, degree-6
// polynomial on [-0.5, 0.5] (max relative error 1.0e-7, the same order as ex2.approx), exponent added as an integer.
__device__ __forceinline__ float poly_ex2_scaled(float x, float s) {
  const float magic = 12582912.f;                           // 1.5 * 2^23
  const float t = fmaf(x, s, magic);
  const float n = t - magic;
  const float f = fmaf(x, s, -n);
  float q = 0.00015461444854736328f;
  q = fmaf(q, f, 0.0013400427997112274f);
  q = fmaf(q, f, 0.009618056938052177f);
  q = fmaf(q, f, 0.05550327152013779f);
  q = fmaf(q, f, 0.24022650718688965f);
  q = fmaf(q, f, 0.6931471824645996f);
  q = fmaf(q, f, 1.0f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}

// One LSTM cell.  NP = number of exponentials evaluated on the FMA pipe (order g, i, f, o, c); NP < 0 = the round-1
// formulation (5 ex2 + 3 rcp).  exp(-x) is clamped to 2^40 so that the product of three denominators stays finite.
template <int NP>
__device__ __forceinline__ float lstm_cell(float xi, float xf, float xg, float xo, float& c) {
  constexpr float kL = 1.4426950408889634f;                 // log2(e)
  if (NP < 0) {
    constexpr float kLim = 57.0f;
    const float ei = fast_ex2(fminf(-kL * xi, kLim));
    const float ef = fast_ex2(fminf(-kL * xf, kLim));
    const float eg = fast_ex2(fminf(-2.f * kL * xg, kLim));
    const float eo = fast_ex2(fminf(-kL * xo, kLim));
    const float ig = (1.f - eg) * fast_rcp((1.f + ei) * (1.f + eg));          // sigmoid(i) * tanh(g)
    const float fg = fast_rcp(1.f + ef);
    const float cn = fmaf(fg, c, ig);
    c = cn;
    const float ec = fast_ex2(fminf(-2.f * kL * cn, kLim));
    return (1.f - ec) * fast_rcp((1.f + eo) * (1.f + ec));                    // sigmoid(o) * tanh(c)
  } else {
    constexpr float kLim = 40.0f;
    constexpr float x1 = kLim / kL, x2 = kLim / (2.f * kL);
    const float eg = NP >= 1 ? poly_ex2_scaled(fminf(fmaxf(xg, -x2), 43.f), -2.f * kL) : fast_ex2(fminf(-2.f * kL * xg, kLim));
    const float ei = NP >= 2 ? poly_ex2_scaled(fminf(fmaxf(xi, -x1), 87.f), -kL) : fast_ex2(fminf(-kL * xi, kLim));
    const float ef = NP >= 3 ? poly_ex2_scaled(fminf(fmaxf(xf, -x1), 87.f), -kL) : fast_ex2(fminf(-kL * xf, kLim));
    const float eo = NP >= 4 ? poly_ex2_scaled(fminf(fmaxf(xo, -x1), 87.f), -kL) : fast_ex2(fminf(-kL * xo, kLim));
    const float d1 = (1.f + ei) * (1.f + eg), d2 = 1.f + ef;
    const float r = fast_rcp(d1 * d2);                      // one reciprocal for sigmoid(i) tanh(g) and sigmoid(f)
    const float ig = (1.f - eg) * d2 * r;
    const float fg = d1 * r;
    const float cn = fmaf(fg, c, ig);
    c = cn;
    const float ec = NP >= 5 ? poly_ex2_scaled(fminf(fmaxf(cn, -x2), 43.f), -2.f * kL) : fast_ex2(fminf(-2.f * kL * cn, kLim));
    return (1.f - ec) * fast_rcp((1.f + eo) * (1.f + ec));
  }
}

struct RecTcParams {
  const float* G;      // gx layout, see gemm_tc_split_gx
  __half* Yh;          // [NB][589][256] fp16 hi
  __half* Yl;          //                      lo
  int NB, ntiles, T;
  int pf;              // Gx L2 prefetch distance in steps (0 = off)
};

template <int NP, bool ASYNC>
// (registers are allocated for 12 warps when 10 are launched: 168 per thread is the limit, __maxnreg__(200) fails to launch)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kRecThreads, 1)
lstm_rec_tc_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl, RecTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  // header: [0] w  [8,24) acc[2]  [24] peer  [32,64) h[4]  [64] tmem slot
  const uint32_t bar_w = base, bar_acc = base + 8, bar_peer = base + 24, bar_h = base + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + 64);
  const uint32_t a_smem = base + 1024, w_smem = a_smem + kRecABytes;

  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1u;
  const int cid = blockIdx.x >> 1;
  const int dir = cid / p.ntiles, tile = cid - dir * p.ntiles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.T;

  {  // h_{-1} = 0
    uint4* a4 = reinterpret_cast<uint4*>(gbase + 1024);
    for (uint32_t i = threadIdx.x; i < kRecABytes / 16; i += kRecThreads) a4[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc + 8, 1);
    mbar_init(bar_peer, 1);
    // ASYNC: one arming arrive (MMA warp) + 16 KB of st.async transactions per phase; else 16 warp arrivals
    for (int c = 0; c < 4; ++c) {
      mbar_init(bar_h + 8 * c, ASYNC ? 1 : 16);
      if (ASYNC) mbar_expect_tx(bar_h + 8 * c, kRecPhaseBytes);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // peer barriers initialised, both A tiles zeroed
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    const bool leader = elect_one_sync();
    if (leader) {
      const int row0 = (dir * 2 + (int)rank) * 256;
      mbar_expect_tx(bar_w, kRecWBytes);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(&tmWh, bar_w, w_smem + kb * 32768u, kb * 64, row0);
        tma_load_2d(&tmWl, bar_w, w_smem + 65536u + kb * 32768u, kb * 64, row0);
      }
    }
    __syncwarp();
    mbar_wait(bar_w, 0);
    const uint32_t dhi = desc_hi(1024u, 2u);                // 128-byte rows, SWIZZLE_128B
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int s = 0; s < T; ++s) {
      const uint32_t d_tmem = tmem_base + (uint32_t)(s & 1) * 256u;
#pragma unroll
      for (uint32_t c = 0; c < 4; ++c) {                    // k-step c of both k-blocks = chunk c of h_{s-1}
        if (s > 0) {
          if (ASYNC) mbar_wait_cluster_acq(bar_h + 8 * c, (uint32_t)(s - 1) & 1u);
          else mbar_wait_cluster(bar_h + 8 * c, (uint32_t)(s - 1) & 1u, true);
          if (ASYNC && leader) mbar_expect_tx(bar_h + 8 * c, kRecPhaseBytes);   // arm the phase of step s
        }
        // the writers' generic-proxy stores -> tcgen05.mma reads (ASYNC: this CTA's shared memory only; the .cluster
        // form adds a MEMBAR.GPU)
        if (ASYNC) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        else fence_proxy_async();
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (uint32_t kb = 0; kb < 2; ++kb) {
            const uint32_t ah = desc_lo(a_smem + kb * 16384u) + 2 * c, al = desc_lo(a_smem + 32768u + kb * 16384u) + 2 * c;
            const uint32_t bh = desc_lo(w_smem + kb * 32768u) + 2 * c, bl = desc_lo(w_smem + 65536u + kb * 32768u) + 2 * c;
            // small cross terms first, the dominant hi*hi term last
            tc_mma_f16(d_tmem, desc_from(dhi, al), desc_from(dhi, bh), idesc, (c | kb) != 0u);
            tc_mma_f16(d_tmem, desc_from(dhi, ah), desc_from(dhi, bl), idesc, 1);
            tc_mma_f16(d_tmem, desc_from(dhi, ah), desc_from(dhi, bh), idesc, 1);
          }
          if (c == 3) tc_commit(bar_acc + 8 * (uint32_t)(s & 1));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;                            // sequence row of the tile = TMEM lane
    const int b = tile * 128 + m;
    const bool live = b < p.NB;
    // Gx: [b/32][t][col/4][b%32] float4; this thread's columns start at dir*512 + rank*256 + half*128
    const size_t g_row = (size_t)(tile * 4 + q) * T;
    const int col4_0 = (dir * 512 + (int)rank * 256 + half * 128) >> 2;
    const float4* G4 = reinterpret_cast<const float4*>(p.G);
    auto load_gx = [&](int t, int chunk, float4 (&dst)[8]) {
      const float4* src = G4 + ((g_row + (size_t)t) * 256 + (size_t)(col4_0 + chunk * 8)) * 32 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = __ldg(src + (size_t)j * 32);
    };
    const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
    const uint32_t a_hi_row = a_smem + rank * 16384u + (uint32_t)m * 128u;   // k-block == rank (64 units per CTA)
    const uint32_t a_lo_row = a_hi_row + 32768u;
    const uint32_t peer_base = map_to_peer(base, peer);
    const uint32_t r_hi_row = peer_base + (a_hi_row - base), r_lo_row = peer_base + (a_lo_row - base);
    const uint32_t r_bar_h = peer_base + 32, r_bar_peer = peer_base + 24;
    const uint32_t own_base = map_to_peer(base, rank);      // this CTA's window in the cluster address space (st.async)
    const uint32_t l_hi_row = own_base + (a_hi_row - base), l_lo_row = own_base + (a_lo_row - base);
    const uint32_t l_bar_h = own_base + 32;
    const size_t y_row = (size_t)b * T;
    const int y_col = dir * 128 + (int)rank * 64 + half * 32;

    // This warp's Gx of one step is 16 KB contiguous ([col4 0..32)[lane] float4).  The recurrence streams Gx from HBM
    // at 3.6 TB/s and the register prefetch reaches only one chunk (~1000 cycles) ahead: long-scoreboard stalls on
    // these loads were the largest stall reason of the kernel (ncu).  One bulk L2 prefetch per warp and step, pf steps
    // ahead, turns them into L2 hits.
    const int pf = p.pf;
    auto prefetch_gx = [&](int t) {
      const float4* src = G4 + ((g_row + (size_t)t) * 256 + (size_t)col4_0) * 32;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(16384u) : "memory");
    };
    if (lane == 0 && pf > 0)
      for (int d = 0; d < pf && d < T; ++d) prefetch_gx(dir ? T - 1 - d : d);

    float c[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = 0.f;
    float4 gx[2][8];
    load_gx(dir ? T - 1 : 0, 0, gx[0]);

    for (int s = 0; s < T; ++s) {
      const int t = dir ? (T - 1 - s) : s;
      const int tn = dir ? (t - 1) : (t + 1);
      const uint32_t taddr = taddr0 + (uint32_t)(s & 1) * 256u;
      if (lane == 0 && pf > 0 && s + pf < T) prefetch_gx(dir ? t - pf : t + pf);
      mbar_wait(bar_acc + 8 * (uint32_t)(s & 1), (uint32_t)(s >> 1) & 1u);
      tc_fence_after();
      if (warp == 2 && lane == 0) {                         // my MMAs no longer read my A tile
        if (ASYNC) mbar_arrive_remote_relaxed(r_bar_peer); else mbar_arrive_remote(r_bar_peer);
      }
#pragma unroll
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint32_t acc[32];
        // next chunk's Gx first (global loads: the longest latency of the step), then the TMEM load and its wait
        if (chunk < 3) load_gx(t, chunk + 1, gx[(chunk + 1) & 1]);
        else if (s + 1 < T) load_gx(tn, 0, gx[0]);
        tc_ld32(taddr + chunk * 32, acc);
        const float4* gxc = gx[chunk & 1];
        uint4 ph, pl;
        __half2* ph2 = reinterpret_cast<__half2*>(&ph);
        __half2* pl2 = reinterpret_cast<__half2*>(&pl);
        float hv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 g4 = gxc[u];
          const float xi = __uint_as_float(acc[4 * u + 0]) + g4.x;
          const float xf = __uint_as_float(acc[4 * u + 1]) + g4.y;
          const float xg = __uint_as_float(acc[4 * u + 2]) + g4.z;
          const float xo = __uint_as_float(acc[4 * u + 3]) + g4.w;
          hv[u] = lstm_cell<NP>(xi, xf, xg, xo, c[chunk * 8 + u]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {                       // packed conversions only (scalar F2F runs on the MUFU pipe)
          const __half2 hh = __floats2half2_rn(hv[2 * e], hv[2 * e + 1]);
          const float2 hb = __half22float2(hh);
          ph2[e] = hh;
          pl2[e] = __floats2half2_rn(hv[2 * e] - hb.x, hv[2 * e + 1] - hb.y);
        }
        if (chunk == 0) mbar_wait_cluster(bar_peer, (uint32_t)s & 1u, false);          // peer's MMAs of this step are done
        // K order inside the k-block: (chunk, half, unit) -> 16-byte piece chunk*2 + half of the swizzled 128-byte row
        const uint32_t coff = (uint32_t)(((chunk * 2 + half) ^ (m & 7)) << 4);
        if (ASYNC) {
          st_async_v4(l_hi_row + coff, ph, l_bar_h + 8 * chunk);
          st_async_v4(l_lo_row + coff, pl, l_bar_h + 8 * chunk);
          st_async_v4(r_hi_row + coff, ph, r_bar_h + 8 * chunk);
          st_async_v4(r_lo_row + coff, pl, r_bar_h + 8 * chunk);
        } else {
          st_shared_v4(a_hi_row + coff, ph);
          st_shared_v4(a_lo_row + coff, pl);
          st_cluster_v4(r_hi_row + coff, ph);
          st_cluster_v4(r_lo_row + coff, pl);
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_release_cluster(bar_h + 8 * chunk);
            mbar_arrive_remote(r_bar_h + 8 * chunk);
          }
        }
        if (live) {                                         // after the exchange: nothing of this waits on HBM
          const size_t o = (y_row + (size_t)t) * 256 + (size_t)(y_col + chunk * 8);
          *reinterpret_cast<uint4*>(p.Yh + o) = ph;
          *reinterpret_cast<uint4*>(p.Yl + o) = pl;
        }
      }
      tc_fence_before();                                    // this step's tcgen05.ld before the next barrier waits
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                       // no CTA exits while its peer may still write into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

static int make_w_map(CUtensorMap* tm, const __half* ptr) {
  PFN_encodeTiled enc = get_encode();
  B200_CHECK(enc != nullptr, B200_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {128, 1024};                         // [dir][rank][256 gate columns] x 128 k
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {64, 256};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled(W_hh) failed: %d", (int)r);
  return B200_OK;
}

int lstm_rec_tc(const float* G, const __half* Whh_hi, const __half* Whh_lo, __half* Yh, __half* Yl, int NB,
                cudaStream_t stream) {
  RecTcParams p{};
  p.G = G; p.Yh = Yh; p.Yl = Yl; p.NB = NB; p.T = kFrames;
  p.ntiles = ceil_div(NB, 128);
  CUtensorMap tmWh, tmWl;
  int rc;
  if ((rc = make_w_map(&tmWh, Whh_hi))) return rc;
  if ((rc = make_w_map(&tmWl, Whh_lo))) return rc;
  const size_t smem = 1024 + 1024 + kRecABytes + kRecWBytes;
  // A/B knobs, read per call (tests flip them inside one process): B200_LSTM_PF = Gx prefetch distance,
  // B200_LSTM_NP = exponentials on the FMA pipe (-1 = round-1 gate math)
  { const char* e = getenv("B200_LSTM_PF"); p.pf = e ? atoi(e) : 0; }
  int np = kRecDefaultNP;
  if (const char* e = getenv("B200_LSTM_NP")) np = atoi(e);
  if (np < -1 || np > 2) np = kRecDefaultNP;
  int async = 1;
  if (const char* e = getenv("B200_LSTM_ASYNC")) async = atoi(e) != 0;
  static bool attr_set[16] = {};
  auto launch = [&](auto kernel) -> int {
    if (!attr_set[(np + 1) * 2 + async]) {
      B200_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set[(np + 1) * 2 + async] = true;
    }
    kernel<<<2 * 2 * p.ntiles, kRecThreads, smem, stream>>>(tmWh, tmWl, p);
    B200_CUDA_OK(cudaGetLastError());
    return B200_OK;
  };
  if (!async) return np < 0 ? launch(lstm_rec_tc_kernel<-1, false>) : launch(lstm_rec_tc_kernel<0, false>);
  switch (np) {
    case -1: return launch(lstm_rec_tc_kernel<-1, true>);
    case 0: return launch(lstm_rec_tc_kernel<0, true>);
    case 1: return launch(lstm_rec_tc_kernel<1, true>);
    default: return launch(lstm_rec_tc_kernel<2, true>);
  }
}

}  // namespace b200
