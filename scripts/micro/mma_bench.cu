// tcgen05 probe (sm_100a, not part of the product): cost of one M=128, K=16 kind::f16 MMA as a function of N with
//   SS: A and B from shared memory            TS: A from tensor memory (written by tcgen05.cp), B from shared memory
// and the cost of the smem -> TMEM copy (tcgen05.cp.128x256b) that TS mode needs, alone and overlapped with MMAs.
// Also checks numerically that an A tile copied with tcgen05.cp.128x256b from the SAME K-major swizzled smem tile the
// SS MMA reads gives the same D (so the conv kernels can switch operand source without changing their layouts).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu ; run on one B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\t"
               "bra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// smem: A tile 128 rows x 64 B (32 fp16 channels, SWIZZLE_64B, 8-row atoms of 512 B) = 8 KB; B tile 256 rows x 64 B
// mode 0: SS MMAs   1: TS MMAs (A copied once)   2: cp only   3: TS MMAs with one cp per 2 MMAs interleaved
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int iters, long long* cycles, float* dout,
                                                const __half* gA, const __half* gB) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - smem_u32(raw));
  const uint32_t bar = base, slot = base + 64, a_s = base + 1024, b_s = base + 1024 + 8192;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // swizzled fill: element (row r, 16-byte chunk c) lives at r*64 + ((c ^ ((r >> 1) & 3)) * 16)
  for (int i = threadIdx.x; i < 128 * 4; i += 128) {
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(g + 1024 + r * 64 + ((c ^ ((r >> 1) & 3)) * 16)) =
        *reinterpret_cast<const uint4*>(gA + r * 32 + c * 8);
  }
  for (int i = threadIdx.x; i < 256 * 4; i += 128) {
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(g + 1024 + 8192 + r * 64 + ((c ^ ((r >> 1) & 3)) * 16)) =
        *reinterpret_cast<const uint4*>(gB + r * 32 + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(g + 64);
  const uint32_t d_t = tmem, a_t = tmem + 256;            // D: columns 0..255, A tiles: columns 256.. (8 per k-step)
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t ad0 = kdesc(a_s, 512, 4), ad1 = kdesc(a_s + 32, 512, 4);
  const uint64_t bd0 = kdesc(b_s, 512, 4), bd1 = kdesc(b_s + 32, 512, 4);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    if (mode == 1 || mode == 3 || mode == 5) {             // A tile (2 k-steps) into TMEM once
      cp_128x256b(a_t, ad0);
      cp_128x256b(a_t + 8, ad1);
      commit(bar);
      mbar_wait(bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    } else {
      commit(bar);
      mbar_wait(bar, 0);
    }
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        mma_ss(d_t, ad0, bd0, idesc, it > 0);
        mma_ss(d_t, ad1, bd1, idesc, 1);
      } else if (mode == 1) {
        mma_ts(d_t, a_t, bd0, idesc, it > 0);
        mma_ts(d_t, a_t + 8, bd1, idesc, 1);
      } else if (mode == 4 || mode == 5) {               // SS / TS rotating over several accumulators
        const int nacc = 256 / N < 1 ? 1 : (256 / N > 4 ? 4 : 256 / N);
        const uint32_t d = d_t + (uint32_t)((it % nacc) * N);
        if (mode == 4) { mma_ss(d, ad0, bd0, idesc, it >= nacc); mma_ss(d, ad1, bd1, idesc, 1); }
        else { mma_ts(d, a_t, bd0, idesc, it >= nacc); mma_ts(d, a_t + 8, bd1, idesc, 1); }
      } else if (mode == 6) {                            // SS, the two k-steps go to different accumulators
        mma_ss(d_t, ad0, bd0, idesc, it > 0);
        mma_ss(d_t + (uint32_t)N, ad1, bd1, idesc, it > 0);
      } else if (mode == 2) {
        cp_128x256b(a_t + 16, ad0);
        cp_128x256b(a_t + 24, ad1);
      } else {
        cp_128x256b(a_t + 16 + (it & 1) * 16, ad0);
        mma_ts(d_t, a_t, bd0, idesc, it > 0);
        mma_ts(d_t, a_t + 8, bd1, idesc, 1);
      }
    }
    commit(bar);
    mbar_wait(bar, 1);
    t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (dout && blockIdx.x == 0) {                           // D[row = lane][col] -> global (first 32 columns)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t r[32];
    const uint32_t ta = d_t + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(ta));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) dout[(warp * 32 + lane) * 32 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

int main() {
  const int smem = 1024 + 1024 + 8192 + 16384 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  __half hA[128 * 32], hB[256 * 32];
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 32; ++k) hA[r * 32 + k] = __float2half((float)((r * 3 + k) % 7 - 3));
  for (int n = 0; n < 256; ++n)
    for (int k = 0; k < 32; ++k) hB[n * 32 + k] = __float2half((float)((n + 2 * k) % 5 - 2));
  __half *dA, *dB;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  long long* dc; cudaMalloc(&dc, 148 * 8);
  float* dd; cudaMalloc(&dd, 128 * 32 * 4);
  float hD[128 * 32];
  // numerics: one iteration, N = 32, modes 0 and 1 against the host product over K = 32
  for (int mode = 0; mode < 2; ++mode) {
    bench<<<1, 128, smem>>>(mode, 32, 1, dc, dd, dA, dB);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD, dd, sizeof(hD), cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        float ref = 0;
        for (int k = 0; k < 32; ++k) ref += __half2float(hA[m * 32 + k]) * __half2float(hB[n * 32 + k]);
        const double err = fabs(ref - hD[m * 32 + n]);
        if (err > maxerr) maxerr = err;
      }
    printf("numerics %s: max |D - ref| = %g\n", mode == 0 ? "SS" : "TS (A via tcgen05.cp)", maxerr);
  }
  const int iters = 2000;
  const char* names[7] = {"SS  mma x2", "TS  mma x2", "cp  x2    ", "TS mma x2 + cp x1", "SS rotating D", "TS rotating D",
                          "SS k-steps to 2 D"};
  for (int mode = 0; mode < 7; ++mode)
    for (int N : {32, 64, 96, 128, 192, 256}) {
      if (mode == 2 && N != 32) continue;
      if (mode >= 4 && N > 128) continue;
      for (int grid : {1}) {
        bench<<<grid, 128, smem>>>(mode, N, iters, dc, nullptr, dA, dB);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N %d failed: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
        long long hc[148];
        cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = hc[i] > mx ? hc[i] : mx;
        printf("%s N=%3d grid=%3d: %.1f cycles per iteration (2 k-steps)\n", names[mode], N, grid, (double)mx / iters);
      }
    }
  return 0;
}
