// LSTM gate-math probe (sm_100a): MUFU throughput and cycles per cell of several formulations of
//   c' = sigmoid(f) c + sigmoid(i) tanh(g),  h = sigmoid(o) tanh(c')
// with W warps per SM (the recurrence kernel has 8 epilogue warps = 2 per SM sub-partition).  Not part of the product.
//   V0: 5 ex2 + 3 rcp on the MUFU pipe (round-1 kernel)        V1: 5 ex2 + 2 rcp (forget gate shares the reciprocal)
//   V2..V6: V1 with 1..5 of the exponentials as an FMA-pipe polynomial (magic-number floor + degree-D Horner + exponent add)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

#ifndef PDEG
#define PDEG 6
#endif
// 2^(s*x) for s*x in [-126, kLim]; x already clamped by the caller
__device__ __forceinline__ float poly_ex2_scaled(float x, float s) {
  const float magic = 12582912.f;                    // 1.5 * 2^23
  const float t = fmaf(x, s, magic);                 // round(s x) in the low mantissa bits
  const float n = t - magic;
  const float f = fmaf(x, s, -n);                    // [-0.5, 0.5]
#if PDEG == 6
  float p = 0.00015461444854736328f;
  p = fmaf(p, f, 0.0013400427997112274f);
  p = fmaf(p, f, 0.009618056938052177f);
  p = fmaf(p, f, 0.05550327152013779f);
  p = fmaf(p, f, 0.24022650718688965f);
  p = fmaf(p, f, 0.6931471824645996f);
  p = fmaf(p, f, 1.0f);
#else
  float p = 0.0013390863314270973f;
  p = fmaf(p, f, 0.009676031768321991f);
  p = fmaf(p, f, 0.055503569543361664f);
  p = fmaf(p, f, 0.2402210682630539f);
  p = fmaf(p, f, 0.6931471824645996f);
  p = fmaf(p, f, 1.0000001192092896f);
#endif
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

constexpr float kL = 1.4426950408889634f;

template <int V>
__device__ __forceinline__ void cell(float xi, float xf, float xg, float xo, float& c, float& h) {
  if (V == 0) {
    constexpr float kLim = 57.f;
    const float ei = mufu_ex2(fminf(-kL * xi, kLim));
    const float ef = mufu_ex2(fminf(-kL * xf, kLim));
    const float eg = mufu_ex2(fminf(-2.f * kL * xg, kLim));
    const float eo = mufu_ex2(fminf(-kL * xo, kLim));
    const float ig = (1.f - eg) * mufu_rcp((1.f + ei) * (1.f + eg));
    const float fg = mufu_rcp(1.f + ef);
    const float cn = fmaf(fg, c, ig);
    c = cn;
    const float ec = mufu_ex2(fminf(-2.f * kL * cn, kLim));
    h = (1.f - ec) * mufu_rcp((1.f + eo) * (1.f + ec));
  } else {
    // exponent limit 40: three denominators multiply to < 2^123
    constexpr float kLim = 40.f;
    constexpr float xlim = kLim / kL, xlim2 = kLim / (2.f * kL);
    constexpr int NP = V - 1;                        // exponentials on the FMA pipe, in the order g, i, f, o, c
    float ei, ef, eg, eo;
    if (NP >= 1) eg = poly_ex2_scaled(fminf(fmaxf(xg, -xlim2), 43.f), -2.f * kL);
    else eg = mufu_ex2(fminf(-2.f * kL * xg, kLim));
    if (NP >= 2) ei = poly_ex2_scaled(fminf(fmaxf(xi, -xlim), 87.f), -kL);
    else ei = mufu_ex2(fminf(-kL * xi, kLim));
    if (NP >= 3) ef = poly_ex2_scaled(fminf(fmaxf(xf, -xlim), 87.f), -kL);
    else ef = mufu_ex2(fminf(-kL * xf, kLim));
    if (NP >= 4) eo = poly_ex2_scaled(fminf(fmaxf(xo, -xlim), 87.f), -kL);
    else eo = mufu_ex2(fminf(-kL * xo, kLim));
    const float d1 = (1.f + ei) * (1.f + eg), d2 = 1.f + ef;
    const float r = mufu_rcp(d1 * d2);
    const float ig = (1.f - eg) * d2 * r;
    const float fg = d1 * r;
    const float cn = fmaf(fg, c, ig);
    c = cn;
    float ec;
    if (NP >= 5) ec = poly_ex2_scaled(fminf(fmaxf(cn, -xlim2), 43.f), -2.f * kL);
    else ec = mufu_ex2(fminf(-2.f * kL * cn, kLim));
    h = (1.f - ec) * mufu_rcp((1.f + eo) * (1.f + ec));
  }
}

template <int V>
__global__ void gate_kernel(const float* __restrict__ x, float* __restrict__ out, long long* cyc, int iters) {
  float c[8], h[8], xin[8][4];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    c[u] = 0.f; h[u] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) xin[u][g] = x[(size_t)tid * 32 + u * 4 + g];
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      // inputs depend on the previous h like the recurrence (keeps the compiler from hoisting anything)
      const float d = h[u] * 0.25f;
      cell<V>(xin[u][0] + d, xin[u][1] - d, xin[u][2] + d, xin[u][3] - d, c[u], h[u]);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 8; ++u) { out[(size_t)tid * 16 + u] = c[u]; out[(size_t)tid * 16 + 8 + u] = h[u]; }
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void mufu_kernel(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.001f * threadIdx.x + i * 0.01f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = OP == 0 ? mufu_ex2(v[i]) : mufu_rcp(v[i]);
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static void host_ref(const float* x, int iters, double* c, double* h) {
  for (int u = 0; u < 8; ++u) { c[u] = 0; h[u] = 0; }
  for (int it = 0; it < iters; ++it)
    for (int u = 0; u < 8; ++u) {
      // the same float32 input arithmetic as the device, exact gates in double
      const float hf = (float)h[u];
      const float d = hf * 0.25f;
      const double xi = x[u * 4 + 0] + d, xf = x[u * 4 + 1] - d, xg = x[u * 4 + 2] + d, xo = x[u * 4 + 3] - d;
      const double si = 1 / (1 + exp(-(double)(float)xi)), sf = 1 / (1 + exp(-(double)(float)xf)),
                   so = 1 / (1 + exp(-(double)(float)xo));
      c[u] = sf * c[u] + si * tanh((double)(float)xg);
      h[u] = so * tanh(c[u]);
    }
}

template <int V>
static void run_gate(const float* dx, const std::vector<float>& hx, float* dout, long long* dcyc, int threads) {
  const int blocks = 148, iters = 2000;
  gate_kernel<V><<<blocks, threads>>>(dx, dout, dcyc, iters);
  cudaDeviceSynchronize();
  gate_kernel<V><<<blocks, threads>>>(dx, dout, dcyc, iters);
  cudaDeviceSynchronize();
  std::vector<long long> cyc(blocks);
  cudaMemcpy(cyc.data(), dcyc, blocks * 8, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (auto v : cyc) mean += (double)v;
  mean /= blocks;
  // accuracy after ONE iteration-set of 8 steps would drift chaotically; check a short run instead
  const int chk_iters = 4;
  gate_kernel<V><<<1, 64>>>(dx, dout, dcyc, chk_iters);
  cudaDeviceSynchronize();
  std::vector<float> o(64 * 16);
  cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
  double emax_c = 0, emax_h = 0;
  for (int t = 0; t < 64; ++t) {
    double c[8], h[8];
    host_ref(hx.data() + (size_t)t * 32, chk_iters, c, h);
    for (int u = 0; u < 8; ++u) {
      emax_c = fmax(emax_c, fabs(o[t * 16 + u] - c[u]));
      emax_h = fmax(emax_h, fabs(o[t * 16 + 8 + u] - h[u]));
    }
  }
  printf("V%d deg%d threads %4d: %7.3f cycles per cell per SM   (%.0f cycles per 8192 cells)   max|dc| %.2e  max|dh| %.2e\n",
         V, PDEG, threads, mean / iters / (8.0 * threads) , mean / iters / (8.0 * threads) * 8192, emax_c, emax_h);
}

int main() {
  const int maxthr = 1024;
  std::vector<float> hx((size_t)148 * maxthr * 32);
  srand(1);
  for (auto& v : hx) v = ((rand() % 20001) - 10000) * 6e-4f;     // gate pre-activations in [-6, 6]
  float *dx, *dout;
  long long* dcyc;
  cudaMalloc(&dx, hx.size() * 4);
  cudaMalloc(&dout, (size_t)148 * maxthr * 16 * 4);
  cudaMalloc(&dcyc, 148 * 8);
  cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
  for (int threads : {256, 512, 1024}) {
    for (int op = 0; op < 2; ++op) {
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) mufu_kernel<0><<<148, threads>>>(dout, dcyc, iters); else mufu_kernel<1><<<148, threads>>>(dout, dcyc, iters);
        cudaDeviceSynchronize();
      }
      std::vector<long long> cyc(148);
      cudaMemcpy(cyc.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost);
      double mean = 0; for (auto v : cyc) mean += (double)v; mean /= 148;
      printf("MUFU.%s threads %4d: %.2f results per cycle per SM\n", op ? "RCP" : "EX2", threads, 16.0 * threads * iters / mean);
    }
  }
  for (int threads : {256, 512}) {
    run_gate<0>(dx, hx, dout, dcyc, threads);
    run_gate<1>(dx, hx, dout, dcyc, threads);
    run_gate<2>(dx, hx, dout, dcyc, threads);
    run_gate<3>(dx, hx, dout, dcyc, threads);
    run_gate<4>(dx, hx, dout, dcyc, threads);
    run_gate<5>(dx, hx, dout, dcyc, threads);
    run_gate<6>(dx, hx, dout, dcyc, threads);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
