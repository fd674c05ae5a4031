// Clustering on the device, fp64 like the reference (numpy/scipy promote to double):
//   * centroid-linkage agglomerative clustering  = scipy.cluster.hierarchy.linkage(X, "centroid", "euclidean")
//       call sites /root/reference/src/pyannote/audio/pipelines/clustering.py:597-603 (VBx) and :371-385 (AHC)
//   * fcluster(Z, t, criterion="distance")        (host, tree walk)      clustering.py:604, 385
//   * cdist(., ., "cosine")                       clustering.py:645-655
//   * VBx variational iterations                  utils/vbx.py:98-136 (called from clustering.py:608-616)
//   * constrained_argmax (3 x K assignment)       clustering.py:127-140
// Arithmetic that decides merges uses explicit round-to-nearest mul/add (no FMA contraction) in the same operation
// order as scipy's C code so that dendrogram heights agree to the last bit on ordinary inputs.
#include "../../include/b200diar.h"
#include "common.cuh"
#include "cluster.cuh"
#include <cfloat>
#include <cmath>

namespace b200 {

// ------------------------------------------------------------------------------------------------------
// pairwise Euclidean distances (full symmetric matrix, diagonal unused)
// ------------------------------------------------------------------------------------------------------
__global__ void normalize_rows_kernel(const double* __restrict__ x, double* __restrict__ y, int n, int dim) {
  const int i = blockIdx.x;
  __shared__ double red[32];
  double s = 0.0;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) s += x[(size_t)i * dim + d] * x[(size_t)i * dim + d];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
    red[0] = sqrt(t);
  }
  __syncthreads();
  const double nrm = red[0];
  for (int d = threadIdx.x; d < dim; d += blockDim.x) y[(size_t)i * dim + d] = x[(size_t)i * dim + d] / nrm;
}

__global__ void pdist_kernel(const double* __restrict__ x, double* __restrict__ D, int n, int dim) {
  // 16x16 tile of pairs per block, operands staged through shared memory in chunks of 32 dims
  __shared__ double xi[16][33], xj[16][33];
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  double s = 0.0;
  for (int d0 = 0; d0 < dim; d0 += 32) {
    const int tid = threadIdx.y * 16 + threadIdx.x;
    for (int e = tid; e < 16 * 32; e += 256) {
      const int r = e >> 5, c = e & 31;
      const int gi = blockIdx.y * 16 + r, gj = blockIdx.x * 16 + r;
      xi[r][c] = (gi < n && d0 + c < dim) ? x[(size_t)gi * dim + d0 + c] : 0.0;
      xj[r][c] = (gj < n && d0 + c < dim) ? x[(size_t)gj * dim + d0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const double df = __dsub_rn(xi[threadIdx.y][c], xj[threadIdx.x][c]);
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    __syncthreads();
  }
  if (i < n && j < n) D[(size_t)i * n + j] = (i == j) ? DBL_MAX : sqrt(s);
}

// ------------------------------------------------------------------------------------------------------
// centroid linkage: single persistent CTA, nearest-neighbour candidates per row (upper triangle)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lw_centroid(double dxi, double dyi, double dxy, double nx, double ny) {
  // scipy/_hierarchy_distance_update.pxi::_centroid
  //   sqrt((((nx*dxi*dxi) + (ny*dyi*dyi)) - (nx*ny*dxy*dxy)/(nx+ny)) / (nx+ny))
  const double a = __dmul_rn(__dmul_rn(nx, dxi), dxi);
  const double b = __dmul_rn(__dmul_rn(ny, dyi), dyi);
  const double c = __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(nx, ny), dxy), dxy), __dadd_rn(nx, ny));
  return sqrt(__ddiv_rn(__dsub_rn(__dadd_rn(a, b), c), __dadd_rn(nx, ny)));
}

struct MinPair {
  double v;
  int i;
};
__device__ __forceinline__ MinPair min_pair(MinPair a, MinPair b) {
  if (b.v < a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ MinPair warp_min(MinPair m) {
  for (int o = 16; o > 0; o >>= 1) {
    MinPair t;
    t.v = __shfl_xor_sync(0xffffffffu, m.v, o);
    t.i = __shfl_xor_sync(0xffffffffu, m.i, o);
    m = min_pair(m, t);
  }
  return m;
}

// nearest alive j > k of row k, computed by one warp
__device__ __forceinline__ void row_nn(const double* __restrict__ D, const unsigned char* __restrict__ alive, int n,
                                       int k, int lane, double* nn_d, int* nn_i) {
  MinPair m{DBL_MAX, n};
  const double* row = D + (size_t)k * n;
  for (int j = k + 1 + lane; j < n; j += 32)
    if (alive[j]) m = min_pair(m, MinPair{row[j], j});
  m = warp_min(m);
  if (lane == 0) { nn_d[k] = m.v; nn_i[k] = m.i; }
}

__global__ void __launch_bounds__(1024) linkage_centroid_kernel(double* __restrict__ D, int n, double* __restrict__ Z,
                                                                double* __restrict__ nn_d, int* __restrict__ nn_i,
                                                                int* __restrict__ size, int* __restrict__ id,
                                                                unsigned char* __restrict__ alive,
                                                                int* __restrict__ todo) {
  __shared__ MinPair s_red[32];
  __shared__ int s_x, s_y, s_ntodo;
  __shared__ double s_dxy;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n; i += 1024) { size[i] = 1; id[i] = i; alive[i] = 1; }
  __syncthreads();
  for (int k = warp; k < n; k += 32) row_nn(D, alive, n, k, lane, nn_d, nn_i);
  __syncthreads();

  for (int m = 0; m < n - 1; ++m) {
    // A. global closest pair
    MinPair best{DBL_MAX, n};
    for (int i = tid; i < n; i += 1024)
      if (alive[i] && nn_i[i] < n) best = min_pair(best, MinPair{nn_d[i], i});
    best = warp_min(best);
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (warp == 0) {
      MinPair b = s_red[lane];
      b = warp_min(b);
      if (lane == 0) {
        s_x = b.i;
        s_y = nn_i[b.i];
        s_dxy = b.v;
        s_ntodo = 0;
      }
    }
    __syncthreads();
    const int x = s_x, y = s_y;
    const double dxy = s_dxy;
    const double nx = size[x], ny = size[y];
    if (tid == 0) {
      const int ia = id[x], ib = id[y];
      Z[m * 4 + 0] = ia < ib ? ia : ib;
      Z[m * 4 + 1] = ia < ib ? ib : ia;
      Z[m * 4 + 2] = dxy;
      Z[m * 4 + 3] = nx + ny;
    }
    // B. Lance-Williams update: merged cluster lives in slot y, slot x dies
    const double* rx = D + (size_t)x * n;
    double* ry = D + (size_t)y * n;
    for (int k = tid; k < n; k += 1024) {
      if (!alive[k] || k == x || k == y) continue;
      const double dn = lw_centroid(rx[k], ry[k], dxy, nx, ny);
      ry[k] = dn;
      D[(size_t)k * n + y] = dn;
    }
    __syncthreads();
    if (tid == 0) { alive[x] = 0; size[y] = (int)(nx + ny); id[y] = n + m; }
    __syncthreads();
    // C. nearest-neighbour maintenance
    for (int k = tid; k < y; k += 1024) {
      if (!alive[k]) continue;
      const int cur = nn_i[k];
      if (cur == x || cur == y) {
        todo[atomicAdd(&s_ntodo, 1)] = k;
      } else {
        const double dn = D[(size_t)k * n + y];
        if (dn < nn_d[k] || (dn == nn_d[k] && y < cur)) { nn_d[k] = dn; nn_i[k] = y; }
      }
    }
    if (tid == 0) todo[atomicAdd(&s_ntodo, 1)] = y;
    __syncthreads();
    const int nt = s_ntodo;
    for (int t = warp; t < nt; t += 32) row_nn(D, alive, n, todo[t], lane, nn_d, nn_i);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// cosine cdist
// ------------------------------------------------------------------------------------------------------
__global__ void cdist_cosine_kernel(const double* __restrict__ a, int m, const double* __restrict__ b, int k, int dim,
                                    double* __restrict__ d) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * k) return;
  const int i = idx / k, j = idx % k;
  const double *u = a + (size_t)i * dim, *v = b + (size_t)j * dim;
  double uv = 0, uu = 0, vv = 0;
  for (int t = 0; t < dim; ++t) {
    uv += u[t] * v[t];
    uu += u[t] * u[t];
    vv += v[t] * v[t];
  }
  double c = uv / (sqrt(uu) * sqrt(vv));
  if (fabs(c) > 1.0) c = copysign(1.0, c);
  d[idx] = 1.0 - c;
}

// ------------------------------------------------------------------------------------------------------
// VBx
// ------------------------------------------------------------------------------------------------------
struct VbxBuf {
  double *rho, *G, *Ng, *alpha, *invL, *cst, *lpx, *elbo;   // elbo[0]=prev, [1]=cur, [2]=done flag (as double)
  int* iters;
};

__global__ void vbx_prep_kernel(const double* __restrict__ X, const double* __restrict__ phi, int n, int D,
                                double* __restrict__ rho, double* __restrict__ G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    const double x = X[(size_t)i * D + d];
    s += x * x;
    rho[(size_t)i * D + d] = x * sqrt(phi[d]);
  }
  G[i] = -0.5 * (s + D * log(2.0 * M_PI));
}

// speaker models: Ng[s] = sum_n gamma[n][s];  alpha[s][d] = FaFb * invL * sum_n gamma[n][s] rho[n][d]
__global__ void vbx_model_kernel(const double* __restrict__ gamma, const double* __restrict__ rho,
                                 const double* __restrict__ phi, int n, int D, int S, double FaFb,
                                 double* __restrict__ alpha, double* __restrict__ invL, double* __restrict__ cst,
                                 const double* __restrict__ elbo) {
  if (elbo[2] != 0.0) return;
  const int s = blockIdx.x;                 // one block per speaker, threads over d
  __shared__ double s_ng;
  __shared__ double red[8];
  double ng = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ng += gamma[(size_t)i * S + s];
  for (int o = 16; o > 0; o >>= 1) ng += __shfl_xor_sync(0xffffffffu, ng, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ng;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < blockDim.x / 32; ++w) t += red[w];
    s_ng = t;
  }
  __syncthreads();
  double c_part = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc += gamma[(size_t)i * S + s] * rho[(size_t)i * D + d];
    const double il = 1.0 / (1.0 + FaFb * s_ng * phi[d]);
    const double al = FaFb * il * acc;
    invL[s * D + d] = il;
    alpha[s * D + d] = al;
    c_part += (il + al * al) * phi[d];
  }
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) c_part += __shfl_xor_sync(0xffffffffu, c_part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c_part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < blockDim.x / 32; ++w) t += red[w];
    cst[s] = -0.5 * t;
  }
}

// responsibilities: one warp per frame n
__global__ void vbx_resp_kernel(const double* __restrict__ rho, const double* __restrict__ G,
                                const double* __restrict__ alpha, const double* __restrict__ cst,
                                const double* __restrict__ pi, int n, int D, int S, double Fa,
                                double* __restrict__ gamma, double* __restrict__ lpx, const double* __restrict__ elbo) {
  if (elbo[2] != 0.0) return;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  double mx = -DBL_MAX;
  for (int s = lane; s < S; s += 32) {
    double dot = 0.0;
    for (int d = 0; d < D; ++d) dot += rho[(size_t)i * D + d] * alpha[s * D + d];
    const double v = Fa * (dot + cst[s] + G[i]) + log(pi[s] + 1e-8);
    gamma[(size_t)i * S + s] = v;
    mx = fmax(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double se = 0.0;
  for (int s = lane; s < S; s += 32) se += exp(gamma[(size_t)i * S + s] - mx);
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  const double lse = log(se) + mx;
  for (int s = lane; s < S; s += 32) gamma[(size_t)i * S + s] = exp(gamma[(size_t)i * S + s] - lse);
  if (lane == 0) lpx[i] = lse;
}

// priors, ELBO, convergence (single block)
__global__ void vbx_finish_kernel(const double* __restrict__ gamma, const double* __restrict__ lpx,
                                  const double* __restrict__ alpha, const double* __restrict__ invL, int n, int D,
                                  int S, double Fb, double epsilon, int it, double* __restrict__ pi,
                                  double* __restrict__ elbo, int* __restrict__ iters) {
  if (elbo[2] != 0.0) return;
  __shared__ double red[32];
  __shared__ double s_tot;
  auto block_sum = [&](double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0) {
      for (int w = 0; w < blockDim.x / 32; ++w) t += red[w];
      s_tot = t;
    }
    __syncthreads();
    return s_tot;
  };
  double tot = 0.0;
  for (int s = 0; s < S; ++s) {
    double c = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += gamma[(size_t)i * S + s];
    c = block_sum(c);
    if (threadIdx.x == 0) pi[s] = c;
    tot += c;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) pi[s] = pi[s] / tot;
  double l = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) l += lpx[i];
  l = block_sum(l);
  double r = 0.0;
  for (int e = threadIdx.x; e < S * D; e += blockDim.x) r += log(invL[e]) - invL[e] - alpha[e] * alpha[e] + 1.0;
  r = block_sum(r);
  if (threadIdx.x == 0) {
    const double E = l + Fb * 0.5 * r;
    const double prev = elbo[1];
    elbo[0] = prev;
    elbo[1] = E;
    *iters = it + 1;
    if (it > 0 && E - prev < epsilon) elbo[2] = 1.0;
  }
}

// ------------------------------------------------------------------------------------------------------
// constrained assignment: maximise sum of soft[c][s][k] over injective maps of (up to 3) speakers to clusters
// ------------------------------------------------------------------------------------------------------
__global__ void assign_kernel(const double* __restrict__ soft, int C, int K, int constrained,
                              signed char* __restrict__ hard) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* p = soft + (size_t)c * 3 * K;
  signed char* h = hard + c * 3;
  if (!constrained) {
    for (int s = 0; s < 3; ++s) {
      int best = 0;
      for (int k = 1; k < K; ++k)
        if (p[s * K + k] > p[s * K + best]) best = k;
      h[s] = (signed char)best;
    }
    return;
  }
  h[0] = h[1] = h[2] = -2;
  double bestv = -DBL_MAX;
  if (K >= 3) {
    for (int k0 = 0; k0 < K; ++k0)
      for (int k1 = 0; k1 < K; ++k1) {
        if (k1 == k0) continue;
        for (int k2 = 0; k2 < K; ++k2) {
          if (k2 == k0 || k2 == k1) continue;
          const double v = p[k0] + p[K + k1] + p[2 * K + k2];
          if (v > bestv) { bestv = v; h[0] = k0; h[1] = k1; h[2] = k2; }
        }
      }
  } else if (K == 2) {
    // two of the three speakers get the two clusters
    for (int s0 = 0; s0 < 3; ++s0)
      for (int s1 = 0; s1 < 3; ++s1) {
        if (s1 == s0) continue;
        const double v = p[s0 * K + 0] + p[s1 * K + 1];
        if (v > bestv) { bestv = v; h[0] = h[1] = h[2] = -2; h[s0] = 0; h[s1] = 1; }
      }
  } else if (K == 1) {
    int best = 0;
    for (int s = 1; s < 3; ++s)
      if (p[s] > p[best]) best = s;
    h[best] = 0;
  }
}

// ------------------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------------------
size_t linkage_workspace_bytes(int n, int dim) {
  return align_up((size_t)n * n * 8, 256) + align_up((size_t)n * dim * 8, 256) + (size_t)n * 64 + 4096;
}

int linkage_centroid(const double* x, int n, int dim, int normalize, double* Z, void* ws, cudaStream_t st) {
  char* p = (char*)ws;
  double* D = (double*)p; p += align_up((size_t)n * n * 8, 256);
  double* xn = (double*)p; p += align_up((size_t)n * dim * 8, 256);
  double* nn_d = (double*)p; p += align_up((size_t)n * 8, 256);
  int* nn_i = (int*)p; p += align_up((size_t)n * 4, 256);
  int* size = (int*)p; p += align_up((size_t)n * 4, 256);
  int* id = (int*)p; p += align_up((size_t)n * 4, 256);
  int* todo = (int*)p; p += align_up((size_t)n * 4 + 4, 256);
  unsigned char* alive = (unsigned char*)p;
  const double* src = x;
  if (normalize) {
    normalize_rows_kernel<<<n, 128, 0, st>>>(x, xn, n, dim);
    src = xn;
  }
  dim3 grid(ceil_div(n, 16), ceil_div(n, 16));
  pdist_kernel<<<grid, dim3(16, 16), 0, st>>>(src, D, n, dim);
  linkage_centroid_kernel<<<1, 1024, 0, st>>>(D, n, Z, nn_d, nn_i, size, id, alive, todo);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

int cdist_cosine(const double* a, int m, const double* b, int k, int dim, double* d, cudaStream_t st) {
  cdist_cosine_kernel<<<ceil_div(m * k, 128), 128, 0, st>>>(a, m, b, k, dim, d);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

size_t vbx_workspace_bytes(int n, int D, int S) {
  return ((size_t)n * D + n + S + 2 * (size_t)S * D + S + n + 8) * 8 + 4096 + 64;
}

int vbx_run(const double* fea, const double* phi, int n, int D, int S, double Fa, double Fb, int max_iters,
            double epsilon, double* gamma, double* pi, int* iters_host, void* ws, cudaStream_t st) {
  double* p = (double*)ws;
  double* rho = p; p += (size_t)n * D;
  double* G = p; p += n;
  double* alpha = p; p += (size_t)S * D;
  double* invL = p; p += (size_t)S * D;
  double* cst = p; p += S;
  double* lpx = p; p += n;
  double* elbo = p; p += 4;
  int* iters = (int*)p;
  B200_CUDA_OK(cudaMemsetAsync(elbo, 0, 4 * 8 + 8, st));
  vbx_prep_kernel<<<ceil_div(n, 128), 128, 0, st>>>(fea, phi, n, D, rho, G);
  // pi init = 1/S (utils/vbx.py:93-94)
  std::vector<double> pi0(S, 1.0 / S);
  B200_CUDA_OK(cudaMemcpyAsync(pi, pi0.data(), S * 8, cudaMemcpyHostToDevice, st));
  B200_CUDA_OK(cudaStreamSynchronize(st));   // pi0 is a stack-lifetime host buffer
  for (int it = 0; it < max_iters; ++it) {
    vbx_model_kernel<<<S, 128, 0, st>>>(gamma, rho, phi, n, D, S, Fa / Fb, alpha, invL, cst, elbo);
    vbx_resp_kernel<<<ceil_div(n * 32, 256), 256, 0, st>>>(rho, G, alpha, cst, pi, n, D, S, Fa, gamma, lpx, elbo);
    vbx_finish_kernel<<<1, 1024, 0, st>>>(gamma, lpx, alpha, invL, n, D, S, Fb, epsilon, it, pi, elbo, iters);
  }
  B200_CUDA_OK(cudaGetLastError());
  if (iters_host) {
    B200_CUDA_OK(cudaMemcpyAsync(iters_host, iters, sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA_OK(cudaStreamSynchronize(st));
  }
  return B200_OK;
}

int assign_clusters(const double* soft, int C, int K, int constrained, signed char* hard, cudaStream_t st) {
  assign_kernel<<<ceil_div(C, 128), 128, 0, st>>>(soft, C, K, constrained, hard);
  B200_CUDA_OK(cudaGetLastError());
  return B200_OK;
}

// fcluster(Z, t, criterion="distance") -- scipy/_hierarchy.pyx cluster_dist -> cluster_monocrit
int fcluster_distance(const double* Z, int n, double t, int* T) {
  if (n == 1) { T[0] = 1; return B200_OK; }
  std::vector<double> MD(n - 1);
  for (int i = 0; i < n - 1; ++i) {
    double m = Z[i * 4 + 2];
    const int l = (int)Z[i * 4 + 0], r = (int)Z[i * 4 + 1];
    if (l >= n) m = std::fmax(m, MD[l - n]);
    if (r >= n) m = std::fmax(m, MD[r - n]);
    MD[i] = m;
  }
  std::vector<int> curr(n);
  std::vector<unsigned char> visited(2 * n, 0);
  int k = 0, n_cluster = 0, leader = -1;
  curr[0] = 2 * n - 2;
  while (k >= 0) {
    const int root = curr[k] - n;
    const int lc = (int)Z[root * 4 + 0], rc = (int)Z[root * 4 + 1];
    if (leader == -1 && MD[root] <= t) { leader = root; ++n_cluster; }
    if (lc >= n && !visited[lc]) { visited[lc] = 1; curr[++k] = lc; continue; }
    if (rc >= n && !visited[rc]) { visited[rc] = 1; curr[++k] = rc; continue; }
    if (lc < n) { if (leader == -1) ++n_cluster; T[lc] = n_cluster; }
    if (rc < n) { if (leader == -1) ++n_cluster; T[rc] = n_cluster; }
    if (leader == root) leader = -1;
    --k;
  }
  return B200_OK;
}

}  // namespace b200
