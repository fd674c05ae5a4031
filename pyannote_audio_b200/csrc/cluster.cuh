#pragma once
#include "common.cuh"
namespace b200 {
size_t linkage_workspace_bytes(int n, int dim);
int linkage_centroid(const double* x, int n, int dim, int normalize, double* Z, void* ws, cudaStream_t st);
int cdist_cosine(const double* a, int m, const double* b, int k, int dim, double* d, cudaStream_t st);
size_t vbx_workspace_bytes(int n, int D, int S);
int vbx_run(const double* fea, const double* phi, int n, int D, int S, double Fa, double Fb, int max_iters,
            double epsilon, double* gamma, double* pi, int* iters_host, void* ws, cudaStream_t st);
int assign_clusters(const double* soft, int C, int K, int constrained, signed char* hard, cudaStream_t st);
int fcluster_distance(const double* Z, int n, double t, int* T);
}
