// Shared helpers for the B200 (sm_100a) diarization kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>

namespace b200 {

// ---- error plumbing: no exceptions cross the C ABI; last error message is thread-local ----
void set_error(const char* fmt, ...);
const char* last_error();

#define B200_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      b200::set_error("%s:%d CUDA error %d (%s) in `%s`", __FILE__, __LINE__, (int)_e,       \
                      cudaGetErrorString(_e), #expr);                                        \
      return (_e == cudaErrorMemoryAllocation) ? B200_ERR_OOM : B200_ERR_CUDA;               \
    }                                                                                        \
  } while (0)

#define B200_CHECK(cond, code, ...)                                                          \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      b200::set_error(__VA_ARGS__);                                                          \
      return (code);                                                                         \
    }                                                                                        \
  } while (0)

enum {
  B200_OK = 0,
  B200_ERR_INVALID = -1,
  B200_ERR_CUDA = -2,
  B200_ERR_OOM = -3,
  B200_ERR_STATE = -4,
};

// ---- fixed geometry of the community-1 hot path (SURVEY.md section 8 constants) ----
constexpr int kChunk = 160000;       // samples per 10 s chunk @16 kHz
constexpr int kFrames = 589;         // segmentation frames per chunk
constexpr int kSincK = 251;
constexpr int kSincStride = 10;
constexpr int kSincLen = 15975;      // (160000-251)/10+1
constexpr int kPool0 = 5325;         // after MaxPool1d(3,3)
constexpr int kConv1Len = 5321;
constexpr int kPool1 = 1773;
constexpr int kConv2Len = 1769;
constexpr int kPool2 = 589;
constexpr int kHidden = 128;
constexpr int kClasses = 7;
constexpr int kSpeakers = 3;
constexpr int kFbankFrames = 998;    // 1 + (160000-400)/160
constexpr int kMel = 80;
constexpr int kEmbT = 125;           // ResNet time frames after 3 stride-2 stages
constexpr int kEmbDim = 256;
constexpr int kStatsDim = 2560;      // 256 channels x 10 freq bins

// Packed fp32x2 FMA (Blackwell FFMA2): two independent IEEE fma.rn per instruction.  The FP32 kernels here are
// limited by instruction issue / operand dispatch (ncu: FMA pipe 50 %, issue 60 %), not by the FMA lanes.
typedef unsigned long long f32x2_t;
#ifdef __CUDACC__
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(f32x2_t& d, f32x2_t a, f32x2_t b) {   // d = a * b + d (lane-wise)
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
#endif

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over one cudaMalloc'd workspace (owned by the ctx; kernels never allocate).
struct Workspace {
  char* base = nullptr;
  size_t cap = 0;
  size_t off = 0;
  void reset() { off = 0; }
  void* take(size_t bytes) {
    size_t o = align_up(off, 1024);
    if (o + bytes > cap) return nullptr;
    off = o + bytes;
    return base + o;
  }
};

}  // namespace b200
