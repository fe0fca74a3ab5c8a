// gemm_4bit.cu -- K4: fused batch>1 4-bit GEMM, out[b, n] = sum_k A[b,k] * w(n,k) (+ bias[n]).
//
// Replaces the reference's batch>1 route MatMul4Bit.forward (python_src_quants/autograd/_functions.py:490-518):
// dequantize_4bit (kDequantizeBlockwise, kernel_quant.cpp:1370-1471) writes the whole weight to HBM in T, F.linear
// reads it back (2 + 2 bytes per weight on top of the 0.5 that are needed).  Here the packed weight is read once and
// fed to the 5th-generation tensor cores, swap-AB: the 128-row weight tile is the UMMA M operand (written by dequant
// warps straight into TENSOR MEMORY), the activations [batch, 64] are the N operand (TMA, zero-filled beyond `batch`),
// D[128 weight rows, batch] accumulates in TMEM (fp32), tcgen05.mma kind::f16, M=128, N=round16(batch), K=16.
//
//   batch <= 32  gemm_4bit_small.cuh: unscaled code values as operand, one TMEM accumulator per quantisation block,
//                fp32 absmax applied to the block sums (more accurate than dequantize-then-matmul)
//   batch > 32   gemm_4bit_wide.cuh:  the reference's operand w = T(fp32 code[q] * fp32 absmax), one accumulator
//
// This file holds what they share: the pair packers, the split-K workspace and finalize kernel, and the dispatcher.
// (Round 1's kernel -- 16-entry table, nine instructions per weight, operand through shared memory -- lived here; it is
// in the history at 84993fb and before.)
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

#include <map>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bnb {

namespace g4 {
constexpr int TK = 64;             // K elements per stage (one quantisation block of blocksize 64)
constexpr int kMaxNB = 256;

// Where the result goes.  ldo = elements between consecutive batch rows of `out` (N for a plain call).  N-sharded stacks
// (bnb_b200/parallel.py) pass the base of this rank's column slice inside the gathered [batch, N_total] buffer, ldo =
// N_total, and the same address in every peer's copy of the buffer (NVLink peer mappings): the epilogue then IS the
// all-gather, as in the batch-1 GEMV (cgemm_4bit_inference_nested_push_*).
struct OutSpec {
  long ldo;
  int npeers;
  void *peer[7];
};

template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
// split-K: out[b, n] = T(sum_s ws[s, b, n] + bias[n]), summed in split order (deterministic)
template <typename T>
__global__ void __launch_bounds__(256) k_gemm4_finalize(const float *__restrict__ ws, const T *__restrict__ bias, T *__restrict__ out,
                                                        int splits, int batch, int N, const OutSpec o) {
  const size_t total = (size_t)batch * N;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    float s = 0.f;
    for (int k = 0; k < splits; k++) s = __fadd_rn(s, ws[(size_t)k * total + i]);
    const size_t b = i / N, n = i % N;
    if (bias != nullptr) s = __fadd_rn(s, to_float<T>(bias[n]));
    const T v = from_float<T>(s);
    out[b * o.ldo + n] = v;
#pragma unroll
    for (int p = 0; p < 7; p++)
      if (p < o.npeers) reinterpret_cast<T *>(o.peer[p])[b * o.ldo + n] = v;
  }
}

// split-K partial sums: one workspace per (device, stream) -- launches on one stream are ordered, two streams running
// the kernel concurrently must not share it.  cudaFree of an outgrown buffer synchronises the device first.
struct WsKey { int dev; cudaStream_t st; bool operator<(const WsKey &o) const { return dev != o.dev ? dev < o.dev : st < o.st; } };
struct WsBuf { float *p; size_t bytes; };
static std::mutex g_ws_mu;
static std::map<WsKey, WsBuf> g_ws;

// *from_pool: stream capture in progress -> the buffer is a cudaMallocAsync allocation of the capturing stream (a node of
// the graph); the caller frees it with cudaFreeAsync after the finalize kernel
static float *workspace(int dev, size_t bytes, cudaStream_t st, bool *from_pool) {
  *from_pool = false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs != cudaStreamCaptureStatusNone) {
    void *p = nullptr;
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *from_pool = true;
    return static_cast<float *>(p);
  }
  std::lock_guard<std::mutex> lk(g_ws_mu);
  WsBuf &b = g_ws[WsKey{dev, st}];
  if (b.bytes >= bytes) return b.p;
  if (b.p) cudaFree(b.p);
  const size_t want = bytes < (size_t)(64u << 20) ? (size_t)(64u << 20) : bytes;
  if (cudaMalloc(&b.p, want) != cudaSuccess) { b.p = nullptr; b.bytes = 0; cudaGetLastError(); return nullptr; }
  b.bytes = want;
  return b.p;
}
}  // namespace g4

static int ilog2_(int v) { int s = 0; while ((1 << s) < v) s++; return s; }

#include "gemm_4bit_small.cuh"
#include "gemm_4bit_wide.cuh"

// returns 0 ok, 1 shape not taken by the fused kernel (caller uses dequantize + matmul), 2 error
template <typename T>
int gemm_4bit(int batch, int N, int K, const T *A, const unsigned char *B, const float *absmax, const float *datatype,
              const T *bias, T *out, int blocksize, long ldo, void *const *peer_outs, int npeers) {
  using namespace g4;
  if (npeers < 0 || npeers > 7 || (ldo != 0 && ldo < N)) { latch_error(cudaErrorInvalidValue, "gemm_4bit: bad output spec"); return 2; }
  OutSpec ospec{};
  ospec.ldo = ldo ? ldo : N;
  ospec.npeers = npeers;
  for (int i = 0; i < npeers; i++) ospec.peer[i] = peer_outs[i];
  if (batch <= 0 || N <= 0) return 0;
  if (batch > kMaxNB || K < TK || (K % TK) != 0 || blocksize < 64 || (blocksize & (blocksize - 1)) != 0 ||
      (reinterpret_cast<uintptr_t>(A) % 16) != 0 || (reinterpret_cast<uintptr_t>(B) % 32) != 0 || ((K / 2) % 32) != 0)
    return 1;
  static int off = -1;
  if (off < 0) { const char *e = getenv("BNB_B200_GEMM4"); off = (e && e[0] == '0') ? 1 : 0; }
  if (off) return 1;
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 64 ? dev : 63;
  static int num_sms[64] = {0};
  if (!num_sms[dev]) cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t st = current_stream();
  static int small_off = -1;   // BNB_B200_GEMM4_SMALL=0: the wide kernel at every batch (A/B measurements)
  if (small_off < 0) { const char *e = getenv("BNB_B200_GEMM4_SMALL"); small_off = (e && e[0] == '0') ? 1 : 0; }
  if (batch <= 32 && !small_off && K / TK >= 8)
    return gemm_4bit_small<T>(batch, N, K, A, B, absmax, datatype, bias, out, ilog2_(blocksize), num_sms[dev], dev, st, ospec);
  if (K / TK < 4) return 1;
  return gemm_4bit_wide<T>(batch, N, K, A, B, absmax, datatype, bias, out, ilog2_(blocksize), num_sms[dev], dev, st, ospec);
}

template int gemm_4bit<__half>(int, int, int, const __half *, const unsigned char *, const float *, const float *, const __half *, __half *, int, long, void *const *, int);
template int gemm_4bit<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const float *, const float *, const __nv_bfloat16 *, __nv_bfloat16 *, int, long, void *const *, int);

}  // namespace bnb
