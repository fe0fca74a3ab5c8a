// common.cuh -- shared host/device helpers for libbitsandbytes_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bnb {

// DataType_t of the reference (sycl/sycl_code/ops.h:87-92)
enum DataType_t { General8bit = 0, FP4 = 1, NF4 = 2 };
// Transform_t subset (sycl/sycl_code/ops.h:78-85)
enum Format_t { COL32 = 0, COL_TURING = 1, COL_AMPERE = 2 };

constexpr int kNumSMs = 148;  // B200

// ---- per-thread launch stream + latched error (the reference ABI has neither)
cudaStream_t current_stream();
void set_current_stream(cudaStream_t s);
void latch_error(cudaError_t e, const char *where);
inline void check_launch(const char *where) { latch_error(cudaGetLastError(), where); }
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of on every launch (c_api.cu)
void ensure_max_dynamic_smem(const void *kernel, int bytes, const char *where);

// ---- 16-bit conversions (all round-to-nearest-even, one rounding)
template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// streaming (read-once / write-once) 128-bit global accesses that stay out of L1
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// 256-bit variant (sm_100+): one lane fetches a whole 32-byte sector
__device__ __forceinline__ void ld_stream_u8(uint32_t (&r)[8], const void *p) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ uint2 ld_stream_u2(const void *p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u1(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void *p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(void *p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_u1(void *p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace bnb
