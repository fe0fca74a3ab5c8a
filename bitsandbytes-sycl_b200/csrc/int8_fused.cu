// int8_fused.cu -- inference fast path of LLM.int8 (Linear8bitLt.forward with has_fp16_weights=False, threshold > 0):
// the whole of MatMul8bitLt.forward (python_src_quants/autograd/_functions.py:292-434) as five stream-ordered launches
// with no host synchronisation -- the reference needs one (nnz, functional.py:2546) plus torch.unique / sort and a
// dozen indexing kernels to find the outlier columns:
//
//   (K <= 4096: k_i8_row_onepass does the statistics AND the quantisation in one pass with the row in registers,
//    k_i8_fix_outliers then zeroes the outlier columns; larger K uses the two passes below)
//   k_i8_rowstats_flags   one pass over A: row absmax with |x| >= thr excluded (kgetColRowStats semantics,
//                         kernel_quant.cpp:3292-3301) + a flag per column that holds any outlier
//   k_i8_compact          flags -> ascending outlier column list idx[], position map pos[], count (device side)
//   k_i8_quant_rows       CA = rint(x * (127 / rowStat)) (kDoubleRowColQuant :3424,:3475), outlier COLUMNS zeroed
//                         (_functions.py:382), subA[i][o] = A[i][idx[o]] gathered on the way
//   k_i8_subB             subB[j][o] = half((CB[j][idx[o]] * SCB[j]) / 127)            (_functions.py:381)
//   k_igemm_tcgen05       int8 GEMM with the mm_dequant epilogue (igemm.cu) + the 16-bit outlier product
//                         half(half(dequant) + half(sum_o subA*subB)) folded into the same epilogue (:431)
// Up to 8 outlier columns ride in the GEMM epilogue; with more, the epilogue adds nothing and a follow-up kernel adds
// the whole product in one rounded piece (it reads the device-side count and exits at once in the common case).
#include "common.cuh"

namespace bnb {

constexpr int kMaxOutliers = 16;   // leading dimension of subA / subB
constexpr int kEpiOutliers = 8;    // outlier columns folded into the GEMM epilogue; the rest go through k_i8_outlier_tail

// every kernel of the fused forward is a link of a programmatic-dependent-launch chain: it lets its successor's
// CTAs be scheduled at once and waits for its predecessor before it touches memory
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = 0; lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  latch_error(cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...), "int8 fused launch");
}

__device__ __forceinline__ int quant_s8_(float x, float scale) {
  int q = __float2int_rn(__fmul_rn(x, scale));
  return max(-128, min(127, q));
}

// one warp per row; 16 bytes per lane per step
__global__ void __launch_bounds__(256) k_i8_rowstats_flags(const __half *__restrict__ A, float *__restrict__ rowStats,
                                                           unsigned char *__restrict__ colflag, float thr, int rows, int cols) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const __half *row = A + (size_t)r * cols;
  float rmax = 0.f;
  for (int c0 = lane * 8; c0 < cols; c0 += 256) {
    const uint4 raw = ld_stream_u4(row + c0);
    const __half *p = reinterpret_cast<const __half *>(&raw);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float v = fabsf(__half2float(p[j]));
      if (v >= thr) { colflag[c0 + j] = 1; v = 0.f; }
      rmax = fmaxf(rmax, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
  if (lane == 0) rowStats[r] = rmax;
}

// single CTA: ordered compaction of the column flags
__global__ void __launch_bounds__(1024) k_i8_compact(unsigned char *__restrict__ colflag, int *__restrict__ idx,
                                                     short *__restrict__ pos, int *__restrict__ count, int cols, int idx_cap) {
  pdl_enter();
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < cols; c0 += 1024) {
    const int c = c0 + tid;
    const int f = (c < cols && colflag[c]) ? 1 : 0;
    if (f) colflag[c] = 0;                 // self-cleaning: the flags are all zero again for the next forward
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const int in_warp = __popc(m & ((1u << lane) - 1));
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    if (c < cols) {
      const int p = before + in_warp;
      pos[c] = f ? (short)min(p, 32767) : (short)-1;
      if (f && p < idx_cap) idx[p] = c;
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 32; w++) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  if (tid == 0) *count = s_base;
}

__global__ void __launch_bounds__(256) k_i8_quant_rows(const __half *__restrict__ A, const float *__restrict__ rowStats,
                                                       const short *__restrict__ pos, signed char *__restrict__ CA,
                                                       __half *__restrict__ subA, int rows, int cols) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const __half *row = A + (size_t)r * cols;
  const float scale = __fdiv_rn(127.0f, rowStats[r]);
  for (int c0 = lane * 8; c0 < cols; c0 += 256) {
    const uint4 raw = ld_stream_u4(row + c0);
    const uint4 praw = __ldg(reinterpret_cast<const uint4 *>(pos + c0));
    const __half *p = reinterpret_cast<const __half *>(&raw);
    const short *ps = reinterpret_cast<const short *>(&praw);
    uint32_t q[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int v = 0;
      if (ps[j] < 0) v = quant_s8_(__half2float(p[j]), scale);
      else if (ps[j] < kMaxOutliers) subA[(size_t)r * kMaxOutliers + ps[j]] = p[j];
      q[j >> 2] |= (uint32_t)(v & 0xFF) << (8 * (j & 3));
    }
    *reinterpret_cast<uint2 *>(CA + (size_t)r * cols + c0) = make_uint2(q[0], q[1]);
  }
}

// K <= NV * 256: the whole row lives in registers (NV 128-bit loads in flight per lane), so statistics and quantisation
// are ONE pass over A.  Outlier COLUMNS are not known yet (they depend on every row): CA is written for all columns and
// k_i8_fix_outliers zeroes the (few) outlier columns afterwards.
// rint + saturate to int8 in one instruction (same result as quant_s8_: rint, then clamp; NaN -> 0)
__device__ __forceinline__ uint32_t f2s8_sat_(float x) {
  int q;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(x));
  return (uint32_t)q;
}
template <int NV>
__global__ void __launch_bounds__(128) k_i8_row_onepass(const __half *__restrict__ A, float *__restrict__ rowStats,
                                                        unsigned char *__restrict__ colflag, signed char *__restrict__ CA,
                                                        float thr, int rows, int cols) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);      // four rows per CTA: 1024 small CTAs backfill the SMs evenly
  if (r >= rows) return;
  const __half *row = A + (size_t)r * cols;
  uint4 raw[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int c0 = (i * 32 + lane) * 8;
    raw[i] = c0 < cols ? ld_stream_u4(row + c0) : make_uint4(0, 0, 0, 0);
  }
  float rmax = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    // |x| of eight halves in four packed operations, their maximum in three more (exact in fp16); only a chunk that
    // holds an outlier (max >= thr, compared in fp32 like the reference) takes the per-element path
    const uint32_t a0 = raw[i].x & 0x7FFF7FFFu, a1 = raw[i].y & 0x7FFF7FFFu, a2 = raw[i].z & 0x7FFF7FFFu, a3 = raw[i].w & 0x7FFF7FFFu;
    const __half2 m01 = __hmax2(*reinterpret_cast<const __half2 *>(&a0), *reinterpret_cast<const __half2 *>(&a1));
    const __half2 m23 = __hmax2(*reinterpret_cast<const __half2 *>(&a2), *reinterpret_cast<const __half2 *>(&a3));
    const __half2 m = __hmax2(m01, m23);
    const float cm = fmaxf(__low2float(m), __high2float(m));
    // (NaN: __hmax2 returns the other operand, fmaxf likewise -- ignored exactly as in the per-element path)
    if (cm >= thr) {
      const int c0 = (i * 32 + lane) * 8;
      const __half *p = reinterpret_cast<const __half *>(&raw[i]);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float v = fabsf(__half2float(p[j]));
        if (v >= thr) { colflag[c0 + j] = 1; v = 0.f; }
        rmax = fmaxf(rmax, v);
      }
    } else {
      rmax = fmaxf(rmax, cm);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
  if (lane == 0) rowStats[r] = rmax;
  const float scale = __fdiv_rn(127.0f, rmax);
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int c0 = (i * 32 + lane) * 8;
    if (c0 >= cols) continue;
    const __half2 *p2 = reinterpret_cast<const __half2 *>(&raw[i]);
    uint32_t q[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float2 f0 = __half22float2(p2[2 * h]), f1 = __half22float2(p2[2 * h + 1]);
      const uint32_t b0 = f2s8_sat_(__fmul_rn(f0.x, scale)), b1 = f2s8_sat_(__fmul_rn(f0.y, scale));
      const uint32_t b2 = f2s8_sat_(__fmul_rn(f1.x, scale)), b3 = f2s8_sat_(__fmul_rn(f1.y, scale));
      q[h] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
    }
    *reinterpret_cast<uint2 *>(CA + (size_t)r * cols + c0) = make_uint2(q[0], q[1]);
  }
}

// zero the outlier columns of CA (_functions.py:382) and gather them into subA (first 16 columns)
__global__ void __launch_bounds__(256) k_i8_fix_outliers(const __half *__restrict__ A, signed char *__restrict__ CA,
                                                         __half *__restrict__ subA, const int *__restrict__ idx,
                                                         const int *__restrict__ count, int rows, int cols, int idx_cap) {
  pdl_enter();
  const int n = min(*count, idx_cap);
  if (n <= 0) return;                       // no outlier column: the GEMM epilogue does not read subA
  const int w = max(n, kMaxOutliers);       // subA is written in full (zeros past the last outlier column): no memset
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < (long)rows * w; i += (long)gridDim.x * 256) {
    const int r = (int)(i / w), o = (int)(i % w);
    if (o < n) {
      const int c = idx[o];
      CA[(size_t)r * cols + c] = 0;
      if (o < kMaxOutliers) subA[(size_t)r * kMaxOutliers + o] = A[(size_t)r * cols + c];
    } else {
      subA[(size_t)r * kMaxOutliers + o] = __float2half(0.f);
    }
  }
}

__global__ void __launch_bounds__(256) k_i8_subB(const signed char *__restrict__ CB, const float *__restrict__ SCB,
                                                 const int *__restrict__ idx, const int *__restrict__ count,
                                                 __half *__restrict__ subB, int n, int k) {
  pdl_enter();
  const int nout = min(*count, kMaxOutliers);
  const long i = (long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long)n * kMaxOutliers) return;
  const int j = (int)(i / kMaxOutliers), o = (int)(i % kMaxOutliers);
  float v = 0.f;
  if (o < nout) v = __fdiv_rn(__fmul_rn((float)CB[(size_t)j * k + idx[o]], SCB[j]), 127.0f);
  subB[i] = __float2half_rn(v);
}

// more than 8 outlier columns (rare): out[i][j] = half(out[i][j] + half(sum_o A[i][idx[o]] * dequant(CB[j][idx[o]])))
__global__ void __launch_bounds__(256) k_i8_outlier_tail(const __half *__restrict__ A, const signed char *__restrict__ CB,
                                                         const float *__restrict__ SCB, const int *__restrict__ idx,
                                                         const int *__restrict__ count, __half *__restrict__ out, int m, int n,
                                                         int k, int idx_cap) {
  pdl_enter();
  const int nout = min(*count, idx_cap);
  if (nout <= kEpiOutliers) return;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < (long)m * n; i += (long)gridDim.x * 256) {
    const int r = (int)(i / n), j = (int)(i % n);
    float u = 0.f;
    for (int o = 0; o < nout; o++) {
      const int c = idx[o];
      const float b = __half2float(__float2half_rn(__fdiv_rn(__fmul_rn((float)CB[(size_t)j * k + c], SCB[j]), 127.0f)));
      u = __fmaf_rn(__half2float(A[(size_t)r * k + c]), b, u);
    }
    out[i] = __float2half_rn(__fadd_rn(__half2float(out[i]), __half2float(__float2half_rn(u))));
  }
}

int igemm_rowmajor_dequant_outliers_fp16(int m, int n, int k, const signed char *A, const signed char *B, const float *rowStats,
                                         const float *colStats, const __half *bias, __half *out, const __half *subA,
                                         const __half *subB, const int *count);

// A fp16 [m,k]; CB int8 [n,k] row-major; SCB fp32[n]; bias fp16[n] or null; out fp16 [m,n].
// workspace (caller-owned, device): CA int8 [m,k], SCA fp32[m], colflag u8[k], pos i16[k], idx i32[idx_cap], count i32[1],
// subA fp16 [m,16], subB fp16 [n,16].
int int8_linear_fused(const __half *A, const signed char *CB, const float *SCB, const __half *bias, __half *out, float thr,
                      int m, int n, int k, signed char *CA, float *SCA, unsigned char *colflag, short *pos, int *idx,
                      int idx_cap, int *count, __half *subA, __half *subB) {
  if (m <= 0 || n <= 0) return 0;
  if (k <= 0 || (k % 16) != 0 || thr <= 0.f || idx_cap < kMaxOutliers || (reinterpret_cast<uintptr_t>(A) % 16) != 0 ||
      (reinterpret_cast<uintptr_t>(CA) % 16) != 0 || (reinterpret_cast<uintptr_t>(pos) % 16) != 0)
    return 1;
  cudaStream_t st = current_stream();
  // colflag must be all zero on entry: the caller allocates it zeroed once, k_i8_compact clears what it consumed
  const dim3 rows8((unsigned)ceil_div(m, 8)), b256(256);
  if (k <= 4096) {
    // one pass over A: the row stays in registers between the statistics and the quantisation
    const dim3 rows4((unsigned)((m + 3) / 4)), b128(128);
    if (k <= 1024) launch_pdl(k_i8_row_onepass<4>, rows4, b128, st, A, SCA, colflag, CA, thr, m, k);
    else if (k <= 2048) launch_pdl(k_i8_row_onepass<8>, rows4, b128, st, A, SCA, colflag, CA, thr, m, k);
    else launch_pdl(k_i8_row_onepass<16>, rows4, b128, st, A, SCA, colflag, CA, thr, m, k);
    launch_pdl(k_i8_compact, dim3(1), dim3(1024), st, colflag, idx, pos, count, k, idx_cap);
    launch_pdl(k_i8_fix_outliers, dim3(kNumSMs), b256, st, A, CA, subA, idx, count, m, k, idx_cap);
  } else {
    launch_pdl(k_i8_rowstats_flags, rows8, b256, st, A, SCA, colflag, thr, m, k);
    launch_pdl(k_i8_compact, dim3(1), dim3(1024), st, colflag, idx, pos, count, k, idx_cap);
    latch_error(cudaMemsetAsync(subA, 0, (size_t)m * kMaxOutliers * sizeof(__half), st), "int8 fused memset");
    launch_pdl(k_i8_quant_rows, rows8, b256, st, A, SCA, pos, CA, subA, m, k);
  }
  launch_pdl(k_i8_subB, dim3((unsigned)ceil_div_ll((long)n * kMaxOutliers, 256)), b256, st, CB, SCB, idx, count, subB, n, k);
  check_launch("int8 fused quantisation");
  const int rc = igemm_rowmajor_dequant_outliers_fp16(m, n, k, CA, CB, SCA, SCB, bias, out, subA, subB, count);
  if (rc != 0) return rc;
  launch_pdl(k_i8_outlier_tail, dim3(2 * kNumSMs), b256, st, A, CB, SCB, idx, count, out, m, n, k, idx_cap);
  check_launch("int8 fused outlier tail");
  return 0;
}

}  // namespace bnb
