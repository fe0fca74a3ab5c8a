// gemv_4bit.cu -- K3: batch-1 (up to 8) 4-bit GEMV, out[r] = sum_k x[k] * code[q(r,k)] * absmax[(r*K+k)/bs].
//
// Replaces kgemm_4bit_inference_naive (reference sycl/sycl_code/kernel_gemm.cpp:1273-1388, launcher
// op_gemm.cpp:893-929: one warp per row, 4 rows per CTA, scalar fp math).
//
// B200 design (DESIGN.md "K3"): the kernel is a pure HBM stream of the packed weight (0.5 B/element);
// at 6.4 TB/s an SM has < 3 issue slots per weight element, so the per-element work is moved off the
// FP32 pipe:
//   * dequant = ONE shared-memory lookup per packed BYTE: a 256-entry table, replicated per lane
//     (bank = lane, conflict-free), returns {T(code[hi]), T(code[lo])} as a ready T x2 register; the
//     lookup address is formed by ONE PRMT (byte << 8 | lane*4);
//   * the multiply-accumulate runs on the tensor pipe (mma.sync m16n8k16, fp32 accumulate) with the
//     16x16 weight fragment built straight from those lookups -- 8 weight elements per lane per MMA;
//   * absmax is applied to the fp32 accumulator once per 64-element block (4 MMAs), in fp32 --
//     code*absmax is never rounded to T, so the result is closer to the exact dot product than the
//     reference's T-arithmetic chain;
//   * a CTA owns 16 output rows, its 8 warps split K in 256-element chunks with register
//     double-buffering (8 x 64-bit loads per lane in flight per chunk), deterministic smem reduction.
// The MMA's n dimension carries the batch (1..8 activations rows) at no extra cost.
#include "common.cuh"

namespace bnb {

constexpr int kGemvWarps = 8;
constexpr int kGemvThreads = kGemvWarps * 32;
constexpr int kGemvLutBytes = 256 * 256;  // entry stride 256 B (128 B used: one word per lane)
constexpr int kChunkK = 256;              // K elements per warp chunk (4 blocks of 64)

template <typename T> struct MmaT;
template <> struct MmaT<__half> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
  }
};
template <> struct MmaT<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
  }
};

struct GemvArgs {
  int N, K, batch, blocksize;
  const void *x;             // [batch, K] T
  const unsigned char *B;    // [N, K/2]
  const float *absmax;       // fp32 [N*K/blocksize]            (plain)
  const unsigned char *qabsmax;  // u8 [N*K/blocksize]          (nested)
  const float *absmax2;      // fp32 [ceil(nblocks/blocksize2)] (nested)
  const float *code2;        // fp32[256]                       (nested)
  float offset;
  int blocksize2;
  const float *code;         // fp32[16]
  void *out;                 // [batch, N] T
};

struct ChunkRegs {
  uint2 w[4][2];   // [step][row half] packed weights, 16 nibbles each
  float am[4][2];  // de-nested absmax per step / row half
};

template <bool NESTED>
__device__ __forceinline__ float load_absmax(const GemvArgs &a, long blk, const float *s_code2) {
  if (NESTED) {
    float v = __fmul_rn(s_code2[a.qabsmax[blk]], __ldg(a.absmax2 + blk / a.blocksize2));
    return __fadd_rn(v, a.offset);
  }
  return __ldg(a.absmax + blk);
}

template <bool NESTED>
__device__ __forceinline__ void load_chunk(ChunkRegs &r, const GemvArgs &a, int k0, long row_lo, long row_hi,
                                           int t, const float *s_code2) {
  const long rows[2] = {row_lo, row_hi};
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = k0 + s * 64;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (k < a.K) {
        r.w[s][h] = ld_stream_u2(a.B + (rows[h] * a.K + k) / 2 + t * 8);
      } else {
        r.w[s][h] = make_uint2(0, 0);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = k0 + s * 64;
#pragma unroll
    for (int h = 0; h < 2; h++)
      r.am[s][h] = (k < a.K) ? load_absmax<NESTED>(a, (rows[h] * a.K + k) / a.blocksize, s_code2) : 0.0f;
  }
}

template <typename T, bool NESTED>
__global__ void __launch_bounds__(kGemvThreads) k_gemv4_mma(const GemvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char *s_lut = smem;                                              // 64 KB
  float *s_red = reinterpret_cast<float *>(smem + kGemvLutBytes);           // [8 warps][16][8]
  float *s_code2 = s_red + kGemvWarps * 128;                                // [256] (nested only)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const long row_lo = min(row0 + g, a.N - 1), row_hi = min(row0 + g + 8, a.N - 1);
  const int nchunks = (a.K + kChunkK - 1) / kChunkK;

  if (NESTED) {
    s_code2[threadIdx.x] = a.code2[threadIdx.x];  // 256 threads
    __syncthreads();
  }
  // first chunk's loads go out before the table is built so the two overlap
  ChunkRegs cur, nxt;
  int c = warp;
  if (c < nchunks) load_chunk<NESTED>(cur, a, c * kChunkK, row_lo, row_hi, t, s_code2);

  {  // byte -> {T(code[hi nibble]), T(code[lo nibble])}, one copy per lane
    const int e = threadIdx.x;
    const uint32_t v = MmaT<T>::pack(__ldg(a.code + (e >> 4)), __ldg(a.code + (e & 15)));
    uint4 *dst = reinterpret_cast<uint4 *>(s_lut + e * 256);
    const uint4 v4 = make_uint4(v, v, v, v);
#pragma unroll
    for (int i = 0; i < 8; i++) dst[i] = v4;
  }
  __syncthreads();

  const uint32_t lane4 = lane * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const T *xrow = reinterpret_cast<const T *>(a.x) + (long)min(g, a.batch - 1) * a.K;
  const bool has_x = g < a.batch;

  for (; c < nchunks; c += kGemvWarps) {
    const int cn = c + kGemvWarps;
    if (cn < nchunks) load_chunk<NESTED>(nxt, a, cn * kChunkK, row_lo, row_hi, t, s_code2);
    const int k0 = c * kChunkK;
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int k = k0 + s * 64;
      if (k < a.K) {
        uint4 xa = make_uint4(0, 0, 0, 0), xb = make_uint4(0, 0, 0, 0);
        if (has_x) {
          const uint4 *xp = reinterpret_cast<const uint4 *>(xrow + k + t * 16);
          xa = __ldg(xp);
          xb = __ldg(xp + 1);
        }
        const uint32_t xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const uint32_t wl[2] = {cur.w[s][0].x, cur.w[s][0].y};
        const uint32_t wh[2] = {cur.w[s][1].x, cur.w[s][1].y};
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          // bytes 2i, 2i+1 of the lane's 8 bytes; selector 0x55b4: byte0 = lane*4, byte1 = packed byte b
          const uint32_t src_l = wl[i >> 1], src_h = wh[i >> 1];
          const uint32_t sel0 = (i & 1) ? 0x5524u : 0x5504u, sel1 = (i & 1) ? 0x5534u : 0x5514u;
          uint32_t af[4];
          af[0] = *reinterpret_cast<const uint32_t *>(s_lut + __byte_perm(src_l, lane4, sel0));
          af[2] = *reinterpret_cast<const uint32_t *>(s_lut + __byte_perm(src_l, lane4, sel1));
          af[1] = *reinterpret_cast<const uint32_t *>(s_lut + __byte_perm(src_h, lane4, sel0));
          af[3] = *reinterpret_cast<const uint32_t *>(s_lut + __byte_perm(src_h, lane4, sel1));
          MmaT<T>::mma(d, af, xr[2 * i], xr[2 * i + 1]);
        }
        acc[0] = __fmaf_rn(d[0], cur.am[s][0], acc[0]);
        acc[1] = __fmaf_rn(d[1], cur.am[s][0], acc[1]);
        acc[2] = __fmaf_rn(d[2], cur.am[s][1], acc[2]);
        acc[3] = __fmaf_rn(d[3], cur.am[s][1], acc[3]);
      }
    }
    cur = nxt;
  }

  // deterministic cross-warp reduction: s_red[warp][row][col]
  float *mine = s_red + warp * 128;
  mine[g * 8 + 2 * t] = acc[0];
  mine[g * 8 + 2 * t + 1] = acc[1];
  mine[(g + 8) * 8 + 2 * t] = acc[2];
  mine[(g + 8) * 8 + 2 * t + 1] = acc[3];
  __syncthreads();
  if (threadIdx.x < 16 * a.batch) {
    const int row = threadIdx.x & 15, col = threadIdx.x >> 4;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kGemvWarps; w++) sum += s_red[w * 128 + row * 8 + col];
    if (row0 + row < a.N) reinterpret_cast<T *>(a.out)[(long)col * a.N + row0 + row] = from_float<T>(sum);
  }
}

// ------------------------------------------------------------------------------------------------
// generic path: any K / ldb / blocksize / dtype (incl. fp32).  One warp per row, 16 packed bytes per
// lane per step, fp32 math.  Semantics of the tail follow kernel_gemm.cpp:1312-1366: bytes at index
// >= K/2 read as 0x77 and activations past K as 0.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_gemv4_simple(int M, int K, const T *__restrict__ A,
                                                      const unsigned char *__restrict__ B,
                                                      const float *__restrict__ absmax,
                                                      const float *__restrict__ datatype, T *__restrict__ out,
                                                      int ldb, int blocksize) {
  __shared__ float s_code[16];
  if (threadIdx.x < 16) s_code[threadIdx.x] = datatype[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  const long offB = (long)ldb * row;
  float acc = 0.f;
  for (int inner = lane * 32; inner < K; inner += 32 * 32) {
    const float am = absmax[(2 * offB + inner) / blocksize];
    float part = 0.f;
    for (int j = 0; j < 16; j++) {
      const int kb = inner / 2 + j;
      const unsigned char byte = (kb < K / 2) ? B[offB + kb] : (unsigned char)0x77;
      const int k = inner + 2 * j;
      const float a0 = (k < K) ? to_float<T>(A[k]) : 0.f;
      const float a1 = (k + 1 < K) ? to_float<T>(A[k + 1]) : 0.f;
      part = __fmaf_rn(a0, s_code[byte >> 4], part);
      part = __fmaf_rn(a1, s_code[byte & 15], part);
    }
    acc = __fmaf_rn(part, am, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = from_float<T>(acc);
}

static bool fast_path_ok(int K, int ldb, int blocksize, const void *A, const void *B) {
  return K > 0 && (K % 64 == 0) && ldb == K / 2 && blocksize >= 64 && (blocksize % 64 == 0) &&
         (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(B) % 8 == 0);
}

template <typename T, bool NESTED>
static void launch_mma(const GemvArgs &a) {
  static bool attr_set = false;
  const size_t smem = kGemvLutBytes + kGemvWarps * 128 * sizeof(float) + (NESTED ? 256 * sizeof(float) : 0);
  if (!attr_set) {
    latch_error(cudaFuncSetAttribute(k_gemv4_mma<T, NESTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                "gemv smem attr");
    attr_set = true;
  }
  k_gemv4_mma<T, NESTED><<<ceil_div(a.N, 16), kGemvThreads, smem, current_stream()>>>(a);
  check_launch("gemv_4bit (mma)");
}

template <typename T>
void gemv_4bit(int m, int n, int k, const T *A, const unsigned char *B, const float *absmax, const float *datatype,
               T *out, int lda, int ldb, int ldc, int blocksize) {
  (void)lda; (void)ldc;
  if (m <= 0 || k <= 0) return;
  if (n != 1 || blocksize <= 0) { latch_error(cudaErrorInvalidValue, "gemv_4bit: n must be 1"); return; }
  if (sizeof(T) == 2 && fast_path_ok(k, ldb, blocksize, A, B)) {
    GemvArgs a{};
    a.N = m; a.K = k; a.batch = 1; a.blocksize = blocksize;
    a.x = A; a.B = B; a.absmax = absmax; a.code = datatype; a.out = out;
    if (sizeof(T) == 2) {
      if (std::is_same<T, __half>::value) launch_mma<__half, false>(a);
      else launch_mma<__nv_bfloat16, false>(a);
    }
    return;
  }
  k_gemv4_simple<T><<<ceil_div(m, 4), 128, 0, current_stream()>>>(m, k, A, B, absmax, datatype, out, ldb, blocksize);
  check_launch("gemv_4bit (generic)");
}

template <typename T>
void gemv_4bit_nested(int m, int n, int k, const T *A, const unsigned char *B, const unsigned char *qabsmax,
                      const float *absmax2, const float *code2, float offset, const float *datatype, T *out,
                      int lda, int ldb, int ldc, int blocksize, int blocksize2) {
  (void)lda; (void)ldc;
  if (m <= 0 || k <= 0) return;
  if (n < 1 || n > 8 || blocksize2 <= 0 || !fast_path_ok(k, ldb, blocksize, A, B)) {
    latch_error(cudaErrorInvalidValue, "gemv_4bit_nested: needs 1<=n<=8, K%64==0, ldb==K/2, blocksize%64==0, aligned A/B");
    return;
  }
  GemvArgs a{};
  a.N = m; a.K = k; a.batch = n; a.blocksize = blocksize;
  a.x = A; a.B = B; a.qabsmax = qabsmax; a.absmax2 = absmax2; a.code2 = code2; a.offset = offset;
  a.blocksize2 = blocksize2; a.code = datatype; a.out = out;
  launch_mma<T, true>(a);
}

template void gemv_4bit<float>(int, int, int, const float *, const unsigned char *, const float *, const float *, float *, int, int, int, int);
template void gemv_4bit<__half>(int, int, int, const __half *, const unsigned char *, const float *, const float *, __half *, int, int, int, int);
template void gemv_4bit<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const float *, const float *, __nv_bfloat16 *, int, int, int, int);
template void gemv_4bit_nested<__half>(int, int, int, const __half *, const unsigned char *, const unsigned char *, const float *, const float *, float, const float *, __half *, int, int, int, int, int);
template void gemv_4bit_nested<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const unsigned char *, const float *, const float *, float, const float *, __nv_bfloat16 *, int, int, int, int, int);

}  // namespace bnb
