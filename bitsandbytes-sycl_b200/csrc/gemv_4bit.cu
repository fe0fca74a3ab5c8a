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
#include <type_traits>

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
  int bs_shift, bs2_shift;   // log2(blocksize), log2(blocksize2)
  const void *x;             // [batch, K] T
  const unsigned char *B;    // [N, K/2]
  const float *absmax;       // fp32 [N*K/blocksize]            (plain)
  const unsigned char *qabsmax;  // u8 [N*K/blocksize]          (nested)
  const float *absmax2;      // fp32 [ceil(nblocks/blocksize2)] (nested)
  const float *code2;        // fp32[256]                       (nested)
  float offset;
  const float *code;         // fp32[16]
  void *out;                 // [batch, N] T
};

struct ChunkRegs {
  uint2 w[4][2];   // [step][row half] packed weights, 16 nibbles each
  float am[4][2];  // de-nested absmax per step / row half
};

// VEC4: blocksize == 64 and K % 256 == 0 -> the four absmax entries of a 256-element chunk are one
// aligned 32-bit (nested, uint8) or 128-bit (fp32) load and share one absmax2 entry.
template <bool NESTED, bool VEC4>
__device__ __forceinline__ void load_chunk(ChunkRegs &r, const GemvArgs &a, int k0, const unsigned char *const (&wrow)[2],
                                           const long (&ebase)[2], const float *s_code2) {
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = k0 + s * 64;
#pragma unroll
    for (int h = 0; h < 2; h++) r.w[s][h] = (k < a.K) ? ld_stream_u2(wrow[h] + (k >> 1)) : make_uint2(0, 0);
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (VEC4) {
      const long blk = (ebase[h] + k0) >> 6;  // multiple of 4
      if (NESTED) {
        const uint32_t q4 = __ldg(reinterpret_cast<const uint32_t *>(a.qabsmax + blk));
        const float am2 = __ldg(a.absmax2 + (blk >> a.bs2_shift));
#pragma unroll
        for (int s = 0; s < 4; s++)
          r.am[s][h] = __fadd_rn(__fmul_rn(s_code2[(q4 >> (8 * s)) & 0xFFu], am2), a.offset);
      } else {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(a.absmax + blk));
        r.am[0][h] = v.x; r.am[1][h] = v.y; r.am[2][h] = v.z; r.am[3][h] = v.w;
      }
    } else {
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const int k = k0 + s * 64;
        float v = 0.0f;
        if (k < a.K) {
          const long blk = (ebase[h] + k) >> a.bs_shift;
          if (NESTED) v = __fadd_rn(__fmul_rn(s_code2[a.qabsmax[blk]], __ldg(a.absmax2 + (blk >> a.bs2_shift))), a.offset);
          else v = __ldg(a.absmax + blk);
        }
        r.am[s][h] = v;
      }
    }
  }
}

template <typename T, bool NESTED, bool VEC4>
__global__ void __launch_bounds__(kGemvThreads) k_gemv4_mma(const GemvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char *s_lut = smem;                                              // 64 KB
  float *s_red = reinterpret_cast<float *>(smem + kGemvLutBytes);           // [8 warps][16][8]
  float *s_code2 = s_red + kGemvWarps * 128;                                // [256]
  uint32_t *s_codeT = reinterpret_cast<uint32_t *>(s_code2 + 256);          // [16] code rounded to T (low 16 bits)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const long row_lo = min(row0 + g, a.N - 1), row_hi = min(row0 + g + 8, a.N - 1);
  const int nchunks = (a.K + kChunkK - 1) / kChunkK;
  const unsigned char *const wrow[2] = {a.B + row_lo * (a.K >> 1) + t * 8, a.B + row_hi * (a.K >> 1) + t * 8};
  const long ebase[2] = {row_lo * a.K, row_hi * a.K};

  if (NESTED) s_code2[threadIdx.x] = a.code2[threadIdx.x];  // 256 threads
  if (threadIdx.x < 16) s_codeT[threadIdx.x] = MmaT<T>::pack(a.code[threadIdx.x], 0.0f) & 0xFFFFu;
  __syncthreads();

  // first chunk's loads go out before the table is built so the two overlap
  ChunkRegs cur, nxt;
  int c = warp;
  if (c < nchunks) load_chunk<NESTED, VEC4>(cur, a, c * kChunkK, wrow, ebase, s_code2);

  // byte -> {T(code[hi nibble]), T(code[lo nibble])}, one copy per lane (bank == lane): each warp fills
  // 32 entries with 128-byte conflict-free stores
#pragma unroll 4
  for (int i = 0; i < 32; i++) {
    const int e = i * kGemvWarps + warp;
    const uint32_t v = s_codeT[e >> 4] | (s_codeT[e & 15] << 16);
    *reinterpret_cast<uint32_t *>(s_lut + e * 256 + lane * 4) = v;
  }
  __syncthreads();

  const uint32_t lane4 = lane * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const T *xrow = reinterpret_cast<const T *>(a.x) + (long)min(g, a.batch - 1) * a.K + t * 16;
  const bool has_x = g < a.batch;

  for (; c < nchunks; c += kGemvWarps) {
    const int cn = c + kGemvWarps;
    if (cn < nchunks) load_chunk<NESTED, VEC4>(nxt, a, cn * kChunkK, wrow, ebase, s_code2);
    const int k0 = c * kChunkK;
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int k = k0 + s * 64;
      if (k < a.K) {
        uint4 xa = make_uint4(0, 0, 0, 0), xb = make_uint4(0, 0, 0, 0);
        if (has_x) {
          const uint4 *xp = reinterpret_cast<const uint4 *>(xrow + k);
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
  return K > 0 && (K % 64 == 0) && ldb == K / 2 && blocksize >= 64 && ((blocksize & (blocksize - 1)) == 0) &&
         (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(B) % 8 == 0);
}

static int ilog2(int v) { int s = 0; while ((1 << s) < v) s++; return s; }

template <typename T, bool NESTED, bool VEC4>
static void launch_mma_inst(const GemvArgs &a) {
  static bool attr_set[64] = {false};
  const size_t smem = kGemvLutBytes + kGemvWarps * 128 * sizeof(float) + 256 * sizeof(float) + 16 * sizeof(uint32_t);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_set[dev]) {
    latch_error(cudaFuncSetAttribute(k_gemv4_mma<T, NESTED, VEC4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                "gemv smem attr");
    attr_set[dev] = true;
  }
  k_gemv4_mma<T, NESTED, VEC4><<<ceil_div(a.N, 16), kGemvThreads, smem, current_stream()>>>(a);
  check_launch("gemv_4bit (mma)");
}

template <typename T, bool NESTED>
static void launch_mma(GemvArgs a, int blocksize2) {
  a.bs_shift = ilog2(a.blocksize);
  a.bs2_shift = ilog2(blocksize2 > 0 ? blocksize2 : 1);
  const bool vec4 = a.blocksize == 64 && (a.K % 256 == 0) &&
                    (NESTED ? (reinterpret_cast<uintptr_t>(a.qabsmax) % 4 == 0) : (reinterpret_cast<uintptr_t>(a.absmax) % 16 == 0));
  if (vec4) launch_mma_inst<T, NESTED, true>(a);
  else launch_mma_inst<T, NESTED, false>(a);
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
      if (std::is_same<T, __half>::value) launch_mma<__half, false>(a, 0);
      else launch_mma<__nv_bfloat16, false>(a, 0);
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
  if (n < 1 || n > 8 || blocksize2 <= 0 || (blocksize2 & (blocksize2 - 1)) != 0 || !fast_path_ok(k, ldb, blocksize, A, B)) {
    latch_error(cudaErrorInvalidValue, "gemv_4bit_nested: needs 1<=n<=8, K%64==0, ldb==K/2, blocksize%64==0, aligned A/B");
    return;
  }
  GemvArgs a{};
  a.N = m; a.K = k; a.batch = n; a.blocksize = blocksize;
  a.x = A; a.B = B; a.qabsmax = qabsmax; a.absmax2 = absmax2; a.code2 = code2; a.offset = offset;
  a.code = datatype; a.out = out;
  launch_mma<T, true>(a, blocksize2);
}

template void gemv_4bit<float>(int, int, int, const float *, const unsigned char *, const float *, const float *, float *, int, int, int, int);
template void gemv_4bit<__half>(int, int, int, const __half *, const unsigned char *, const float *, const float *, __half *, int, int, int, int);
template void gemv_4bit<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const float *, const float *, __nv_bfloat16 *, int, int, int, int);
template void gemv_4bit_nested<__half>(int, int, int, const __half *, const unsigned char *, const unsigned char *, const float *, const float *, float, const float *, __half *, int, int, int, int, int);
template void gemv_4bit_nested<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const unsigned char *, const float *, const float *, float, const float *, __nv_bfloat16 *, int, int, int, int, int);

}  // namespace bnb
