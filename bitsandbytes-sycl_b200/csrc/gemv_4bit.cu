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
#include <stdlib.h>
#include <type_traits>

#include "codebooks.cuh"
#include "common.cuh"
#include "tcgen05.cuh"

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
  int flags;                 // bit 1: phase probe (debug)
  const void *x;             // [batch, K] T
  const unsigned char *B;    // [N, K/2]
  const float *absmax;       // fp32 [N*K/blocksize]            (plain)
  const unsigned char *qabsmax;  // u8 [N*K/blocksize]          (nested)
  const float *absmax2;      // fp32 [ceil(nblocks/blocksize2)] (nested)
  const float *code2;        // fp32[256]                       (nested)
  float offset;
  const float *code;         // fp32[16]
  // 2: the host verified (cbnb_set_gemv_host_tables) that code == the NF4 table -> the kernel builds its lookup table
  // from immediates.  A small global load issued at kernel start queues behind the weight stream of the CTAs already
  // running on the SM and takes 0.6-2.5 us to come back -- on the critical path of every launch.  (Passing the
  // tables by value was measured too: kernel parameters come through the same memory system, no gain.)
  int tables_in_args;
  // several weight matrices that share x, K and the code tables in ONE launch (q/k/v, gate/up): the tile index space
  // is the concatenation of the matrices' 16-row tiles; mt[i] = first tile of matrix i (mt[nmat..4] = INT_MAX)
  int nmat;
  int mt[5];
  int mN[4];
  float moff[4];
  const unsigned char *mB[4];
  const unsigned char *mq[4];
  const float *mam2[4];
  void *mout[4];
  void *mpeer[4][7];         // MULTI + N-sharding: peer-mapped copies of each matrix's output slice
  void *out;                 // [batch, N] T
  // multi-GPU (N-sharded linear): the same output slice is also stored into the peers' copies of the full
  // output vector through NVLink peer mappings -- the all-gather happens in the GEMV epilogue
  void *peer_out[7];
  int npeers;
  // cross-GPU ordering of the gathered vectors, folded into the kernels (no barrier launch): the kernel that ends
  // a consumer group publishes sequence number seq+1 into every peer's signal slot once ALL its CTAs have stored
  // (and fenced) their slices; the kernel that starts a group waits -- after griddepcontrol.wait, before it reads
  // x -- until every peer has published >= its own seq.  seq = *epoch * ngroups + gidx; epoch is bumped once per
  // step by k_epoch_bump so that replayed CUDA graphs keep counting.
  unsigned int *sig_local;      // [world] slots on this GPU, slot p written by peer p (nullptr: no sync)
  unsigned int *sig_peer[7];    // peers' slot for this rank
  const unsigned int *epoch;
  unsigned int *cta_counter;    // per-GPU scratch, zero at rest
  int gidx, ngroups, do_signal, do_wait;
};

// host-side mirror of bnb_gemv_sync_t (include/bnb_b200.h)
struct GemvSync {
  unsigned int *sig_local;
  unsigned int *sig_peer[7];
  const unsigned int *epoch;
  unsigned int *cta_counter;
  int gidx, ngroups, do_signal, do_wait;
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
// Fast path: blocksize == 64, K % 256 == 0 (every shape of BASELINE configs 2 and 5).
// Persistent CTA per SM (512 threads = 2 groups of 8 warps; a group owns one 16-row tile at a time and its
// warps split K), hand-scheduled:
//   * the byte LUT sits at a 64 KB-ALIGNED shared address, so ONE PRMT yields the complete lookup address
//     (bytes 2,3 = table base, byte 1 = packed byte, byte 0 = lane*4) -- no add, no shift;
//   * running pointers, no tail predicates, ping-pong register buffers, loads of the next (tile, chunk)
//     issued before the current one is consumed -- also across tile boundaries;
//   * absmax de-nested right before use; per-tile reduction through a 256-thread named barrier.
// ------------------------------------------------------------------------------------------------
constexpr int kFastGroups = 2;
constexpr int kFastThreads = kFastGroups * kGemvThreads;      // 512
constexpr int kFastSmall = (kFastGroups * 2 * kGemvWarps * 128 + 256 + 16) * 4;  // s_red + s_code2 + s_codeT bytes
constexpr int kFastSmem = 2 * 65536;                         // small arrays, then the table at the next 64 KB boundary

template <bool NESTED> struct FastBuf {
  uint2 w[4][2];
  uint32_t q[2];   // nested: four uint8 absmax codes per row half
  float am2[2];    // nested: absmax2 of the 256-group
  float4 am4[2];   // plain: four fp32 absmax per row half
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

template <typename T, bool NESTED>
__global__ void __launch_bounds__(kFastThreads, 1) k_gemv4_fast(const GemvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t lut_s = (smem_base + kFastSmall + 0xFFFFu) & ~0xFFFFu;   // 64 KB aligned shared address
  unsigned char *s_lut = smem + (lut_s - smem_base);
  float *s_red = reinterpret_cast<float *>(smem);             // [groups][2 parities][8 warps][128]
  float *s_code2 = s_red + kFastGroups * 2 * kGemvWarps * 128;
  uint32_t *s_codeT = reinterpret_cast<uint32_t *>(s_code2 + 256);
  if (lut_s + 65536u > smem_base + kFastSmem) __trap();   // shared window base moved: layout no longer fits

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = warp >> 3, gw = warp & 7, gtid = tid & 255;
  const int g = lane >> 2, t = lane & 3;
  const int nchunks = a.K >> 8;
  const int kb = a.K >> 6;  // absmax blocks per row
  const int ntiles = (a.N + 15) >> 4;
  const int tile_stride = gridDim.x * kFastGroups;
  const float offset = a.offset;
  const int bs2_shift = a.bs2_shift;
  const T *xrow = reinterpret_cast<const T *>(a.x) + (size_t)min(g, a.batch - 1) * a.K + t * 16;

  const uint2 *wp0, *wp1;
  size_t blk0, blk1;
  auto set_tile = [&](int tile, int c) {
    const int row_lo = min(tile * 16 + g, a.N - 1), row_hi = min(tile * 16 + g + 8, a.N - 1);
    wp0 = reinterpret_cast<const uint2 *>(a.B + (size_t)row_lo * (a.K >> 1)) + t + c * 16;
    wp1 = reinterpret_cast<const uint2 *>(a.B + (size_t)row_hi * (a.K >> 1)) + t + c * 16;
    blk0 = (size_t)row_lo * kb + c * 4;
    blk1 = (size_t)row_hi * kb + c * 4;
  };
  // pull a whole 16-row tile (packed rows + its absmax run) into L2 with bulk prefetches: one instruction
  // per row, no registers held -- this is what keeps tens of KB per SM in flight towards HBM
  auto prefetch_tile = [&](int ptile) {
    if (gw == 0 && ptile < ntiles && !(a.flags & 1)) {
      const int prow = ptile * 16 + lane;
      const uint32_t row_bytes = (uint32_t)(a.K >> 1);
      if (lane < 16 && prow < a.N)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.B + (size_t)prow * row_bytes), "r"(row_bytes) : "memory");
      if (lane == 16) {
        const int rows = min(16, a.N - ptile * 16);
        const size_t first = (size_t)ptile * 16 * kb;
        if (NESTED) {
          const uint32_t bytes = (uint32_t)((rows * kb) & ~15);
          if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.qabsmax + first), "r"(bytes) : "memory");
        } else {
          const uint32_t bytes = (uint32_t)(rows * kb * 4);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.absmax + first), "r"(bytes) : "memory");
        }
      }
    }
  };
  auto load = [&](FastBuf<NESTED> &b) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
      b.w[s][0] = ld_stream_u2(wp0 + s * 4);
      b.w[s][1] = ld_stream_u2(wp1 + s * 4);
    }
    if (NESTED) {
      b.q[0] = __ldg(reinterpret_cast<const uint32_t *>(a.qabsmax + blk0));
      b.q[1] = __ldg(reinterpret_cast<const uint32_t *>(a.qabsmax + blk1));
      b.am2[0] = __ldg(a.absmax2 + (blk0 >> bs2_shift));
      b.am2[1] = __ldg(a.absmax2 + (blk1 >> bs2_shift));
    } else {
      b.am4[0] = __ldg(reinterpret_cast<const float4 *>(a.absmax + blk0));
      b.am4[1] = __ldg(reinterpret_cast<const float4 *>(a.absmax + blk1));
    }
    wp0 += kGemvWarps * 16;
    wp1 += kGemvWarps * 16;
    blk0 += kGemvWarps * 4;
    blk1 += kGemvWarps * 4;
  };

  if (NESTED && tid < 256) s_code2[tid] = a.code2[tid];
  if (tid < 16) s_codeT[tid] = MmaT<T>::pack(a.code[tid], 0.0f) & 0xFFFFu;

  FastBuf<NESTED> bufA, bufB;
  int tile = blockIdx.x * kFastGroups + grp, c = gw;
  prefetch_tile(tile);
  prefetch_tile(tile + tile_stride);
  if (tile < ntiles && c < nchunks) { set_tile(tile, c); load(bufA); }  // in flight while the table is built
  __syncthreads();
  {
    // byte e -> {T(code[e>>4]), T(code[e&15])}, replicated for the 32 lanes: entry stride 256 B, 128 B used.
    // lanes 8i..8i+7 of a warp write one entry (8 x 16 B = its 128 B): conflict-free 128-bit stores.
    const int j = tid & 7, esub = tid >> 3;  // 64 entries per pass
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int e = i * 64 + esub;
      const uint32_t v = s_codeT[e >> 4] | (s_codeT[e & 15] << 16);
      *reinterpret_cast<uint4 *>(s_lut + e * 256 + j * 16) = make_uint4(v, v, v, v);
    }
  }
  __syncthreads();

  const uint32_t lutlane = lut_s | (uint32_t)(lane * 4);   // PRMT operand b: bytes 0,2,3 of every lookup address
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int parity = 0;
  const bool has_x = g < a.batch;
  uint4 xa = make_uint4(0, 0, 0, 0), xb = make_uint4(0, 0, 0, 0);

  auto compute = [&](const FastBuf<NESTED> &b, int cc) {
    float am[4][2];
    if (NESTED) {
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int s = 0; s < 4; s++)
          am[s][h] = __fadd_rn(__fmul_rn(s_code2[(b.q[h] >> (8 * s)) & 0xFFu], b.am2[h]), offset);
    } else {
#pragma unroll
      for (int h = 0; h < 2; h++) { am[0][h] = b.am4[h].x; am[1][h] = b.am4[h].y; am[2][h] = b.am4[h].z; am[3][h] = b.am4[h].w; }
    }
    const uint4 *x4 = reinterpret_cast<const uint4 *>(xrow + (size_t)cc * 256);
#pragma unroll
    for (int s = 0; s < 4; s++) {
      // only the lanes that own a real activation row (g < batch) load x: the others feed ignored MMA
      // columns with whatever their registers hold, and cost no L1 write-back bandwidth
      if (has_x) { xa = __ldg(x4 + s * 8); xb = __ldg(x4 + s * 8 + 1); }
      const uint32_t xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      const uint32_t wl[2] = {b.w[s][0].x, b.w[s][0].y};
      const uint32_t wh[2] = {b.w[s][1].x, b.w[s][1].y};
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        // selector 0x76b4: result = {lutlane.b0, w.byte b, lutlane.b2, lutlane.b3} = table + byte*256 + lane*4
        const uint32_t src_l = wl[i >> 1], src_h = wh[i >> 1];
        const uint32_t sel0 = (i & 1) ? 0x7624u : 0x7604u, sel1 = (i & 1) ? 0x7634u : 0x7614u;
        uint32_t af[4];
        af[0] = lds_u32(__byte_perm(src_l, lutlane, sel0));
        af[2] = lds_u32(__byte_perm(src_l, lutlane, sel1));
        af[1] = lds_u32(__byte_perm(src_h, lutlane, sel0));
        af[3] = lds_u32(__byte_perm(src_h, lutlane, sel1));
        MmaT<T>::mma(d, af, xr[2 * i], xr[2 * i + 1]);
      }
      acc[0] = __fmaf_rn(d[0], am[s][0], acc[0]);
      acc[1] = __fmaf_rn(d[1], am[s][0], acc[1]);
      acc[2] = __fmaf_rn(d[2], am[s][1], acc[2]);
      acc[3] = __fmaf_rn(d[3], am[s][1], acc[3]);
    }
  };

  // finish a tile: deterministic cross-warp reduction inside the group (double-buffered by parity)
  auto finish_tile = [&](int done_tile) {
    float *red = s_red + ((grp * 2 + parity) * kGemvWarps) * 128;
    float *mine = red + gw * 128;
    mine[g * 8 + 2 * t] = acc[0];
    mine[g * 8 + 2 * t + 1] = acc[1];
    mine[(g + 8) * 8 + 2 * t] = acc[2];
    mine[(g + 8) * 8 + 2 * t + 1] = acc[3];
    asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
    if (gtid < 16 * a.batch) {
      const int row = gtid & 15, col = gtid >> 4;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kGemvWarps; w++) sum += red[w * 128 + row * 8 + col];
      const int r = done_tile * 16 + row;
      if (r < a.N) reinterpret_cast<T *>(a.out)[(size_t)col * a.N + r] = from_float<T>(sum);
    }
    parity ^= 1;
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
  };

  // one pipeline step: prefetch the next (tile, chunk) into `nxt`, consume `cur`
  auto step = [&](const FastBuf<NESTED> &cur, FastBuf<NESTED> &nxt) {
    int nc = c + kGemvWarps, ntile = tile;
    if (nc >= nchunks) { nc = gw; ntile = tile + tile_stride; }
    if (ntile < ntiles && nc < nchunks) {
      if (ntile != tile) { set_tile(ntile, nc); prefetch_tile(ntile + tile_stride); }
      load(nxt);
    }
    if (c < nchunks) compute(cur, c);
    if (ntile != tile) finish_tile(tile);
    tile = ntile;
    c = nc;
  };

  while (tile < ntiles) {
    step(bufA, bufB);
    if (tile >= ntiles) break;
    step(bufB, bufA);
  }
}

// ------------------------------------------------------------------------------------------------
// Block-column path (batch 1, blocksize 64, K % 256 == 0): the headline kernel.
//
// The MMA's eight output columns are not wasted on a single activation row: column j of the 16x8
// accumulator collects the partial dot product of quantisation block j of a 512-element K chunk.  The k
// slots of an MMA are spread over four blocks (lane t's slots <-> block t for "even" MMAs, block t + 4 for
// "odd" ones), and the B fragment holds x only where its column equals the block of its k slots (lane
// (g,t) feeds real activations to even MMAs iff g == t, to odd MMAs iff g == t + 4; every other lane feeds
// zeros that are set once and never change).  One accumulator therefore
// runs through all 32 MMAs of a (16-row tile, 512-K chunk) item, and afterwards every lane owns four
// distinct (row, block) partial sums: the absmax scale -- and its de-nesting -- is applied ONCE per
// (row, block) by exactly one lane (no 4x redundancy, no per-block accumulator reset).
// Per item and lane: 32 x (4 PRMT + 4 LDS + 1 HMMA) + 8 LDG.128 (weights) + 16 one-wavefront LDS.128 (x)
// + 4 de-nests = ~1.4 instructions per weight element, against ~3.3 for the per-block-accumulator kernels.
//
// A persistent CTA per SM owns a contiguous range of 16-row tiles; its 16 warps take (tile, chunk) items
// round-robin, each warp keeps its weights in a register ring that is refilled half an item at a time
// (3-4 KB per warp, ~56 KB per SM in flight towards HBM), partial sums of a tile meet in shared memory in a
// fixed order (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int kBcXPitch = 144;                 // bytes per 64-element block of x in shared memory (128 + 16: conflict-free)
constexpr int kBcHead = 1024 + 128;            // code2 | codeT
constexpr int kBcSmemMax = 227 * 1024;

// activations may have been produced by the kernel just before this one (PDL): coherent load, not .nc
__device__ __forceinline__ uint4 ld_x_u4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void lds_x4_pred(uint32_t (&b)[4], uint32_t saddr, uint32_t active) {
  asm volatile("{\n .reg .pred P;\n setp.ne.u32 P, %5, 0;\n @P ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n}"
               : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]) : "r"(saddr), "r"(active));
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__global__ void k_epoch_bump(unsigned int *epoch) { *epoch += 1; }
void epoch_bump(unsigned int *epoch) {
  k_epoch_bump<<<1, 1, 0, current_stream()>>>(epoch);
  check_launch("epoch_bump");
}

// Cross-GPU barrier of an N-sharded stack as a link of the programmatic-dependent-launch chain: it lets the NEXT
// kernel's CTAs start (tables, weight prefetch) at once, waits for the previous kernel of the stream -- the GEMV whose
// epilogue stored into the peers -- and then exchanges one sequence number with every peer through symmetric memory.
// The library barrier of torch's symmetric memory is an ordinary launch: it serialises both neighbours (~8 us per
// consumer group on the 7B stack).  seq lives on the device, so a replayed CUDA graph keeps counting.
struct PeerSlots { unsigned int *p[7]; };
__global__ void __launch_bounds__(32) k_peer_barrier(unsigned int *counter, const unsigned int *sig_local, PeerSlots peers, int npeers) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned int seq = 0;
  if (threadIdx.x == 0) { seq = *counter + 1u; *counter = seq; }
  seq = __shfl_sync(0xffffffffu, seq, 0);
  if ((int)threadIdx.x < npeers) {
    __threadfence_system();                           // the previous kernel's peer stores are ordered before the flag
    st_release_sys(peers.p[threadIdx.x], seq);
    const unsigned int *slot = sig_local + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(slot) - seq) < 0) {
      if (clock64() - t0 > 4000000000ll) __trap();    // bounded spin: trap, never hang
    }
  }
}
void peer_barrier(unsigned int *counter, const unsigned int *sig_local, unsigned int *const *sig_peer, int npeers) {
  if (npeers < 0 || npeers > 7) { latch_error(cudaErrorInvalidValue, "peer_barrier: at most 7 peers"); return; }
  PeerSlots ps{};
  for (int i = 0; i < npeers; i++) ps.p[i] = sig_peer[i];
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(1); lc.blockDim = dim3(32); lc.dynamicSmemBytes = 0; lc.stream = current_stream();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  latch_error(cudaLaunchKernelEx(&lc, k_peer_barrier, counter, sig_local, ps, npeers), "peer_barrier launch");
}

// debug probe (flags bit 1): SM cycles and nanoseconds spent by CTA 0 -> effective SM clock under this kernel's load
__device__ unsigned long long g_gemv_probe[12];
// per-CTA trace (flags bit 2, debug): absolute globaltimer ns of {entry, loads issued, previous kernel complete, x ready,
// all warps done, exit} + SM id, for the last two launches (slot = flags bit 3) -- tools/gemv_trace.py reads them
__device__ unsigned long long g_gemv_trace[2][320][8];   // [0] cycles, [1] ns of the probed CTA; [2..6] phase timestamps (ns since entry)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

template <typename T, bool NESTED, int DEPTH = 2, int WARPS = 16, bool MULTI = false>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 8 ? 2 : 1) k_gemv4_bc(const __grid_constant__ GemvArgs a, int x_blocks_padded, int tiles_total) {
  // shared memory: [0, 64 KB) byte LUT (entry stride 256 B, one word per lane) | code2 | x | partial sums.
  // The lookup address is  LUT base (uniform register) + PRMT(byte << 8 | lane * 4): no alignment requirement.
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  float *s_code2 = reinterpret_cast<float *>(smem + 65536);
  unsigned char *s_x = smem + 65536 + kBcHead;
  float *s_part = reinterpret_cast<float *>(s_x + (size_t)x_blocks_padded * kBcXPitch);   // [tile_local][warp][16]
  const uint32_t x_s = smem_base + 65536u + kBcHead;
  // programmatic dependent launch: let the next kernel of the stream start its own prologue (tables, first
  // weight loads) while this one is still running; it waits (griddepcontrol.wait) before touching x / out
  asm volatile("griddepcontrol.launch_dependents;");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // table constants first: ONE round of small global loads (the 16 code values as four 128-bit loads, this thread's
  // low-nibble value, its code2 entries), issued AHEAD of the weight stream so that they are not queued behind it.
  // (A load per LUT entry behind the weight prefetch cost eight serial round trips, ~2 us of every launch on the
  // critical path -- measured with the phase probe.)
  constexpr int CT0 = WARPS * 32;
  float c2v[(256 + CT0 - 1) / CT0];
  float cv[16];
  float clo;
  if (a.tables_in_args == 2) {     // the host verified code == the NF4 table: immediates, no memory access at all
    constexpr float nf4[16] = BNB_NF4_TABLE;
#pragma unroll
    for (int u = 0; u < (256 + CT0 - 1) / CT0; u++) c2v[u] = (NESTED && tid + u * CT0 < 256) ? __ldg(a.code2 + tid + u * CT0) : 0.f;
#pragma unroll
    for (int u = 0; u < 16; u++) cv[u] = nf4[u];
    clo = nf4[0];
#pragma unroll
    for (int u = 1; u < 16; u++) clo = (((tid >> 3) & 15) == u) ? nf4[u] : clo;
  } else {
#pragma unroll
    for (int u = 0; u < (256 + CT0 - 1) / CT0; u++) c2v[u] = (NESTED && tid + u * CT0 < 256) ? __ldg(a.code2 + tid + u * CT0) : 0.f;
    if ((reinterpret_cast<uintptr_t>(a.code) & 15) == 0) {
      const float4 *cg = reinterpret_cast<const float4 *>(a.code);
#pragma unroll
      for (int u = 0; u < 4; u++) { const float4 f = __ldg(cg + u); cv[4 * u] = f.x; cv[4 * u + 1] = f.y; cv[4 * u + 2] = f.z; cv[4 * u + 3] = f.w; }
    } else {
#pragma unroll
      for (int u = 0; u < 16; u++) cv[u] = __ldg(a.code + u);
    }
    clo = __ldg(a.code + ((tid >> 3) & 15));
  }
  unsigned long long probe_c = 0, probe_t = 0;
  const bool probing = (a.flags & 2) && blockIdx.x == gridDim.x / 2 && tid == 0;
  if (probing) { probe_c = clock64(); probe_t = globaltimer_ns(); }
  const bool tracing = (a.flags & 4) && tid == 0 && blockIdx.x < 320;
  unsigned long long *tr = g_gemv_trace[(a.flags >> 3) & 1][blockIdx.x < 320 ? blockIdx.x : 0];
  if (tracing) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    tr[0] = globaltimer_ns(); tr[6] = smid;
  }
  const int kb = a.K >> 6;                 // blocks per row
  const int nch = (a.K + 511) >> 9;        // 512-element chunks per row (the last one may be half)
  const int row_bytes = a.K >> 1;
  // Tile ranges, HEAVY FIRST: CTAs are dispatched in blockIdx order as slots free up, so in a chain of launches the
  // highest block indices enter last (they inherit the slots of the previous kernel's stragglers).  Giving the extra
  // tile to the lowest indices keeps the late entrants light (per-CTA trace: with the interleaved split the last CTA
  // of 11008x4096 entered 3.3 us late AND owned 3 tiles instead of 2 -- 3 us of an 8.5 us period was that tail).
  const int tq = tiles_total / (int)gridDim.x, trem = tiles_total % (int)gridDim.x;
  const int bx = (int)blockIdx.x;
  const int t_begin = bx < trem ? bx * (tq + 1) : trem * (tq + 1) + (bx - trem) * tq;
  const int t_end = t_begin + tq + (bx < trem ? 1 : 0);
  const int ntl = t_end - t_begin;

  uint32_t w[DEPTH][2][2][8];   // [ring slot][block t / t+4][row half][32 bytes = one sector per lane]
  struct Abs { uint32_t q[2]; float am2[2]; float2 am[2]; float off; };
  Abs ab[DEPTH];
  // MULTI: which matrix a tile belongs to, and that matrix's operands (select chains: kernel parameters cannot be
  // indexed dynamically without a copy to local memory)
  auto mat_of = [&](int tile) { return MULTI ? (int)(tile >= a.mt[1]) + (int)(tile >= a.mt[2]) + (int)(tile >= a.mt[3]) : 0; };
#define BNB_MSEL(arr, m) ((m) == 0 ? a.arr[0] : (m) == 1 ? a.arr[1] : (m) == 2 ? a.arr[2] : a.arr[3])

  // one 256-bit load per (row, block): a lane owns a whole 32-byte sector, 4 lanes one 128-byte line
  auto load_w = [&](uint32_t (&dst)[2][8], int j, int tile, int c) {
    const int m = mat_of(tile);
    const int lt = MULTI ? tile - BNB_MSEL(mt, m) : tile;
    const int Nm = MULTI ? BNB_MSEL(mN, m) : a.N;
    const unsigned char *Bm = MULTI ? BNB_MSEL(mB, m) : a.B;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int row = min(lt * 16 + g + 8 * h, Nm - 1);
      const unsigned char *p = Bm + (size_t)row * row_bytes + c * 256 + (t + 4 * j) * 32;
      if (c * 8 + t + 4 * j < kb) ld_stream_u8(dst[h], p);
      else {
#pragma unroll
        for (int i = 0; i < 8; i++) dst[h][i] = 0;
      }
    }
  };
  auto load_abs = [&](Abs &d, int tile, int c) {
    const int m = mat_of(tile);
    const int lt = MULTI ? tile - BNB_MSEL(mt, m) : tile;
    const int Nm = MULTI ? BNB_MSEL(mN, m) : a.N;
    const unsigned char *qm = MULTI ? BNB_MSEL(mq, m) : a.qabsmax;
    const float *am2m = MULTI ? BNB_MSEL(mam2, m) : a.absmax2;
    if (MULTI) d.off = BNB_MSEL(moff, m);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int row = min(lt * 16 + g + 8 * h, Nm - 1);
      const size_t idx = (size_t)row * kb + min(c * 8 + 2 * t, kb - 2);
      if (NESTED) {
        d.q[h] = __ldg(reinterpret_cast<const unsigned short *>(qm + idx));
        d.am2[h] = __ldg(am2m + (idx >> a.bs2_shift));
      } else {
        d.am[h] = __ldg(reinterpret_cast<const float2 *>(a.absmax + idx));
      }
    }
  };
  auto advance = [&](int &tl_, int &c_) {
    c_ += WARPS;
    while (c_ >= nch) { c_ -= nch; tl_++; }
  };

  // compute cursor (tl, c) and load cursor (ltl, lc): the register ring keeps DEPTH items per warp in flight
  int tl = 0, c = warp;
  while (c >= nch) { c -= nch; tl++; }
  int ltl = tl, lc = c;
#pragma unroll
  for (int s = 0; s < DEPTH; s++) {
    if (ltl < ntl) {
      load_w(w[s][0], 0, t_begin + ltl, lc);
      load_w(w[s][1], 1, t_begin + ltl, lc);
      load_abs(ab[s], t_begin + ltl, lc);
    }
    advance(ltl, lc);
  }
  // ---- prologue (overlaps the first weight loads): tables and partial-sum slots; everything here reads only
  // constants (code, code2), so it may run before the previous kernel of the stream has finished
  {
    constexpr int CT = WARPS * 32;
    if (probing) g_gemv_probe[7] = globaltimer_ns() - probe_t;      // first weight loads issued
    if (tracing) { tr[1] = globaltimer_ns(); tr[7] = (unsigned long long)ntl; }
    // byte e -> {T(code[e >> 4]), T(code[e & 15])}, replicated for the 32 lanes (bank == lane): 8 threads write
    // one 128-byte entry with conflict-free 128-bit stores.  Thread (tid >> 3) + it * CT/8 handles entries whose low
    // nibble is fixed ((tid >> 3) & 15) and whose high nibble is a compile-time function of `it` plus bits of tid.
    const int j = tid & 7;
    const uint32_t lo16 = MmaT<T>::pack(clo, 0.0f) << 16;
    constexpr int EPI = CT / 8;                      // entries per iteration (32 for 8 warps, 64 for 16)
    const int hsel = (tid >> 3) >> 4;                // 0 .. EPI/16 - 1
#pragma unroll
    for (int it = 0; it < (256 + EPI - 1) / EPI; it++) {
      const int e = (tid >> 3) + it * EPI;
      constexpr int HB = EPI / 16;
      float chi = cv[it * HB < 15 ? it * HB : 15];
#pragma unroll
      for (int h = 1; h < HB; h++) chi = (hsel == h) ? cv[it * HB + h < 15 ? it * HB + h : 15] : chi;
      const uint32_t v = (MmaT<T>::pack(chi, 0.0f) & 0xFFFFu) | lo16;
      if (e < 256) *reinterpret_cast<uint4 *>(smem + e * 256 + j * 16) = make_uint4(v, v, v, v);
    }
    if (probing) g_gemv_probe[9] = globaltimer_ns() - probe_t;      // LUT stored
    for (int i = tid; i < ntl * WARPS * 16; i += CT) s_part[i] = 0.f;
    if (probing) g_gemv_probe[2] = globaltimer_ns() - probe_t;      // prologue (tables) done, about to wait
    asm volatile("griddepcontrol.wait;" ::: "memory");     // x (and out) belong to the previous kernel until here
    if (probing) g_gemv_probe[3] = globaltimer_ns() - probe_t;      // previous kernel complete
    if (tracing) tr[2] = globaltimer_ns();
    if (a.sig_local != nullptr && a.do_wait) {
      // first kernel of a consumer group on an N-sharded stack: the gathered vectors of the previous group must
      // be complete on this GPU, i.e. every peer has published a sequence number >= ours (bounded spin: trap, never hang)
      if (tid < a.npeers) {
        const unsigned int target = ld_acquire_sys(a.epoch) * (unsigned int)a.ngroups + (unsigned int)a.gidx;
        const unsigned int *slot = a.sig_local + tid;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(slot) - target) < 0) {
          if (clock64() - t0 > 4000000000ll) __trap();
        }
      }
      __syncthreads();
    }
    const uint4 *xg = reinterpret_cast<const uint4 *>(a.x);
    const int pieces = x_blocks_padded * 8, valid = a.K >> 3;
    for (int p0 = tid; p0 < pieces; p0 += 4 * CT) {   // four independent loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int p = p0 + u * CT;
        v[u] = p < valid ? ld_x_u4(xg + p) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int p = p0 + u * CT;
        if (p < pieces) *reinterpret_cast<uint4 *>(s_x + (p >> 3) * kBcXPitch + (p & 7) * 16) = v[u];
      }
    }
  }
  // code2 is first needed by the de-nest at the end of the first item: stored last, its load had the whole prologue
#pragma unroll
  for (int u = 0; u < (256 + CT0 - 1) / CT0; u++) if (NESTED && tid + u * CT0 < 256) s_code2[tid + u * CT0] = c2v[u];
  if (probing) g_gemv_probe[8] = globaltimer_ns() - probe_t;        // code2 arrived and stored
  __syncthreads();
  if (probing) g_gemv_probe[4] = globaltimer_ns() - probe_t;        // x in shared memory
  if (tracing) tr[3] = globaltimer_ns();

  const uint32_t lane4 = (uint32_t)(lane * 4);
  const uint32_t act0 = (g == t), act1 = (g == t + 4);
  const uint32_t xlane = x_s + g * kBcXPitch;
  uint32_t b0[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0};
  float acc0 = 0.f, acc1 = 0.f;
  const float offset = a.offset;

  while (tl < ntl) {
#pragma unroll
    for (int s = 0; s < DEPTH; s++) {
      if (tl >= ntl) break;
      const bool lhave = ltl < ntl;
      const uint32_t xc = xlane + c * (8 * kBcXPitch);
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 2; j++) {          // j = 0: blocks 0-3 (columns 0-3), j = 1: blocks 4-7 (columns 4-7)
#pragma unroll
        for (int mg = 0; mg < 8; mg++) {     // one 32-bit word of each row = 8 elements = 2 MMAs
          if (j == 0) lds_x4_pred(b0, xc + mg * 16, act0);
          else lds_x4_pred(b1, xc + mg * 16, act1);
          const uint32_t s0 = w[s][j][0][mg], s1 = w[s][j][1][mg];
#pragma unroll
          for (int mm = 0; mm < 2; mm++) {
            const uint32_t selA = 0x7604u | ((2 * mm) << 4), selB = 0x7604u | ((2 * mm + 1) << 4);
            uint32_t af[4];
            af[0] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s0, lane4, selA));
            af[1] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s1, lane4, selA));
            af[2] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s0, lane4, selB));
            af[3] = *reinterpret_cast<const uint32_t *>(smem + __byte_perm(s1, lane4, selB));
            if (j == 0) MmaT<T>::mma(d, af, b0[2 * mm], b0[2 * mm + 1]);
            else MmaT<T>::mma(d, af, b1[2 * mm], b1[2 * mm + 1]);
          }
        }
        if (lhave) load_w(w[s][j], j, t_begin + ltl, lc);   // refill this half of the slot: item DEPTH ahead
      }
      float am00, am01, am10, am11;
      if (NESTED) {
        const float off = MULTI ? ab[s].off : offset;
        am00 = __fadd_rn(__fmul_rn(s_code2[ab[s].q[0] & 0xFFu], ab[s].am2[0]), off);
        am01 = __fadd_rn(__fmul_rn(s_code2[ab[s].q[0] >> 8], ab[s].am2[0]), off);
        am10 = __fadd_rn(__fmul_rn(s_code2[ab[s].q[1] & 0xFFu], ab[s].am2[1]), off);
        am11 = __fadd_rn(__fmul_rn(s_code2[ab[s].q[1] >> 8], ab[s].am2[1]), off);
      } else {
        am00 = ab[s].am[0].x; am01 = ab[s].am[0].y; am10 = ab[s].am[1].x; am11 = ab[s].am[1].y;
      }
      if (lhave) load_abs(ab[s], t_begin + ltl, lc);
      advance(ltl, lc);
      acc0 = __fmaf_rn(d[0], am00, acc0);
      acc0 = __fmaf_rn(d[1], am01, acc0);
      acc1 = __fmaf_rn(d[2], am10, acc1);
      acc1 = __fmaf_rn(d[3], am11, acc1);
      int ntl_ = tl, nc = c;
      advance(ntl_, nc);
      if (ntl_ != tl) {   // this warp is done with the tile: park its partial sums
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
        if (t == 0) {
          float *slot = s_part + (tl * WARPS + warp) * 16;
          slot[g] = acc0;
          slot[g + 8] = acc1;
        }
        acc0 = acc1 = 0.f;
      }
      tl = ntl_; c = nc;
    }
  }
  if (probing) g_gemv_probe[5] = globaltimer_ns() - probe_t;        // warp 0 finished its items
  __syncthreads();
  if (probing) g_gemv_probe[6] = globaltimer_ns() - probe_t;        // every warp finished
  if (tracing) tr[4] = globaltimer_ns();
  for (int i = tid; i < ntl * 16; i += (WARPS * 32)) {
    const int tile_l = i >> 4, row = i & 15;
    const float *p = s_part + tile_l * WARPS * 16 + row;
    float sum = 0.f;
#pragma unroll
    for (int wq = 0; wq < WARPS; wq++) sum += p[wq * 16];
    if (MULTI) {
      const int m = mat_of(t_begin + tile_l);
      const int rr = (t_begin + tile_l - BNB_MSEL(mt, m)) * 16 + row;
      if (rr < BNB_MSEL(mN, m)) {
        const T v = from_float<T>(sum);
        reinterpret_cast<T *>(BNB_MSEL(mout, m))[rr] = v;
#pragma unroll
        for (int pr = 0; pr < 7; pr++)                                       // NVLink P2P stores
          if (pr < a.npeers)
            reinterpret_cast<T *>(m == 0 ? a.mpeer[0][pr] : m == 1 ? a.mpeer[1][pr] : m == 2 ? a.mpeer[2][pr] : a.mpeer[3][pr])[rr] = v;
      }
      continue;
    }
    const int r = (t_begin + tile_l) * 16 + row;
    if (r < a.N) {
      const T v = from_float<T>(sum);
      reinterpret_cast<T *>(a.out)[r] = v;
      for (int pr = 0; pr < a.npeers; pr++) reinterpret_cast<T *>(a.peer_out[pr])[r] = v;   // NVLink P2P stores
    }
  }
  if (a.sig_local != nullptr && a.do_signal) {
    __threadfence_system();                       // this thread's peer stores are visible system-wide
    __syncthreads();
    if (tid == 0) {
      const unsigned int done = atomicAdd(a.cta_counter, 1u);
      if (done == gridDim.x - 1) {                // last CTA of the grid: the whole slice is out
        *a.cta_counter = 0u;
        __threadfence_system();
        const unsigned int seq = ld_acquire_sys(a.epoch) * (unsigned int)a.ngroups + (unsigned int)a.gidx + 1u;
        for (int pr = 0; pr < a.npeers; pr++) st_release_sys(a.sig_peer[pr], seq);
      }
    }
  }
  if (tracing) tr[5] = globaltimer_ns();
  if (probing) {
    g_gemv_probe[0] = clock64() - probe_c;
    g_gemv_probe[1] = globaltimer_ns() - probe_t;
  }
}

#undef BNB_MSEL
void gemv_trace(unsigned long long *out) { cudaMemcpyFromSymbol(out, g_gemv_trace, sizeof(unsigned long long) * 2 * 320 * 8); }
// host: last probe of the block-column kernel -> {cycles, ns}
void gemv_probe(unsigned long long *out2) {
  cudaMemcpyFromSymbol(out2, g_gemv_probe, sizeof(unsigned long long) * 12);
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

// smem of the block-column kernel for `warps` warps per CTA on `grid` CTAs
static size_t bc_smem_need(int K, int tiles, int warps, int grid) {
  const int xblocks = ceil_div(K, 512) * 8;
  return (size_t)65536 + kBcHead + (size_t)xblocks * kBcXPitch + (size_t)ceil_div(tiles, grid) * warps * 16 * sizeof(float);
}
static cudaLaunchConfig_t pdl_config(int grid, int threads, size_t smem, cudaLaunchAttribute *attr) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem; lc.stream = current_stream();
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  return lc;
}
// BNB_B200_GEMV_PROBE=1: the middle CTA records its phase timestamps (tools/kbench.py reads them); =2: every CTA of the last
// two launches records absolute timestamps (tools/gemv_trace.py)
static int probe_flag() {
  static int v = -1;
  static unsigned int seq = 0;
  if (v < 0) { const char *pr = getenv("BNB_B200_GEMV_PROBE"); v = pr ? (pr[0] == '1' ? 2 : pr[0] == '2' ? 4 : 0) : 0; }
  if (v == 4) return 4 | (int)((seq++ & 1u) << 3);
  return v;
}

template <typename T, bool NESTED, bool VEC4>
static void launch_mma_inst(const GemvArgs &a) {
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  GemvArgs a2 = a;
  a2.flags = probe_flag();
  if (VEC4 && a.batch == 1) {
    // block-column kernel, register ring.  8-warp CTAs, two per SM (<= 113 KB of shared memory and <= 128 registers
    // each), so consecutive GEMVs of a stream overlap through programmatic dependent launch; 16-warp CTAs, one per
    // SM, when x + partial sums do not fit twice (K > ~16K).
    const int tiles = ceil_div(a.N, 16);
    int warps = 8;
    int grid = tiles < sms * 2 ? tiles : sms * 2;
    if (bc_smem_need(a.K, tiles, warps, grid) > (size_t)(113 * 1024)) {
      warps = 16;
      grid = tiles < sms ? tiles : sms;
    }
    const size_t need = bc_smem_need(a.K, tiles, warps, grid);
    if (need <= (size_t)kBcSmemMax) {
      cudaLaunchAttribute attr[1];
      cudaLaunchConfig_t lc = pdl_config(grid, warps * 32, need, attr);
      const int xblocks = ceil_div(a.K, 512) * 8;
      if (warps == 8) {
        auto kfn = k_gemv4_bc<T, NESTED, 2, 8>;
        ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), kBcSmemMax, "gemv bc smem attr");
        latch_error(cudaLaunchKernelEx(&lc, kfn, a2, xblocks, tiles), "gemv_4bit (block-column) launch");
      } else {
        auto kfn = k_gemv4_bc<T, NESTED, 2, 16>;
        ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), kBcSmemMax, "gemv bc smem attr");
        latch_error(cudaLaunchKernelEx(&lc, kfn, a2, xblocks, tiles), "gemv_4bit (block-column) launch");
      }
      check_launch("gemv_4bit (block-column)");
      return;
    }
  }
  if (a.npeers > 0) { latch_error(cudaErrorInvalidValue, "gemv_4bit: peer outputs are only available in the block-column kernel"); return; }
  // batch 2..8 (the batch rides in the MMA n dimension), blocksize != 64, K % 256 != 0: the per-block-accumulator kernels
  if (VEC4) {
    auto kfn = k_gemv4_fast<T, NESTED>;
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), (int)kFastSmem, "gemv smem attr");
    const int ctas = ceil_div(ceil_div(a.N, 16), kFastGroups);
    kfn<<<ctas < sms ? ctas : sms, kFastThreads, kFastSmem, current_stream()>>>(a2);
  } else {
    auto kfn = k_gemv4_mma<T, NESTED, false>;
    const size_t smem = kGemvLutBytes + kGemvWarps * 128 * sizeof(float) + 256 * sizeof(float) + 16 * sizeof(uint32_t);
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), (int)smem, "gemv smem attr");
    kfn<<<ceil_div(a.N, 16), kGemvThreads, smem, current_stream()>>>(a);
  }
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

// host copies of code[16] / code2[256] for the NEXT nested GEMV of this thread (consumed by that call)
static thread_local const float *tl_code_host = nullptr;
void set_gemv_host_tables(const float *code16, const float *code2_256) { (void)code2_256; tl_code_host = code16; }
// consumed (and cleared) at the very top of the next nested GEMV of this thread, whatever path that call then takes:
// the pointer belongs to the caller and must not outlive the call
static bool take_host_code_is_nf4() {
  const float *c = tl_code_host;
  tl_code_host = nullptr;
  if (c == nullptr) return false;
  static const float nf4[16] = BNB_NF4_TABLE;
  for (int i = 0; i < 16; i++) if (c[i] != nf4[i]) return false;
  return true;
}

template <typename T>
void gemv_4bit_nested(int m, int n, int k, const T *A, const unsigned char *B, const unsigned char *qabsmax,
                      const float *absmax2, const float *code2, float offset, const float *datatype, T *out,
                      int lda, int ldb, int ldc, int blocksize, int blocksize2, void *const *peer_outs, int npeers,
                      const GemvSync *sync) {
  (void)lda; (void)ldc;
  const bool code_is_nf4 = take_host_code_is_nf4();
  if (m <= 0 || k <= 0) return;
  if (n < 1 || n > 8 || blocksize2 <= 0 || (blocksize2 & (blocksize2 - 1)) != 0 || !fast_path_ok(k, ldb, blocksize, A, B)) {
    latch_error(cudaErrorInvalidValue, "gemv_4bit_nested: needs 1<=n<=8, K%64==0, ldb==K/2, blocksize%64==0, aligned A/B");
    return;
  }
  GemvArgs a{};
  a.N = m; a.K = k; a.batch = n; a.blocksize = blocksize;
  a.x = A; a.B = B; a.qabsmax = qabsmax; a.absmax2 = absmax2; a.code2 = code2; a.offset = offset;
  a.code = datatype; a.out = out;
  if (code_is_nf4) a.tables_in_args = 2;
  if (npeers > 0) {
    // peer stores exist only in the block-column kernel (batch 1, blocksize 64, K % 256 == 0, K small enough for shared memory)
    if (npeers > 7 || n != 1 || blocksize != 64 || (k % 256) != 0 || k > 28672 || !peer_outs) {
      latch_error(cudaErrorInvalidValue, "gemv_4bit_nested: peer outputs need batch 1, blocksize 64, K % 256 == 0, <= 7 peers");
      return;
    }
    a.npeers = npeers;
    for (int i = 0; i < npeers; i++) a.peer_out[i] = peer_outs[i];
    if (sync && sync->sig_local) {
      a.sig_local = sync->sig_local; a.epoch = sync->epoch; a.cta_counter = sync->cta_counter;
      a.gidx = sync->gidx; a.ngroups = sync->ngroups; a.do_signal = sync->do_signal; a.do_wait = sync->do_wait;
      for (int i = 0; i < npeers; i++) a.sig_peer[i] = sync->sig_peer[i];
    }
  }
  launch_mma<T, true>(a, blocksize2);
}

// Several nested-absmax NF4/FP4 weight matrices that share x (q/k/v, gate/up of a decoder layer) in ONE launch of the
// block-column kernel: the per-launch constant (table build, dependency wait, x fetch, drain: ~2.7 us) is
// paid once.  Same arithmetic per output element as the single-matrix call (bit-identical results).
// returns 0 ok, 1 shape not taken (caller issues the single-matrix calls)
template <typename T>
int gemv_4bit_nested_multi(int count, const int *ms, int k, const T *A, const unsigned char *const *Bs,
                           const unsigned char *const *qabs, const float *const *am2s, const float *code2,
                           const float *offsets, const float *datatype, T *const *outs, int blocksize, int blocksize2,
                           void *const *peer_outs, int npeers) {
  const bool code_is_nf4 = take_host_code_is_nf4();
  if (npeers < 0 || npeers > 7 || (npeers > 0 && peer_outs == nullptr)) return 1;
  if (count < 1 || count > 4 || k <= 0 || blocksize != 64 || (k % 256) != 0 || blocksize2 <= 0 || (blocksize2 & (blocksize2 - 1)) != 0 ||
      (reinterpret_cast<uintptr_t>(A) % 16) != 0)
    return 1;
  GemvArgs a{};
  a.K = k; a.batch = 1; a.blocksize = blocksize; a.bs_shift = 6; a.bs2_shift = ilog2(blocksize2);
  a.x = A; a.code2 = code2; a.code = datatype; a.nmat = count;
  int tiles = 0, nsum = 0;
  for (int i = 0; i < 5; i++) a.mt[i] = 0x7fffffff;
  for (int i = 0; i < count; i++) {
    if (ms[i] <= 0 || (reinterpret_cast<uintptr_t>(Bs[i]) % 8) != 0 || (reinterpret_cast<uintptr_t>(qabs[i]) % 4) != 0) return 1;
    a.mt[i] = tiles; a.mN[i] = ms[i]; a.moff[i] = offsets[i];
    a.mB[i] = Bs[i]; a.mq[i] = qabs[i]; a.mam2[i] = am2s[i]; a.mout[i] = outs[i];
    tiles += ceil_div(ms[i], 16); nsum += ms[i];
  }
  for (int i = count; i < 4; i++) { a.mN[i] = 1; a.mB[i] = Bs[0]; a.mq[i] = qabs[0]; a.mam2[i] = am2s[0]; a.mout[i] = outs[0]; }
  a.npeers = npeers;
  for (int i = 0; i < count; i++)
    for (int p = 0; p < npeers; p++) a.mpeer[i][p] = peer_outs[i * npeers + p];     // matrix-major
  a.N = nsum; a.B = Bs[0]; a.qabsmax = qabs[0]; a.absmax2 = am2s[0]; a.offset = offsets[0]; a.out = outs[0];
  if (code_is_nf4) a.tables_in_args = 2;
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int xblocks = ceil_div(k, 512) * 8;
  int warps = 8, grid = tiles < sms * 2 ? tiles : sms * 2;
  if (bc_smem_need(k, tiles, warps, grid) > (size_t)(113 * 1024)) { warps = 16; grid = tiles < sms ? tiles : sms; }
  const size_t need = bc_smem_need(k, tiles, warps, grid);
  if (need > (size_t)kBcSmemMax) return 1;
  a.flags = probe_flag();
  cudaLaunchAttribute attr[1];
  cudaLaunchConfig_t lc = pdl_config(grid, warps * 32, need, attr);
  if (warps == 8) {
    auto kfn = k_gemv4_bc<T, true, 2, 8, true>;
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), kBcSmemMax, "gemv multi smem attr");
    latch_error(cudaLaunchKernelEx(&lc, kfn, a, xblocks, tiles), "gemv_4bit (multi) launch");
  } else {
    auto kfn = k_gemv4_bc<T, true, 2, 16, true>;
    ensure_max_dynamic_smem(reinterpret_cast<const void *>(kfn), kBcSmemMax, "gemv multi smem attr");
    latch_error(cudaLaunchKernelEx(&lc, kfn, a, xblocks, tiles), "gemv_4bit (multi) launch");
  }
  check_launch("gemv_4bit (multi)");
  return 0;
}
template int gemv_4bit_nested_multi<__half>(int, const int *, int, const __half *, const unsigned char *const *, const unsigned char *const *, const float *const *, const float *, const float *, const float *, __half *const *, int, int, void *const *, int);
template int gemv_4bit_nested_multi<__nv_bfloat16>(int, const int *, int, const __nv_bfloat16 *, const unsigned char *const *, const unsigned char *const *, const float *const *, const float *, const float *, const float *, __nv_bfloat16 *const *, int, int, void *const *, int);

template void gemv_4bit<float>(int, int, int, const float *, const unsigned char *, const float *, const float *, float *, int, int, int, int);
template void gemv_4bit<__half>(int, int, int, const __half *, const unsigned char *, const float *, const float *, __half *, int, int, int, int);
template void gemv_4bit<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const float *, const float *, __nv_bfloat16 *, int, int, int, int);
template void gemv_4bit_nested<__half>(int, int, int, const __half *, const unsigned char *, const unsigned char *, const float *, const float *, float, const float *, __half *, int, int, int, int, int, void *const *, int, const GemvSync *);
template void gemv_4bit_nested<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const unsigned char *, const float *, const float *, float, const float *, __nv_bfloat16 *, int, int, int, int, int, void *const *, int, const GemvSync *);

}  // namespace bnb
