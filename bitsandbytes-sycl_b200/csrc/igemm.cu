// igemm.cu -- K6: int8 GEMM C[i,j] = sum_k A[i,k] * B[j,k] with exact int32 accumulators on the
// sm_100a tensor pipe (tcgen05.mma kind::i8, accumulators in TMEM), optional fused mm_dequant epilogue.
//
// Replaces igemmlt<FORMATB,32,0> (reference op_gemm.cpp:541-603 -> vendored blas_utils.h:459-724 ->
// oneDNN s8*s8->s32; originally cublasLtMatmul on Turing/Ampere IMMA layouts) and, when the epilogue
// is fused, kdequant_mm_int32_fp16 (kernel_quant.cpp:3848-3987).
//
// Design (DESIGN.md "K6"): persistent warp-specialised CTA per SM --
//   warp 0   : TMA producer, 128x128 B (A) + 256x128 B (B) K-major tiles, SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1   : TMEM owner + single-thread tcgen05.mma issuer (M=128, N=256, K=32 per instruction),
//              two 256-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-5: epilogue, tcgen05.ld 32x32b -> registers -> (int32 | dequant -> fp16) -> global
// Operands are ROW-MAJOR int8 (K contiguous) -- the layout TMA + UMMA consume directly.  The
// reference's col32 / col_turing / col_ampere operands (Turing/Ampere IMMA layouts) are accepted by the
// cigemmlt_* ABI wrappers, which re-layout through scratch (correct, slower; the B200-native entry
// points are cigemm_rowmajor_*).
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bnb {

// implemented in int8_quant.cu
void untransform_s8(int fmt, const signed char *A, signed char *out, int rows, int cols);
template <typename E> void to_col32(const E *A, E *out, int rows, int cols);

// ------------------------------------------------------------------------------------------------
// tensor map helper
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

bool make_tmap_2d(CUtensorMap *map, const void *base, int elem_bytes, uint64_t rows, uint64_t cols, uint32_t box_rows,
                  uint32_t box_cols, bool is_16bit_float, bool is_bf16, bool swizzle128) {
  auto fn = get_encode_fn();
  if (!fn) { latch_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled unavailable"); return false; }
  CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  if (is_16bit_float) dt = is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  else if (elem_bytes == 4) dt = CU_TENSOR_MAP_DATA_TYPE_INT32;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { latch_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed"); return false; }
  return true;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 256, BK = 128;  // BK in bytes == int8 elements
// CG = CTAs per MMA (tcgen05 cta_group).  CG = 2: a CTA pair (cluster of two, one TPC) computes a 256 x 256 tile; each CTA
// loads ITS 128 rows of A and ITS 128 of the 256 rows of B (32 KB per stage instead of 48), the leader issues
// tcgen05.mma.cta_group::2 with M = 256 and each CTA's TMEM receives its own 128 x 256 half of the result.  Per stage and
// SM the operand traffic through shared memory drops from 96 KB (768 cycles at 128 B/cycle, more than the 528 cycles the
// four MMAs take) to 64 KB.
constexpr int kStageA = BM * BK;   // 16 KB
__host__ __device__ constexpr int stage_b_bytes(int CG) { return (BN / CG) * BK; }          // 32 KB / 16 KB
__host__ __device__ constexpr int stage_bytes(int CG) { return kStageA + stage_b_bytes(CG); }
__host__ __device__ constexpr int stages_for(int CG) { return CG == 2 ? 6 : 4; }           // 192 KB of operand ring either way
constexpr int kEpiThreads = 256;                 // 8 epilogue warps: two per TMEM lane quarter, each takes half the columns
constexpr int kIgemmThreads = 64 + kEpiThreads;
constexpr int kTmemCols = 512;
constexpr int kIgemmSmem = 4 * (kStageA + BN * BK) + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BN * (4 + 4) /*col stats + bias (fp32)*/ +
                           2 * BN * 8 * 4 /*outlier columns of the weight as fp32, per accumulator buffer*/;

enum { EPI_INT32 = 0, EPI_DEQUANT_FP16 = 1, EPI_INT32_COL32 = 2, EPI_S8_COL32 = 3 };   // the last two: the reference ABI's C layout, written by the epilogue

struct IgemmArgs {
  int M, N, K;
  int *C;                  // EPI_INT32: row-major [M,N]; EPI_INT32_COL32: col32 [ceil(N/32)][M][32]
  signed char *C8;         // EPI_S8_COL32: col32 int8, saturated rint(acc * alpha)
  const float *row_scale;  // EPI_S8_COL32: per-row alpha (null: 1.0f)
  const float *rowStats;   // EPI_DEQUANT_FP16
  const float *colStats;
  const __half *bias;      // may be null
  __half *out;             // row-major [M,N]
  // 16-bit outlier product folded into the dequant epilogue (int8_fused.cu): out = half(half(dequant) + half(sum_o subA*subB))
  const __half *subA;      // [M,16] (null: none)
  const __half *subB;      // [N,16]
  const int *nout;         // device-side number of outlier columns
};

template <int EPI, int CG>
__global__ void __launch_bounds__(kIgemmThreads, 1)
k_igemm_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ IgemmArgs a) {
  constexpr int kStages = stages_for(CG), kStageBytes = stage_bytes(CG);
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
  uint64_t *full = bars, *empty = bars + kStages, *tfull = bars + 2 * kStages, *tempty = bars + 2 * kStages + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);
  float *s_cs = reinterpret_cast<float *>(smem + kStages * kStageBytes + 256);  // [2][BN]
  float *s_bias = s_cs + 2 * BN;                                                // [2][BN]
  float4 *s_sb = reinterpret_cast<float4 *>(s_bias + 2 * BN);                   // [2][BN][8 floats] = 2 float4 per column

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CG == 2 ? (int)cg2::cta_rank() : 0;          // 0 = leader of the pair (issues the MMAs)
  const int unit = (int)blockIdx.x / CG, units = (int)gridDim.x / CG;   // a unit = a CTA or a CTA pair
  const int num_m = (a.M + BM * CG - 1) / (BM * CG), num_n = (a.N + BN - 1) / BN;
  const int tiles = num_m * num_n, kblocks = (a.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&tmA); tc::prefetch_tmap(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; s++) { tc::mbar_init(tc::smem_u32(full + s), 1); tc::mbar_init(tc::smem_u32(empty + s), 1); }
      // the accumulator is free again when the epilogue warps of BOTH CTAs have read their half (they arrive on the leader's barrier)
      for (int i = 0; i < 2; i++) { tc::mbar_init(tc::smem_u32(tfull + i), 1); tc::mbar_init(tc::smem_u32(tempty + i), CG * kEpiThreads / 32); }
      tc::fence_barrier_init();
    }
    __syncwarp();
    if (CG == 2) cg2::tmem_alloc(tc::smem_u32(tmem_slot), kTmemCols); else tc::tmem_alloc(tc::smem_u32(tmem_slot), kTmemCols);
  }
  tc::fence_before_sync();
  if (CG == 2) cg2::cluster_sync(); else __syncthreads();   // the peer's mbarriers must exist before anything signals them
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (whole warp runs the loop, one elected lane issues) =================
    int stage = 0; uint32_t phase = 0;
    for (int tile = unit; tile < tiles; tile += units) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      for (int kb = 0; kb < kblocks; kb++) {
        tc::mbar_wait(tc::smem_u32(empty + stage), phase ^ 1);
        if (tc::elect_one()) {
          const uint32_t fb = tc::smem_u32(full + stage);
          const uint32_t sa = tc::smem_u32(smem + stage * kStageBytes);
          if (CG == 2) {
            if (rank == 0) tc::mbar_arrive_expect_tx(fb, 2 * kStageBytes);   // the bytes of both CTAs land on the leader's barrier
            cg2::tma_load_2d(sa, &tmA, fb, kb * BK, (m_blk * 2 + rank) * BM);
            cg2::tma_load_2d(sa + kStageA, &tmB, fb, kb * BK, n_blk * BN + rank * (BN / 2));
          } else {
            tc::mbar_arrive_expect_tx(fb, kStageBytes);
            tc::tma_load_2d(sa, &tmA, fb, kb * BK, m_blk * BM);
            tc::tma_load_2d(sa + kStageA, &tmB, fb, kb * BK, n_blk * BN);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (the leader CTA of a pair only) =================
    if (rank == 0) {
      constexpr uint32_t idesc = tc::umma_idesc(tc::kCFormatS32, 1u, BM * CG, BN);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = unit; tile < tiles; tile += units, it++) {
        const int acc = it & 1; const uint32_t use = (uint32_t)(it >> 1);
        tc::mbar_wait(tc::smem_u32(tempty + acc), (use & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; kb++) {
          tc::mbar_wait(tc::smem_u32(full + stage), phase);
          tc::fence_after_sync();
          const uint32_t sa = tc::smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = tc::umma_desc_sw128_kmajor(sa);
          const uint64_t bdesc = tc::umma_desc_sw128_kmajor(sa + kStageA);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 32; k++) {
              if (CG == 2) cg2::umma_i8(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              else tc::umma_i8(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
            if (CG == 2) cg2::umma_commit_both(tc::smem_u32(empty + stage)); else tc::umma_commit(tc::smem_u32(empty + stage));
            if (kb == kblocks - 1) {
              if (CG == 2) cg2::umma_commit_both(tc::smem_u32(tfull + acc)); else tc::umma_commit(tc::smem_u32(tfull + acc));
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (warps 2..9: TMEM lane quarter = warp & 3, column half = (warp - 2) >> 2) =================
    const int q = warp & 3;
    const int half_n = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;  // 0..255
    int it = 0;
    for (int tile = unit; tile < tiles; tile += units, it++) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      const int acc = it & 1; const uint32_t use = (uint32_t)(it >> 1);
      const int row = (m_blk * CG + rank) * BM + q * 32 + lane;
      float rs = 0.f;
      float arow[8];
      int nout = 0;
#pragma unroll
      for (int o = 0; o < 8; o++) arow[o] = 0.f;
      if (EPI == EPI_DEQUANT_FP16) {
        // stage this tile's column stats / bias; the buffer `acc` was last read two tiles ago
        for (int j = et; j < BN; j += kEpiThreads) {
          const int col = n_blk * BN + j;
          s_cs[acc * BN + j] = col < a.N ? a.colStats[col] : 0.f;
          s_bias[acc * BN + j] = (a.bias != nullptr && col < a.N) ? __half2float(a.bias[col]) : 0.f;
        }
        rs = row < a.M ? a.rowStats[row] : 0.f;
        if (a.subA != nullptr) {
          nout = *a.nout;
          if (nout > 8) nout = 0;   // more than 8 outlier columns: k_i8_outlier_tail adds the whole product instead
          if (nout > 0) {
            for (int j = et; j < BN; j += kEpiThreads) {
              const int col = n_blk * BN + j;
              float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              if (col < a.N) {
                const uint4 b = __ldg(reinterpret_cast<const uint4 *>(a.subB + (size_t)col * 16));
                const __half *hb = reinterpret_cast<const __half *>(&b);
#pragma unroll
                for (int o = 0; o < 8; o++) f[o] = __half2float(hb[o]);
              }
              s_sb[(acc * BN + j) * 2] = make_float4(f[0], f[1], f[2], f[3]);
              s_sb[(acc * BN + j) * 2 + 1] = make_float4(f[4], f[5], f[6], f[7]);
            }
            if (row < a.M) {
              const uint4 a0 = __ldg(reinterpret_cast<const uint4 *>(a.subA + (size_t)row * 16));
              const __half *h0 = reinterpret_cast<const __half *>(&a0);
#pragma unroll
              for (int o = 0; o < 8; o++) arow[o] = __half2float(h0[o]);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      }
      tc::mbar_wait(tc::smem_u32(tfull + acc), use & 1);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half_n * (BN / 2);
#pragma unroll 1
      for (int c = 0; c < BN / 64; c++) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
        const int cl = half_n * (BN / 2) + c * 32;   // column inside the tile
        const int col0 = n_blk * BN + cl;
        if (row < a.M && col0 < a.N) {
          if (EPI == EPI_INT32_COL32) {
            // col32 (blas_utils.h:263-266): element (r, c) at (c/32)*32*M + r*32 + c%32 -- the 32 columns this thread holds
            // are one contiguous 128-byte run; the padding columns of the last panel are left as the caller zeroed them
            int *dst = a.C + ((size_t)(col0 >> 5) * a.M + row) * 32;
            if (col0 + 32 <= a.N) {
#pragma unroll
              for (int j = 0; j < 8; j++)
                reinterpret_cast<int4 *>(dst)[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col0 + j < a.N) dst[j] = (int)v[j];
            }
          } else if (EPI == EPI_S8_COL32) {
            const float alpha = a.row_scale != nullptr ? a.row_scale[row] : 1.0f;
            signed char *dst = a.C8 + ((size_t)(col0 >> 5) * a.M + row) * 32;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
              uint32_t w4 = 0;
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const int q8 = max(-128, min(127, __float2int_rn(__fmul_rn(__int2float_rn((int)v[4 * j + b]), alpha))));
                w4 |= (uint32_t)(q8 & 0xFF) << (8 * b);
              }
              pk[j] = w4;
            }
            if (col0 + 32 <= a.N) {
              reinterpret_cast<uint4 *>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              reinterpret_cast<uint4 *>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col0 + j < a.N) dst[j] = (signed char)((pk[j >> 2] >> (8 * (j & 3))) & 0xFF);
            }
          } else if (EPI == EPI_INT32) {
            int *dst = a.C + (long)row * a.N + col0;
            if (col0 + 32 <= a.N && (a.N & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; j++)
                reinterpret_cast<int4 *>(dst)[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col0 + j < a.N) dst[j] = (int)v[j];
            }
          } else {
            __half h[32];
            const float *cs = s_cs + acc * BN + cl, *sb = s_bias + acc * BN + cl;
            // 32 independent chains, no branch inside: the single warp of a scheduler hides latency with ILP
#pragma unroll
            for (int j = 0; j < 32; j++) {
              float t = __fmul_rn(__int2float_rn((int)v[j]), 6.200012e-05f);
              t = __fmul_rn(t, rs);
              t = __fmul_rn(t, cs[j]);
              t = __fadd_rn(t, sb[j]);
              h[j] = __float2half_rn(t);
            }
            if (nout > 0) {   // + the 16-bit product over the outlier columns (warp-uniform branches, broadcast smem reads)
              const float4 *b = s_sb + (acc * BN + cl) * 2;
              if (nout <= 4) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                  const float4 b0 = b[2 * j];
                  float u = __fmul_rn(arow[0], b0.x);
                  u = __fmaf_rn(arow[1], b0.y, u);
                  u = __fmaf_rn(arow[2], b0.z, u);
                  u = __fmaf_rn(arow[3], b0.w, u);
                  h[j] = __float2half_rn(__fadd_rn(__half2float(h[j]), __half2float(__float2half_rn(u))));
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                  const float4 b0 = b[2 * j], b1 = b[2 * j + 1];
                  float u = __fmul_rn(arow[0], b0.x);
                  u = __fmaf_rn(arow[1], b0.y, u);
                  u = __fmaf_rn(arow[2], b0.z, u);
                  u = __fmaf_rn(arow[3], b0.w, u);
                  u = __fmaf_rn(arow[4], b1.x, u);
                  u = __fmaf_rn(arow[5], b1.y, u);
                  u = __fmaf_rn(arow[6], b1.z, u);
                  u = __fmaf_rn(arow[7], b1.w, u);
                  h[j] = __float2half_rn(__fadd_rn(__half2float(h[j]), __half2float(__float2half_rn(u))));
                }
              }
            }
            __half *dst = a.out + (long)row * a.N + col0;
            if (col0 + 32 <= a.N && (a.N & 7) == 0) {
#pragma unroll
              for (int j = 0; j < 4; j++) reinterpret_cast<uint4 *>(dst)[j] = reinterpret_cast<const uint4 *>(h)[j];
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (col0 + j < a.N) dst[j] = h[j];
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) cg2::mbar_arrive_remote(tc::smem_u32(tempty + acc), 0); else tc::mbar_arrive(tc::smem_u32(tempty + acc));
      }
    }
  }

  tc::fence_before_sync();
  if (CG == 2) cg2::cluster_sync(); else __syncthreads();    // nobody of the pair may still be reading TMEM or signalling a peer barrier
  if (warp == 1) { if (CG == 2) cg2::tmem_dealloc(tmem_base, kTmemCols); else tc::tmem_dealloc(tmem_base, kTmemCols); }
}

// ------------------------------------------------------------------------------------------------
// SIMT path (dp4a) for shapes TMA cannot describe (K % 16 != 0 or unaligned bases).  Exact as well.
// ------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) k_igemm_simt(const signed char *__restrict__ A, const signed char *__restrict__ B,
                                                    const IgemmArgs a) {
  __shared__ signed char sA[64][68], sB[64][68];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  int acc[4][4] = {};
  for (int k0 = 0; k0 < a.K; k0 += 64) {
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
      const int r = i / 64, c = i % 64;
      sA[r][c] = (m0 + r < a.M && k0 + c < a.K) ? A[(long)(m0 + r) * a.K + k0 + c] : 0;
      sB[r][c] = (n0 + r < a.N && k0 + c < a.K) ? B[(long)(n0 + r) * a.K + k0 + c] : 0;
    }
    __syncthreads();
    for (int kk = 0; kk < 64; kk += 4) {
      int av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        av[i] = *reinterpret_cast<const int *>(&sA[ty * 4 + i][kk]);
        bv[i] = *reinterpret_cast<const int *>(&sB[tx * 4 + i][kk]);
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = __dp4a(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
      if (r < a.M && c < a.N) {
        if (EPI == EPI_INT32) a.C[(long)r * a.N + c] = acc[i][j];
        else if (EPI == EPI_INT32_COL32) a.C[((size_t)(c >> 5) * a.M + r) * 32 + (c & 31)] = acc[i][j];
        else if (EPI == EPI_S8_COL32) {
          const float alpha = a.row_scale != nullptr ? a.row_scale[r] : 1.0f;
          a.C8[((size_t)(c >> 5) * a.M + r) * 32 + (c & 31)] =
              (signed char)max(-128, min(127, __float2int_rn(__fmul_rn(__int2float_rn(acc[i][j]), alpha))));
        } else {
          float t = __fmul_rn(__int2float_rn(acc[i][j]), 6.200012e-05f);
          t = __fmul_rn(t, a.rowStats[r]);
          t = __fmul_rn(t, a.colStats[c]);
          t = __fadd_rn(t, a.bias ? __half2float(a.bias[c]) : 0.f);
          a.out[(long)r * a.N + c] = __float2half_rn(t);
        }
      }
    }
}

static int env_force_simt() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("BNB_B200_IGEMM_IMPL"); v = (e && e[0] == 's') ? 1 : 0; }
  return v;
}

template <int EPI>
static int igemm_rowmajor(const signed char *A, const signed char *B, const IgemmArgs &a) {
  if (a.M <= 0 || a.N <= 0) return 0;
  if (a.K <= 0) { latch_error(cudaErrorInvalidValue, "igemm: k must be > 0"); return 2; }
  cudaStream_t st = current_stream();
  const bool tma_ok = (a.K % 16 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(B) % 16 == 0) && !env_force_simt();
  if (tma_ok) {
    CUtensorMap tmA, tmB;
    if (!make_tmap_2d(&tmA, A, 1, (uint64_t)a.M, (uint64_t)a.K, BM, BK, false, false)) return 2;
    int sms = kNumSMs, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static int one_cta = -1;   // BNB_B200_IGEMM_CG=1: one CTA per MMA everywhere (A/B measurements)
    if (one_cta < 0) { const char *e = getenv("BNB_B200_IGEMM_CG"); one_cta = (e && e[0] == '1') ? 1 : 0; }
    const bool pair = !one_cta && a.M > BM && sms >= 2;
    if (!make_tmap_2d(&tmB, B, 1, (uint64_t)a.N, (uint64_t)a.K, pair ? BN / 2 : BN, BK, false, false)) return 2;
    if (pair) {
      // CTA pairs: a cluster of two computes 256 x 256 tiles with tcgen05 cta_group::2
      ensure_max_dynamic_smem(reinterpret_cast<const void *>(k_igemm_tcgen05<EPI, 2>), kIgemmSmem, "igemm smem attr");
      const int tiles = ceil_div(a.M, 2 * BM) * ceil_div(a.N, BN);
      const int pairs = tiles < sms / 2 ? tiles : sms / 2;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * pairs);
      cfg.blockDim = dim3(kIgemmThreads);
      cfg.dynamicSmemBytes = kIgemmSmem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_igemm_tcgen05<EPI, 2>, tmA, tmB, a);
    } else {
      ensure_max_dynamic_smem(reinterpret_cast<const void *>(k_igemm_tcgen05<EPI, 1>), kIgemmSmem, "igemm smem attr");
      const int tiles = ceil_div(a.M, BM) * ceil_div(a.N, BN);
      const int grid = tiles < sms ? tiles : sms;
      k_igemm_tcgen05<EPI, 1><<<grid, kIgemmThreads, kIgemmSmem, st>>>(tmA, tmB, a);
    }
  } else {
    dim3 grid(ceil_div(a.N, 64), ceil_div(a.M, 64));
    k_igemm_simt<EPI><<<grid, 256, 0, st>>>(A, B, a);
  }
  cudaError_t e = cudaGetLastError();
  latch_error(e, "igemm launch");
  return e == cudaSuccess ? 0 : 2;
}

int igemm_rowmajor_32(int m, int n, int k, const signed char *A, const signed char *B, int *C) {
  IgemmArgs a{};
  a.M = m; a.N = n; a.K = k; a.C = C;
  return igemm_rowmajor<EPI_INT32>(A, B, a);
}
int igemm_rowmajor_dequant_fp16(int m, int n, int k, const signed char *A, const signed char *B, const float *rowStats,
                                const float *colStats, const __half *bias, __half *out) {
  IgemmArgs a{};
  a.M = m; a.N = n; a.K = k; a.rowStats = rowStats; a.colStats = colStats; a.bias = bias; a.out = out;
  return igemm_rowmajor<EPI_DEQUANT_FP16>(A, B, a);
}

int igemm_rowmajor_dequant_outliers_fp16(int m, int n, int k, const signed char *A, const signed char *B, const float *rowStats,
                                         const float *colStats, const __half *bias, __half *out, const __half *subA,
                                         const __half *subB, const int *count) {
  IgemmArgs a{};
  a.M = m; a.N = n; a.K = k; a.rowStats = rowStats; a.colStats = colStats; a.bias = bias; a.out = out;
  a.subA = subA; a.subB = subB; a.nout = count;
  // the outlier term exists in the tcgen05 epilogue only: refuse (caller takes the step-by-step route) whenever the SIMT
  // kernel would run, including when it is forced for debugging
  if ((k % 16) != 0 || (reinterpret_cast<uintptr_t>(A) % 16) != 0 || (reinterpret_cast<uintptr_t>(B) % 16) != 0 || env_force_simt()) return 1;
  return igemm_rowmajor<EPI_DEQUANT_FP16>(A, B, a);
}

// ------------------------------------------------------------------------------------------------
// reference ABI: A col32, B col_turing / col_ampere, C col32 (int32, or saturated int8 with fp32 alpha)
// ------------------------------------------------------------------------------------------------
// scratch for the un-permuted operands: one buffer per (device, stream), grown on demand (launches on a stream are
// ordered, so consecutive calls may reuse it; cudaFree of an outgrown buffer synchronises the device first)
struct LtKey { int dev; cudaStream_t st; bool operator<(const LtKey &o) const { return dev != o.dev ? dev < o.dev : st < o.st; } };
struct LtBuf { signed char *p; size_t bytes; };
static std::mutex g_lt_mu;
static std::map<LtKey, LtBuf> g_lt;
// *from_pool: the buffer came from cudaMallocAsync (stream capture in progress: the allocation becomes a node of the graph)
// and must be handed back with cudaFreeAsync after the last kernel that uses it
static signed char *lt_scratch(size_t bytes, cudaStream_t st, bool *from_pool) {
  *from_pool = false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs != cudaStreamCaptureStatusNone) {
    void *p = nullptr;
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *from_pool = true;
    return static_cast<signed char *>(p);
  }
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_lt_mu);
  LtBuf &b = g_lt[LtKey{dev, st}];
  if (b.bytes >= bytes) return b.p;
  if (b.p) cudaFree(b.p);
  if (cudaMalloc(&b.p, bytes) != cudaSuccess) { b.p = nullptr; b.bytes = 0; cudaGetLastError(); return nullptr; }
  b.bytes = bytes;
  return b.p;
}

// cigemmlt_<fmt>_<32|8|8_rowscale> (reference op_gemm.cpp:541-638): A col32, B col_turing / col_ampere, C col32.
// The tensor-core kernel wants K-major rows for TMA, so the two int8 operands are un-permuted into per-stream scratch
// (17 + 67 MB for config 3: ~10 % of the GEMM's time); C is written in col32 -- int32, or int8 saturated from
// rint(acc * alpha) -- straight from the epilogue: no int32 round trip through a third buffer, no allocation per call.
int igemmlt(int fmtB, int dtype_out, bool scale_rows, int m, int n, int k, const signed char *A, const signed char *B,
            void *C, const float *row_scale, int lda, int ldb, int ldc) {
  if (m <= 0 || n <= 0) return 0;
  if (k <= 0 || lda != m * 32 || ldc != m * 32) { latch_error(cudaErrorInvalidValue, "igemmlt: lda/ldc must be m*32 (col32)"); return 2; }
  (void)ldb;
  if (scale_rows && row_scale == nullptr) return 2;
  cudaStream_t st = current_stream();
  const size_t szA = ((size_t)m * k + 255) & ~(size_t)255, szB = ((size_t)n * k + 255) & ~(size_t)255;
  bool from_pool = false;
  signed char *scratch = lt_scratch(szA + szB, st, &from_pool);
  if (scratch == nullptr) { latch_error(cudaErrorMemoryAllocation, "igemmlt scratch"); return 2; }
  signed char *Arm = scratch, *Brm = scratch + szA;
  untransform_s8(COL32, A, Arm, m, k);
  untransform_s8(fmtB, B, Brm, n, k);
  IgemmArgs a{};
  a.M = m; a.N = n; a.K = k;
  int rc;
  if (dtype_out == 32) {
    a.C = reinterpret_cast<int *>(C);
    rc = igemm_rowmajor<EPI_INT32_COL32>(Arm, Brm, a);
  } else {
    a.C8 = reinterpret_cast<signed char *>(C);
    a.row_scale = scale_rows ? row_scale : nullptr;
    rc = igemm_rowmajor<EPI_S8_COL32>(Arm, Brm, a);
  }
  if (from_pool) cudaFreeAsync(scratch, st);
  return rc;
}

}  // namespace bnb
