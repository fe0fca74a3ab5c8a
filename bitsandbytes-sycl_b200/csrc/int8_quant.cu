// int8_quant.cu -- K5/K7 and the layout helpers of the LLM.int8 path for sm_100a:
//   get_col_row_stats   (reference kgetColRowStats,      kernel_quant.cpp:3214-3379)
//   double_rowcol_quant (reference kDoubleRowColQuant,   kernel_quant.cpp:3384-3512)
//   transform_row2fmt   (reference kTransformRowToFormat,kernel_quant.cpp:3516-3841)
//   dequant_mm_int32    (reference kdequant_mm_int32_fp16,kernel_quant.cpp:3848-3987)
//   extract_outliers    (reference kExtractOutliers,     kernel_quant.cpp:3992-4053)
// All are HBM streams.  Work decomposition: one WARP owns a (band of rows) x (256-column segment),
// i.e. the reference's 16x256 nnz tiles stacked; every lane moves 16 bytes per access, the column
// statistics live in registers across the band, row statistics are warp-shuffle reductions, and the
// float max is merged with integer atomicMax (all values are >= 0, so the int order is the float order).
#include "common.cuh"

namespace bnb {

constexpr int kBandRows = 32;   // rows per warp band (2 of the reference's 16-row nnz tiles; 16 measured the same 20.7 us)
constexpr float kMMDequantConst = 6.200012e-05f;  // kernel_quant.cpp:3846

__device__ __forceinline__ void atomic_max_nonneg(float *addr, float v) {
  // v >= +0; existing value may be negative (caller initialises to -50000, functional.py:2413-2415)
  if (*addr < v) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
}

// ------------------------------------------------------------------------------------------------
// K5a: row/col absmax (+ per (tile,row) outlier counts when thr > 0)
// ------------------------------------------------------------------------------------------------
// butterfly reduce-scatter of 8 per-lane values: afterwards v[0] of lane l is the warp-wide reduction of everybody's
// v[l & 7] (9 shuffles for 8 rows instead of 5 per row)
template <typename V, typename Op>
__device__ __forceinline__ void warp_reduce_scatter8(V (&v)[8], int lane, Op op) {
#pragma unroll
  for (int o = 4; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; i++) {
      const V send = up ? v[i] : v[i + o];
      const V keep = up ? v[i + o] : v[i];
      v[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
  v[0] = op(v[0], __shfl_xor_sync(0xffffffffu, v[0], 8));
  v[0] = op(v[0], __shfl_xor_sync(0xffffffffu, v[0], 16));
}

constexpr int kStatWarpRows = 8;    // rows per warp, all loads in flight at once (16 rows = 64 registers held the SM at 30 % of its warps)
constexpr int kStatCtaRows = 8 * kStatWarpRows;

// CTA = 8 warps stacked on ONE 256-column segment (128 rows): column maxima meet in shared memory first, so the global
// atomics are 256 per CTA; row maxima: one atomic per (row, segment), issued by 8 lanes at a time.
// Persistent: four CTAs per SM walk the (row band, column tile) units, so the whole matrix is requested in one wave of
// resident CTAs instead of 3.5 one-shot CTAs per SM with a ragged tail.
template <bool VEC, bool SPARSE>
__global__ void __launch_bounds__(256, 4) k_col_row_stats(const __half *__restrict__ A, float *rowStats, float *colStats,
                                                          int *nnz_count_row, float thr, int rows, int cols, int col_tiles, int units) {
  __shared__ int s_cmax[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows16 = ((rows + 15) / 16) * 16;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
  const int ct = unit % col_tiles, rband = unit / col_tiles;
  const int c0 = ct * 256 + lane * 8;
  const int r0 = rband * kStatCtaRows + warp * kStatWarpRows;
  s_cmax[threadIdx.x] = 0;
  __syncthreads();

  uint4 raw[kStatWarpRows];
#pragma unroll
  for (int u = 0; u < kStatWarpRows; u++) {
    const int r = r0 + u;
    raw[u] = make_uint4(0, 0, 0, 0);
    if (r < rows) {
      if (VEC) {
        if (c0 < cols) raw[u] = ld_stream_u4(A + (long)r * cols + c0);
      } else {
        __half *p = reinterpret_cast<__half *>(&raw[u]);
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (c0 + j < cols) p[j] = A[(long)r * cols + c0 + j];
      }
    }
  }
  // |x| >= 0, and padded / outlier entries count as 0.  Column maxima are kept as packed fp16 pairs (a maximum of halves is
  // exact in fp16): |x| is a mask, eight columns advance with four HMNMX2, and a chunk goes element by element only when
  // its own maximum reaches the threshold (compared in fp32 like the reference) -- ~2 instructions per element instead of 6.
  __half2 cmax2[4];
#pragma unroll
  for (int k = 0; k < 4; k++) cmax2[k] = __float2half2_rn(0.0f);
#pragma unroll
  for (int h = 0; h < kStatWarpRows / 8; h++) {
    float rm[8];
    int cn[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint4 rw = raw[h * 8 + u];
      const uint32_t aw[4] = {rw.x & 0x7FFF7FFFu, rw.y & 0x7FFF7FFFu, rw.z & 0x7FFF7FFFu, rw.w & 0x7FFF7FFFu};
      const __half2 *a2 = reinterpret_cast<const __half2 *>(aw);
      const __half2 m = __hmax2(__hmax2(a2[0], a2[1]), __hmax2(a2[2], a2[3]));
      const float chunk_max = fmaxf(__low2float(m), __high2float(m));     // NaN ignored by __hmax2 / fmaxf, as below
      float rmax = 0.0f;
      int cnt = 0;
      if (SPARSE && chunk_max >= thr) {
        const __half *p = reinterpret_cast<const __half *>(aw);
        uint32_t kept[4];
        __half *kp = reinterpret_cast<__half *>(kept);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float v = __half2float(p[j]);
          if (v >= thr) { cnt++; v = 0.0f; }
          kp[j] = __float2half_rn(v);                                     // exact: v is a half or zero
          rmax = fmaxf(rmax, v);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) cmax2[k] = __hmax2(cmax2[k], *reinterpret_cast<const __half2 *>(&kept[k]));
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) cmax2[k] = __hmax2(cmax2[k], a2[k]);
        rmax = fmaxf(chunk_max, 0.0f);                                    // a chunk of NaNs only: 0, as the element loop gives
      }
      rm[u] = rmax;
      cn[u] = cnt;
    }
    warp_reduce_scatter8(rm, lane, [](float a, float b) { return fmaxf(a, b); });
    const int r = r0 + h * 8 + (lane & 7);
    if (lane < 8 && r < rows) atomicMax(reinterpret_cast<int *>(rowStats + r), __float_as_int(rm[0]));   // values >= +0: int order == float order
    if (SPARSE) {
      warp_reduce_scatter8(cn, lane, [](int a, int b) { return a + b; });
      // tile id = (r/16)*col_tiles + ct; slot +1 (slot 0 stays 0 for the caller's cumsum)
      if (lane < 8 && nnz_count_row && r < rows16)
        nnz_count_row[((long)(r / 16) * col_tiles + ct) * 16 + (r % 16) + 1] = (r < rows) ? cn[0] : 0;
    }
  }
  float cmax[8];
#pragma unroll
  for (int k = 0; k < 4; k++) { cmax[2 * k] = __low2float(cmax2[k]); cmax[2 * k + 1] = __high2float(cmax2[k]); }
#pragma unroll
  for (int j = 0; j < 8; j++) atomicMax(&s_cmax[lane * 8 + j], __float_as_int(cmax[j]));
  __syncthreads();
  const int c = ct * 256 + threadIdx.x;
  if (c < cols) atomicMax(reinterpret_cast<int *>(colStats + c), s_cmax[threadIdx.x]);
  __syncthreads();                                   // s_cmax is reset by the next unit
  }
}

void get_col_row_stats(const __half *A, float *rowStats, float *colStats, int *nnz_count_row, float thr, int rows,
                       int cols) {
  if (rows <= 0 || cols <= 0) return;
  const int col_tiles = ceil_div(cols, 256);
  const int rbands = ceil_div(ceil_div(rows, 16) * 16, kStatCtaRows);
  const long units_l = (long)rbands * col_tiles;
  if (units_l > 0x7fffffffL) { latch_error(cudaErrorInvalidValue, "get_col_row_stats: matrix too large"); return; }
  const int units = (int)units_l;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const unsigned grid = (unsigned)(units < sms * 4 ? units : sms * 4);
  const bool vec = (cols % 8 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0);
  cudaStream_t st = current_stream();
#define STATS_LAUNCH(V_, S_) k_col_row_stats<V_, S_><<<grid, 256, 0, st>>>(A, rowStats, colStats, nnz_count_row, thr, rows, cols, col_tiles, units)
  if (thr > 0.0f) { if (vec) STATS_LAUNCH(true, true); else STATS_LAUNCH(false, true); }
  else { if (vec) STATS_LAUNCH(true, false); else STATS_LAUNCH(false, false); }
#undef STATS_LAUNCH
  check_launch("get_col_row_stats");
}

// ------------------------------------------------------------------------------------------------
// K5b: double quant.  out_row = (int8)rint(x * (127/rowStat)), out_col = (int8)rint(x * (127/colStat));
// IEEE divide, one multiply, round-half-even, saturating convert, NaN -> 0.  thr > 0: outliers get
// out_row = 0 and a COO entry at nnz_row_ptr[tile*16 + r%16] + (rank of the column inside the segment)
// -- ascending column order inside each (tile,row) segment, so the COO is deterministic.
// ------------------------------------------------------------------------------------------------
// == quant_s8 of an already scaled value: rint, saturate to int8, NaN -> 0, in one instruction; the byte sits in bits [0, 8)
__device__ __forceinline__ uint32_t f2s8_sat(float x) {
  int q;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(x));
  return (uint32_t)q;
}
__device__ __forceinline__ int quant_s8(float x, float scale) {
  int q = __float2int_rn(__fmul_rn(x, scale));  // NaN -> 0, saturates at int32
  return max(-128, min(127, q));
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_double_rowcol_quant(const __half *__restrict__ A, const float *__restrict__ rowStats,
                                                             const float *__restrict__ colStats, signed char *out_col,
                                                             signed char *out_row, int *rowidx, int *colidx, __half *val,
                                                             const int *__restrict__ nnz_row_ptr, float thr, int rows,
                                                             int cols, int col_tiles, int nbands) {
  const int lane = threadIdx.x & 31;
  const long wid = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= (long)nbands * col_tiles) return;
  const int band = (int)(wid / col_tiles), ct = (int)(wid % col_tiles);
  const int c0 = ct * 256 + lane * 8;
  const int r0 = band * kBandRows;
  const bool sparse = thr > 0.0f;

  float cscale[8];
#pragma unroll
  for (int j = 0; j < 8; j++) cscale[j] = (c0 + j < cols) ? __fdiv_rn(127.0f, colStats[c0 + j]) : 0.0f;
  // the band's 32 row scales: one load + one IEEE divide per lane up front, broadcast by shuffle in the loop (a load and a
  // divide per row inside the loop put an L2 round trip in front of every four rows)
  static_assert(kBandRows == 32, "one row scale per lane");
  const float my_rscale = (r0 + lane < rows) ? __fdiv_rn(127.0f, __ldg(rowStats + r0 + lane)) : 0.0f;

  constexpr int kRB = 4;                          // rows in flight per warp (8 measured the same 20.7 us)
  for (int rb = 0; rb < kBandRows; rb += kRB) {
    uint4 raw[kRB];
#pragma unroll
    for (int u = 0; u < kRB; u++) {
      const int r = r0 + rb + u;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (r < rows) {
        if (VEC) {
          if (c0 < cols) raw[u] = ld_stream_u4(A + (long)r * cols + c0);
        } else {
          __half *p = reinterpret_cast<__half *>(&raw[u]);
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (c0 + j < cols) p[j] = A[(long)r * cols + c0 + j];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kRB; u++) {
      const int r = r0 + rb + u;
      if (r >= rows) continue;  // warp-uniform
      const __half *p = reinterpret_cast<const __half *>(&raw[u]);
      const float rscale = __shfl_sync(0xffffffffu, my_rscale, rb + u);
      uint32_t qr[2], qc[2];
      uint32_t outlier_mask = 0;
      uint32_t br[8], bc[8];                      // rint + saturate in one conversion each, bytes gathered with three PRMT per word
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float x = __half2float(p[j]);
        const bool outl = sparse && fabsf(x) >= thr && c0 + j < cols;
        if (outl) outlier_mask |= 1u << j;
        br[j] = outl ? 0u : f2s8_sat(__fmul_rn(x, rscale));
        bc[j] = f2s8_sat(__fmul_rn(x, cscale[j]));
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
        qr[h] = __byte_perm(__byte_perm(br[4 * h], br[4 * h + 1], 0x0040), __byte_perm(br[4 * h + 2], br[4 * h + 3], 0x0040), 0x5410);
        qc[h] = __byte_perm(__byte_perm(bc[4 * h], bc[4 * h + 1], 0x0040), __byte_perm(bc[4 * h + 2], bc[4 * h + 3], 0x0040), 0x5410);
      }
      const long o = (long)r * cols + c0;
      if (VEC) {
        if (c0 < cols) {
          *reinterpret_cast<uint2 *>(out_row + o) = make_uint2(qr[0], qr[1]);
          *reinterpret_cast<uint2 *>(out_col + o) = make_uint2(qc[0], qc[1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (c0 + j < cols) {
            out_row[o + j] = (signed char)(qr[j >> 2] >> (8 * (j & 3)));
            out_col[o + j] = (signed char)(qc[j >> 2] >> (8 * (j & 3)));
          }
      }
      if (sparse && rowidx != nullptr) {
        // exclusive prefix of the per-lane outlier counts -> ascending-column slots
        const int mine = __popc(outlier_mask);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        if (mine) {
          int slot = nnz_row_ptr[((long)(r / 16) * col_tiles + ct) * 16 + (r % 16)] + incl - mine;
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (outlier_mask & (1u << j)) {
              rowidx[slot] = r;
              colidx[slot] = c0 + j;
              val[slot] = p[j];
              slot++;
            }
        }
      }
    }
  }
}

void double_rowcol_quant(const __half *A, const float *rowStats, const float *colStats, signed char *out_col,
                         signed char *out_row, int *rowidx, int *colidx, __half *val, const int *nnz_row_ptr,
                         float thr, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return;
  const int col_tiles = ceil_div(cols, 256);
  const int nbands = ceil_div(rows, kBandRows);
  const unsigned grid = (unsigned)ceil_div_ll((long)nbands * col_tiles, 8);
  const bool vec = (cols % 8 == 0) && (reinterpret_cast<uintptr_t>(A) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out_row) % 8 == 0) && (reinterpret_cast<uintptr_t>(out_col) % 8 == 0);
  if (vec) k_double_rowcol_quant<true><<<grid, 256, 0, current_stream()>>>(A, rowStats, colStats, out_col, out_row, rowidx, colidx, val, nnz_row_ptr, thr, rows, cols, col_tiles, nbands);
  else k_double_rowcol_quant<false><<<grid, 256, 0, current_stream()>>>(A, rowStats, colStats, out_col, out_row, rowidx, colidx, val, nnz_row_ptr, thr, rows, cols, col_tiles, nbands);
  check_launch("double_rowcol_quant");
}

// ------------------------------------------------------------------------------------------------
// layouts (kernel_quant.cpp:3673-3675, :3740-3755, :3822-3832 == blas_utils.h:263-325)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ long layout_out_rows(int fmt, int rows) {
  if (fmt == COL_TURING) return ((rows + 7) / 8) * 8L;
  if (fmt == COL_AMPERE) return ((rows + 31) / 32) * 32L;
  return rows;
}
__host__ __device__ __forceinline__ long layout_offset(int fmt, long out_rows, int r, int c) {
  const int c32 = c & 31;
  if (fmt == COL32) return (long)(c >> 5) * 32 * out_rows + (long)r * 32 + c32;
  if (fmt == COL_TURING) {
    long off = (long)(c >> 5) * out_rows * 32 + (long)(r >> 3) * 256 + (c32 >> 2) * 16 + (c32 & 3);
    return off + ((r & 1) ? 128 + ((r & 7) - 1) * 2 : (r & 7) * 2);
  }
  const int lr = r & 31;
  const int ar = ((lr & 7) >> 1) * 8 + (lr >> 3) * 2 + (lr & 1);
  return (long)(c >> 5) * out_rows * 32 + (long)(r >> 5) * 1024 + ar * 32 + c32;
}

// row-major [rows, cols] -> fmt.  Each thread moves 4 consecutive columns of one row (4 consecutive
// columns are contiguous in all three layouts); a warp covers 128 columns of a row.
template <int FMT, typename E>
__global__ void __launch_bounds__(256) k_transform(const E *__restrict__ A, E *__restrict__ out, int rows, int cols,
                                                   long out_rows) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cgroups = (cols + 3) / 4;
  if (idx >= (long)rows * cgroups) return;
  const int r = (int)(idx / cgroups), c = (int)(idx % cgroups) * 4;
  const E *src = A + (long)r * cols + c;
  E *dst = out + layout_offset(FMT, out_rows, r, c);
  if (sizeof(E) == 1 && c + 4 <= cols && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(src);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (c + j < cols) dst[j] = src[j];
  }
}
// transposed: out = fmt(A^T).  Thread per element of A^T with the A^T column (= A row) fastest, so the
// writes are the coalesced side.
template <int FMT, typename E>
__global__ void __launch_bounds__(256) k_transform_T(const E *__restrict__ A, E *__restrict__ out, int rows, int cols,
                                                     long out_rows) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)rows * cols) return;
  const int c = (int)(idx / rows), r = (int)(idx % rows);  // element A[r][c] == A^T[c][r]
  out[layout_offset(FMT, out_rows, c, r)] = A[(long)r * cols + c];
}


// Tiled int8 layout kernels (rows % 32 == 0, cols % 32 == 0, 16-byte aligned buffers: every LLM.int8 shape).  In all three
// layouts the block of (one 32-column panel) x (32 rows starting at a multiple of 32) is ONE contiguous kilobyte: col32 has
// the 32 bytes of a row back to back, col_turing four 256-byte tiles of 8 rows, col_ampere one 1024-byte tile.  A CTA
// moves 32 rows x 256 columns through shared memory: 16-byte accesses, fully coalesced on the row-major side (256 bytes
// of a row per half-warp) AND on the layout side (a kilobyte per 64 threads), where the element-per-thread kernels
// above wrote 4 bytes at a time into scattered sectors.  TO_ROW = false: row-major -> layout; true: layout -> row-major.
constexpr int kLtRows = 32, kLtCols = 256, kLtPitch = kLtCols + 16;   // pitch keeps 16-byte alignment, spreads the banks
template <int FMT, bool TO_ROW>
__global__ void __launch_bounds__(256) k_layout_tiled(const signed char *__restrict__ src, signed char *__restrict__ dst, int rows,
                                                      int cols, long lay_rows) {
  __shared__ __align__(16) signed char tile[kLtRows][kLtPitch];
  const int t = threadIdx.x;
  const int c_base = blockIdx.x * kLtCols, r_base = blockIdx.y * kLtRows;
  const int npanels = min(kLtCols, cols - c_base) >> 5;               // valid 32-column panels of this tile
  const signed char *rm_c = TO_ROW ? nullptr : src;
  signed char *rm = TO_ROW ? dst : const_cast<signed char *>(rm_c);    // the row-major buffer
  // ---- row-major side: thread = (row t / 16 (+16), 16-byte chunk t % 16)
  auto row_major_pass = [&](bool load) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = (t >> 4) + h * 16, ch = t & 15;
      if ((ch >> 1) < npanels) {
        signed char *g = rm + (long)(r_base + r) * cols + c_base + ch * 16;
        uint4 *sm = reinterpret_cast<uint4 *>(&tile[r][ch * 16]);
        if (load) *sm = *reinterpret_cast<const uint4 *>(g); else *reinterpret_cast<uint4 *>(g) = *sm;
      }
    }
  };
  // ---- layout side: panel p = piece / 64, 16-byte piece o16 = piece % 64 of its kilobyte
  auto layout_pass = [&](bool load) {
    const signed char *lay_c = TO_ROW ? src : nullptr;
    signed char *lay = TO_ROW ? const_cast<signed char *>(lay_c) : dst;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int piece = t + h * 256;
      const int p = piece >> 6, o = (piece & 63) * 16;
      if (p >= npanels) continue;
      signed char *g = lay + ((long)((c_base >> 5) + p) * lay_rows + r_base) * 32 + o;
      if (FMT == COL_TURING) {
        // 256-byte tile of 8 rows: bytes [0,128) even rows, [128,256) odd rows; 16 bytes = 4 rows x 4 columns
        const int tile8 = o >> 8, o2 = o & 255, half = o2 >> 7, grp = (o2 & 127) >> 4;
        uint32_t w[4];
        if (load) {
          const uint4 v = *reinterpret_cast<const uint4 *>(g);
          w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint32_t *sm = reinterpret_cast<uint32_t *>(&tile[tile8 * 8 + 2 * j + half][p * 32 + grp * 4]);
          if (load) *sm = w[j]; else w[j] = *sm;
        }
        if (!load) *reinterpret_cast<uint4 *>(g) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        int r = o >> 5;                                                // col32: rows back to back
        if (FMT == COL_AMPERE) r = (r >> 3) * 2 + (r & 1) + ((r >> 1) & 3) * 8;   // inverse of ar = ((lr & 7) >> 1) * 8 + (lr >> 3) * 2 + (lr & 1)
        uint4 *sm = reinterpret_cast<uint4 *>(&tile[r][p * 32 + (o & 31)]);
        if (load) *sm = *reinterpret_cast<const uint4 *>(g); else *reinterpret_cast<uint4 *>(g) = *sm;
      }
    }
  };
  if (TO_ROW) layout_pass(true); else row_major_pass(true);
  __syncthreads();
  if (TO_ROW) row_major_pass(false); else layout_pass(false);
}
static bool layout_tiled_ok(const void *a, const void *b, int rows, int cols) {
  return rows % 32 == 0 && cols % 32 == 0 && (reinterpret_cast<uintptr_t>(a) % 16) == 0 && (reinterpret_cast<uintptr_t>(b) % 16) == 0 &&
         rows / 32 <= 65535;
}
template <int FMT, bool TO_ROW>
static void launch_layout_tiled(const signed char *src, signed char *dst, int rows, int cols) {
  const dim3 grid((unsigned)ceil_div(cols, kLtCols), (unsigned)(rows / kLtRows));
  k_layout_tiled<FMT, TO_ROW><<<grid, 256, 0, current_stream()>>>(src, dst, rows, cols, layout_out_rows(FMT, rows));
}

template <int FMT>
void transform_row2fmt(const signed char *A, signed char *out, int rows, int cols, bool transpose) {
  if (rows <= 0 || cols <= 0) return;
  cudaStream_t st = current_stream();
  if (!transpose && layout_tiled_ok(A, out, rows, cols)) {
    launch_layout_tiled<FMT, false>(A, out, rows, cols);
  } else if (!transpose) {
    const long n = (long)rows * ((cols + 3) / 4);
    k_transform<FMT, signed char><<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(A, out, rows, cols, layout_out_rows(FMT, rows));
  } else {
    const long n = (long)rows * cols;
    k_transform_T<FMT, signed char><<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(A, out, rows, cols, layout_out_rows(FMT, cols));
  }
  check_launch("transform_row2fmt");
}
template void transform_row2fmt<COL32>(const signed char *, signed char *, int, int, bool);
template void transform_row2fmt<COL_TURING>(const signed char *, signed char *, int, int, bool);
template void transform_row2fmt<COL_AMPERE>(const signed char *, signed char *, int, int, bool);

// fmt -> row-major (used by the cigemmlt_* ABI wrappers to feed the row-major tcgen05 GEMM)
template <int FMT, typename E>
__global__ void __launch_bounds__(256) k_untransform(const E *__restrict__ A, E *__restrict__ out, int rows, int cols,
                                                     long in_rows) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cgroups = (cols + 3) / 4;
  if (idx >= (long)rows * cgroups) return;
  const int r = (int)(idx / cgroups), c = (int)(idx % cgroups) * 4;
  const E *src = A + layout_offset(FMT, in_rows, r, c);
  E *dst = out + (long)r * cols + c;
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (c + j < cols) dst[j] = src[j];
}
void untransform_s8(int fmt, const signed char *A, signed char *out, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return;
  if (layout_tiled_ok(A, out, rows, cols)) {
    if (fmt == COL32) launch_layout_tiled<COL32, true>(A, out, rows, cols);
    else if (fmt == COL_TURING) launch_layout_tiled<COL_TURING, true>(A, out, rows, cols);
    else launch_layout_tiled<COL_AMPERE, true>(A, out, rows, cols);
    check_launch("untransform_s8 (tiled)");
    return;
  }
  const long n = (long)rows * ((cols + 3) / 4);
  const unsigned grid = (unsigned)ceil_div_ll(n, 256);
  cudaStream_t st = current_stream();
  if (fmt == COL32) k_untransform<COL32, signed char><<<grid, 256, 0, st>>>(A, out, rows, cols, layout_out_rows(COL32, rows));
  else if (fmt == COL_TURING) k_untransform<COL_TURING, signed char><<<grid, 256, 0, st>>>(A, out, rows, cols, layout_out_rows(COL_TURING, rows));
  else k_untransform<COL_AMPERE, signed char><<<grid, 256, 0, st>>>(A, out, rows, cols, layout_out_rows(COL_AMPERE, rows));
  check_launch("untransform_s8");
}
// row-major int32 / int8 -> col32 (C operand of the cigemmlt_* ABI)
template <typename E>
void to_col32(const E *A, E *out, int rows, int cols) {
  const long n = (long)rows * ((cols + 3) / 4);
  k_transform<COL32, E><<<(unsigned)ceil_div_ll(n, 256), 256, 0, current_stream()>>>(A, out, rows, cols, rows);
  check_launch("to_col32");
}
template void to_col32<int>(const int *, int *, int, int);
template void to_col32<signed char>(const signed char *, signed char *, int, int);

// ------------------------------------------------------------------------------------------------
// K7: outlier column gather out[row, j] = A_fmt(row, idx[j])
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) k_extract_outliers(const signed char *__restrict__ A, const int *__restrict__ idx,
                                                          signed char *__restrict__ out, int idx_size, int rows,
                                                          long fmt_rows) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)rows * idx_size) return;
  const int r = (int)(i / idx_size), j = (int)(i % idx_size);
  out[i] = A[layout_offset(FMT, fmt_rows, r, idx[j])];
}
template <int FMT>
void extract_outliers(const signed char *A, const int *idx, signed char *out, int idx_size, int rows, int cols) {
  (void)cols;
  if (rows <= 0 || idx_size <= 0) return;
  const long n = (long)rows * idx_size;
  k_extract_outliers<FMT><<<(unsigned)ceil_div_ll(n, 256), 256, 0, current_stream()>>>(A, idx, out, idx_size, rows, layout_out_rows(FMT, rows));
  check_launch("extract_outliers");
}
template void extract_outliers<COL_TURING>(const signed char *, const int *, signed char *, int, int, int);
template void extract_outliers<COL_AMPERE>(const signed char *, const int *, signed char *, int, int, int);

// ------------------------------------------------------------------------------------------------
// a9: int32 (col32) -> fp16 row-major, out = half(((float(c) * K) * rowStat) * colStat + bias);
// four fp32 operations in the reference's order, no contraction.
// A lane owns 4 consecutive columns of one row: 128-bit load from the col32 tile, 64-bit store.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dequant_one(int c, float rs, float cs, float b) {
  float t = __fmul_rn(__int2float_rn(c), kMMDequantConst);
  t = __fmul_rn(t, rs);
  t = __fmul_rn(t, cs);
  return __fadd_rn(t, b);
}

__global__ void __launch_bounds__(256) k_dequant_mm_int32_fp16(const int *__restrict__ A, const float *__restrict__ rowStats,
                                                               const float *__restrict__ colStats, __half *__restrict__ out,
                                                               const __half *__restrict__ bias, int numRows, int numCols) {
  // work item = (col group of 32, row, quad of 4 columns); quads fastest, then rows: a warp covers
  // 4 rows x 32 columns = 4 x 128 contiguous bytes of the col32 tile
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cgroups = (numCols + 31) / 32;
  if (idx >= (long)cgroups * numRows * 8) return;
  const int quad = (int)(idx & 7);
  const long t = idx >> 3;
  const int r = (int)(t % numRows);
  const int cg = (int)(t / numRows);
  const int c = cg * 32 + quad * 4;
  if (c >= numCols) return;
  const int4 v = *reinterpret_cast<const int4 *>(A + ((long)cg * numRows + r) * 32 + quad * 4);
  const float rs = __ldg(rowStats + r);
  const int vals[4] = {v.x, v.y, v.z, v.w};
  __half h[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int cc = c + j;
    const float cs = cc < numCols ? __ldg(colStats + cc) : 0.0f;
    const float b = (bias != nullptr && cc < numCols) ? __half2float(bias[cc]) : 0.0f;
    h[j] = __float2half_rn(dequant_one(vals[j], rs, cs, b));
  }
  __half *dst = out + (long)r * numCols + c;
  if (c + 4 <= numCols && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(h);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (c + j < numCols) dst[j] = h[j];
  }
}

void dequant_mm_int32_fp16(const int *A, const float *rowStats, const float *colStats, __half *out, const __half *bias,
                           int numRows, int numCols) {
  if (numRows <= 0 || numCols <= 0) return;
  const long n = (long)((numCols + 31) / 32) * numRows * 8;
  k_dequant_mm_int32_fp16<<<(unsigned)ceil_div_ll(n, 256), 256, 0, current_stream()>>>(A, rowStats, colStats, out, bias, numRows, numCols);
  check_launch("dequant_mm_int32_fp16");
}

}  // namespace bnb
