/*
 * bnb_oracle.c -- CPU restatement of the reference's quantized-linear hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (bitsandbytes-sycl_b200/) never links, imports or calls anything in oracle/.
 *
 * Strict IEEE fp32, scalar, single-threaded (unless the caller threads it).
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (see oracle/Makefile).
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Where the reference port is visibly broken (SURVEY.md section 8a "WIP artefacts") the
 * kernel-body arithmetic / upstream intent is followed, never the broken launcher.
 *
 * Parity pinning: the reference holds no golden vectors for this path.
 * The oracle is pinned by (i) the literal constants copied from the kernels (tables,
 * thresholds, MM_DEQUANT_CONST), cross-checked in tests against tests/golden/ fixtures
 * generated from the reference's own Python sources, (ii) the reference's own
 * sycl/cpu_ops.cpp compiled into oracle/_ref/ (8-bit blockwise path, >=99.99% code
 * agreement + exact absmax + exact dequantize), (iii) exact integer arithmetic for
 * igemmlt and the layout permutations (checked against blas_utils.h's index maps).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* dtype tags shared with oracle/oracle.py */
enum { ORC_F32 = 0, ORC_F16 = 1, ORC_BF16 = 2 };
/* DataType_t of sycl/sycl_code/ops.h:87-92 */
enum { ORC_GENERAL8BIT = 0, ORC_FP4 = 1, ORC_NF4 = 2 };
/* Transform_t subset of sycl/sycl_code/ops.h:78-85 */
enum { ORC_COL32 = 0, ORC_COL_TURING = 1, ORC_COL_AMPERE = 2 };

/* ------------------------------------------------------------------------------------
 * 16-bit float helpers (round-to-nearest-even, like sycl::half / bfloat16 conversions)
 * ---------------------------------------------------------------------------------- */
static inline float bf16_bits_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u); /* qNaN */
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}
static inline float f16_bits_to_f32(uint16_t h) {
  _Float16 x;
  memcpy(&x, &h, 2);
  return (float)x;
}
static inline uint16_t f32_to_f16_bits(float f) {
  _Float16 x = (_Float16)f;
  uint16_t h;
  memcpy(&h, &x, 2);
  return h;
}
static inline float load_as_f32(const void *p, int dtype, long i) {
  switch (dtype) {
    case ORC_F16: return f16_bits_to_f32(((const uint16_t *)p)[i]);
    case ORC_BF16: return bf16_bits_to_f32(((const uint16_t *)p)[i]);
    default: return ((const float *)p)[i];
  }
}
static inline void store_from_f32(void *p, int dtype, long i, float v) {
  switch (dtype) {
    case ORC_F16: ((uint16_t *)p)[i] = f32_to_f16_bits(v); break;
    case ORC_BF16: ((uint16_t *)p)[i] = f32_to_bf16_bits(v); break;
    default: ((float *)p)[i] = v; break;
  }
}
/* round an fp32 value to T and back (T-arithmetic emulation) */
static inline float round_to_T(float v, int dtype) {
  switch (dtype) {
    case ORC_F16: return f16_bits_to_f32(f32_to_f16_bits(v));
    case ORC_BF16: return bf16_bits_to_f32(f32_to_bf16_bits(v));
    default: return v;
  }
}
ORC_API float orc_round_to_dtype(float v, int dtype) { return round_to_T(v, dtype); }

/* ------------------------------------------------------------------------------------
 * Codebooks.  sycl/sycl_code/kernel_quant.cpp:650-703 (dDequantizeNF4),
 * :705-756 (dQuantizeNF4), :547-594 (dQuantizeFP4), :520-545 (dDequantizeFP4Tree).
 * ---------------------------------------------------------------------------------- */
static const float NF4_TABLE[16] = {
    -1.0f,
    -0.6961928009986877f,
    -0.5250730514526367f,
    -0.39491748809814453f,
    -0.28444138169288635f,
    -0.18477343022823334f,
    -0.09105003625154495f,
    0.0f,
    0.07958029955625534f,
    0.16093020141124725f,
    0.24611230194568634f,
    0.33791524171829224f,
    0.44070982933044434f,
    0.5626170039176941f,
    0.7229568362236023f,
    1.0f,
};
ORC_API void orc_nf4_table(float *out) { memcpy(out, NF4_TABLE, sizeof(NF4_TABLE)); }

/* decision tree exactly as written at kernel_quant.cpp:705-756 (strict '>' compares) */
static inline unsigned char quantize_nf4(float x) {
  if (x > 0.03979014977812767f)
    if (x > 0.3893125355243683f)
      if (x > 0.6427869200706482f)
        if (x > 0.8614784181118011f) return 15; else return 14;
      else
        if (x > 0.5016634166240692f) return 13; else return 12;
    else
      if (x > 0.2035212516784668f)
        if (x > 0.2920137718319893f) return 11; else return 10;
      else
        if (x > 0.1202552504837513f) return 9; else return 8;
  else
    if (x > -0.33967943489551544f)
      if (x > -0.13791173323988914f)
        if (x > -0.045525018125772476f) return 7; else return 6;
      else
        if (x > -0.23460740596055984f) return 5; else return 4;
    else
      if (x > -0.6106329262256622f)
        if (x > -0.4599952697753906f) return 3; else return 2;
      else
        if (x > -0.8480964004993439f) return 1; else return 0;
}
ORC_API unsigned char orc_quantize_nf4_scalar(float x) { return quantize_nf4(x); }

/* kernel_quant.cpp:547-594 */
static inline unsigned char quantize_fp4(float x) {
  int sign = x < 0 ? 8 : 0;
  x = fabsf(x);
  if (x > 0.29166667f)
    if (x > 0.583333f)
      if (x > 0.8333333f) return 3 + sign; else return 2 + sign;
    else
      if (x > 0.4166667f) return 5 + sign; else return 4 + sign;
  else
    if (x > 0.0859375f)
      if (x > 0.20833333f) return 7 + sign; else return 6 + sign;
    else
      if (x > 0.00260417f) return 1 + sign; else return 0 + sign;
}
ORC_API unsigned char orc_quantize_fp4_scalar(float x) { return quantize_fp4(x); }

/* kernel_quant.cpp:520-545: value = (c * absmax) * sign, left to right */
static const float FP4_MAG[8] = {0.00000000f, 5.208333333e-03f, 0.66666667f, 1.00000000f,
                                 0.33333333f, 0.50000000f,      0.16666667f, 0.25000000f};
static inline float dequantize_fp4_tree(unsigned char val, float absmax) {
  float sign = (val & 8) ? -1.0f : 1.0f;
  return FP4_MAG[val & 7] * absmax * sign;
}
ORC_API void orc_fp4_table(float *out) {
  for (int i = 0; i < 16; i++) out[i] = dequantize_fp4_tree((unsigned char)i, 1.0f);
}

/* kernel_quant.cpp:765-819, dQuantize<STOCHASTIC=0>: 7-step pivot search + midpoint */
static inline unsigned char quantize_8bit(const float *code, float x) {
  int pivot = 127, upper_pivot = 255, lower_pivot = 0;
  float lower = -1.0f, upper = 1.0f;
  float val = code[pivot];
  for (int i = 64; i > 0; i >>= 1) {
    if (x > val) { lower_pivot = pivot; lower = val; pivot += i; }
    else         { upper_pivot = pivot; upper = val; pivot -= i; }
    val = code[pivot];
  }
  if (upper_pivot == 255) upper = code[upper_pivot];
  if (lower_pivot == 0) lower = code[lower_pivot];
  if (x > val) {
    float midpoint = (upper + val) * 0.5f;
    return (unsigned char)(x > midpoint ? upper_pivot : pivot);
  } else {
    float midpoint = (lower + val) * 0.5f;
    return (unsigned char)(x < midpoint ? lower_pivot : pivot);
  }
}
ORC_API unsigned char orc_quantize_8bit_scalar(const float *code, float x) {
  return quantize_8bit(code, x);
}

/* ------------------------------------------------------------------------------------
 * a1: kQuantizeBlockwise, kernel_quant.cpp:1229-1365 (arithmetic), op_quant.cpp:431-655.
 * absmax = max|x| (fp32, init -FLT_MAX :1267), inv = 1.0f/absmax (:1304),
 * q = code(float(x) * inv); 4-bit packs (q(x[2j]) << 4) | q(x[2j+1]) (:1346-1348).
 * Out-of-range elements of a partial last block read as 0.0 (upstream BlockLoad default;
 * the port computes valid_items :1266 and then ignores it -- WIP artefact).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_quantize_blockwise(const float *code, const void *A, int dtype, float *absmax,
                                    unsigned char *out, int blocksize, long n, int qtype) {
  long nblocks = (n + blocksize - 1) / blocksize;
  for (long b = 0; b < nblocks; b++) {
    long start = b * (long)blocksize;
    long end = start + blocksize < n ? start + blocksize : n;
    float m = -FLT_MAX;
    for (long i = start; i < end; i++) m = fmaxf(m, fabsf(load_as_f32(A, dtype, i)));
    if (end - start < blocksize) m = fmaxf(m, 0.0f); /* zero-filled tail */
    absmax[b] = m;
    float inv = 1.0f / m;
    if (qtype == ORC_GENERAL8BIT) {
      for (long i = start; i < end; i++)
        out[i] = quantize_8bit(code, load_as_f32(A, dtype, i) * inv);
    } else {
      for (long i = start; i < end; i += 2) {
        float x0 = load_as_f32(A, dtype, i) * inv;
        float x1 = (i + 1 < n ? load_as_f32(A, dtype, i + 1) : 0.0f) * inv;
        unsigned char hi = qtype == ORC_NF4 ? quantize_nf4(x0) : quantize_fp4(x0);
        unsigned char lo = qtype == ORC_NF4 ? quantize_nf4(x1) : quantize_fp4(x1);
        out[i / 2] = (unsigned char)((hi << 4) | lo);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * a2: kDequantizeBlockwise, kernel_quant.cpp:1370-1471; op_quant.cpp:659-703.
 * out[i] = T(table[q_i] * absmax[i / blocksize]); fp32 multiply, one rounding to T.
 * 4-bit: n = number of OUTPUT elements, high nibble = even element (:1449).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_dequantize_blockwise(const float *code, const unsigned char *A,
                                      const float *absmax, void *out, int out_dtype,
                                      int blocksize, long n, int qtype) {
  for (long i = 0; i < n; i++) {
    float am = absmax[i / blocksize];
    float v;
    if (qtype == ORC_GENERAL8BIT) {
      v = code[A[i]] * am;
    } else {
      unsigned char byte = A[i / 2];
      unsigned char nib = (i & 1) ? (byte & 0x0F) : (byte >> 4);
      v = qtype == ORC_NF4 ? NF4_TABLE[nib] * am : dequantize_fp4_tree(nib, am);
    }
    store_from_f32(out, out_dtype, i, v);
  }
}

/* Nested absmax inverse: functional.py:1346-1350 / :1982-1984.
 * absmax = dequantize_blockwise(qabsmax, state2) ; absmax += offset  (separate fp32 mul, add) */
ORC_API void orc_denest_absmax(const float *code2, const unsigned char *qabsmax,
                               const float *absmax2, float offset, float *absmax_out,
                               int blocksize2, long nblocks) {
  for (long i = 0; i < nblocks; i++) {
    float v = code2[qabsmax[i]] * absmax2[i / blocksize2];
    absmax_out[i] = v + offset;
  }
}

/* ------------------------------------------------------------------------------------
 * a3: kgemm_4bit_inference_naive, kernel_gemm.cpp:1273-1388 (">= 800" T-arithmetic chain).
 * One warp per output row; lane l owns k in [l*32 + 1024*t, +32); quant_map and absmax are
 * rounded to T (:1294,:1305); B = qm *_T absmax (:1337-1338); p = A *_T B; local_C += float(p)
 * (:1374); warp sum (:1383) done here as an xor-butterfly.  K tail padded with nibble 7, A = 0.
 * mode 0: reference-faithful T-arithmetic chain; mode 1: fp32 products ("#else" branch
 * :1340-1342,:1377: B rounded to T, A*B in fp32).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_gemv_4bit(int M, int K, const void *A, const unsigned char *B,
                           const float *absmax, const float *datatype, void *out, int dtype,
                           int ldb, int blocksize, int mode) {
  float qm[16];
  for (int i = 0; i < 16; i++) qm[i] = round_to_T(datatype[i], dtype);
  for (int row = 0; row < M; row++) {
    float lane_acc[32];
    long offset_B = (long)ldb * row;
    for (int lane = 0; lane < 32; lane++) {
      float c = 0.0f;
      for (int inner = lane * 32; inner < K; inner += 32 * 32) {
        long absidx = (2 * offset_B + inner) / blocksize;
        float la = round_to_T(absmax[absidx], dtype);
        for (int j = 0; j < 32; j++) {
          int k = inner + j;
          unsigned char nib;
          float a;
          if (k < K) {
            unsigned char byte = (k / 2 < K / 2) ? B[offset_B + k / 2] : 0x77;
            nib = (k & 1) ? (byte & 0x0F) : (byte >> 4);
            a = load_as_f32(A, dtype, k);
          } else {
            nib = 7;
            a = 0.0f;
          }
          float b;
          if (mode == 0) {
            b = round_to_T(qm[nib] * la, dtype);
            c += round_to_T(a * b, dtype);
          } else {
            b = round_to_T(qm[nib] * la, dtype);
            c += a * b;
          }
        }
      }
      lane_acc[lane] = c;
    }
    for (int s = 16; s > 0; s >>= 1)
      for (int l = 0; l < 32; l++)
        if ((l & s) == 0) { /* pair (l, l^s): both end with the same sum */
          float t = lane_acc[l] + lane_acc[l ^ s];
          lane_acc[l] = t;
          lane_acc[l ^ s] = t;
        }
    store_from_f32(out, dtype, row, lane_acc[0]);
  }
}

/* fp64-exact GEMV/GEMM of the fp32-dequantized weight:
 * out[b, r] = sum_k double(A[b,k]) * double(fp32(code[nib] * absmax[(r*K+k)/bs])) */
ORC_API void orc_gemm_4bit_exact(int batch, int N, int K, const void *A, int dtype,
                                 const unsigned char *B, const float *absmax,
                                 const float *datatype, double *out, int blocksize) {
  float *w = (float *)malloc(sizeof(float) * (size_t)K);
  for (int r = 0; r < N; r++) {
    for (int k = 0; k < K; k++) {
      long e = (long)r * K + k;
      unsigned char byte = B[e / 2];
      unsigned char nib = (e & 1) ? (byte & 0x0F) : (byte >> 4);
      w[k] = datatype[nib] * absmax[e / blocksize];
    }
    for (int b = 0; b < batch; b++) {
      double acc = 0.0;
      for (int k = 0; k < K; k++) acc += (double)load_as_f32(A, dtype, (long)b * K + k) * (double)w[k];
      out[(long)b * N + r] = acc;
    }
  }
  free(w);
}

/* a4: MatMul4Bit.forward, autograd/_functions.py:490-518 = dequantize_4bit (W rounded to T)
 * then F.linear with fp32 (here fp64) accumulation, one rounding of the output to T. */
ORC_API void orc_gemm_4bit_dequant_ref(int batch, int N, int K, const void *A, int dtype,
                                       const unsigned char *B, const float *absmax,
                                       const float *datatype, void *out, int blocksize) {
  float *w = (float *)malloc(sizeof(float) * (size_t)K);
  for (int r = 0; r < N; r++) {
    for (int k = 0; k < K; k++) {
      long e = (long)r * K + k;
      unsigned char byte = B[e / 2];
      unsigned char nib = (e & 1) ? (byte & 0x0F) : (byte >> 4);
      w[k] = round_to_T(datatype[nib] * absmax[e / blocksize], dtype);
    }
    for (int b = 0; b < batch; b++) {
      double acc = 0.0;
      for (int k = 0; k < K; k++) acc += (double)load_as_f32(A, dtype, (long)b * K + k) * (double)w[k];
      store_from_f32(out, dtype, (long)b * N + r, (float)acc);
    }
  }
  free(w);
}

/* ------------------------------------------------------------------------------------
 * a5: kgetColRowStats, kernel_quant.cpp:3214-3379; launcher op_quant.cpp:1356-1427.
 * Tiles 16 rows x 256 cols; tile id = row_tile * col_tiles + col_tile.  |x| taken in fp16,
 * outliers (|x| >= thr, thr > 0) zeroed before the max and counted per (tile,row) into
 * nnz_count_row[tile*16 + row_in_tile + 1] (:3375-3377).  Stats merge with the caller's
 * initial values via atomicMax (:3358-3371) == max(existing, computed).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_get_col_row_stats(const uint16_t *A, float *rowStats, float *colStats,
                                   int *nnz_count_row, float thr, int rows, int cols) {
  int col_tiles = (cols + 255) / 256;
  int row_tiles = (rows + 15) / 16;
  if (thr > 0.0f && nnz_count_row)
    for (long t = 0; t < (long)row_tiles * col_tiles * 16; t++) nnz_count_row[t + 1] = 0;
  for (int r = 0; r < rows; r++) {
    for (int c = 0; c < cols; c++) {
      float v = fabsf(f16_bits_to_f32(A[(long)r * cols + c]));
      if (thr > 0.0f && v >= thr) {
        if (nnz_count_row) {
          long tile = (long)(r / 16) * col_tiles + c / 256;
          nnz_count_row[tile * 16 + (r % 16) + 1] += 1;
        }
        v = 0.0f;
      }
      if (rowStats[r] < v) rowStats[r] = v;
      if (colStats[c] < v) colStats[c] = v;
    }
  }
}

/* float -> int8 the way the device does it: rint (half-to-even) then a saturating convert;
 * NaN -> 0.  (The reference's `(char)rint(NaN or >127)` is implementation-defined) */
static inline signed char f32_to_s8_rn_sat(float v) {
  if (v != v) return 0;
  float r = rintf(v);
  if (r > 127.0f) return 127;
  if (r < -128.0f) return -128;
  return (signed char)r;
}

/* ------------------------------------------------------------------------------------
 * a6: kDoubleRowColQuant, kernel_quant.cpp:3384-3512; launcher op_quant.cpp:1430-1534.
 * out_row = (int8)rint(float(x) * (127.0f / rowStats[r])), out_col likewise with colStats[c]
 * (:3424,:3453,:3475,:3496).  thr > 0: |x| >= thr -> out_row = 0 and a COO entry appended at
 * nnz_block_ptr[tile*16 + row_in_tile] + (running count) (:3461-3472); out_col NOT zeroed.
 * The reference appends in atomic (arbitrary) order inside a (tile,row) segment; the oracle
 * emits ascending column order -- compare as a set, or exactly against a kernel that also
 * orders by column.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_double_rowcol_quant(const uint16_t *A, const float *rowStats,
                                     const float *colStats, signed char *out_col,
                                     signed char *out_row, int *rowidx, int *colidx,
                                     uint16_t *val, const int *nnz_block_ptr, float thr,
                                     int rows, int cols) {
  int col_tiles = (cols + 255) / 256;
  for (int r = 0; r < rows; r++) {
    float rs = 127.0f / rowStats[r];
    for (int ct = 0; ct < col_tiles; ct++) {
      long tile = (long)(r / 16) * col_tiles + ct;
      int cursor = (thr > 0.0f && nnz_block_ptr) ? nnz_block_ptr[tile * 16 + (r % 16)] : 0;
      int cend = (ct + 1) * 256 < cols ? (ct + 1) * 256 : cols;
      for (int c = ct * 256; c < cend; c++) {
        uint16_t h = A[(long)r * cols + c];
        float x = f16_bits_to_f32(h);
        float cs = 127.0f / colStats[c];
        if (thr > 0.0f && fabsf(x) >= thr) {
          out_row[(long)r * cols + c] = 0;
          if (rowidx) {
            rowidx[cursor] = r;
            colidx[cursor] = c;
            val[cursor] = h;
            cursor++;
          }
        } else {
          out_row[(long)r * cols + c] = f32_to_s8_rn_sat(x * rs);
        }
        out_col[(long)r * cols + c] = f32_to_s8_rn_sat(x * cs);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * a7: layouts.  kTransformRowToFormat, kernel_quant.cpp:3516-3841 (index math :3673-3675,
 * :3740-3755, :3822-3832); same maps as the vendored blas_utils.h:244-346.
 * layout_offset(fmt, R, r, c): linear offset of element (r, c) of an R-row matrix.
 * ---------------------------------------------------------------------------------- */
static inline long layout_out_rows(int fmt, int rows) {
  if (fmt == ORC_COL_TURING) return ((rows + 7) / 8) * 8L;
  if (fmt == ORC_COL_AMPERE) return ((rows + 31) / 32) * 32L;
  return rows;
}
static inline long layout_offset(int fmt, long out_rows, int r, int c) {
  int c32 = c % 32;
  switch (fmt) {
    case ORC_COL32:
      return (long)(c / 32) * 32 * out_rows + (long)r * 32 + c32;
    case ORC_COL_TURING: {
      long off = (long)(c / 32) * out_rows * 32 + (long)(r / 8) * 256;
      if (r % 2 == 1) off += 128 + (c32 / 4) * 16 + (c32 % 4) + ((r % 8) - 1) * 2;
      else            off += 0 + (c32 / 4) * 16 + (c32 % 4) + (r % 8) * 2;
      return off;
    }
    default: { /* ORC_COL_AMPERE */
      int lr = r % 32;
      int ar = ((lr % 8) / 2) * 8 + (lr / 8) * 2 + (lr % 2);
      return (long)(c / 32) * out_rows * 32 + (long)(r / 32) * 1024 + ar * 32 + c32;
    }
  }
}
ORC_API long orc_layout_offset(int fmt, int rows, int r, int c) {
  return layout_offset(fmt, layout_out_rows(fmt, rows), r, c);
}
/* the vendored dpct map (blas_utils.h:263-325), restated independently for cross-checking */
ORC_API long orc_layout_offset_blasutils(int fmt, long ld, int r, int c) {
  if (fmt == ORC_COL32) return ld * (c / 32) + 32L * r + c % 32;
  if (fmt == ORC_COL_TURING) {
    int fr = r % 8, fc = c % 32;
    int tr = 4 * (fr % 2) + fc / 8;
    int tc = 16 * ((fc / 4) % 2) + 4 * (fr / 2) + fc % 4;
    return ld * (c / 32) + (long)(r / 8) * 256 + tr * 32 + tc;
  }
  int fr = r % 32, fc = c % 32;
  int tr = 8 * ((fr % 8) / 2) + (fr / 8) * 2 + fr % 2;
  return ld * (c / 32) + (long)(r / 32) * 1024 + tr * 32 + fc;
}
ORC_API long orc_layout_size(int fmt, int rows, int cols) {
  return layout_out_rows(fmt, rows) * (((cols + 31) / 32) * 32L);
}
/* transform row-major int8 [rows, cols] -> fmt (of A, or of A^T when transpose != 0).
 * Only valid elements are written (the reference relies on a pre-zeroed buffer,
 * functional.py:482-518). elem_size 1 (int8) or 4 (int32). */
ORC_API void orc_transform_row2fmt(const void *A, void *out, int rows, int cols, int fmt,
                                   int transpose, int elem_size) {
  int R = transpose ? cols : rows;
  long out_rows = layout_out_rows(fmt, R);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      long dst = transpose ? layout_offset(fmt, out_rows, c, r) : layout_offset(fmt, out_rows, r, c);
      memcpy((char *)out + dst * elem_size, (const char *)A + ((long)r * cols + c) * elem_size,
             (size_t)elem_size);
    }
}
ORC_API void orc_transform_fmt2row(const void *A, void *out, int rows, int cols, int fmt,
                                   int elem_size) {
  long out_rows = layout_out_rows(fmt, rows);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++)
      memcpy((char *)out + ((long)r * cols + c) * elem_size,
             (const char *)A + layout_offset(fmt, out_rows, r, c) * elem_size, (size_t)elem_size);
}

/* ------------------------------------------------------------------------------------
 * a8: igemmlt<FORMATB,32,0>, op_gemm.cpp:541-603 -> blas_utils.h:459-724 (-> oneDNN s8*s8->s32,
 * third-party, not in tree).  C[i,j] = sum_k int32(A[i,k]) * int32(B[j,k]); A col32 (m rows),
 * B col_turing / col_ampere (n rows), C int32 col32 (m rows); alpha = 1, beta = 0.
 * Integer matmul has one exact answer; that is the oracle.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_igemmlt_32(int m, int n, int k, const signed char *A, const signed char *B,
                            int *C, int fmtB) {
  long brows = layout_out_rows(fmtB, n);
  /* un-permute into row-major scratch for speed */
  signed char *a = (signed char *)malloc((size_t)m * k);
  signed char *b = (signed char *)malloc((size_t)n * k);
  for (int i = 0; i < m; i++)
    for (int kk = 0; kk < k; kk++) a[(long)i * k + kk] = A[layout_offset(ORC_COL32, m, i, kk)];
  for (int j = 0; j < n; j++)
    for (int kk = 0; kk < k; kk++) b[(long)j * k + kk] = B[layout_offset(fmtB, brows, j, kk)];
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      int acc = 0;
      const signed char *pa = a + (long)i * k, *pb = b + (long)j * k;
      for (int kk = 0; kk < k; kk++) acc += (int)pa[kk] * (int)pb[kk];
      C[layout_offset(ORC_COL32, m, i, j)] = acc;
    }
  free(a);
  free(b);
}
/* plain row-major exact int8 GEMM: C[i,j] = sum_k A[i,k]*B[j,k] */
ORC_API void orc_igemm_rowmajor(int m, int n, int k, const signed char *A, const signed char *B,
                                int *C) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      int acc = 0;
      const signed char *pa = A + (long)i * k, *pb = B + (long)j * k;
      for (int kk = 0; kk < k; kk++) acc += (int)pa[kk] * (int)pb[kk];
      C[(long)i * n + j] = acc;
    }
}

/* ------------------------------------------------------------------------------------
 * a9: kdequant_mm_int32_fp16, kernel_quant.cpp:3848-3987 (formula :3969, const :3846).
 * out[r,c] = half(((float(C[r,c]) * 6.200012e-05f) * rowStats[r]) * colStats[c] + float(bias[c]))
 * left-to-right fp32, no fma contraction; C read from col32; out row-major.
 * a_is_col32 = 0 lets tests feed a row-major int32 matrix (B200-native fused epilogue).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_dequant_mm_int32_fp16(const int *A, const float *rowStats, const float *colStats,
                                       uint16_t *out, const uint16_t *bias, int numRows,
                                       int numCols, int a_is_col32) {
  const float MM_DEQUANT_CONST = 6.200012e-05f;
  for (int r = 0; r < numRows; r++)
    for (int c = 0; c < numCols; c++) {
      int v = a_is_col32 ? A[layout_offset(ORC_COL32, numRows, r, c)] : A[(long)r * numCols + c];
      float b = bias ? f16_bits_to_f32(bias[c]) : 0.0f;
      float t = (float)v * MM_DEQUANT_CONST;
      t = t * rowStats[r];
      t = t * colStats[c];
      t = t + b;
      out[(long)r * numCols + c] = f32_to_f16_bits(t);
    }
}

/* ------------------------------------------------------------------------------------
 * a10: kExtractOutliers, kernel_quant.cpp:3992-4053.
 * out[row, j] = A_fmt(row, idx[j]) for an int8 weight in col_turing / col_ampere layout.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_extract_outliers(const signed char *A, const int *idx, signed char *out,
                                  int idx_size, int rows, int cols, int fmt) {
  (void)cols;
  long out_rows = layout_out_rows(fmt, rows);
  for (int r = 0; r < rows; r++)
    for (int j = 0; j < idx_size; j++)
      out[(long)r * idx_size + j] = A[layout_offset(fmt, out_rows, r, idx[j])];
}

/* ------------------------------------------------------------------------------------
 * CPU 8-bit blockwise path restated: sycl/cpu_ops.cpp:7-63 + sycl/common.cpp:4-35.
 * (The real thing is compiled into oracle/_ref/libref_cpu.so; this restatement exists so the
 * algorithm can run where /root/reference does not exist and to cross-check the port.)
 * quantize_block: absmax, x / absmax (true divide), lower-bound search on the strictly
 * increasing 256-entry code, then pick the nearer neighbour (ties -> left).
 * quantize_cpu overwrites code[0] = -1.0f in the caller's buffer (cpu_ops.cpp:20).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_quantize_cpu(float *code, const float *A, float *absmax, unsigned char *out,
                              long long blocksize, long long n) {
  code[0] = -1.0f;
  for (long long start = 0; start < n; start += blocksize) {
    long long end = start + blocksize < n ? start + blocksize : n;
    float m = -FLT_MAX;
    for (long long i = start; i < end; i++) m = fmaxf(m, fabsf(A[i]));
    absmax[start / blocksize] = m;
    for (long long i = start; i < end; i++) {
      float v = A[i] / m;
      /* BinSearch scalar(): index of the last code entry <= v, clamped to [0, 255] */
      int lo = 0, hi = 255;
      while (lo < hi) {
        int mid = (lo + hi + 1) / 2;
        if (code[mid] <= v) lo = mid; else hi = mid - 1;
      }
      int idx = lo;
      if (idx < 255) {
        float dl = fabsf(v - code[idx]);
        float dr = fabsf(v - code[idx + 1]);
        if (dr < dl) idx += 1;
      }
      out[i] = (unsigned char)idx;
    }
  }
}
ORC_API void orc_dequantize_cpu(const float *code, const unsigned char *A, const float *absmax,
                                float *out, long long blocksize, long long n) {
  for (long long i = 0; i < n; i++) out[i] = code[A[i]] * absmax[i / blocksize];
}
