// c_api.cu -- the extern "C" boundary of libbitsandbytes_b200.so (declared in include/bnb_b200.h).
// Thin wrappers, exactly like the reference's sycl/pythonInterface.cpp:192-422: typed template
// instantiations behind C names; no logic lives here besides the per-thread stream / error latch.
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <set>
#include <utility>

#include "../../include/bnb_b200.h"
#include "common.cuh"

namespace bnb {

// ---- per-thread state ----------------------------------------------------------------------------
static thread_local cudaStream_t tl_stream = nullptr;  // nullptr == legacy default stream (== torch default)
static thread_local cudaError_t tl_error = cudaSuccess;
static thread_local char tl_error_msg[256] = {0};

cudaStream_t current_stream() { return tl_stream; }
void set_current_stream(cudaStream_t s) { tl_stream = s; }
void latch_error(cudaError_t e, const char *where) {
  if (e == cudaSuccess || tl_error != cudaSuccess) return;
  tl_error = e;
  snprintf(tl_error_msg, sizeof(tl_error_msg), "%s: %s", where, cudaGetErrorString(e));
  if (getenv("BNB_B200_VERBOSE")) fprintf(stderr, "bnb_b200 error: %s\n", tl_error_msg);
}

void ensure_max_dynamic_smem(const void *kernel, int bytes, const char *where) {
  static std::mutex mu;
  static std::set<std::pair<int, const void *>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, kernel})) return;
  latch_error(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), where);
  done.insert({dev, kernel});
}

// ---- typed entry points implemented in the kernel TUs -------------------------------------------
template <typename T, int QT> void quantize_blockwise(const float *, const T *, float *, unsigned char *, int, long);
template <typename T, int QT> void dequantize_blockwise(const float *, const unsigned char *, const float *, T *, int, long);
long long selftest_quant_lut(int qtype);
void gemv_probe(unsigned long long *out2);
void gemv_trace(unsigned long long *out);
void set_gemv_host_tables(const float *code16, const float *code2_256);
template <typename T> int gemv_4bit_nested_multi(int, const int *, int, const T *, const unsigned char *const *, const unsigned char *const *, const float *const *, const float *, const float *, const float *, T *const *, int, int, void *const *, int);
template <typename T> void gemv_4bit(int, int, int, const T *, const unsigned char *, const float *, const float *, T *, int, int, int, int);
struct GemvSync;
void epoch_bump(unsigned int *epoch);
void peer_barrier(unsigned int *counter, const unsigned int *sig_local, unsigned int *const *sig_peer, int npeers);
template <typename T> void gemv_4bit_nested(int, int, int, const T *, const unsigned char *, const unsigned char *, const float *, const float *, float, const float *, T *, int, int, int, int, int, void *const *, int, const GemvSync *);
template <typename T> int gemm_4bit(int, int, int, const T *, const unsigned char *, const float *, const float *, const T *, T *, int, long, void *const *, int);
void get_col_row_stats(const __half *, float *, float *, int *, float, int, int);
void double_rowcol_quant(const __half *, const float *, const float *, signed char *, signed char *, int *, int *, __half *, const int *, float, int, int);
template <int FMT> void transform_row2fmt(const signed char *, signed char *, int, int, bool);
void untransform_s8(int fmt, const signed char *A, signed char *out, int rows, int cols);
template <int FMT> void extract_outliers(const signed char *, const int *, signed char *, int, int, int);
void dequant_mm_int32_fp16(const int *, const float *, const float *, __half *, const __half *, int, int);
int igemmlt(int, int, bool, int, int, int, const signed char *, const signed char *, void *, const float *, int, int, int);
int igemm_rowmajor_32(int, int, int, const signed char *, const signed char *, int *);
int igemm_rowmajor_dequant_fp16(int, int, int, const signed char *, const signed char *, const float *, const float *, const __half *, __half *);
int int8_linear_fused(const __half *, const signed char *, const float *, const __half *, __half *, float, int, int, int, signed char *, float *, unsigned char *, short *, int *, int, int *, __half *, __half *);

}  // namespace bnb

using namespace bnb;
typedef __half half_t;
typedef __nv_bfloat16 bf16_t;

extern "C" {

// ---------------------------------------------------------------- additive: stream / errors / version
void cbnb_set_stream(void *stream) { set_current_stream(reinterpret_cast<cudaStream_t>(stream)); }
void *cbnb_get_stream(void) { return reinterpret_cast<void *>(current_stream()); }
int cbnb_last_error(void) {
  int e = (int)tl_error;
  tl_error = cudaSuccess;
  return e;
}
const char *cbnb_last_error_string(void) { return tl_error_msg; }
const char *cbnb_version(void) { return "bnb_b200 sm_100a r1"; }
long long cbnb_selftest_quant_lut(int qtype) { return selftest_quant_lut(qtype); }
void cbnb_debug_gemv_probe(unsigned long long *cycles_ns) { gemv_probe(cycles_ns); }
void cbnb_debug_gemv_trace(unsigned long long *out_2x320x8) { gemv_trace(out_2x320x8); }
int cgemm_4bit_inference_nested_multi_fp16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2) {
  return gemv_4bit_nested_multi<half_t>(count, m, k, (half_t *)A, B, qabsmax, absmax2, code2, offsets, datatype, (half_t *const *)outs, blocksize, blocksize2, nullptr, 0); }
int cgemm_4bit_inference_nested_multi_bf16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2) {
  return gemv_4bit_nested_multi<bf16_t>(count, m, k, (bf16_t *)A, B, qabsmax, absmax2, code2, offsets, datatype, (bf16_t *const *)outs, blocksize, blocksize2, nullptr, 0); }
int cgemm_4bit_inference_nested_multi_push_fp16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2, void **peer_outs, int npeers) {
  return gemv_4bit_nested_multi<half_t>(count, m, k, (half_t *)A, B, qabsmax, absmax2, code2, offsets, datatype, (half_t *const *)outs, blocksize, blocksize2, peer_outs, npeers); }
int cgemm_4bit_inference_nested_multi_push_bf16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2, void **peer_outs, int npeers) {
  return gemv_4bit_nested_multi<bf16_t>(count, m, k, (bf16_t *)A, B, qabsmax, absmax2, code2, offsets, datatype, (bf16_t *const *)outs, blocksize, blocksize2, peer_outs, npeers); }
void cbnb_set_gemv_host_tables(const float *code16_host, const float *code2_256_host) { set_gemv_host_tables(code16_host, code2_256_host); }

// ---------------------------------------------------------------- blockwise quantize (pythonInterface.cpp:203-217)
#define QUANT_FN(name, T, QT) \
  void name(float *code, T *A, float *absmax, unsigned char *out, int blocksize, const int n) { \
    quantize_blockwise<T, QT>(code, A, absmax, out, blocksize, (long)n); }
void cquantize_blockwise_fp32(float *code, float *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<float, General8bit>(code, A, absmax, out, blocksize, n); }
void cquantize_blockwise_fp32_fp4(float *code, float *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<float, FP4>(code, A, absmax, out, blocksize, n); }
void cquantize_blockwise_fp32_nf4(float *code, float *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<float, NF4>(code, A, absmax, out, blocksize, n); }
void cquantize_blockwise_fp16(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<half_t, General8bit>(code, (half_t *)A, absmax, out, blocksize, n); }
void cquantize_blockwise_fp16_fp4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<half_t, FP4>(code, (half_t *)A, absmax, out, blocksize, n); }
void cquantize_blockwise_fp16_nf4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<half_t, NF4>(code, (half_t *)A, absmax, out, blocksize, n); }
void cquantize_blockwise_bf16(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<bf16_t, General8bit>(code, (bf16_t *)A, absmax, out, blocksize, n); }
void cquantize_blockwise_bf16_fp4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<bf16_t, FP4>(code, (bf16_t *)A, absmax, out, blocksize, n); }
void cquantize_blockwise_bf16_nf4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n) { quantize_blockwise<bf16_t, NF4>(code, (bf16_t *)A, absmax, out, blocksize, n); }

// ---------------------------------------------------------------- blockwise dequantize (pythonInterface.cpp:199-221)
void cdequantize_blockwise_fp32(float *code, unsigned char *A, float *absmax, float *out, int blocksize, const int n) { dequantize_blockwise<float, General8bit>(code, A, absmax, out, blocksize, n); }
void cdequantize_blockwise_fp32_fp4(float *code, unsigned char *A, float *absmax, float *out, int blocksize, const int n) { dequantize_blockwise<float, FP4>(code, A, absmax, out, blocksize, n); }
void cdequantize_blockwise_fp32_nf4(float *code, unsigned char *A, float *absmax, float *out, int blocksize, const int n) { dequantize_blockwise<float, NF4>(code, A, absmax, out, blocksize, n); }
void cdequantize_blockwise_fp16(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n) { dequantize_blockwise<half_t, General8bit>(code, A, absmax, (half_t *)out, blocksize, n); }
void cdequantize_blockwise_fp16_fp4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n) { dequantize_blockwise<half_t, FP4>(code, A, absmax, (half_t *)out, blocksize, n); }
void cdequantize_blockwise_fp16_nf4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n) { dequantize_blockwise<half_t, NF4>(code, A, absmax, (half_t *)out, blocksize, n); }
void cdequantize_blockwise_bf16(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n) { dequantize_blockwise<bf16_t, General8bit>(code, A, absmax, (bf16_t *)out, blocksize, n); }
void cdequantize_blockwise_bf16_fp4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n) { dequantize_blockwise<bf16_t, FP4>(code, A, absmax, (bf16_t *)out, blocksize, n); }
void cdequantize_blockwise_bf16_nf4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n) { dequantize_blockwise<bf16_t, NF4>(code, A, absmax, (bf16_t *)out, blocksize, n); }

// ---------------------------------------------------------------- 4-bit GEMV (pythonInterface.cpp:408-415)
void cgemm_4bit_inference_naive_fp16(int m, int n, int k, void *A, unsigned char *B, float *absmax, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize) {
  gemv_4bit<half_t>(m, n, k, (half_t *)A, B, absmax, datatype, (half_t *)out, lda, ldb, ldc, blocksize); }
void cgemm_4bit_inference_naive_bf16(int m, int n, int k, void *A, unsigned char *B, float *absmax, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize) {
  gemv_4bit<bf16_t>(m, n, k, (bf16_t *)A, B, absmax, datatype, (bf16_t *)out, lda, ldb, ldc, blocksize); }
void cgemm_4bit_inference_naive_fp32(int m, int n, int k, float *A, unsigned char *B, float *absmax, float *datatype, float *out, int lda, int ldb, int ldc, int blocksize) {
  gemv_4bit<float>(m, n, k, A, B, absmax, datatype, out, lda, ldb, ldc, blocksize); }
void cgemm_4bit_inference_nested_fp16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2) {
  gemv_4bit_nested<half_t>(m, n, k, (half_t *)A, B, qabsmax, absmax2, code2, offset, datatype, (half_t *)out, lda, ldb, ldc, blocksize, blocksize2, nullptr, 0, nullptr); }
void cgemm_4bit_inference_nested_bf16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2) {
  gemv_4bit_nested<bf16_t>(m, n, k, (bf16_t *)A, B, qabsmax, absmax2, code2, offset, datatype, (bf16_t *)out, lda, ldb, ldc, blocksize, blocksize2, nullptr, 0, nullptr); }

void cgemm_4bit_inference_nested_push_fp16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2, void **peer_outs, int npeers, const bnb_gemv_sync_t *sync) {
  gemv_4bit_nested<half_t>(m, n, k, (half_t *)A, B, qabsmax, absmax2, code2, offset, datatype, (half_t *)out, lda, ldb, ldc, blocksize, blocksize2, peer_outs, npeers, reinterpret_cast<const GemvSync *>(sync)); }
void cgemm_4bit_inference_nested_push_bf16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2, void **peer_outs, int npeers, const bnb_gemv_sync_t *sync) {
  gemv_4bit_nested<bf16_t>(m, n, k, (bf16_t *)A, B, qabsmax, absmax2, code2, offset, datatype, (bf16_t *)out, lda, ldb, ldc, blocksize, blocksize2, peer_outs, npeers, reinterpret_cast<const GemvSync *>(sync)); }
void cbnb_epoch_bump(unsigned int *epoch) { epoch_bump(epoch); }
void cbnb_peer_barrier(unsigned int *counter, const unsigned int *sig_local, unsigned int **sig_peer, int npeers) { peer_barrier(counter, sig_local, sig_peer, npeers); }

// ---------------------------------------------------------------- fused 4-bit GEMM (additive)
int cgemm_4bit_fp16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize) {
  return gemm_4bit<half_t>(batch, N, K, (half_t *)A, B, absmax, datatype, (half_t *)bias, (half_t *)out, blocksize, 0, nullptr, 0); }
int cgemm_4bit_bf16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize) {
  return gemm_4bit<bf16_t>(batch, N, K, (bf16_t *)A, B, absmax, datatype, (bf16_t *)bias, (bf16_t *)out, blocksize, 0, nullptr, 0); }
// N-sharded form: `out` is the base of this rank's column slice inside the gathered [batch, ldo] buffer, peer_outs the same
// address in every peer's copy (NVLink peer mappings); the epilogue stores into all of them
int cgemm_4bit_push_fp16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize, long ldo, void **peer_outs, int npeers) {
  return gemm_4bit<half_t>(batch, N, K, (half_t *)A, B, absmax, datatype, (half_t *)bias, (half_t *)out, blocksize, ldo, peer_outs, npeers); }
int cgemm_4bit_push_bf16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize, long ldo, void **peer_outs, int npeers) {
  return gemm_4bit<bf16_t>(batch, N, K, (bf16_t *)A, B, absmax, datatype, (bf16_t *)bias, (bf16_t *)out, blocksize, ldo, peer_outs, npeers); }

// ---------------------------------------------------------------- LLM.int8 (pythonInterface.cpp:333-369)
void cget_col_row_stats(void *A, float *rowStats, float *colStats, int *nnz_count_row, float nnz_threshold, int rows, int cols) {
  get_col_row_stats((half_t *)A, rowStats, colStats, nnz_count_row, nnz_threshold, rows, cols); }
void cdouble_rowcol_quant(void *A, float *rowStats, float *colStats, char *out_col_normed, char *out_row_normed, int *rowidx, int *colidx, void *val, int *nnz_row_ptr, float threshold, int rows, int cols) {
  double_rowcol_quant((half_t *)A, rowStats, colStats, (signed char *)out_col_normed, (signed char *)out_row_normed, rowidx, colidx, (half_t *)val, nnz_row_ptr, threshold, rows, cols); }
void ctransform_row2col32(char *A, char *out, int rows, int cols) { transform_row2fmt<COL32>((signed char *)A, (signed char *)out, rows, cols, false); }
void ctransform_row2col32T(char *A, char *out, int rows, int cols) { transform_row2fmt<COL32>((signed char *)A, (signed char *)out, rows, cols, true); }
void ctransform_row2turing(char *A, char *out, int rows, int cols) { transform_row2fmt<COL_TURING>((signed char *)A, (signed char *)out, rows, cols, false); }
void ctransform_row2turingT(char *A, char *out, int rows, int cols) { transform_row2fmt<COL_TURING>((signed char *)A, (signed char *)out, rows, cols, true); }
void ctransform_row2ampere(char *A, char *out, int rows, int cols) { transform_row2fmt<COL_AMPERE>((signed char *)A, (signed char *)out, rows, cols, false); }
void ctransform_row2ampereT(char *A, char *out, int rows, int cols) { transform_row2fmt<COL_AMPERE>((signed char *)A, (signed char *)out, rows, cols, true); }
// inverse layouts: the reference's Python calls them (functional.py:2645-2647, maybe_rearrange_weight at checkpoint load)
// although its own library never exported them (SURVEY 8a "WIP artefacts"); same argument convention as row2*
void ctransform_turing2row(char *A, char *out, int rows, int cols) { untransform_s8(COL_TURING, (signed char *)A, (signed char *)out, rows, cols); }
void ctransform_ampere2row(char *A, char *out, int rows, int cols) { untransform_s8(COL_AMPERE, (signed char *)A, (signed char *)out, rows, cols); }
void ctransform_col322row(char *A, char *out, int rows, int cols) { untransform_s8(COL32, (signed char *)A, (signed char *)out, rows, cols); }

int cigemmlt_turing_32(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc) { return igemmlt(COL_TURING, 32, false, m, n, k, A, B, C, row_scale, lda, ldb, ldc); }
int cigemmlt_turing_8(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc) { return igemmlt(COL_TURING, 8, false, m, n, k, A, B, C, row_scale, lda, ldb, ldc); }
int cigemmlt_turing_8_rowscale(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc) { return igemmlt(COL_TURING, 8, true, m, n, k, A, B, C, row_scale, lda, ldb, ldc); }
int cigemmlt_ampere_32(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc) { return igemmlt(COL_AMPERE, 32, false, m, n, k, A, B, C, row_scale, lda, ldb, ldc); }
int cigemmlt_ampere_8(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc) { return igemmlt(COL_AMPERE, 8, false, m, n, k, A, B, C, row_scale, lda, ldb, ldc); }
int cigemmlt_ampere_8_rowscale(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc) { return igemmlt(COL_AMPERE, 8, true, m, n, k, A, B, C, row_scale, lda, ldb, ldc); }

void cdequant_mm_int32_fp16(int *A, float *rowStats, float *colStats, void *out, float *newRowStats, float *newcolStats, void *bias, int numRows, int numCols) {
  (void)newRowStats; (void)newcolStats;  // accepted, never written -- as in the reference kernel
  dequant_mm_int32_fp16(A, rowStats, colStats, (half_t *)out, (half_t *)bias, numRows, numCols); }
void cextractOutliers_turing(char *A, int *idx, char *out, int idx_size, int rows, int cols) { extract_outliers<COL_TURING>((signed char *)A, idx, (signed char *)out, idx_size, rows, cols); }
void cextractOutliers_ampere(char *A, int *idx, char *out, int idx_size, int rows, int cols) { extract_outliers<COL_AMPERE>((signed char *)A, idx, (signed char *)out, idx_size, rows, cols); }

int cigemm_rowmajor_32(int m, int n, int k, const int8_t *A, const int8_t *B, int *C) { return igemm_rowmajor_32(m, n, k, A, B, C); }
int cigemm_rowmajor_dequant_fp16(int m, int n, int k, const int8_t *A, const int8_t *B, float *rowStats, float *colStats, void *bias, void *out) {
  return igemm_rowmajor_dequant_fp16(m, n, k, A, B, rowStats, colStats, (half_t *)bias, (half_t *)out); }

int cint8_linear_fp16(void *A, const int8_t *CB, float *SCB, void *bias, void *out, float threshold, int m, int n, int k,
                      int8_t *CA, float *SCA, unsigned char *colflag, short *pos, int *idx, int idx_cap, int *count, void *subA, void *subB) {
  return int8_linear_fused((half_t *)A, CB, SCB, (half_t *)bias, (half_t *)out, threshold, m, n, k, CA, SCA, colflag, pos, idx,
                           idx_cap, count, (half_t *)subA, (half_t *)subB); }

// ---------------------------------------------------------------- context (pythonInterface.cpp:295)
struct Context { int device; };
void *get_context(void) {
  Context *c = new Context();  // leaked on purpose, exactly like the reference
  cudaGetDevice(&c->device);
  return c;
}
// The reference loader sets .restype on these two unconditionally (python_src_quants/cextension.py:82-84), so an
// UNMODIFIED loader needs the symbols to exist.  Neither is on the hot path (sparse handle: spmm; managed memory:
// paged optimizers -- SURVEY 8b "out of scope"): both return NULL.  (The reference's own cget_managed_ptr returns an
// uninitialised pointer, pythonInterface.cpp:380-387.)
void *get_cusparse(void) { return nullptr; }
void *cget_managed_ptr(size_t bytes) { (void)bytes; return nullptr; }

}  // extern "C"
