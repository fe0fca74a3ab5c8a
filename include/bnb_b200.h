/*
 * bnb_b200.h -- C-ABI of libbitsandbytes_b200.so, the B200 (sm_100a) drop-in for the quantized-linear
 * hot path of abhilash1910/bitsandbytes-SYCL.
 *
 * Every entry point in section 1 has the EXACT name, argument order and argument meaning of the
 * symbol the reference exports from sycl/pythonInterface.cpp (file:line cited per group) and that
 * python_src_quants/functional.py binds through ctypes (python_src_quants/cextension.py:88-110).
 * Conventions, kept as in the reference:
 *   - plain C, raw DEVICE pointers, 32-bit ints / floats by value, NULL allowed where the reference
 *     passes None; every buffer is allocated and owned by the caller; the library never frees or
 *     allocates user-visible memory;
 *   - launches are asynchronous on the caller's current device; the ABI has no stream slot, so the
 *     launch stream is a thread-local set with cbnb_set_stream() (default: the legacy default
 *     stream 0 == PyTorch's default stream, so an un-patched caller stays correct);
 *   - `void` functions report nothing (errors are latched, see cbnb_last_error()); cigemmlt_*
 *     return int: 0 ok, 1 -> the reference's Python turns it into NotImplementedError,
 *     anything else -> "cublasLt ran into an error!" (functional.py:2341-2348).
 * `half` / `bf16` pointers are 16-bit IEEE binary16 / bfloat16 device arrays (typed `void*` here so
 * the header is plain C).
 *
 * Section 2 lists the ADDITIVE symbols (same convention) that exist only in this library.
 */
#ifndef BNB_B200_H
#define BNB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNB_B200_API __attribute__((visibility("default")))

/* ============================== 1. reference ABI (hot-path subset) ============================== */

/* --- blockwise quantize: sycl/pythonInterface.cpp:203-217 (kQuantizeBlockwise, kernel_quant.cpp:1229)
 * code: 256-entry fp32 map (General8bit) or NULL (fp4/nf4); A: n elements of T; absmax: ceil(n/blocksize)
 * fp32; out: n bytes (8-bit) or (n+1)/2 bytes (4-bit, even element in the high nibble);
 * blocksize in {64,128,256,512,1024,2048,4096}. */
BNB_B200_API void cquantize_blockwise_fp32(float *code, float *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_fp32_fp4(float *code, float *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_fp32_nf4(float *code, float *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_fp16(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_fp16_fp4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_fp16_nf4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_bf16(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_bf16_fp4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n);
BNB_B200_API void cquantize_blockwise_bf16_nf4(float *code, void *A, float *absmax, unsigned char *out, int blocksize, const int n);

/* --- blockwise dequantize: sycl/pythonInterface.cpp:199-221 (kDequantizeBlockwise, kernel_quant.cpp:1370)
 * n = number of OUTPUT elements (4-bit: A holds (n+1)/2 bytes). */
BNB_B200_API void cdequantize_blockwise_fp32(float *code, unsigned char *A, float *absmax, float *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_fp32_fp4(float *code, unsigned char *A, float *absmax, float *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_fp32_nf4(float *code, unsigned char *A, float *absmax, float *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_fp16(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_fp16_fp4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_fp16_nf4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_bf16(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_bf16_fp4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n);
BNB_B200_API void cdequantize_blockwise_bf16_nf4(float *code, unsigned char *A, float *absmax, void *out, int blocksize, const int n);

/* --- batch-1 4-bit GEMV: sycl/pythonInterface.cpp:408-415 (kgemm_4bit_inference_naive, kernel_gemm.cpp:1273)
 * m = output features N, n = 1, k = K; A: [1,K] T; B: packed [N, ldb] bytes (ldb = (K+1)/2);
 * absmax: fp32 [N*K/blocksize] (already de-nested); datatype: fp32[16] code; out: [1,N] T. */
BNB_B200_API void cgemm_4bit_inference_naive_fp16(int m, int n, int k, void *A, unsigned char *B, float *absmax, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize);
BNB_B200_API void cgemm_4bit_inference_naive_bf16(int m, int n, int k, void *A, unsigned char *B, float *absmax, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize);
BNB_B200_API void cgemm_4bit_inference_naive_fp32(int m, int n, int k, float *A, unsigned char *B, float *absmax, float *datatype, float *out, int lda, int ldb, int ldc, int blocksize);

/* --- LLM.int8 statistics + double quant: sycl/pythonInterface.cpp:335-339
 * (kgetColRowStats kernel_quant.cpp:3214, kDoubleRowColQuant :3384). A: fp16 [rows, cols] row-major. */
BNB_B200_API void cget_col_row_stats(void *A, float *rowStats, float *colStats, int *nnz_count_row, float nnz_threshold, int rows, int cols);
BNB_B200_API void cdouble_rowcol_quant(void *A, float *rowStats, float *colStats, char *out_col_normed, char *out_row_normed, int *rowidx, int *colidx, void *val, int *nnz_row_ptr, float threshold, int rows, int cols);

/* --- int8 layout transforms: sycl/pythonInterface.cpp:341-357 (kTransformRowToFormat, kernel_quant.cpp:3516).
 * Only valid elements are written; padding comes from the caller's zeroed buffer (functional.py:482-518). */
BNB_B200_API void ctransform_row2col32(char *A, char *out, int rows, int cols);
BNB_B200_API void ctransform_row2col32T(char *A, char *out, int rows, int cols);
BNB_B200_API void ctransform_row2turing(char *A, char *out, int rows, int cols);
BNB_B200_API void ctransform_row2turingT(char *A, char *out, int rows, int cols);
BNB_B200_API void ctransform_row2ampere(char *A, char *out, int rows, int cols);
BNB_B200_API void ctransform_row2ampereT(char *A, char *out, int rows, int cols);
/* inverse layouts -> row-major.  python_src_quants/functional.py:2645-2647 calls ctransform_turing2row / ctransform_ampere2row
 * (checkpoint load: nn/modules.py:635-654 maybe_rearrange_weight); the reference library itself never exported them.
 * out is [rows, cols] row-major int8; A holds the padded layout of a [rows, cols] matrix (blas_utils.h:244-346). */
BNB_B200_API void ctransform_turing2row(char *A, char *out, int rows, int cols);
BNB_B200_API void ctransform_ampere2row(char *A, char *out, int rows, int cols);
BNB_B200_API void ctransform_col322row(char *A, char *out, int rows, int cols);

/* --- int8 GEMM C = A * B^T: sycl/pythonInterface.cpp:298-316 (igemmlt, op_gemm.cpp:541-655).
 * A: int8 col32 [m,k] (lda = m*32); B: int8 col_turing / col_ampere [n,k]; C: col32 [m,n] (ldc = m*32),
 * int32 for *_32, saturated int8 for *_8 (alpha = 1.0f) and *_8_rowscale (alpha = row_scale[i]). */
BNB_B200_API int cigemmlt_turing_32(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc);
BNB_B200_API int cigemmlt_turing_8(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc);
BNB_B200_API int cigemmlt_turing_8_rowscale(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc);
BNB_B200_API int cigemmlt_ampere_32(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc);
BNB_B200_API int cigemmlt_ampere_8(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc);
BNB_B200_API int cigemmlt_ampere_8_rowscale(int m, int n, int k, const int8_t *A, const int8_t *B, void *C, float *row_scale, int lda, int ldb, int ldc);

/* --- int32 -> fp16 dequant: sycl/pythonInterface.cpp:333 (kdequant_mm_int32_fp16, kernel_quant.cpp:3848).
 * A: int32 col32 [numRows, numCols]; out: fp16 row-major; newRowStats/newcolStats accepted, never written. */
BNB_B200_API void cdequant_mm_int32_fp16(int *A, float *rowStats, float *colStats, void *out, float *newRowStats, float *newcolStats, void *bias, int numRows, int numCols);

/* --- outlier column gather: sycl/pythonInterface.cpp:368-369 (kExtractOutliers, kernel_quant.cpp:3992) */
BNB_B200_API void cextractOutliers_turing(char *A, int *idx, char *out, int idx_size, int rows, int cols);
BNB_B200_API void cextractOutliers_ampere(char *A, int *idx, char *out, int idx_size, int rows, int cols);

/* --- context: sycl/pythonInterface.cpp:295; presence of this symbol marks the library GPU-capable
 * (python_src_quants/cextension.py:103). Returns a leaked opaque handle, as the reference does. */
BNB_B200_API void *get_context(void);
/* touched by the reference loader only (python_src_quants/cextension.py:82-84 sets their restype unconditionally):
 * sycl/pythonInterface.cpp:296 and :380-387.  Off the hot path (spmm handle / managed memory): both return NULL. */
BNB_B200_API void *get_cusparse(void);
BNB_B200_API void *cget_managed_ptr(size_t bytes);
/* NOT exported on purpose: cquantize_blockwise_cpu_fp32 / cdequantize_blockwise_cpu_fp32 (sycl/pythonInterface.cpp:419-420,
 * SURVEY 8b).  They are the reference's CPU fallback; this library has no CPU path (the Python layer raises for CPU
 * tensors), and the reference's CPU implementation lives on as test infrastructure only (oracle/_ref). */

/* ============================== 2. additive symbols (this library only) ============================== */

/* launch stream for the calling thread (cudaStream_t as void*); NULL = legacy default stream */
BNB_B200_API void cbnb_set_stream(void *stream);
BNB_B200_API void *cbnb_get_stream(void);
/* last CUDA error latched by any entry point on this thread (0 = cudaSuccess); clears the latch */
BNB_B200_API int cbnb_last_error(void);
BNB_B200_API const char *cbnb_last_error_string(void);
/* library identification: "bnb_b200 sm_100a <build tag>" */
BNB_B200_API const char *cbnb_version(void);
/* device self-test: number of fp32 bit patterns (all 2^32 are tried) for which the shared-memory-LUT
 * 4-bit quantiser disagrees with the reference decision tree; qtype 1 = FP4, 2 = NF4. Must return 0. */
BNB_B200_API long long cbnb_selftest_quant_lut(int qtype);

/* debug: out[12] = {SM cycles, nanoseconds} the middle CTA of the last block-column GEMV spent, then the nanoseconds
 * since its entry at which it (2) finished its tables, (3) saw the previous kernel complete, (4) had x in shared
 * memory, (5) finished warp 0's items, (6) finished every warp's items; (7) first weight loads issued, (8) code2 stored, (9) LUT stored; out[10..11] unused.  Recorded only when the
 * environment has BNB_B200_GEMV_PROBE=1. */
BNB_B200_API void cbnb_debug_gemv_probe(unsigned long long *cycles_ns);
BNB_B200_API void cbnb_debug_gemv_trace(unsigned long long *out_2x320x8);

/* GEMV with the NESTED (double-quantised) absmax consumed directly: qabsmax uint8 [N*K/blocksize],
 * absmax2 fp32 [ceil(nblocks/blocksize2)], code2 fp32[256], offset scalar. De-nesting is
 * fl(fl(code2[q] * absmax2[i / blocksize2]) + offset), identical to functional.py:1982-1984. */
/* Up to four nested-absmax 4-bit weight matrices that SHARE the activation vector, K and the code tables (q/k/v or
 * gate/up of a decoder layer) in one launch: out_i[0..m_i) = B_i[m_i, k] . A.  m, offsets: HOST arrays of `count`
 * entries; B, qabsmax, absmax2, outs: HOST arrays of `count` device pointers; code2 / datatype: device, shared by all
 * matrices.  Results are bit-identical to `count` calls of cgemm_4bit_inference_nested_*; the per-launch constant
 * (~2.7 us on a B200) is paid once.  Returns 0 ok, 1 shape not taken (issue the single calls instead). */
BNB_B200_API int cgemm_4bit_inference_nested_multi_fp16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2);
BNB_B200_API int cgemm_4bit_inference_nested_multi_bf16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2);

/* Same, N-sharded: every output slice is also stored into the peers' copies (peer_outs: HOST array of count * npeers
 * peer-mapped device addresses, matrix-major), as cgemm_4bit_inference_nested_push_* does for one matrix. */
BNB_B200_API int cgemm_4bit_inference_nested_multi_push_fp16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2, void **peer_outs, int npeers);
BNB_B200_API int cgemm_4bit_inference_nested_multi_push_bf16(int count, const int *m, int k, void *A, unsigned char **B, unsigned char **qabsmax, float **absmax2, float *code2, const float *offsets, float *datatype, void **outs, int blocksize, int blocksize2, void **peer_outs, int npeers);

/* Optional, per thread, consumed by the NEXT cgemm_4bit_inference_nested[_push]_* call: HOST copies of that call's
 * `datatype` (16 floats) and `code2` (256 floats; may be NULL).  When the host copy equals the NF4 table the kernel
 * builds its lookup table from immediates instead of waiting for a global load (which queues behind the weight
 * stream of a busy GPU).  The device pointers must still be passed. */
BNB_B200_API void cbnb_set_gemv_host_tables(const float *code16_host, const float *code2_256_host);
BNB_B200_API void cgemm_4bit_inference_nested_fp16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2);
BNB_B200_API void cgemm_4bit_inference_nested_bf16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2);

/* same GEMV for an N-sharded (column-parallel) linear on several GPUs of one NVLink box: `out` is this rank's
 * slice inside its own copy of the full output vector; the kernel also stores the slice at the same offset into
 * the peers' copies (peer_outs: HOST array of npeers <= 7 peer-mapped DEVICE pointers) -- the output all-gather of
 * SURVEY 8e fused into the GEMV epilogue.  Needs n == 1, blocksize 64, K % 256 == 0.
 * `sync` (may be NULL) folds the cross-GPU ordering into the kernels, no barrier launch: with do_signal the kernel
 * publishes sequence number (*epoch * ngroups + gidx + 1) into every peer's slot (sig_peer[i] = address of THIS
 * rank's slot inside peer i's signal array) once all its CTAs have stored and fenced; with do_wait it waits, before
 * reading A, until every slot of sig_local (one per peer) holds >= (*epoch * ngroups + gidx).  cbnb_epoch_bump()
 * increments *epoch on the launch stream (once per pass over the stack, so replayed CUDA graphs keep counting). */
typedef struct {
  unsigned int *sig_local;      /* [npeers] slots on this GPU, slot i written by peer i */
  unsigned int *sig_peer[7];
  const unsigned int *epoch;
  unsigned int *cta_counter;    /* per-GPU scratch word, zero at rest */
  int gidx, ngroups, do_signal, do_wait;
} bnb_gemv_sync_t;
BNB_B200_API void cgemm_4bit_inference_nested_push_fp16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2, void **peer_outs, int npeers, const bnb_gemv_sync_t *sync);
BNB_B200_API void cgemm_4bit_inference_nested_push_bf16(int m, int n, int k, void *A, unsigned char *B, unsigned char *qabsmax, float *absmax2, float *code2, float offset, float *datatype, void *out, int lda, int ldb, int ldc, int blocksize, int blocksize2, void **peer_outs, int npeers, const bnb_gemv_sync_t *sync);
BNB_B200_API void cbnb_epoch_bump(unsigned int *epoch);
/* Cross-GPU barrier of an N-sharded stack as one link of the programmatic-dependent-launch chain (stream-ordered,
 * graph-capturable): waits for the previous kernel of the stream, then publishes ++*counter into every peer's slot
 * (sig_peer: HOST array of npeers peer-mapped addresses) and waits until each of the npeers slots of sig_local holds
 * at least that value.  counter: device int, zero-initialised, private to the caller; every rank must call it the
 * same number of times.  The kernels after it may rely on all peers' earlier stores into symmetric memory. */
BNB_B200_API void cbnb_peer_barrier(unsigned int *counter, const unsigned int *sig_local, unsigned int **sig_peer, int npeers);

/* batch > 1 fused 4-bit GEMM (replaces dequantize_4bit + F.linear, autograd/_functions.py:490-518):
 * out[b, j] = sum_k A[b,k] * T(code[q(j,k)] * absmax[(j*K+k)/blocksize]) (+ bias[j]); A: [batch, K] T row-major,
 * B packed [N, K/2], out [batch, N] T. tcgen05 kind::f16, TMEM fp32 accumulators. */
BNB_B200_API int cgemm_4bit_fp16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize);
BNB_B200_API int cgemm_4bit_bf16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize);
/* N-sharded form of the fused GEMM (bnb_b200/parallel.py; SURVEY 8e, BASELINE config 5 at batch > 1): `out` is the base of
 * this rank's column slice inside the gathered [batch, ldo] buffer, peer_outs[i] the same address in peer i's copy of that
 * buffer (NVLink peer mappings, <= 7 peers); the epilogue stores the slice into all of them, i.e. it IS the output
 * all-gather (the reference has no multi-GPU path; parallel.py's fallback is one NCCL all-gather per linear). */
BNB_B200_API int cgemm_4bit_push_fp16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize, long ldo, void **peer_outs, int npeers);
BNB_B200_API int cgemm_4bit_push_bf16(int batch, int N, int K, void *A, unsigned char *B, float *absmax, float *datatype, void *bias, void *out, int blocksize, long ldo, void **peer_outs, int npeers);

/* B200-native int8 GEMM on ROW-MAJOR operands (the layout tcgen05 + TMA consume directly):
 * C[i,j] = sum_k A[i,k] * B[j,k]; A [m,k], B [n,k] int8 row-major (K-major), C int32 row-major [m,n]. */
BNB_B200_API int cigemm_rowmajor_32(int m, int n, int k, const int8_t *A, const int8_t *B, int *C);
/* same GEMM with the mm_dequant epilogue fused: out fp16 row-major [m,n] =
 * half(((float(acc) * 6.200012e-05f) * rowStats[i]) * colStats[j] + bias[j]) */
BNB_B200_API int cigemm_rowmajor_dequant_fp16(int m, int n, int k, const int8_t *A, const int8_t *B, float *rowStats, float *colStats, void *bias, void *out);

/* The whole LLM.int8 inference forward (MatMul8bitLt.forward, autograd/_functions.py:292-434, has_fp16_weights=False,
 * threshold > 0) as stream-ordered launches with NO host synchronisation: row statistics + outlier-column flags,
 * device-side compaction, row quantisation with the outlier columns zeroed, int8 tcgen05 GEMM with the mm_dequant
 * epilogue and the 16-bit outlier product folded into it.  A fp16 [m,k]; CB int8 [n,k] row-major; SCB fp32[n]; bias
 * fp16[n] or NULL; out fp16 [m,n].  Caller-owned device workspace: CA int8 [m,k], SCA fp32[m] (both are outputs too:
 * the quantised activations and their row statistics), colflag u8[k] (ALL ZERO on entry; left all zero on return), pos i16[k] (16-byte aligned), idx i32[idx_cap
 * >= 16] (ascending outlier columns), count i32[1], subA fp16 [m,16], subB fp16 [n,16].  Returns 0 ok, 1 shape not
 * taken (caller runs the step-by-step path), 2 error. */
BNB_B200_API int cint8_linear_fp16(void *A, const int8_t *CB, float *SCB, void *bias, void *out, float threshold, int m, int n, int k, int8_t *CA, float *SCA, unsigned char *colflag, short *pos, int *idx, int idx_cap, int *count, void *subA, void *subB);

#ifdef __cplusplus
}
#endif
#endif /* BNB_B200_H */
