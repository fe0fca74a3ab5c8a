// ref_shim.cpp -- extern "C" entry points over the reference's OWN cpu_ops (compiled in place
// from /root/reference/sycl/{cpu_ops,common}.cpp; no reference source is copied into this repo).
// Mirrors the two wrappers the reference keeps in its SYCL-only TU
// (sycl/pythonInterface.cpp:419-420), which cannot be compiled here.
// TEST INFRASTRUCTURE ONLY (see oracle/bnb_oracle.c header).
#include <cpu_ops.h>

extern "C" {
void cquantize_blockwise_cpu_fp32(float *code, float *A, float *absmax, unsigned char *out,
                                  long long blocksize, long long n) {
  quantize_cpu(code, A, absmax, out, blocksize, n);
}
void cdequantize_blockwise_cpu_fp32(float *code, unsigned char *A, float *absmax, float *out,
                                    long long blocksize, long long n) {
  dequantize_cpu(code, A, absmax, out, blocksize, n);
}
}
