// gemm_4bit.cu -- K4 placeholder until the fused tcgen05 kernel lands (returns 1 == not implemented;
// the Python layer then takes the reference's own batch>1 route: dequantize_4bit + F.linear on the GPU).
#include "common.cuh"
namespace bnb {
template <typename T>
int gemm_4bit(int, int, int, const T *, const unsigned char *, const float *, const float *, const T *, T *, int) { return 1; }
template int gemm_4bit<__half>(int, int, int, const __half *, const unsigned char *, const float *, const float *, const __half *, __half *, int);
template int gemm_4bit<__nv_bfloat16>(int, int, int, const __nv_bfloat16 *, const unsigned char *, const float *, const float *, const __nv_bfloat16 *, __nv_bfloat16 *, int);
}
