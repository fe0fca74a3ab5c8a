// codebooks.cuh -- the 4-bit codebooks and decision thresholds of the reference, as data.
// Values are the literals of sycl/sycl_code/kernel_quant.cpp: dDequantizeNF4 (:650-703),
// dQuantizeNF4 (:705-756), dQuantizeFP4 (:547-594), dDequantizeFP4Tree (:520-545); tests compare them
// with tests/golden/ref_kernel_constants.json (parsed out of the reference source).
#pragma once

namespace bnb {

#define BNB_NF4_TABLE                                                                                   \
  {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f, -0.28444138169288635f,   \
   -0.18477343022823334f, -0.09105003625154495f, 0.0f, 0.07958029955625534f, 0.16093020141124725f,     \
   0.24611230194568634f, 0.33791524171829224f, 0.44070982933044434f, 0.5626170039176941f,              \
   0.7229568362236023f, 1.0f}

// code = number of thresholds strictly below x (the reference's tree of strict '>' compares)
#define BNB_NF4_THRESHOLDS                                                                              \
  {-0.8480964004993439f, -0.6106329262256622f, -0.4599952697753906f, -0.33967943489551544f,            \
   -0.23460740596055984f, -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f,         \
   0.1202552504837513f, 0.2035212516784668f, 0.2920137718319893f, 0.3893125355243683f,                 \
   0.5016634166240692f, 0.6427869200706482f, 0.8614784181118011f}

// FP4: magnitude for the low three bits 0..7; bit 3 is the sign
#define BNB_FP4_MAGNITUDES                                                                              \
  {0.00000000f, 5.208333333e-03f, 0.66666667f, 1.00000000f, 0.33333333f, 0.50000000f, 0.16666667f,     \
   0.25000000f}
// FP4 quantize: ascending thresholds on |x| and the code of each of the 8 buckets they delimit
#define BNB_FP4_THRESHOLDS {0.00260417f, 0.0859375f, 0.20833333f, 0.29166667f, 0.4166667f, 0.583333f, 0.8333333f}
#define BNB_FP4_BUCKET_CODES {0, 1, 6, 7, 4, 5, 2, 3}

// the reference's trees, verbatim in behaviour (device + host): used by the self-test and the slow paths
__host__ __device__ inline unsigned char quantize_nf4_tree(float x) {
  if (x > 0.03979014977812767f)
    if (x > 0.3893125355243683f)
      if (x > 0.6427869200706482f)
        return x > 0.8614784181118011f ? 15 : 14;
      else
        return x > 0.5016634166240692f ? 13 : 12;
    else if (x > 0.2035212516784668f)
      return x > 0.2920137718319893f ? 11 : 10;
    else
      return x > 0.1202552504837513f ? 9 : 8;
  else if (x > -0.33967943489551544f)
    if (x > -0.13791173323988914f)
      return x > -0.045525018125772476f ? 7 : 6;
    else
      return x > -0.23460740596055984f ? 5 : 4;
  else if (x > -0.6106329262256622f)
    return x > -0.4599952697753906f ? 3 : 2;
  else
    return x > -0.8480964004993439f ? 1 : 0;
}

__host__ __device__ inline unsigned char quantize_fp4_tree(float x) {
  int sign = x < 0 ? 8 : 0;
  x = fabsf(x);
  if (x > 0.29166667f)
    if (x > 0.583333f)
      return (x > 0.8333333f ? 3 : 2) + sign;
    else
      return (x > 0.4166667f ? 5 : 4) + sign;
  else if (x > 0.0859375f)
    return (x > 0.20833333f ? 7 : 6) + sign;
  else
    return (x > 0.00260417f ? 1 : 0) + sign;
}

// dQuantize<0> (kernel_quant.cpp:765-819): 7-step pivot search on the 256-entry code + midpoint rounding
__device__ __forceinline__ unsigned char quantize_8bit_search(const float *code, float x) {
  int pivot = 127, upper_pivot = 255, lower_pivot = 0;
  float lower = -1.0f, upper = 1.0f;
  float val = code[pivot];
#pragma unroll
  for (int i = 64; i > 0; i >>= 1) {
    if (x > val) { lower_pivot = pivot; lower = val; pivot += i; }
    else         { upper_pivot = pivot; upper = val; pivot -= i; }
    val = code[pivot];
  }
  if (upper_pivot == 255) upper = code[upper_pivot];
  if (lower_pivot == 0) lower = code[lower_pivot];
  if (x > val) {
    float midpoint = __fmul_rn(__fadd_rn(upper, val), 0.5f);
    return (unsigned char)(x > midpoint ? upper_pivot : pivot);
  } else {
    float midpoint = __fmul_rn(__fadd_rn(lower, val), 0.5f);
    return (unsigned char)(x < midpoint ? lower_pivot : pivot);
  }
}

}  // namespace bnb
