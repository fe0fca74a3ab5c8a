#!/bin/bash
# GPU session 10: per-CTA traces of the GEMV chain; ncu --set full of the int8 GEMM, dequantize, double quant, GEMV
for shp in "4096 4096" "11008 4096" "4096 11008" "28672 8192"; do
  set -- $shp
  BNB_B200_GEMV_PROBE=2 timeout 120 python tools/gemv_trace.py $1 $2 > gpurun_out/r10_trace_$1x$2.json 2> gpurun_out/r10_trace_$1x$2.err
  cat gpurun_out/r10_trace_$1x$2.json
done
python tools/run_one.py igemm > gpurun_out/r10_plain_igemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_igemm_tcgen05 -c 2 -o gpurun_out/prof_igemm_r2 python tools/run_one.py igemm > gpurun_out/r10_ncu_igemm.log 2>&1
python tools/run_one.py dequant bf16 > gpurun_out/r10_plain_dq.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_dequantize|k_quantize4_bulk" -s 8 -c 4 -o gpurun_out/prof_quant_r2 python tools/run_one.py dequant bf16 > gpurun_out/r10_ncu_dq.log 2>&1
python tools/run_one.py stats > gpurun_out/r10_plain_stats.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_double_rowcol_quant|k_col_row_stats" -s 4 -c 4 -o gpurun_out/prof_int8quant_r2 python tools/run_one.py stats > gpurun_out/r10_ncu_stats.log 2>&1
python tools/run_one.py gemv 11008 4096 > gpurun_out/r10_plain_gemv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gemv4_bc -s 4 -c 2 -o gpurun_out/prof_gemv_bc_r2 python tools/run_one.py gemv 11008 4096 > gpurun_out/r10_ncu_gemv.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
