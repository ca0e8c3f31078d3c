#!/bin/bash
# Round-2 GPU call 9 (1 GPU): ncu --set full (+ source) of the small-shape conv kernels that sink the class average:
# conv3_kernel<32> (3x3 32->32 @96^2 fprop / dgrad), conv_gemm2_kernel<64> and <32> (1x1), conv_wgrad3_kernel<64>.
tag=r2c9
mkdir -p gpurun_out
cap () {  # name regex skip count
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -s $3 -c $4 \
      -o gpurun_out/${tag}_$1 -f python tools/profile_step.py > gpurun_out/${tag}_ncu_$1.log 2>&1
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page details > gpurun_out/${tag}_$1_details.txt 2>/dev/null
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_$1_source.csv.gz
  rm -f gpurun_out/${tag}_$1.ncu-rep
  grep -E "^  void|^  [a-z_:]*kernel|Duration  " gpurun_out/${tag}_$1_details.txt | head -8 | cut -c1-200
}
cap conv3_32 'conv3_kernel<32>' 4 2
cap gemm2_64 'conv_gemm2_kernel<64>' 6 2
cap gemm2_32 'conv_gemm2_kernel<32>' 2 2
cap wgrad3_64 'conv_wgrad3_kernel<64>' 2 2
( time timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_baseline_config.py tests/test_gpu_labels.py tests/test_gpu_engine.py -q -x ) > gpurun_out/${tag}_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/${tag}_pytest.log | tail -5 | cut -c1-200
