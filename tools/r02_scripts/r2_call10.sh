#!/bin/bash
# Round-2 GPU call 10 (1 GPU): two task streams on two CUDA streams (tests + A/B bench), ncu of the small-shape convs.
tag=r2c10
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_two_streams.py -q -s -x ) > gpurun_out/${tag}_pytest_two.log 2>&1
grep -E "passed|failed|FAILED|two streams|one stream" gpurun_out/${tag}_pytest_two.log | tail -6 | cut -c1-400
NPP_TWO_STREAMS=1 timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_two.json 2> gpurun_out/${tag}_bench_two.err
echo "bench two exit $?"; grep '^{' gpurun_out/${tag}_bench_two.json | cut -c1-260; tail -3 gpurun_out/${tag}_bench_two.err | cut -c1-300
timeout 400 python bench.py --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_one.json 2> gpurun_out/${tag}_bench_one.err
echo "bench one exit $?"; grep '^{' gpurun_out/${tag}_bench_one.json | cut -c1-260
( time NPP_TWO_STREAMS=1 timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_engine.py tests/test_gpu_golden.py tests/test_gpu_checkpoint.py -q -x ) > gpurun_out/${tag}_pytest_suite_two.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/${tag}_pytest_suite_two.log | tail -5 | cut -c1-200
cap () {  # name regex skip count
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"$2" -s $3 -c $4 \
      -o gpurun_out/${tag}_$1 -f python tools/profile_step.py > gpurun_out/${tag}_ncu_$1.log 2>&1
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page details > gpurun_out/${tag}_$1_details.txt 2>/dev/null
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_$1_source.csv.gz
  rm -f gpurun_out/${tag}_$1.ncu-rep
  grep -E "^  void|^  [a-z_:]*kernel|Duration  " gpurun_out/${tag}_$1_details.txt | head -6 | cut -c1-160
}
cap conv3_32 'conv3_kernel<\(int\)32>' 2 2
cap gemm2_64 'conv_gemm2_kernel<\(int\)64>' 6 2
cap wgrad3_64 'conv_wgrad3_kernel<\(int\)64>' 2 2
