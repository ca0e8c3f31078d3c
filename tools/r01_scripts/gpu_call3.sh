#!/bin/bash
tag=${1:-c3}
mkdir -p gpurun_out
out=gpurun_out/${tag}_conv3_variants.txt
: > $out
for cs in 17 21 23 32; do
  for v in "NPP_CONV3_2PROD=0" "NPP_CONV3_2PROD=1" "NPP_CONV3_2PROD=1 NPP_CONV3_TILE=1"; do
    echo "### $v" >> $out
    env $v timeout 90 tests/csrc/_bin/test_conv $cs 2>&1 | grep "case\|time\|FAIL" | grep -v "OK$" >> $out
  done
done
cat $out | cut -c1-200
for cs in 17 18 33; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm2_kernel|conv3_kernel" -c 2 \
      -o gpurun_out/${tag}_case$cs -f tests/csrc/_bin/test_conv $cs > gpurun_out/${tag}_case$cs.log 2>&1
  ncu -i gpurun_out/${tag}_case$cs.ncu-rep --page raw --csv > gpurun_out/${tag}_case${cs}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${tag}_case$cs.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_case${cs}_source.csv.gz
  rm -f gpurun_out/${tag}_case$cs.ncu-rep
done
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?" >> gpurun_out/${tag}_bench.err
timeout 300 python tools/shape_table.py --top 150 > gpurun_out/${tag}_shape_table.txt 2> gpurun_out/${tag}_shape_table.err
cut -c1-700 gpurun_out/${tag}_bench.json
