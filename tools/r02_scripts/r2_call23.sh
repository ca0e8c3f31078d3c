#!/bin/bash
# Round-2 GPU call 23 (1 GPU): batched statistics loads in the conv epilogue: parity cases, timings, train/infer bench.
tag=r2c23
mkdir -p gpurun_out
bash tools/r02_scripts/r2_epi2_ab.sh ${tag} > gpurun_out/${tag}_stdout.txt 2>&1
grep -E "FAIL|exit code|run_conv_cases exit|timed out" gpurun_out/${tag}_stdout.txt | head
grep -A1 "EPI2=1" gpurun_out/${tag}_ab.txt | grep time | cut -c1-140
run () {  # name workload env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 400 python bench.py --workload $wl --steps 10 --no-cpu-baseline --no-gpu-reference --no-kernel-table > gpurun_out/${tag}_bench_$name.json 2> gpurun_out/${tag}_bench_$name.err
  echo "bench $name exit $?: $(python -c "import json;d=[json.loads(l) for l in open('gpurun_out/${tag}_bench_$name.json') if l.startswith('{\"')][0];print(d['ms_per_step'], d['value'], d['e2e']['value'])" 2>/dev/null)"
  tail -2 gpurun_out/${tag}_bench_$name.err | cut -c1-300
}
run train train NPP_PDL=1
run infer512 infer512 NPP_PDL=1
run search search NPP_PDL=1
