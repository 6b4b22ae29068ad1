#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --side-configs none --steps 3 --warmup 3 > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.log
  echo "== $name: $*"
  grep -E "dsc upload|host PBVH" gpurun_out/exp_$name.log | head -12
  python tools/show_bench.py gpurun_out/exp_$name.json | grep -E "^c3|radius sweep"
}
run base DSC_TIMING=1
run pdl DSC_PDL=1
