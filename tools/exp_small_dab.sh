#!/bin/bash
# small-dab routing experiment: the C3 sweep under different settings (per-radius us/dab from show_bench)
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --side-configs none --steps 3 --warmup 3 > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.log
  echo "== $name: $*"
  python tools/show_bench.py gpurun_out/exp_$name.json | grep -E "^c3|radius sweep"
}
run lazy X=1
run eager DSC_EAGER_REFIT=1
run g32 DSC_BATCH_KERNEL=1 DSC_SMALL_DAB_FRAC=0.06 DSC_BATCH_GRID=32
run g96 DSC_BATCH_KERNEL=1 DSC_SMALL_DAB_FRAC=0.06 DSC_BATCH_GRID=96
run g148w DSC_BATCH_KERNEL=1 DSC_SMALL_DAB_FRAC=0.11 DSC_BATCH_GRID=148
run g296w DSC_BATCH_KERNEL=1 DSC_SMALL_DAB_FRAC=0.11 DSC_BATCH_GRID=296
