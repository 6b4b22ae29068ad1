#!/bin/bash
# usage: run_mgpu.sh WORLD SCENARIO
W=$1; S=$2
rm -f /tmp/nccl_id_$S
for r in $(seq 0 $((W-1))); do
  python tests/mgpu_worker.py $W $r /tmp/nccl_id_$S $S > gpurun_out/mgpu_${S}_$r.log 2>&1 &
done
wait
for r in $(seq 0 $((W-1))); do echo "== rank $r"; tail -4 gpurun_out/mgpu_${S}_$r.log; done
