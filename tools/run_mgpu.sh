#!/bin/bash
# one multi-GPU parity scenario, one process per GPU, every rank under its own timeout: tools/run_mgpu.sh WORLD SCENARIO [SECONDS]
W=$1; S=$2; T=${3:-90}
rm -f /tmp/nccl_id_$S
mkdir -p gpurun_out
for r in $(seq 0 $((W-1))); do
  timeout $T python tests/mgpu_worker.py $W $r /tmp/nccl_id_$S $S > gpurun_out/mgpu_${S}_$r.log 2>&1 &
done
wait
for r in $(seq 0 $((W-1))); do echo "== rank $r"; tail -5 gpurun_out/mgpu_${S}_$r.log | cut -c1-400; done
