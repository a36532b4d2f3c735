#!/bin/bash
# usage: tools/gpu_room_sweep.sh NGPUS [extra args]; writes gpurun_out/room_sweep_nN.json
N=${1:-1}; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python tools/room_sweep.py "$@" > gpurun_out/room_sweep_n$N.json 2> gpurun_out/room_sweep_n$N.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/room_sweep.py "$@" > gpurun_out/room_sweep_n$N.json 2> gpurun_out/room_sweep_n$N.err
fi
echo "exit $?"; tail -3 gpurun_out/room_sweep_n$N.err; cat gpurun_out/room_sweep_n$N.json
