#!/bin/bash
# ncu evidence of round 2: (1) every kernel of ONE network evaluation at the bench shapes (64 patches), (2) the kernels around the
# network.  Metric set kept to what the roofline needs so that ~200 launches finish in minutes.
mkdir -p gpurun_out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum
timeout 1500 ncu --metrics $MET --clock-control none -s 1500 -c 420 --csv --log-file gpurun_out/ncu_eval.csv \
  python bench.py --steps 1 --warmup 0 --no-roofline --no-cpu-baseline --no-extra --no-graph > gpurun_out/ncu_eval.log 2>&1
echo "eval exit $?"; grep -c "gpu__time_duration" gpurun_out/ncu_eval.csv
timeout 900 ncu --metrics $MET --clock-control none --csv --log-file gpurun_out/ncu_aux.csv python tools/prof_aux.py > gpurun_out/ncu_aux.log 2>&1
echo "aux exit $?"; tail -2 gpurun_out/ncu_aux.log; grep -c "gpu__time_duration" gpurun_out/ncu_aux.csv
