# ncu passes over ONE replay of the headline step (148 clips, fused-row layout): metrics per launch, launch list, full-set captures
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
python tools/profile_step.py > gpurun_out/prof0.log 2>&1 || exit 1
ncu --metrics $M --clock-control none --profile-from-start off -c 80 --csv --log-file gpurun_out/ncu_metrics_r2.csv python tools/profile_step.py > gpurun_out/ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py > gpurun_out/ncu2.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'select_rows_kernel|rows_gather|nms_matrix|nms_kernel|attn_pv|attn_round2' -c 9 -o gpurun_out/full_r2 python tools/profile_step.py > gpurun_out/ncu3.log 2>&1
ncu -i gpurun_out/full_r2.ncu-rep --page raw --csv > gpurun_out/ncu_full_r2.csv
ls -la gpurun_out | tail -8
