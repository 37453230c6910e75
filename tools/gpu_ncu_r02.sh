#!/usr/bin/env bash
# round-2 profiles of the headline bench workload (TactilePush 32x13, B=4096, T=200), one GPU:
#   1. launch list of one bench step (gpu__time_duration.sum per launch; cold-cache, serialised: compare SHARES)
#   2. counters of fwd_kernel at the bench size: DRAM bytes, fp64 thread-instruction counts, pipe / issue / stall figures
#   3. ncu --set full of fwd_kernel at T=20 (report kept for the source page)
set -u
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv $BENCH > $OUT/r02_launches.log 2>&1
echo "launch list rc=$?"; grep -c tsimns $OUT/r02_launches.csv
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
M=$M,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
M=$M,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
M=$M,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
M=$M,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,sm__cycles_elapsed.max
M=$M,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:'fwd_kernel|tape_kernel|tac_kernel|vjp_kernel|bwd_kernel' --launch-skip 6 -c 6 --csv --log-file $OUT/r02_counters.csv $BENCH > $OUT/r02_counters.log 2>&1
echo "counters rc=$?"; grep -c tsimns $OUT/r02_counters.csv
ncu --set full --clock-control none --import-source on -k regex:'fwd_kernel' -c 1 -f -o $OUT/r02_fwd_full \
    python tools/perf_probe.py --B 4096 --T 20 --lanes 8 --reps 1 --grad-only > $OUT/r02_full.log 2>&1
echo "full rc=$?"; ls -la $OUT/r02_fwd_full.ncu-rep
