# Round-2 profiling call for the z pass (0.77 of the HBM roofline at 512 planes against 0.90 at 128):
# which counters move with the number of planes -- address translation, DRAM row locality or
# shared-memory wavefronts (bank conflicts of the tile reads: 37-43 % of the wavefronts in round 1)?
set -x
mkdir -p gpurun_out
ncu --query-metrics 2>/dev/null | grep -i -E "tlb|mmu|pte|translation" > gpurun_out/r2_ncu_tlb_metric_names.txt
head -40 gpurun_out/r2_ncu_tlb_metric_names.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_active.avg.pct_of_peak_sustained_elapsed"
M="$M,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
M="$M,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"
M="$M,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum"
for shape in "2048 512 128" "1024 512 256" "512 512 512"; do
  set -- $shape
  timeout 120 python tools/prof_lapl.py --shape $1 $2 $3 > gpurun_out/r2_zpass_plain_$3.log 2>&1 && \
  timeout 300 ncu --metrics $M --clock-control none -k regex:yz_tma_kernel --csv --log-file gpurun_out/r2_zpass_ncu_$3.csv \
    python tools/prof_lapl.py --shape $1 $2 $3 > gpurun_out/r2_zpass_ncu_$3.log 2>&1
  tail -3 gpurun_out/r2_zpass_plain_$3.log
done
# the same with the bank-conflict-free tile reads (PBX_YZ_ROT=1: swizzled tiles + register swaps)
for rot in 0 1; do
  PBX_YZ_ROT=$rot timeout 120 python tools/prof_lapl.py --n 512 > gpurun_out/r2_rot${rot}_plain.log 2>&1 && \
  PBX_YZ_ROT=$rot timeout 300 ncu --metrics $M --clock-control none -k regex:yz_tma_kernel --csv --log-file gpurun_out/r2_rot${rot}_ncu.csv \
    python tools/prof_lapl.py --n 512 > gpurun_out/r2_rot${rot}_ncu.log 2>&1
  tail -2 gpurun_out/r2_rot${rot}_plain.log
done
