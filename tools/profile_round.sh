#!/bin/bash
# The profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo, run on the GPU box under gpurun:
#   bash tools/profile_round.sh <tag>     -> gpurun_out/<tag>_launches.csv, <tag>_kernels.csv, <tag>_<kernel>.ncu-rep
# Numbers printed by bench.py under ncu are NOT bench values; only the per-launch ncu metrics are used.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio
# every launch with its device time (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.out 2>&1
# per-kernel counters of the steady state (the first batches grow the arenas and re-launch the emitters)
ncu --metrics $M --clock-control none -s 60 -c 80 --csv --log-file $OUT/${TAG}_kernels.csv $B > $OUT/${TAG}_kernels.out 2>&1
# full-set captures of the top kernels
for k in k_chunk_emit k_smooth_chunks k_chunk_count k_terrain2d_sheet k_scan_chunks; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $OUT/${TAG}_$k $B > $OUT/${TAG}_$k.out 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_seam_pass -s 2 -c 1 -f -o $OUT/${TAG}_k_seam_pass python tools/seam_probe.py > $OUT/${TAG}_k_seam_pass.out 2>&1
# the 3-D noise kernel (the bounding stage of the fractal-noise configs) on a 512-chunk batch
ncu --set full --clock-control none --import-source on -k regex:k_terrain3d -s 1 -c 1 -f -o $OUT/${TAG}_k_terrain3d python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --sampler terrain3d_pert --chunks-per-axis 8 > $OUT/${TAG}_k_terrain3d.out 2>&1
# summaries are made here, on the box: only 64 MiB of gpurun_out/ travel back, so most of the .ncu-rep files stay behind
python tools/ncu_summary.py $OUT/${TAG}_ncu_full_summary.txt $OUT/${TAG}_k_*.ncu-rep
for k in k_chunk_emit k_chunk_count k_smooth_chunks; do
  python tools/ncu_lines.py $OUT/${TAG}_$k.ncu-rep $k 40 --phases $( [ $k = k_smooth_chunks ] && echo smooth.cuh || echo fused.cuh ) > $OUT/${TAG}_${k}_lines.txt 2>&1
done
for k in k_chunk_count k_terrain2d_sheet k_scan_chunks k_seam_pass k_terrain3d; do rm -f $OUT/${TAG}_$k.ncu-rep; done
ls -la $OUT | grep $TAG
