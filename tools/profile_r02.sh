#!/bin/bash
# Round-2 ncu captures, run on the GPU box:  gpurun --timeout 1800 -- 'bash tools/profile_r02.sh'
# (ncu times are cold-cache and serialised: shares and DRAM bytes are what they are for.)
O=gpurun_out
FULL="--set full --clock-control none --import-source on"
# 1. every launch of the bench command with its device time
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > $O/r02_prof_plain.json 2> $O/r02_prof_plain.err &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --no-extras --no-cpu > $O/r02_launches_bench.log 2>&1
echo "launch list rc=$?"
# 2. the step kernel in steady state, ONE step per launch (tools/cfg3_step.py), full set
timeout 600 ncu $FULL -k regex:k_step_fused -s 305 -c 2 -o $O/r02_step_fused_grid python tools/cfg3_step.py 310 > $O/r02_ncu_fused.log 2>&1
echo "fused rc=$?"
# 3. the same step as one shard of one (k_step_shard: the sharded phase sequence against the rank's own region)
BH_FUSED=shard timeout 600 ncu $FULL -k regex:k_step_shard -s 305 -c 2 -o $O/r02_step_shard python tools/cfg3_step.py 310 > $O/r02_ncu_shard.log 2>&1
echo "shard rc=$?"
# 4. every per-stage kernel of one step at the same size
timeout 900 ncu $FULL -k 'regex:^k_(sp_|topk|tm_|duty|rng_|boost|inhibit)' -c 90 -o $O/r02_cfg3_kernels python tools/cfg3_kernels.py 65536 16384 300 1 > $O/r02_ncu_kernels.log 2>&1
echo "kernels rc=$?"
# 5. the shared-mask batched overlap (tcgen05 / mma.sync / popcount)
timeout 600 ncu $FULL -k regex:k_sp_overlap_batched_t5 -c 1 -o $O/r02_overlap_t5 python tools/batched_overlap.py 256 65536 16384 5 > $O/r02_ncu_batched_t5.log 2>&1
timeout 600 ncu $FULL -k regex:k_sp_overlap_batched_tc -c 1 -o $O/r02_overlap_mma python tools/batched_overlap.py 256 65536 16384 5 > $O/r02_ncu_batched_mma.log 2>&1
echo "batched rc=$?"
# what travels back is capped at 64 MiB: raw pages as CSV for every capture, the reports themselves only for the
# step kernel and the tcgen05 kernel
for r in r02_step_fused_grid r02_step_shard r02_cfg3_kernels r02_overlap_t5 r02_overlap_mma; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2> /dev/null
done
ncu -i $O/r02_step_fused_grid.ncu-rep --page source --csv > $O/r02_step_fused_grid.source.csv 2> /dev/null
rm -f $O/r02_step_shard.ncu-rep $O/r02_cfg3_kernels.ncu-rep $O/r02_overlap_mma.ncu-rep
ls -la $O/
