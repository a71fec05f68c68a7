#!/bin/bash
# Everything profiles/ quotes, in one call on the GPU box (1 GPU):  gpurun -- 'bash profiles/refresh.sh TAG'
# Each ncu pass runs only after the same command has exited 0 without ncu.
set -u
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err || { echo "bench failed"; tail -5 $O/${TAG}_bench_n1.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench_n1.err || echo "reference arm failed"
python bench.py --steps 20 --warmup 3 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches_bench.csv \
      python bench.py --steps 20 --warmup 3 > $O/ncu_launches.log 2>&1
python profiles/tune_scenarios.py > $O/${TAG}_scenarios.txt 2>&1
python profiles/tune.py >> $O/${TAG}_scenarios.txt 2>&1
python profiles/time_fullgame.py 1 60 >> $O/${TAG}_scenarios.txt 2>&1
echo "--- fresh commands every cycle (8 command tensors in turn: the bench workload)" >> $O/${TAG}_scenarios.txt
FG_POOL=8 python profiles/time_fullgame.py 1 30 >> $O/${TAG}_scenarios.txt 2>&1
echo "--- every player the same kind of command: 0 none, 1 dash, 4 go-to-point" >> $O/${TAG}_scenarios.txt
for c in 0 1 4; do FG_CMD=$c python profiles/time_fullgame.py 1 10 2>&1 | grep "^flush" >> $O/${TAG}_scenarios.txt; done
python profiles/time_step_phases.py >> $O/${TAG}_scenarios.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fg_layout profiles/micro/fg_layout.cu && /tmp/fg_layout > $O/${TAG}_fg_layout_probe.txt 2>&1
python profiles/time_fullgame.py 16 6 >> $O/${TAG}_scenarios.txt 2>&1
FG_MATCHES=32768 python profiles/time_fullgame.py 1 30 >> $O/${TAG}_scenarios.txt 2>&1
python profiles/tune_rollout.py >> $O/${TAG}_scenarios.txt 2>&1
python profiles/prof_step.py both 2 > /dev/null 2>&1 && {
  ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 150 -c 1 -o $O/prof_${TAG}_k16 -f python profiles/prof_step.py k16 2 150 > $O/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o $O/prof_${TAG}_k1 -f python profiles/prof_step.py k1 2 >> $O/ncu_full.log 2>&1
}
python profiles/prof_fullgame.py > /dev/null 2>&1 && {
  ncu --set full --clock-control none --import-source on -k regex:fullgame_step -s 3 -c 1 -o $O/prof_${TAG}_fg_k1 -f python profiles/prof_fullgame.py >> $O/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:fullgame_step -s 8 -c 1 -o $O/prof_${TAG}_fg_k16 -f python profiles/prof_fullgame.py >> $O/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:fullgame_step -s 45 -c 1 -o $O/prof_${TAG}_fg_k1_late -f python profiles/prof_fullgame.py 48 >> $O/ncu_full.log 2>&1
}
python profiles/prof_rollout.py tf32 > /dev/null 2>&1 && {
  ncu --set full --clock-control none --import-source on -k regex:rollout_mlp -s 2 -c 1 -o $O/prof_${TAG}_rollout_tc5 -f python profiles/prof_rollout.py tf32 >> $O/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rollout_mlp -s 2 -c 1 -o $O/prof_${TAG}_rollout_mma -f python profiles/prof_rollout.py tf32_mma_sync >> $O/ncu_full.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rollout_mlp -s 2 -c 1 -o $O/prof_${TAG}_actor_tc5 -f python profiles/prof_rollout.py actor >> $O/ncu_full.log 2>&1
}
# summaries here (the captures are ~18 MB each; gpurun brings back at most 64 MiB): CSV + traffic.json + source profiles
python profiles/summarize_ncu.py $O/${TAG}_kernels_ncu_full.csv $O/${TAG}_traffic.json \
  k16_envs_1048576:16777216=$O/prof_${TAG}_k16.ncu-rep k1_envs_8388608:8388608=$O/prof_${TAG}_k1.ncu-rep \
  fullgame_k1_envs_262144:262144=$O/prof_${TAG}_fg_k1.ncu-rep fullgame_k16_envs_262144:4194304=$O/prof_${TAG}_fg_k16.ncu-rep \
  fullgame_k1_late_envs_262144:262144=$O/prof_${TAG}_fg_k1_late.ncu-rep \
  rollout_tc5_k16_envs_1048576:16777216=$O/prof_${TAG}_rollout_tc5.ncu-rep rollout_mma_k16_envs_1048576:16777216=$O/prof_${TAG}_rollout_mma.ncu-rep actor_tc5_k16_envs_1048576:16777216=$O/prof_${TAG}_actor_tc5.ncu-rep \
  > $O/summarize.log 2>&1
for c in fg_k1:262144 fg_k1_late:262144 rollout_tc5:524288 rollout_mma:524288 k16:524288; do
  n=${c%%:*}; u=${c##*:}
  python profiles/source_profile.py $O/prof_${TAG}_$n.ncu-rep $u 40 > $O/${TAG}_source_profile_$n.txt 2>&1
done
rm -f $O/prof_${TAG}_actor_tc5.ncu-rep $O/prof_${TAG}_k16.ncu-rep $O/prof_${TAG}_k1.ncu-rep $O/prof_${TAG}_fg_k16.ncu-rep $O/prof_${TAG}_fg_k1_late.ncu-rep $O/prof_${TAG}_rollout_mma.ncu-rep
ls -la $O | tail -24
