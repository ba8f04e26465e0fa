#!/bin/bash
# tools/final_measure.sh -- the single-GPU evidence of a round, one gpurun call:
#   gpurun --timeout 2400 -- 'bash tools/final_measure.sh r02'
# writes gpurun_out/<tag>_*.{json,csv,log,ncu-rep}; copy what is to be judged into profiles/.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log
# the bench lines: default (C2, with the CPU arm), the reference arm, the other workloads
timeout 600 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; tail -1 $out/${tag}_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_c2_reference.json 2> $out/${tag}_bench_c2_reference.err
for w in C1 C3 C4 C5; do
  timeout 900 python bench.py --workload $w --steps 5 --cpu-budget 8 > $out/${tag}_bench_${w}.json 2> $out/${tag}_bench_${w}.err; tail -1 $out/${tag}_bench_${w}.err
done
# ncu: launch list of the default command, then the top kernels in full (single-pass forced: launch numbering is fixed)
PFNAV_TWO_PHASE=2 timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 1500 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_list.log 2>&1
PFNAV_TWO_PHASE=2 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_agent_velocity|k_cohesion|k_desired_velocity|k_flow_unit|k_cell_sort" -s 30 -c 8 \
    -o $out/${tag}_c2_tick python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_tick.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_los -s 3 -c 1 \
    -o $out/${tag}_c2_los python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_los.log 2>&1
PFNAV_TWO_PHASE=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_agent_velocity -s 24 -c 1 \
    -o $out/${tag}_c2_moving python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_moving.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cohesion_window|k_blockers|k_chunks" -c 6 \
    -o $out/${tag}_c5 python bench.py --workload C5 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_c5.log 2>&1
ls -la $out/${tag}_*
