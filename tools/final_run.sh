set -x
python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference2.json 2>> gpurun_out/bench_r1_final2.err
python bench.py --workload obstacles --batch 8192 --steps 10 --no-cpu-baseline > gpurun_out/bench_r1_obstacles2.json 2>> gpurun_out/bench_r1_final2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_solve_pair -s 6 -c 1 -f -o gpurun_out/prof_pair_c python tools/ab_step.py --steps 4 > gpurun_out/ncu_pair2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 6 -c 1 -f -o gpurun_out/prof_raycast_r1h python tools/ab_step.py --steps 4 > gpurun_out/ncu_ray2.log 2>&1
tail -3 gpurun_out/bench_r1_final2.err
