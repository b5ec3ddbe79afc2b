# tools/final_run_r2.sh -- the round-2 record on one B200: bench lines, launch list, ncu captures (run under gpurun)
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_4096cars.json 2> gpurun_out/r2_final.err
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_final.err
python bench.py --workload obstacles > gpurun_out/r2_bench_obstacles_8192.json 2>> gpurun_out/r2_final.err
python bench.py --workload timeopt > gpurun_out/r2_bench_timeopt_32768_1gpu.json 2>> gpurun_out/r2_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity-setting --sustain-seconds 0 > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_solve_pair -s 6 -c 1 -f -o gpurun_out/r2_prof_pair python tools/ab_step.py --steps 4 > gpurun_out/r2_ncu_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:localize_gather -s 6 -c 1 -f -o gpurun_out/r2_prof_gather python tools/ab_step.py --steps 4 > gpurun_out/r2_ncu_gather.log 2>&1
MPC_WIDTH_MEMO=off ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 6 -c 1 -f -o gpurun_out/r2_prof_raycast1 python tools/ab_step.py --steps 4 > gpurun_out/r2_ncu_ray1.log 2>&1
tail -3 gpurun_out/r2_final.err
