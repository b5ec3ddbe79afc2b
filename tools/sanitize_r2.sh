# tools/sanitize_r2.sh -- compute-sanitizer memcheck over smoke() and the tests of the kernels changed in round 2
set -x
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_san_smoke.log 2>&1; echo "smoke rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -x -q -k "more_free_segments or width_table or static_width_batched or final_checks or teacher_forced or raycast_modes or host_step or solve_order" > gpurun_out/r2_san_tests.log 2>&1; echo "tests rc=$?"
MPC_ADMM_KERNEL=tm compute-sanitizer --tool memcheck --error-exitcode 1 python tools/ab_step.py --batch 512 --steps 2 > gpurun_out/r2_san_tm.log 2>&1; echo "tm rc=$?"
MPC_ADMM_KERNEL=quad compute-sanitizer --tool memcheck --error-exitcode 1 python tools/ab_step.py --batch 512 --steps 2 > gpurun_out/r2_san_quad.log 2>&1; echo "quad rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_race_smoke.log 2>&1; echo "racecheck smoke rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests -m gpu -x -q -k "host_step_without_copy_nodes and 0-False" > gpurun_out/r2_race_hoststep.log 2>&1; echo "racecheck host step rc=$?"
for f in gpurun_out/r2_san_*.log gpurun_out/r2_race_hoststep.log gpurun_out/r2_race_smoke.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $f | tail -3; done
