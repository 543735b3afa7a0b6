#!/bin/bash
# Round-2 measurement artifacts on one B200 (run under gpurun): outputs under gpurun_out/, copied to profiles/ here.
export LAGB_HEAD=$(cat .lagb_head 2>/dev/null || echo unknown)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/a_gpu_tests.log 2>&1; tail -2 gpurun_out/a_gpu_tests.log
# 1. headline bench lines
python bench.py > gpurun_out/a_bench_1gpu.json 2> gpurun_out/a_bench_1gpu.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/a_bench_1gpu_20steps.json 2>> gpurun_out/a_bench_1gpu.err
python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/a_bench_reference_arm.json 2>> gpurun_out/a_bench_1gpu.err
python bench.py --problem 0 --no-cpu > gpurun_out/a_bench_taylor_green.json 2>> gpurun_out/a_bench_1gpu.err
for ok in 2 3 4 5; do python bench.py --workload box01 --ok $ok --no-cpu --no-e2e > gpurun_out/a_bench_box01_ok$ok.json 2>> gpurun_out/a_bench_1gpu.err; done
# 2. ncu launch list of the bench command (share of the step per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/a_launches_bench.log 2>&1
# 3. ncu --set full of the dominant kernel and of the PCG vector kernels
tools/ncu_capture.sh a_mass3d 'mass3d' --op pcg --reps 1
tools/ncu_capture.sh a_update_r 'update_r' --op pcg --reps 1
tools/ncu_capture.sh a_update_dx 'update_dx' --op pcg --reps 1
tools/ncu_capture.sh a_l2inv_apply 'l2inv_apply' --op cgl2 --reps 2
tools/ncu_capture.sh a_force3d '^force3d' --op force --reps 2
tools/ncu_capture.sh a_forcet3d 'forcet3d' --op forcet --reps 2
tools/ncu_capture.sh a_qupdate3d 'qupdate3d' --op q --reps 2
# 4. sanitizers at this commit
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/a_sanitize_memcheck.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py --quick > gpurun_out/a_sanitize_racecheck.txt 2>&1
tail -4 gpurun_out/a_sanitize_memcheck.txt gpurun_out/a_sanitize_racecheck.txt
ls -la gpurun_out | tail -30
