#!/bin/bash
# last measurement batch of round 2 (run under gpurun): bench lines, ncu launch list, ncu --set full of the 64->64 kernels,
# reference arm, smoke, and the full GPU suite as the LAST action
R=${1:-r2i}
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${R}_native.json 2> gpurun_out/bench_${R}_native.err; tail -c 300 gpurun_out/bench_${R}_native.json
for w in cgan_moons simple_moons kc dcgan moons_cf cwgan; do python bench.py --workload $w --steps 300 --warmup 30 2>/dev/null | tail -1; done > gpurun_out/bench_${R}_other_configs.jsonl
cut -c1-160 gpurun_out/bench_${R}_other_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/ncu_launch_${R}.log 2>&1; tail -1 gpurun_out/ncu_launch_${R}.log | cut -c1-160
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc64 -s 40 -c 8 -o gpurun_out/prof_tc64_${R} -f python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/ncu_full_${R}.log 2>&1; tail -1 gpurun_out/ncu_full_${R}.log | cut -c1-160
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${R}_reference_arm.json; cut -c1-300 gpurun_out/bench_${R}_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r2_full.log 2>&1; tail -3 gpurun_out/gputest_r2_full.log
