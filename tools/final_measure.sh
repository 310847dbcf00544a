#!/bin/bash
# round-end measurement batch (run under gpurun): bench lines, ncu launch list, ncu --set full of the dominant kernel
R=${1:-r1e}
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${R}_native.json 2> gpurun_out/bench_${R}_native.err; tail -c 600 gpurun_out/bench_${R}_native.json
for w in cgan_moons simple_moons kc dcgan moons_cf cwgan; do python bench.py --workload $w --steps 300 --warmup 30 2>/dev/null | tail -1; done > gpurun_out/bench_${R}_other_configs.jsonl
PCG_WGAN_BREAKDOWN=0 python tools/bench_wgan.py 2>/dev/null | grep -v Warning > gpurun_out/bench_${R}_wgan_modes_and_torch_eager.txt
for w in mnist_infer mnist_loader mnist_clf_train kc_clf_train mnist_eval; do python bench.py --workload $w --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/bench_${R}_widened_rows.jsonl
python bench.py --precision fp32 --steps 20 --warmup 5 --skip-cpu 2>/dev/null | tail -1 > gpurun_out/bench_${R}_native_fp32.json
python tools/bench_graph_floor.py 444 > gpurun_out/graph_floor_${R}.txt 2>&1
python tools/bench_torch_eager.py > gpurun_out/bench_${R}_torch_eager_b200.jsonl 2>/dev/null
cut -c1-200 gpurun_out/bench_${R}_other_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/ncu_launch_${R}.log 2>&1; tail -2 gpurun_out/ncu_launch_${R}.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc64s_fprop -s 30 -c 3 -o gpurun_out/prof_tc64s_fprop_${R} -f python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/ncu_full_${R}.log 2>&1; tail -2 gpurun_out/ncu_full_${R}.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${R}_reference_arm.json; cut -c1-300 gpurun_out/bench_${R}_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
