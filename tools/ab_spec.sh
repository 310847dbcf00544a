# A/B of kernel switches on one box: parity first, then interleaved bench lines (ms/step, samples/s, roofline, detail)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc64_gpu.py -x -q 2>&1 | tail -3
line() { python bench.py --steps 50 --warmup 5 --skip-cpu 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_detail_ms_per_step']
print(d['ms_per_step'], d['value'], d['roofline']['frac'], 'wgrad', k.get('conv_tc64_wgrad:g.res.wgrad'), 'reduce', k.get('wgrad_reduce_tc:g.res.wgrad'), 'bnred', k.get('conv_tc64_fprop:g.res.dgrad_bnred'), d['clocks']['sm_mhz'])
"; }
for v in 2048 0 2048 0; do echo "VARIANT=$v"; PCG_TC64_VARIANT=$v line; done
