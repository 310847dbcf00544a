# A/B of kernel switches on one box: interleaved bench lines (ms/step, samples/s, e2e, roofline, launches)
mkdir -p gpurun_out
line() { python bench.py --steps 50 --warmup 5 --skip-cpu 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
k = d['kernel_detail_ms_per_step']
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['roofline']['frac'], 'wgrad', k.get('conv_tc64_wgrad:g.res.wgrad'), 'bnapply', k.get('bn_bwd_apply'), d['clocks']['sm_mhz'])
"; }
for cfg in 7 15 23 31 7 15 23 31; do echo "L2_HINTS=$cfg"; PCG_L2_HINTS=$cfg line; done
