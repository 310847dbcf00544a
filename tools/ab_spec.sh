# A/B of kernel switches on one box: parity first, then interleaved bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_tc64_gpu.py tests/test_mnist_step_gpu.py tests/test_mnist_eval.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
for cfg in 0 3 0 3; do echo "TC64_SPEC=$cfg"; PCG_TC64_SPEC=$cfg python bench.py --workload mnist_infer --steps 100 --warmup 10 2>/dev/null | tail -1 | cut -c1-260; done
python bench.py --steps 50 --warmup 5 --skip-cpu 2>/dev/null | tail -1 | cut -c1-200
