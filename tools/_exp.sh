timeout -s KILL 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile 2>&1 | tail -1 | cut -c1-330
timeout -s KILL 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -x -p no:cacheprovider 2>&1 | tail -3
