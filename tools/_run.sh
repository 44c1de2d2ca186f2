mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_elementwise_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py tests/test_trainer.py -m gpu -q -x 2>&1 | tail -4
for i in 1 2; do
 (cd _prev && timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PREV train', d['ms_per_step'], d['value'])")
 timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NEW train', d['ms_per_step'], d['value'])"
 (cd _prev && timeout 200 python bench.py --workload detect --no-cpu-baseline --no-e2e --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PREV detect', d['ms_per_step'], d['value'], d['stages']['forward']['ms'])")
 timeout 200 python bench.py --workload detect --no-cpu-baseline --no-e2e --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NEW detect', d['ms_per_step'], d['value'], d['stages']['forward']['ms'])"
done
