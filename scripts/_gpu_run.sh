mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sketch.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2b_sketch_tests.log
cat gpurun_out/r2b_sketch_tests.log
for i in 1 2; do timeout 600 python scripts/sketch_bench.py 2>&1 | tail -1; done > gpurun_out/r2b_sketch_bench.jsonl
cat gpurun_out/r2b_sketch_bench.jsonl | cut -c1-700
