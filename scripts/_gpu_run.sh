mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sketch.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r2b_sketch_tests.log
cat gpurun_out/r2b_sketch_tests.log
