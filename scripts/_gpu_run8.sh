mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r4d_topo.txt 2>&1; nproc >> gpurun_out/r4d_topo.txt; free -g >> gpurun_out/r4d_topo.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r4d_topo.txt
echo "== h2d"; : > gpurun_out/r4d_h2d.jsonl
for n in 1 2 4 8; do timeout 200 $TR --nproc-per-node $n --master-port 2951$n scripts/diag_h2d_multi.py 2>/dev/null | grep '^{' >> gpurun_out/r4d_h2d.jsonl; done; cat gpurun_out/r4d_h2d.jsonl | cut -c1-300
echo "== weak 8"; MLG_BENCH_SKIP_CPU=1 timeout 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r4d_bench_8gpu_weak.json 2> gpurun_out/r4d_weak.err; tail -2 gpurun_out/r4d_weak.err | cut -c1-300
echo "== strong 8 (configs[2], parity)"; timeout 600 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --workload strong > gpurun_out/r4d_bench_8gpu_strong.json 2> gpurun_out/r4d_strong.err; tail -2 gpurun_out/r4d_strong.err | cut -c1-300
echo "== stream 8 (configs[3])"; MLG_BENCH_SKIP_CPU=1 timeout 600 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --workload stream > gpurun_out/r4d_bench_8gpu_stream.json 2> gpurun_out/r4d_stream.err; tail -2 gpurun_out/r4d_stream.err | cut -c1-300
echo "== multi tests"; timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r4d_multi_tests.log 2>&1; tail -3 gpurun_out/r4d_multi_tests.log
python - <<'PY'
import json
for w in ("weak","strong","stream"):
    try:
        d=json.load(open("gpurun_out/r4d_bench_8gpu_%s.json"%w))
        print(w, "value %.1f G ms %.3f K1 %.3f nonprobe %.3f e2e %.1f G (%.2f ms) parity %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["non_probe_ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d.get("parity_checked")))
    except Exception as e: print(w, "failed", e)
PY
