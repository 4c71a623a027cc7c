mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== stress 8 (configs[4])"; MLG_BENCH_SKIP_CPU=1 timeout 600 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --steps 10 --warmup 3 --workload stress > gpurun_out/r4g_bench_8gpu_stress.json 2> gpurun_out/r4g_stress.err; tail -2 gpurun_out/r4g_stress.err | cut -c1-300
for m in dense sparse; do
echo "== weak 8 $m"; MLG_EXCHANGE=$m MLG_BENCH_SKIP_CPU=1 timeout 400 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r4g_bench_8gpu_weak_$m.json 2> gpurun_out/r4g_weak_$m.err; tail -1 gpurun_out/r4g_weak_$m.err | cut -c1-300
done
python - <<'PY'
import json
for w in ("stress","weak_dense","weak_sparse"):
    try:
        txt=open("gpurun_out/r4g_bench_8gpu_%s.json"%w).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(w, "value %.1f G ms %.3f K1 %.3f nonprobe %.3f e2e %.1f G (%.2f ms) build %.1f s; %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["non_probe_ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["config"]["db_build_s"], d["config"]["parallelism"][:90]))
    except Exception as e: print(w, "failed", e)
PY
