mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4; do
echo "== weak $n"; MLG_BENCH_SKIP_CPU=1 timeout 400 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r4q_bench_${n}gpu_weak.json 2> gpurun_out/r4q_weak_$n.err; tail -1 gpurun_out/r4q_weak_$n.err | cut -c1-200
done
echo "== reference arm under torchrun (rank 0 only)"; timeout 400 $TR --nproc-per-node 2 --master-port 29539 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 2>/dev/null | cut -c1-300
python - <<'PY'
import json
for n in (2,4):
    try:
        txt=open("gpurun_out/r4q_bench_%dgpu_weak.json"%n).read()
        print("stdout lines:", len(txt.strip().splitlines()))
        d=json.loads(txt.strip().splitlines()[-1])
        print(n, "value %.1f G ms %.3f K1 %.3f nonprobe %.3f e2e %.1f G (%.2f ms); %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["non_probe_ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["config"]["parallelism"][:80]))
    except Exception as e: print(n, "failed", e)
PY
