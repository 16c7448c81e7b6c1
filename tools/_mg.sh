timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_bench_n8_full.json 2> gpurun_out/r2_bench_n8_full.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['usher']['trials_per_step'], d.get('slab_check'))
for k in ('n1_same_workload','weak_1m_per_gpu'):
    if k in d: print(k, d[k]['ms_per_step'], d[k]['value'], d[k]['trials_per_step'], d[k].get('slab_check'))
print(d.get('e2e'))
PY
tail -3 gpurun_out/r2_bench_n8_full.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 2>/dev/null | cut -c1-300
