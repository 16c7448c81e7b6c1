timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2c.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['usher']['trials_per_step'], d.get('slab_check'))
for k in ('n1_same_workload','shear_4m','weak_1m_per_gpu'):
    if k in d: print(k, d[k]['ms_per_step'], d[k]['value'], d[k]['trials_per_step'])
PY
tail -3 gpurun_out/r2_bench_n2c.err
