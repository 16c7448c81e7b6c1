timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/r2_mgpu2b.log 2>&1; echo "mgpu_check rc=$?"; tail -14 gpurun_out/r2_mgpu2b.log
timeout 300 python -m pytest tests -m gpu -x -q -k "slab" 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --equil 300 --secondary-steps 20 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2b.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d.get('slab_check'))
for k in ('shear_4m','weak_1m_per_gpu'):
    if k in d: print(k, d[k]['ms_per_step'], d[k]['value'], d[k].get('slab_check'))
PY
