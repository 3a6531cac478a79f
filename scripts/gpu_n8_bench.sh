mkdir -p gpurun_out
for i in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_8.json 2> gpurun_out/r2_scale_8.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_scale_8.json').read().strip().splitlines()[-1])
print('N=8 value %.0f Mrays/s  ms %.3f | e2e %.0f ms %.3f | bands %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['band_rows']))
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_4.json 2> gpurun_out/r2_scale_4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_2.json 2> gpurun_out/r2_scale_2.err
python - <<PY
import json
for n in (2,4):
    d=json.loads(open('gpurun_out/r2_scale_%d.json'%n).read().strip().splitlines()[-1])
    print('N=%d value %.0f Mrays/s  ms %.3f | e2e %.0f ms %.3f | bands %s' % (n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['band_rows']))
PY
