# N-GPU validation: the 2-GPU parity test + the bench under torchrun (as the driver launches it)
N=${N:-2}
mkdir -p gpurun_out
(free -g; nproc; nvidia-smi topo -m | head -14; df -h /dev/shm | cat) > gpurun_out/box_n$N.txt 2>&1
[ -n "$SKIP_TEST" ] || python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
grep "^\[bench\]" gpurun_out/bench_n$N.err | head
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'n_gpus',d['n_gpus'],'frac',d['roofline']['frac'])
e=d['e2e']; print('e2e',e['value'],e['ms_per_step'],'ceiling',e.get('platform_ceiling'),'cold',e.get('cold'), e.get('rank_affinity'))
for k in ('strings','regex','chunk_index'):
    x=d.get(k)
    if x: print(k, json.dumps({kk:vv for kk,vv in x.items() if kk not in ('workload','roofline','cpu_baseline')})[:900])
PY
