mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "two_gpus" 2>&1 | tail -6
for ar in 1 0; do
TNL_PEER_SCALAR_AR=$ar timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$ar bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-sweep --no-decaying > gpurun_out/r02m_bench_n2_ar$ar.json 2> gpurun_out/r02m_bench_n2_ar$ar.err
tail -2 gpurun_out/r02m_bench_n2_ar$ar.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02m_bench_n2_ar$ar.json').read().splitlines() if l.startswith('{')][-1])
print("peer scalar AR=$ar", {k:d.get(k) for k in ('value','ms_per_step','phase_ms_per_step','energies','parity','step_ms_list')})
print(d.get('collectives_rank0_per_step'))
PY
done
