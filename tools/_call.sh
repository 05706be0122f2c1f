mkdir -p gpurun_out
for n in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_n$n.json 2> gpurun_out/r02o_bench_n$n.err
tail -2 gpurun_out/r02o_bench_n$n.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02o_bench_n$n.json').read().splitlines() if l.startswith('{')][-1])
print("N=$n", {k:d.get(k) for k in ('value','ms_per_step','phase_ms_per_step','energies','parity','sweep_time_s','step_ms_list')})
print(d.get('collectives_rank0_per_step'))
print(d.get('device_ms_per_step_by_phase_and_kernel_class',{}).get('solver'))
PY
done
