mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_n8.json 2> gpurun_out/r02i_bench_n8.err
tail -5 gpurun_out/r02i_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench_n8.json').read().splitlines() if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','phase_ms_per_step','energies','parity','sweep_time_s')})
print(d.get('collectives_rank0_per_step'))
print(d.get('device_ms_per_step_by_phase_and_kernel_class'))
print(d.get('device_ms_per_step_by_rank'))
PY
