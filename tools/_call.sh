mkdir -p gpurun_out
for mode in 1 0; do
TNL_FUSED_RS=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --no-sweep --no-decaying > gpurun_out/r02k_bench_n8_fused$mode.json 2> gpurun_out/r02k_bench_n8_fused$mode.err
tail -3 gpurun_out/r02k_bench_n8_fused$mode.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02k_bench_n8_fused$mode.json').read().splitlines() if l.startswith('{')][-1])
print("fused=$mode", {k:d.get(k) for k in ('value','ms_per_step','phase_ms_per_step','parity','step_ms_list')})
print(d.get('collectives_rank0_per_step'))
print(d.get('device_ms_per_step_by_phase_and_kernel_class',{}).get('solver'))
print(d.get('phase_ms_max_over_steps'))
PY
done
