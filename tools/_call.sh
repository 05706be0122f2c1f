mkdir -p gpurun_out
TNL_EIGH_DEBUG=1 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
grep "tnl" gpurun_out/r02h_bench.err | sort | uniq -c | sort -rn | head -12
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench.json').read().splitlines() if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','phase_ms_per_step','energies','sweep_time_s','truncation_decaying_spectrum')})
PY
for parts in "0" "4,3,2" "3,3,3" "5,4"; do
TNL_EIGH_PARTS=$parts timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sweep --no-decaying > gpurun_out/r02h_bench_p.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench_p.json').read().splitlines() if l.startswith('{')][-1])
print("$parts", d['ms_per_step'], d['phase_ms_per_step']['replacebond'])
PY
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
