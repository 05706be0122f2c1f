mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err
tail -2 gpurun_out/r02n_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','phase_ms_per_step','energies','sweep_time_s','step_ms_list','phase_ms_max_over_steps')})
print(d['roofline'], d['e2e'], d['cpu_baseline'], d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
