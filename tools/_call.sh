mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py 2>&1 | tail -9
timeout 300 python tools/_gemm_shapes.py 2>&1 | grep "65536)\|, 256, 0, 0)\|, 512, 0, 0)" | head -8
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --workload ttn_tfi --chi 256 --steps 4 --warmup 3 > gpurun_out/r02p_ttn_chi256.json 2> gpurun_out/r02p_ttn.err
timeout 900 python bench.py --workload hubbard_tdvp --chi 2048 --steps 3 --warmup 3 > gpurun_out/r02p_hubbard_chi2048.json 2> gpurun_out/r02p_hubbard.err
python - <<'PY'
import json
for f in ('ttn_chi256','hubbard_chi2048'):
    try:
        d=json.loads([l for l in open('gpurun_out/r02p_%s.json'%f).read().splitlines() if l.startswith('{')][-1])
        print(f, d.get('value'), d.get('ms_per_step'), d.get('phase_ms_per_step'), d['roofline']['frac'], d.get('device_ms_per_step_by_kernel_class'), d.get('apply_gflop'), d.get('applies_per_step'))
    except Exception as e:
        print(f, 'failed', e)
PY
