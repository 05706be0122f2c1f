mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02c_gpu.txt
tools/_build/eigh_bench > gpurun_out/r02c_eigh_bench.json 2> gpurun_out/r02c_eigh_bench.err
cat gpurun_out/r02c_eigh_bench.json
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02c_gputests.txt
cat gpurun_out/r02c_gputests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -c 1500 gpurun_out/r02c_bench.json
