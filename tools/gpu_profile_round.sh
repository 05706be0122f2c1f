#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu --set full of the apply kernels.
# usage (from the repo root, under gpurun):  bash tools/gpu_profile_round.sh <tag>
# Costs ~12 GPU-minutes on one B200 (ncu replays every captured kernel ~40 times); the two ncu passes can be run alone.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 3000 $out/${tag}_bench.json
# launch list of one timed bond update (cold-cache, serialised: compare SHARES) -> tools/summarize_launches.py
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-region \
    > $out/${tag}_ncu_list.log 2>&1
# --set full of the kernels of two H_eff applies (the first 9 matching launches belong to make_phi / position)
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"gemm_tma_ws_kernel|gemm_kernel|transform_kernel|relayout_kernel" --launch-skip 9 -c 10 -f -o $out/${tag}_apply \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-region > $out/${tag}_ncu_apply.log 2>&1
# Krylov vector kernels
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"dot_kernel|axpy_kernel|axpy_dot_kernel|lincomb_kernel|scale_kernel" --launch-skip 4 -c 8 -f -o $out/${tag}_vec \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-region > $out/${tag}_ncu_vec.log 2>&1
ls -la $out
# then, here:  python tools/summarize_launches.py gpurun_out/${tag}_launches.csv ; python tools/ncu_summary.py gpurun_out/${tag}_apply.ncu-rep
