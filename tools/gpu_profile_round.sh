#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu --set full of the dominant kernels.
# usage: tools/gpu_profile_round.sh <tag>
tag=${1:-r01b}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -c 6000 $out/${tag}_bench.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-region \
    > $out/${tag}_ncu_list.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"gemm_kernel" -c 3 -f -o $out/${tag}_gemm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-region > $out/${tag}_ncu_gemm.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"transform_kernel|dot_kernel|axpy_kernel|lincomb_kernel" --launch-skip 8 -c 10 -f -o $out/${tag}_hbm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-region > $out/${tag}_ncu_hbm.log 2>&1
ls -la $out
