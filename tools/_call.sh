out=gpurun_out; tag=r02e; mkdir -p $out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sweep --no-decaying --profile-region \
    > $out/${tag}_ncu_list.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"gemm_tma_ws_kernel|transform_kernel|relayout_kernel" --launch-skip 9 -c 12 -f -o $out/${tag}_apply \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sweep --no-decaying --profile-region > $out/${tag}_ncu_apply.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"dot_kernel|axpy_kernel|axpy_dot_kernel|lincomb_kernel|scale_kernel" --launch-skip 4 -c 8 -f -o $out/${tag}_vec \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sweep --no-decaying --profile-region > $out/${tag}_ncu_vec.log 2>&1
tail -3 $out/${tag}_ncu_list.log $out/${tag}_ncu_apply.log $out/${tag}_ncu_vec.log
ls -la $out | tail -12
