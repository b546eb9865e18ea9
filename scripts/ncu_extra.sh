# Extra ncu captures of the final build (one B200): C3 'same' mode, the whole C5 volume on one GPU, the next-row kernels.
# Raw pages only (the reports are dropped: gpurun copies at most 64 MiB back).
O=gpurun_out/r02
mkdir -p $O
K='regex:rows_fwd|cols_pipe|cols_fast|rows_inv'
ncu --set full --clock-control none -k "$K" -s 13 -c 5 -o $O/x_c3same python scripts/profile_c3.py same 4 > $O/ncu_x.log 2>&1
ncu -i $O/x_c3same.ncu-rep --page raw --csv > $O/ncu_c3same_full_raw.csv 2>/dev/null; rm -f $O/x_c3same.ncu-rep
ncu --set full --clock-control none -k "$K" -s 13 -c 5 -o $O/x_c5 python scripts/profile_c3.py reference 4 1024x1024x800 51x51x51 >> $O/ncu_x.log 2>&1
ncu -i $O/x_c5.ncu-rep --page raw --csv > $O/ncu_c5_full_raw.csv 2>/dev/null; rm -f $O/x_c5.ncu-rep
ncu --set full --clock-control none -k 'regex:weighted_combine|ct_prepare|dvh_hist|roi_minmax|monoexp_fit' -c 60 -o $O/x_next python scripts/bench_next_rows.py >> $O/ncu_x.log 2>&1
ncu -i $O/x_next.ncu-rep --page raw --csv > $O/ncu_next_rows_full_raw.csv 2>/dev/null; rm -f $O/x_next.ncu-rep
tail -3 $O/ncu_x.log; ls -la $O | tail -5
