# Round-2 evidence run on ONE B200 (under gpurun): tests, bench lines, ncu launch list + full captures.  Outputs -> gpurun_out/r02/
set -x
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -q -m gpu > $O/gputest.log 2>&1; tail -3 $O/gputest.log
python bench.py > $O/bench_c3.json 2> $O/bench.err
python bench.py --workload c2 --no-cpu-baseline > $O/bench_c2.json 2>> $O/bench.err
python bench.py --workload c3 --boundary same --no-cpu-baseline --no-extras > $O/bench_c3_same.json 2>> $O/bench.err
python bench.py --workload c1 --no-cpu-baseline > $O/bench_c1.json 2>> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c3_reference.json 2>> $O/bench.err
python scripts/bench_next_rows.py > $O/next_rows_bench.json 2>> $O/bench.err
python scripts/direct_vs_fft.py > $O/direct_vs_fft.txt 2>> $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rows_fwd_pipe|cols_pipe|cols_fast|rows_inv_pipe" -s 10 -c 5 -o $O/prof_c3 python scripts/profile_c3.py reference 4 > $O/ncu_full.log 2>&1
ncu -i $O/prof_c3.ncu-rep --page raw --csv > $O/prof_c3_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"direct_conv_cubic" -s 2 -c 1 -o $O/prof_direct python scripts/profile_c3.py same 4 512x512x400 5x5x5 > $O/ncu_direct.log 2>&1
ncu -i $O/prof_direct.ncu-rep --page raw --csv > $O/prof_direct_raw.csv 2>/dev/null
python -c "from pyvoxeldosimetry_b200._capi import get_lib; print(get_lib().build_id())" > $O/build_id.txt
rm -f $O/prof_c3.ncu-rep.tmp
ls -la $O
