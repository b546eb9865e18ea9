set -x
mkdir -p gpurun_out/f
python bench.py > gpurun_out/f/bench_c3.json 2> gpurun_out/f/bench_c3.err
python bench.py --workload c2 > gpurun_out/f/bench_c2.json 2>> gpurun_out/f/bench_c3.err
python bench.py --workload c3 --boundary same --no-cpu-baseline > gpurun_out/f/bench_c3_same.json 2>> gpurun_out/f/bench_c3.err
python bench.py --workload c1 --no-cpu-baseline > gpurun_out/f/bench_c1.json 2>> gpurun_out/f/bench_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f/bench_c3_reference.json 2>> gpurun_out/f/bench_c3.err
python scripts/bench_next_rows.py > gpurun_out/f/next_rows_bench.json 2>> gpurun_out/f/bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rows_fwd_pipe|cols_pipe|cols_fast|rows_inv_pipe" -s 10 -c 5 -o gpurun_out/f/prof_r1f python scripts/profile_c3.py reference 4 > gpurun_out/f/ncu_full.log 2>&1
ncu -i gpurun_out/f/prof_r1f.ncu-rep --page raw --csv > gpurun_out/f/prof_r1f_raw.csv 2>/dev/null
ls -la gpurun_out/f
