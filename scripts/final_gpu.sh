# Round-2 evidence run on ONE B200 (under gpurun): tests, bench lines, ncu launch list + full captures.  Outputs -> gpurun_out/r02/
set -x
O=gpurun_out/r02
mkdir -p $O
python -m pytest tests -q -m gpu > $O/gputest.log 2>&1; tail -3 $O/gputest.log
BID=$(python -c "from pyvoxeldosimetry_b200._capi import get_lib; print(get_lib().build_id())")
K='regex:rows_fwd|cols_pipe|cols_fast|rows_inv'
# launches 1-3 = the spectrum build of set_kernel, 4-13 = two warm executes, 14-18 = the five passes of the third execute
ncu --set full --clock-control none --import-source on -k "$K" -s 13 -c 5 -o $O/prof_c3 python scripts/profile_c3.py reference 4 > $O/ncu_full.log 2>&1
ncu -i $O/prof_c3.ncu-rep --page raw --csv > $O/prof_c3_raw.csv 2>/dev/null
python scripts/ncu_traffic.py $O/prof_c3_raw.csv $O/dram_traffic_c3.json "ncu --set full --clock-control none, C3 reference mode, the five launches of one execute (scripts/final_gpu.sh)" $BID > /dev/null
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k "$K" -s 13 -c 5 -o $O/prof_c3same python scripts/profile_c3.py same 4 > $O/ncu_c3same.log 2>&1
ncu -i $O/prof_c3same.ncu-rep --page raw --csv > $O/prof_c3same_raw.csv 2>/dev/null
python scripts/ncu_traffic.py $O/prof_c3same_raw.csv $O/dram_traffic_c3_same.json "ncu --metrics dram bytes + duration, C3 'same' mode, the five launches of one execute" $BID > /dev/null
ncu --metrics $M --clock-control none -k "$K" -s 13 -c 5 -o $O/prof_c2 python scripts/profile_c3.py reference 4 256x256x256 31x31x31 4 0 > $O/ncu_c2.log 2>&1
ncu -i $O/prof_c2.ncu-rep --page raw --csv > $O/prof_c2_raw.csv 2>/dev/null
python scripts/ncu_traffic.py $O/prof_c2_raw.csv $O/dram_traffic_c2_reference.json "ncu --metrics dram bytes + duration, C2 (256^3, 4 time points, 31^3), the five launches of one execute" $BID > /dev/null
# bench.py reports roofline.traffic only from a capture of the build it runs: put the fresh captures where it looks
cp $O/dram_traffic_c3.json profiles/r02_dram_traffic_c3.json; cp $O/dram_traffic_c3_same.json profiles/r02_dram_traffic_c3_same.json; cp $O/dram_traffic_c2_reference.json profiles/r02_dram_traffic_c2_reference.json
python bench.py > $O/bench_c3.json 2> $O/bench.err
python bench.py --workload c2 --no-cpu-baseline > $O/bench_c2.json 2>> $O/bench.err
python bench.py --workload c3 --boundary same --no-cpu-baseline --no-extras > $O/bench_c3_same.json 2>> $O/bench.err
python bench.py --workload c1 --no-cpu-baseline > $O/bench_c1.json 2>> $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c3_reference.json 2>> $O/bench.err
python scripts/bench_next_rows.py > $O/next_rows_bench.json 2>> $O/bench.err
python scripts/direct_vs_fft.py > $O/direct_vs_fft.txt 2>> $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:direct_conv_cubic -s 2 -c 1 -o $O/prof_direct python scripts/profile_c3.py same 4 512x512x400 5x5x5 > $O/ncu_direct.log 2>&1
ncu -i $O/prof_direct.ncu-rep --page raw --csv > $O/prof_direct_raw.csv 2>/dev/null
python -c "from pyvoxeldosimetry_b200._capi import get_lib; print(get_lib().build_id())" > $O/build_id.txt
# gpurun copies at most 64 MiB back: keep the full C3 report, drop the others once their raw pages are extracted
rm -f $O/prof_c3.ncu-rep.tmp $O/prof_direct.ncu-rep $O/prof_c3same.ncu-rep $O/prof_c2.ncu-rep
du -sh gpurun_out
ls -la $O
