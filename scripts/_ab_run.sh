for rep in 1 2; do
for v in base ra rb rc rd; do PVDOSE_LIB=build_variants/libpvdose_$v.so python scripts/ab_time.py $v c3same,c2,c2same >> gpurun_out/r02_ab_rows_small.jsonl 2>>gpurun_out/ab.err; done
done
