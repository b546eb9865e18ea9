# Round-2 multi-GPU evidence (under gpurun --gpus 8): bare link ceilings and the driver's bench line at N = 8 and N = 4.
set -x
O=gpurun_out/r02
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in ${NS:-8 4}; do
  export CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((N-1)))
  if [ -z "$SKIP_LINK" ]; then
    $TR --nproc-per-node $N --master-port 2960$N scripts/link_peak.py > $O/r02_link_peak_${N}gpu.json 2>> $O/multi.err
    cp $O/r02_link_peak_${N}gpu.json profiles/r02_link_peak_${N}gpu.json
  fi
  $TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3_${N}gpu.json 2>> $O/multi.err
done
tail -5 $O/multi.err
ls -la $O
