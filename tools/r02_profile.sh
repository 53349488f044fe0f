# round-2 profile captures (one GPU): launch lists + one `--set full` capture per hot kernel; outputs under gpurun_out/
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the default bench command (device times are cold-cache and serialised: shares, not absolutes)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --replay-blocks 2048 > gpurun_out/r02_bench_under_ncu.log 2>&1
# 2. the verify kernel
timeout 300 $NCU --set full --import-source on -k regex:k_schnorr_verify -s 1 -c 1 -o gpurun_out/r02_schnorr -f python tools/prof_schnorr.py 1048576 2 > gpurun_out/r02_ncu_schnorr.log 2>&1
# 3. the UTXO table kernels
timeout 300 $NCU --set full --import-source on -k regex:k_utxo_ -c 6 -o gpurun_out/r02_utxo -f python tools/prof_utxo.py > gpurun_out/r02_ncu_utxo.log 2>&1
# 4. the replay window: sources, static rules, walk, finish kernels of the second 1024-block window
timeout 400 $NCU --set full --import-source on -k regex:k_replay_ -s 10 -c 10 -o gpurun_out/r02_replay -f python tools/prof_replay.py 1 1024 > gpurun_out/r02_ncu_replay.log 2>&1
for f in schnorr utxo replay; do
  ncu -i gpurun_out/r02_$f.ncu-rep --page raw --csv > gpurun_out/r02_${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | grep r02_
tail -2 gpurun_out/r02_ncu_schnorr.log gpurun_out/r02_ncu_utxo.log gpurun_out/r02_ncu_replay.log gpurun_out/r02_bench_under_ncu.log
