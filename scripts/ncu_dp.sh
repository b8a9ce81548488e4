# ncu captures for profiles/ (one GPU): full-set profile of one launch of the persistent alignment
# kernel on configs[1] graphs (pruning on), after the same command ran clean without ncu.
export NWIN=148 DEPTH=12 WORKERS=12 STREAMS=1 ED=0
python scripts/perf_probe.py > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:poa_persistent -s 20 -c 1 -o gpurun_out/dp_c2_pruned -f python scripts/perf_probe.py > gpurun_out/ncu_c2.log 2>&1
tail -n 5 gpurun_out/plain_c2.log; tail -n 2 gpurun_out/ncu_c2.log
