# ncu captures for profiles/ (one GPU).  Each profiled command first runs clean without ncu.
# (1) launch list of a short bench run (shares of the step per kernel)
python bench.py --windows 24 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_b24.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b24.csv python bench.py --windows 24 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b24.log 2>&1
tail -c 200 gpurun_out/plain_b24.log; wc -l gpurun_out/launches_b24.csv
# (2) full-set profile of one launch of the default persistent kernel (256 threads x 8 columns, two
#     CTAs per SM, pruning on): 296 configs[1] windows at depth 8+8, one stream -> 296 alignments per launch
export NWIN=296 DEPTH=8 WORKERS=12 STREAMS=1 ED=0
python scripts/perf_probe.py > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:poa_persistent -s 13 -c 1 -o gpurun_out/dp_final -f python scripts/perf_probe.py > gpurun_out/ncu_c2.log 2>&1
tail -n 4 gpurun_out/plain_c2.log; tail -n 2 gpurun_out/ncu_c2.log
