# ncu captures for profiles/: (1) one full-occupancy launch of the persistent alignment kernel on
# realistic window-MSA graphs, (2) the launch list of a short bench run.
export NWIN=148 WORKERS=12 STREAMS=1 ED=0
python scripts/perf_probe.py > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:poa_persistent -s 45 -c 1 -o gpurun_out/dp_c2 -f python scripts/perf_probe.py > gpurun_out/ncu_c2.log 2>&1
tail -n 6 gpurun_out/plain_c2.log; tail -n 3 gpurun_out/ncu_c2.log
python bench.py --windows 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_b64.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b64.csv python bench.py --windows 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b64.log 2>&1
tail -c 300 gpurun_out/plain_b64.log; wc -l gpurun_out/launches_b64.csv
