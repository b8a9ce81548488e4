export C1=1 NWIN=148 DEPTH=4 WORKERS=1 LANE_JOBS=148 INFLIGHT=148 ED=0 POA_THREADS=512
python scripts/perf_probe.py > gpurun_out/plain_512.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:poa_dp -s 6 -c 1 -o gpurun_out/dp512 -f python scripts/perf_probe.py > gpurun_out/ncu_512.log 2>&1
export POA_THREADS=256 NWIN=296 LANE_JOBS=296 INFLIGHT=296
python scripts/perf_probe.py > gpurun_out/plain_256.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:poa_dp -s 6 -c 1 -o gpurun_out/dp256 -f python scripts/perf_probe.py > gpurun_out/ncu_256.log 2>&1
tail -4 gpurun_out/plain_512.log gpurun_out/plain_256.log; tail -3 gpurun_out/ncu_512.log gpurun_out/ncu_256.log; ls -la gpurun_out
