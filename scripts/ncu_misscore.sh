# ncu full-set capture of one misscore_kernel launch (296 consensus pairs = one CTA per resident
# slot at two CTAs per SM).  The probe first runs clean without ncu.
python tests/tools/misscore_probe.py --pairs 296 --cpu-pairs 1 > gpurun_out/misscore_probe_296.json 2> gpurun_out/misscore_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:misscore_kernel -s 1 -c 1 -o gpurun_out/misscore -f \
    python tests/tools/misscore_probe.py --pairs 296 --cpu-pairs 0 > gpurun_out/ncu_misscore.log 2>&1
cat gpurun_out/misscore_probe_296.json; tail -n 3 gpurun_out/ncu_misscore.log
