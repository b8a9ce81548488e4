# MSA stage of the configs[1] batch (1000 windows) for several CTA shapes
for cfg in "512,8,2" "384,8,2" "256,10,2"; do NWIN=1000 CFGS="$cfg" timeout 200 python scripts/poa_probe.py 2>&1 | grep -E "cfg|warps in"; done
