# DP time per alignment against resident CTAs per SM (plain runs; SVS_SM_LIMIT keeps the work on 8 SMs)
export SVS_SM_LIMIT=8 BODY=8000 DEPTH=16 ARENA_MB=6000
for cfg in "128,10,2 8" "128,10,2 16" "128,10,2 24" "128,10,2 32" "256,10,2 8" "256,10,2 16"; do
  set -- $cfg
  NWIN=$2 CFGS="$1" timeout 120 python scripts/poa_probe.py 2>&1 | grep -E "cfg|warps in" | sed "s/^/[NWIN=$2] /"
done
