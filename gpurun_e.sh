for ctas in 17 24 32; do
  NPB_GLS_ARENA_GLOBAL=1 NPB_GLS_CTAS_PER_SM=$ctas python tools/run_once.py tet 69 gls 3 2>&1 | tail -1 | sed "s/^/global ctas=$ctas /"
done
python tools/run_once.py tet 69 gls 3 2>&1 | tail -1 | sed "s/^/smem /"
