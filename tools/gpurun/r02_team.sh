# A/B: GLS with the warps of an SM in one CTA, loosely in step (NPB_GLS_TEAM=slack,spins) against one-warp CTAs (off)
run() {  # kind n chunks team
  export NPB_GLS_TEAM=$4
  echo "== $1 $2 chunks=$3 NPB_GLS_TEAM=$4"
  RUN_ONCE_CHUNKS=$3 timeout 120 python tools/run_once.py $1 $2 gls 3 2>&1 | grep -v Warning | sed -e "s/load_mesh_process.*'streamed_ms'/'streamed_ms'/" | cut -c1-250
}
for t in off 1,64 1,16 1,8; do run tet 100 1 $t; done
for t in off 1,16; do run tet 100 8 $t; done
for t in off 1,16 1,64; do run mixed 100 1 $t; done
for t in off 1,16; do run hex 128 1 $t; done
