# A/B of GLS kernel builds: every ninpol_b200/_variants/lib_*.so takes the library's place for one timed run.
# usage: bash tools/gpurun/r02_variants.sh [KIND N REPEAT]
kind=${1:-tet}; n=${2:-100}; rep=${3:-3}
cp ninpol_b200/libninpol_b200.so /tmp/lib_keep.so
for v in ninpol_b200/_variants/lib_*.so; do
  cp $v ninpol_b200/libninpol_b200.so
  echo "== $v"
  RUN_ONCE_CHUNKS=1 timeout 300 python tools/run_once.py $kind $n gls $rep 2>&1 | grep -v Warning | cut -c1-200 | sed -e 's/load_mesh_process.*//' 
done
cp /tmp/lib_keep.so ninpol_b200/libninpol_b200.so
