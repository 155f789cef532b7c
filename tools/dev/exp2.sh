# usage (under gpurun): QLIBS="vB.so vC.so" C5LIBS="vB.so vC.so vD.so" C1LIBS="vB.so" bash tools/dev/exp2.sh
mkdir -p gpurun_out
for lib in $QLIBS; do
  ZMESH_B200_LIB=$PWD/build_ab/$lib timeout 120 python tools/quick_check.py > gpurun_out/quick_$lib.log 2>&1; echo "quick $lib rc=$?"; tail -1 gpurun_out/quick_$lib.log
done
bash tools/dev/ab.sh "c5" $C5LIBS
bash tools/dev/ab.sh "c1" $C1LIBS
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_v*_c5.json')):
  d=json.load(open(f)); print(f, d['config'].get('tiles'))
P
