# usage (under gpurun): VLIBS="vK2.so" BASE=base.so WLS="c5 c1" [FULLTEST=vK2.so] bash tools/dev/r2b.sh
mkdir -p gpurun_out
for lib in $VLIBS; do
  ZMESH_B200_LIB=$PWD/build_ab/$lib timeout 120 python tools/quick_check.py > gpurun_out/quick_$lib.log 2>&1; echo "quick $lib rc=$?"; tail -1 gpurun_out/quick_$lib.log
done
if [ -n "$FULLTEST" ]; then
  ZMESH_B200_LIB=$PWD/build_ab/$FULLTEST timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$FULLTEST.log 2>&1; echo "pytest $FULLTEST rc=$?"; tail -3 gpurun_out/pytest_gpu_$FULLTEST.log
fi
bash tools/dev/ab.sh "${WLS:-c5 c1}" $BASE $VLIBS
