# usage (under gpurun): VLIB=vF.so BASE=vD2.so bash tools/dev/exp3.sh -- full pytest on a variant, reduced pytest on the default build, A/B
mkdir -p gpurun_out
ZMESH_B200_LIB=$PWD/build_ab/$VLIB timeout 60 python tools/quick_check.py > gpurun_out/quick_$VLIB.log 2>&1; echo "quick $VLIB rc=$?"
ZMESH_B200_LIB=$PWD/build_ab/$VLIB timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$VLIB.log 2>&1; echo "pytest $VLIB rc=$?"; tail -1 gpurun_out/pytest_gpu_$VLIB.log
timeout 150 python -m pytest tests -m gpu -x -q -k "not connectomics_full and not config4_full" > gpurun_out/pytest_gpu_default_reduced.log 2>&1; echo "pytest default (reduced) rc=$?"; tail -1 gpurun_out/pytest_gpu_default_reduced.log
bash tools/dev/ab.sh "c1 c5 c4" $BASE $VLIB
