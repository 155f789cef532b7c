# round 2, first GPU session: A/B of ZM_S3_ATOMIC_RANK, then ncu --set full (source counters) of the current kernels
mkdir -p gpurun_out
ZMESH_B200_LIB=$PWD/build_ab/vAR.so timeout 120 python tools/quick_check.py > gpurun_out/quick_vAR.log 2>&1; echo "quick vAR rc=$?"; tail -1 gpurun_out/quick_vAR.log
bash tools/dev/ab.sh "c5 c1" base.so vAR.so
bash tools/dev/prof.sh r02base c5s c1
