# ncu --set full (source counters) of the current default build on c5s and c1
bash tools/dev/prof.sh ${TAG:-r02k2} c5s c1
