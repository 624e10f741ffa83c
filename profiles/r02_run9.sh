# round 2, ninth GPU pass: ncu of the loop kernel under load in the POOL configuration (64-register build, clusters of 4,
# 74 pairs = 296 CTAs = 2 per SM) and at 37 pairs (148 CTAs)
export APD_LM_CLUSTER=4 APD_LM_MINB=2
timeout 300 python profiles/multi_lm.py --jobs 74 --repeat 3
timeout 300 python profiles/multi_lm.py --jobs 37 --repeat 3
timeout 300 python profiles/multi_lm.py --jobs 148 --repeat 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 1 -c 1 -o gpurun_out/prof_lm_load74_r02 -f python profiles/multi_lm.py --jobs 74 --repeat 1 > gpurun_out/ncu_lm_load74.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 1 -c 1 -o gpurun_out/prof_lm_load37_r02 -f python profiles/multi_lm.py --jobs 37 --repeat 1 > gpurun_out/ncu_lm_load37.log 2>&1; echo rc=$?
