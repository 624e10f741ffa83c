# A/B on one box: kNN flush / query as one out-of-line copy (current) vs inlined at every site (previous)
P="python profiles/pool_probe.py --no-launch-rate --pairs 4096 --steps 3"
for i in 1 2; do
  APD_LIB=$PWD/go-rio_b200/libapdgicp_prev.so timeout 300 $P 2>&1 | cut -c1-200
  timeout 300 $P 2>&1 | cut -c1-200
done
