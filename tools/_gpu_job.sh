set -x
for env in "SCTDA_NO_OVERLAP=1" "SCTDA_NO_OVERLAP=0"; do
for rk in 0 3; do
env $env python tools/mapper_rank.py $rk 8 2>&1 | tail -3
done
done
