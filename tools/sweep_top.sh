set -e
for mb in 2 3 4; do
  RPF_NVCC_EXTRA="-DRELABEL_MINB=$mb" python rp-tree_b200/_build.py --force > /dev/null 2>&1
  echo "RELABEL_MINB=$mb"; python tools/prof_step.py 4 2>&1 | tail -3 | cut -c1-600
done
