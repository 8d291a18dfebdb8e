#!/bin/bash
# A/B: K1 pass time of an older checkout vs the current tree, same box, interleaved
for rep in 1 2; do
  for tree in scratch/old_ae73 .; do
    echo "== tree $tree"
    (cd $tree && NS=1e8,1.25e7 PYTHONPATH=$PWD timeout 100 python /root/repo/scratch/sweep_k1.py)
  done
done
