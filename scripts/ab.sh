#!/bin/bash
# A/B on the GPU box: scripts/ab.sh cfg2 "VAR=1" "VAR=2 OTHER=3" ...  (one quick bench per environment setting)
wl=$1; shift
for envs in "$@"; do
  echo "--- $envs"
  env $envs scripts/quick.sh $wl auto
done
