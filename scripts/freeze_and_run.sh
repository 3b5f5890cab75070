#!/bin/bash
# usage: scripts/freeze_and_run.sh <name> <timeout_s> <script relative to the repo root> [gpus]
# Copies the working tree into frozen_<name>/ (git-ignored) and runs the script from there on the GPU box, retrying
# while the pod answers "busy": the live tree can keep changing while the call waits for a GPU slot.
name=$1; tmo=$2; script=$3; gpus=${4:-1}
cd /root/repo
rm -rf frozen_$name
mkdir -p frozen_$name && tar -cf - --exclude='./frozen_*' --exclude=./gpurun_out --exclude=./.git --exclude=./ab_r1 --exclude='__pycache__' --exclude=.pytest_cache --exclude=.hypothesis . | tar -xf - -C frozen_$name
[ -d ab_r1 ] && ln -sfn ../ab_r1 frozen_$name/ab_r1
ln -sfn ../gpurun_out frozen_$name/gpurun_out
extra=""
[ "$gpus" != "1" ] && extra="--gpus $gpus"
for try in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $extra --timeout $tmo -- "cd frozen_$name && bash $script" > /tmp/gpurun_$name.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 45
done
echo "gpurun rc=$rc after $try tries" >> /tmp/gpurun_$name.log
rm -rf frozen_$name
