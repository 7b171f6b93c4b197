#!/bin/bash
# Stage the unmodified reference tree where the GPU box can see it: baseline/_ref/ is git-ignored
# (never part of the repository or its history) but travels with gpurun snapshots.
set -e
cd "$(dirname "$0")/../.."
mkdir -p baseline/_ref/reference
cp -r /root/reference/src /root/reference/scripts baseline/_ref/reference/
echo "staged $(find baseline/_ref/reference -name '*.py' | wc -l) reference files under baseline/_ref/reference"
