#!/bin/bash
# build.sh NAME SOURCE [nvcc flags...]: compile a development harness of this directory for sm_100a (output: ./NAME)
cd "$(dirname "$0")" || exit 1
name=$1; src=$2; shift 2
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I../../fast-cu-decision-hevc_b200/csrc "$@" "$src" -o "$name" -Xptxas -v 2>&1 | grep -E "error|spill" | grep -v " 0 bytes spill" | head
