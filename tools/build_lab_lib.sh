#!/bin/sh
# Lab build of the engine (-DRT_LAB: exports rt_lab_skip, see rt_engine.cu) into tools/librtb200_lab.so -- timing experiments only.
set -e
cd "$(dirname "$0")/../pyradiotracking_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DRT_LAB ${RT_LAB_DEFS} -o ../../tools/librtb200_lab${1:+_$1}.so rt_engine.cu rt_matcher.cpp
