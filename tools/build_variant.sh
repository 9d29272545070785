#!/bin/bash
# usage: build_variant.sh <name> <extra nvcc flags...>  -> build/variants/lib_<name>.so
set -e
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -o build/variants/lib_$name.so contouring_uncertainty_b200/csrc/*.cu 2>&1 | grep -v "deprecated" || true
