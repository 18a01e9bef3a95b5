#!/bin/bash
# tools/ab_build.sh <name> [extra nvcc flags...]: builds the f16 variant of the working tree into audio-to-midi_b200/_build/ab/<name>.so
# (A/B experiments: `A2M_LIB_F16=<path> python bench.py ...` loads it instead of the in-tree library)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p audio-to-midi_b200/_build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DA2M_OP_F16 "$@" \
  audio-to-midi_b200/csrc/a2m_api.cu audio-to-midi_b200/csrc/modelutil.cpp -ldl -o audio-to-midi_b200/_build/ab/$name.so
ls -la audio-to-midi_b200/_build/ab/$name.so
