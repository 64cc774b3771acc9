#!/bin/bash
# Builds copies of the library with other register tiles / occupancy targets of count_rect_kernel
# into variants/ (for TAXI2_B200_LIB=... python tools/count_perf.py).  Usage: tools/build_count_variants.sh "4 2" "2 4" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/variants" /tmp/v
for cfg in "$@"; do
  set -- $cfg
  rm -rf /tmp/v/src && mkdir -p /tmp/v/src
  cp "$ROOT"/taxi2_b200/csrc/*.cu "$ROOT"/taxi2_b200/csrc/*.cuh "$ROOT"/taxi2_b200/csrc/*.cpp /tmp/v/src/
  sed -i -e "s/constexpr int COUNT_RX = [0-9]*;/constexpr int COUNT_RX = $1;/" \
         -e "s/__launch_bounds__(COUNT_TY, [0-9]*) count_rect_kernel/__launch_bounds__(COUNT_TY, $2) count_rect_kernel/" /tmp/v/src/count_planes.cuh
  sed -i "s#\"../../include/taxi2_b200.h\"#\"$ROOT/include/taxi2_b200.h\"#" /tmp/v/src/taxi_abi.cu /tmp/v/src/host_format.cpp
  (cd /tmp/v/src && nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -shared \
      -o "$ROOT/variants/lib_rx$1_b$2.so" taxi_abi.cu host_format.cpp -lcudart 2> /tmp/v/log_$1_$2.txt)
  echo "RX=$1 blocks=$2: $(grep -A2 count_rect_kernel /tmp/v/log_$1_$2.txt | grep -o 'Used [0-9]* registers') $(grep -A1 count_rect_kernel /tmp/v/log_$1_$2.txt | grep -o '[0-9]* bytes spill stores')"
done
