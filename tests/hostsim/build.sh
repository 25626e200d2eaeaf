#!/bin/bash
# TEST INFRASTRUCTURE: host build of the device math headers for the CPU-only parity tests.
set -e
cd "$(dirname "$0")"
mkdir -p _build
if [ ! -f _build/libhostsim.so ] || [ hostsim.cc -nt _build/libhostsim.so ] || \
   [ -n "$(find ../../airiceraytracing_b200/csrc -newer _build/libhostsim.so -name '*.c*' -o -newer _build/libhostsim.so -name '*.hpp')" ]; then
  g++ -O2 -ffp-contract=off -fPIC -shared -x c++ -I../../airiceraytracing_b200/csrc hostsim.cc \
      ../../airiceraytracing_b200/csrc/atmosphere.cc -o _build/libhostsim.so
fi
