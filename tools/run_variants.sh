#!/bin/bash
# run a command once per variants/*.so (swapped in as the product library), then restore
L=nori-ray-tracer_b200/csrc/libnori_gpu.so
cp $L /tmp/orig.so
echo "== base"; eval "$1"
for v in variants/*.so; do
  cp $v $L; echo "== $v"; eval "$1"
done
cp /tmp/orig.so $L
