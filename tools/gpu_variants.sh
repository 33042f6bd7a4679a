set -e
cp nori-ray-tracer_b200/csrc/libnori_gpu.so /tmp/orig.so
for mb in 7 8; do
  cp tools/libnori_gpu_mb$mb.so nori-ray-tracer_b200/csrc/libnori_gpu.so
  echo "minblocks $mb"; python tools/gpu_tp.py 1024 4194304 | tail -1
done
cp /tmp/orig.so nori-ray-tracer_b200/csrc/libnori_gpu.so
