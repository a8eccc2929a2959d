#!/bin/bash
# tools/build_variant.sh <out.so> <extra nvcc flags...> : builds an experimental libdis_b200 variant (A/B kernel tuning)
set -e
out=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/depthinspace_b200/csrc
tmp=$(mktemp -d)
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -I $src -I $root/include $*"
pids=()
for f in api lcn misc smooth flow_warp flow_consistency conv3d_gather resize ext_misc point_loss; do nvcc $flags -c $src/$f.cu -o $tmp/$f.o & pids+=($!); done
for r in 0 1 2 3 4 5 6 7; do nvcc $flags -DDIS_R=$r -c $src/photometric_inst.cu -o $tmp/p$r.o & pids+=($!); done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $out $tmp/*.o -gencode arch=compute_100a,code=sm_100a
rm -rf $tmp
echo built $out
