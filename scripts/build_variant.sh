#!/bin/bash
# build an alternative libqgcm_b200.so for A/B runs (scripts/ab_bench.sh):
#   scripts/build_variant.sh spread "-DDST3_SPREAD=1"   ->  q-gcm_b200/csrc/alt/libqgcm_spread.so
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=/tmp/qgcm_variant_$name
rm -rf $tmp && mkdir -p $tmp/q-gcm_b200 $tmp/include
cp -r $root/q-gcm_b200/csrc $tmp/q-gcm_b200/ && cp $root/include/qgcm_b200.h $tmp/include/
cd $tmp/q-gcm_b200/csrc && rm -rf alt *.o *.so
make -s -j8 EXTRA="$extra"
mkdir -p $root/q-gcm_b200/csrc/alt
cp libqgcm_b200.so $root/q-gcm_b200/csrc/alt/libqgcm_$name.so
grep -A2 "${3:-k_dst3}" $tmp/q-gcm_b200/csrc/helmholtz.ptxas.log | grep -i "registers\|spill" | sort | uniq -c
