#!/bin/bash
# usage: tools/sweep.sh <label:lib|-> ...   (lib relative to vk-renderer_b200/, "-" = the default build)
# one bench_brief line per variant; extra bench args via SWEEP_ARGS
for v in "$@"; do
  label=${v%%:*}; lib=${v#*:}
  if [ "$lib" = "-" ]; then unset VKRT_LIB; else export VKRT_LIB=$PWD/vk-renderer_b200/$lib; fi
  tools/bench_brief.sh $label $SWEEP_ARGS
done
