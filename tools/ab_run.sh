#!/bin/bash
# A/B: the headline device-resident figure (bench.py --device-only) for each libohs_cuda_<tag>.so given, and the default
# library.  usage: ab_run.sh tag1 tag2 ...
for tag in default "$@"; do
  if [ "$tag" = default ]; then unset OHS_LIB_OVERRIDE; else export OHS_LIB_OVERRIDE=$PWD/open-headstage_b200/libohs_cuda_$tag.so; fi
  for rep in 1 2; do
    echo -n "$tag: "; python bench.py --device-only --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'])"
  done
done
