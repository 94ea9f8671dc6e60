#!/bin/bash
# Full SASS listing of every kernel (cuobjdump -sass of each object of the current build, encodings stripped), gzipped under profiles/<round>_sass/,
# plus the per-kernel mnemonic summary.  usage: tools/dump_sass.sh r02     (runs on the CPU box after `python nerf-vo_b200/build.py`)
tag=${1:-r02}
out=profiles/${tag}_sass
mkdir -p $out
for o in build/nvo_b200/*.o; do
  f=$(basename $o .o)
  cuobjdump -sass $o | grep -v '^\s*/\* 0x' | grep -v '^\s*$' | c++filt | gzip -9 > $out/$f.sass.gz
done
python tools/sass_summary.py $tag
ls -la $out
