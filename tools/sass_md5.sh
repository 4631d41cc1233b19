#!/bin/bash
# md5 of the SASS instruction stream (addresses and encodings stripped) of every tcgen05 kernel instantiation in
# engine.o: a refactoring that must not touch a GPU-verified kernel is checked by comparing these before / after.
#   tools/sass_md5.sh [path/to/engine.o]
OBJ=${1:-$(dirname "$0")/../vae_npvc_b200/csrc/engine.o}
for f in $(cuobjdump -sass "$OBJ" | grep -oE "Function : \S*umma\S*" | awk '{print $3}' | sort); do
  n=$(cuobjdump -sass -fun "$f" "$OBJ" | grep -cE "^\s+/\*[0-9a-f]{4,6}\*/")
  m=$(cuobjdump -sass -fun "$f" "$OBJ" | grep -E "^\s+/\*[0-9a-f]{4,6}\*/" | sed -E 's#^\s+/\*[0-9a-f]{4,6}\*/\s+##; s#/\* 0x[0-9a-f]+ \*/##' | md5sum | cut -d' ' -f1)
  echo "$m  $n instr  $(echo $f | c++filt | sed 's/(.*//')"
done
