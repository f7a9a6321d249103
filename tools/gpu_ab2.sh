#!/bin/bash
for v in "$@"; do
  if [ "$v" = default ]; then unset FFN_B200_LIB; else export FFN_B200_LIB=$PWD/variants/libffn_b200_$v.so; fi
  echo "== $v"
  timeout 200 python tools/profile_roles.py 2>&1 | grep -E '^\{|rror' | cut -c1-330
done
