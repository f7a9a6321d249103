#!/bin/bash
# usage: tools/run_variants.sh "<profile_modes args>" variant1 variant2 ...   ('default' = the product build)
args="$1"; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset FFN_B200_LIB; else export FFN_B200_LIB=$PWD/variants/libffn_b200_$v.so; fi
  timeout 100 python tools/profile_modes.py $args 2>&1 | grep -E '^\{|Error|error' 
done
