"""Builds experiment variants of the engine library: tools/build_variants.py name:DEF=1,DEF2=3 ..."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ffn_b200 import build
os.makedirs(os.path.join(REPO, 'variants'), exist_ok=True)
for spec in sys.argv[1:]:
  name, _, defs = spec.partition(':')
  out = os.path.join(REPO, 'variants', 'libffn_b200_%s.so' % name)
  build.build(force=True, defines=[d for d in defs.split(',') if d], out=out)
  print(out)
