"""Schema fixture: every message / field / enum of the reference's .proto SOURCES (ffn/inference/inference.proto,
ffn/utils/bounding_box.proto, ffn/utils/vector.proto — the authoritative schema; the shipped *_pb2.py are stale and
protoc is not available), extracted with a small proto2 text parser.

    python tests/golden/make_golden_schema.py   ->   proto_schema_ref.json

{"package.Message": {"fields": {name: [number, label, type, default|null, oneof|null]}, "enums": {...}}, ...}
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/ffn'
FILES = ['inference/inference.proto', 'utils/bounding_box.proto', 'utils/vector.proto']

TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[A-Za-z_][\w.]*|-?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?|[{}=;\[\],]')


def tokens(text):
  text = re.sub(r'//[^\n]*', '', text)
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return TOKEN.findall(text)


def parse(path):
  toks = tokens(open(path).read())
  pos = 0
  package = ''
  out = {}

  def block(prefix):
    nonlocal pos
    name = toks[pos]
    pos += 1
    assert toks[pos] == '{', (name, toks[pos])
    pos += 1
    full = prefix + '.' + name if prefix else name
    msg = {'fields': {}, 'enums': {}}
    out[full] = msg
    body(full, msg, None)

  def enum(prefix, holder):
    nonlocal pos
    name = toks[pos]
    pos += 2                      # name {
    values = {}
    while toks[pos] != '}':
      vname = toks[pos]
      assert toks[pos + 1] == '='
      values[vname] = int(toks[pos + 2])
      pos += 3
      while toks[pos] != ';':
        pos += 1
      pos += 1
    pos += 1
    holder[name] = values

  def body(full, msg, oneof):
    nonlocal pos
    while toks[pos] != '}':
      t = toks[pos]
      if t == 'message':
        pos += 1
        block(full)
      elif t == 'enum':
        pos += 1
        enum(full, msg['enums'])
      elif t == 'oneof':
        name = toks[pos + 1]
        pos += 3                  # oneof name {
        body(full, msg, name)
      elif t in ('reserved', 'option', 'extensions'):
        while toks[pos] != ';':
          pos += 1
        pos += 1
      elif t == ';':
        pos += 1
      else:
        label = 'optional'
        if t in ('optional', 'repeated', 'required'):
          label = t
          pos += 1
        ftype = toks[pos]
        fname = toks[pos + 1]
        assert toks[pos + 2] == '=', (full, ftype, fname, toks[pos:pos + 4])
        number = int(toks[pos + 3])
        pos += 4
        default = None
        if toks[pos] == '[':
          while toks[pos] != ']':
            if toks[pos] == 'default':
              default = toks[pos + 2].strip('"')
            pos += 1
          pos += 1
        assert toks[pos] == ';', (full, fname, toks[pos])
        pos += 1
        msg['fields'][fname] = [number, label, ftype, default, oneof]
    pos += 1

  while pos < len(toks):
    t = toks[pos]
    if t == 'package':
      package = toks[pos + 1]
      pos += 3
    elif t == 'message':
      pos += 1
      block(package)
    elif t == 'enum':
      pos += 1
      holder = out.setdefault(package + '.<file>', {'fields': {}, 'enums': {}})['enums']
      enum(package, holder)
    else:
      while pos < len(toks) and toks[pos] != ';':
        pos += 1
      pos += 1
  return out


def main():
  schema = {}
  for rel in FILES:
    schema.update(parse(os.path.join(REF, rel)))
  n_fields = sum(len(m['fields']) for m in schema.values())
  json.dump(schema, open(os.path.join(HERE, 'proto_schema_ref.json'), 'w'), indent=1, sort_keys=True)
  print('wrote proto_schema_ref.json: %d messages, %d fields' % (len(schema), n_fields))
  print(sorted(schema)[:40])


if __name__ == '__main__':
  sys.exit(main())
