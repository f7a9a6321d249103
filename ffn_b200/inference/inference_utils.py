"""Counters and timers with the reference's names and semantics.

Mirrors ffn/inference/inference_utils.py:32-199 (StatCounter, Counters, timer_counter, TimedIter):
the counter names written to counters.txt and embedded in result files are part of the public
surface.  Histogram-matching helpers (CLAHE) are outside the inference hot path and not provided.
"""

import contextlib
import json
import threading
import time

MSEC_IN_SEC = 1000


class StatCounter:
  """Counter with the MR-style interface (Increment / IncrementBy / Set / value)."""

  def __init__(self, update, name, parent=None):
    self._counter = 0
    self._update = update
    self._lock = threading.Lock()
    self._parent = parent
    self._name = name

  def Increment(self):  # pylint: disable=invalid-name
    self.IncrementBy(1)

  def IncrementBy(self, x, export=True):  # pylint: disable=invalid-name
    del export
    with self._lock:
      self._counter += int(x)
      self._update()
    if self._parent is not None:
      self._parent.IncrementBy(x)

  def Set(self, x, export=True):  # pylint: disable=invalid-name
    self.IncrementBy(x - self._counter, export=export)

  def __repr__(self):
    return 'StatCounter(total=%g)' % self.value

  @property
  def value(self):
    return self._counter


class Counters:
  """Container of named counters; `get_sub_counters` children propagate into their parent."""

  def __init__(self, parent=None):
    self._lock = threading.Lock()
    self.reset()
    self.parent = parent

  def reset(self):
    with self._lock:
      self._counters = {}
    self._last_update = 0

  def __getitem__(self, name):
    return self.get(name)

  def get(self, name, **kwargs):
    with self._lock:
      if name not in self._counters:
        self._counters[name] = self._make_counter(name, **kwargs)
      return self._counters[name]

  def __iter__(self):
    return iter(self._counters.items())

  def _make_counter(self, name, **kwargs):
    del kwargs
    # The reference intends sub-counters to feed their parent (inference_utils.py:133-134) but
    # never links them; here the link is made so Runner.counters aggregates its canvases.
    parent = self.parent.get(name) if self.parent is not None else None
    return StatCounter(self.update_status, name, parent)

  def update_status(self):
    pass

  def get_sub_counters(self):
    return Counters(self)

  def dump(self, filename):
    from . import storage
    with storage.atomic_file(filename, 'w') as fd:
      for name, counter in sorted(self._counters.items()):
        fd.write('%s: %d\n' % (name, counter.value))

  def dumps(self):
    return json.dumps({name: c.value for name, c in self._counters.items()})

  def loads(self, encoded_state):
    for name, value in json.loads(encoded_state).items():
      self[name].Set(value, export=False)


@contextlib.contextmanager
def timer_counter(counters, name, export=True, increment=1):
  """Adds `<name>-calls` and `<name>-time-ms` around the context."""
  assert isinstance(counters, Counters)
  counter = counters.get(name + '-calls', export=export)
  timer = counters.get(name + '-time-ms', export=export)
  start = time.time()
  try:
    yield timer, counter
  finally:
    counter.IncrementBy(increment)
    timer.IncrementBy((time.time() - start) * MSEC_IN_SEC)


class TimedIter:
  """Iterator wrapper charging the time spent in `next` to a timer counter."""

  def __init__(self, it, counters, counter_name):
    self.it = it
    self.counters = counters
    self.counter_name = counter_name

  def __iter__(self):
    return self

  def __next__(self):
    with timer_counter(self.counters, self.counter_name):
      return next(self.it)

  def next(self):
    return self.__next__()
