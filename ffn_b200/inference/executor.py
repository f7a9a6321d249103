"""Executor: the operator interface the FFN model sits behind.

Mirrors ffn/inference/executor.py: `ExecutorInterface` (:46-82), `ExecutorClient` (:85-108),
`BatchExecutor` (:142-204) with `start_server/stop_server/get_client/num_devices`.  The
TensorFlow/JAX server thread and its request queue are gone: `B200Executor` owns one CUDA
engine and `B200ExecutorClient.predict` calls straight into `ffn_predict` (host buffers in,
host buffers out), serialised by a lock, so any caller of `ExecutorClient.predict(seed, image,
fetches)` keeps working.  `Canvas` additionally recognises this client and runs its whole
flood-fill loop on the device instead of calling `predict` per step.
"""

import threading
from typing import Sequence

import numpy as np

from .. import _lib
from .. import engine as engine_lib
from .. import tf_checkpoint
from . import inference_utils
from .errors import TerminationException
from .inference_utils import timer_counter


class ExecutorInterface:
  """Registration book-keeping shared by clients and the executor."""

  def __init__(self):
    self.lock = threading.Lock()
    self.outputs = {}
    self.exit_request = threading.Event()


class ExecutorClient:
  """Client interface for the FFN executor (executor.py:85-108)."""

  def __init__(self, counters: inference_utils.Counters, interface: ExecutorInterface):
    self._client_id = None
    self.counters = counters
    self._interface = interface

  def start(self) -> int:
    raise NotImplementedError()

  def finish(self):
    raise NotImplementedError()

  def predict(self, seed: np.ndarray, image: np.ndarray, fetches: Sequence[str]):
    raise NotImplementedError()


class B200ExecutorClient(ExecutorClient):
  """Client bound to one engine of a `B200Executor`; `predict` runs the conv stack on the GPU."""

  def __init__(self, counters, interface, executor, slot):
    super().__init__(counters, interface)
    self._executor = executor
    self._slot = slot

  @property
  def engine(self) -> engine_lib.Engine:
    return self._executor.engines[self._slot]

  @property
  def engine_lock(self):
    return self._executor.locks[self._slot]

  def start(self) -> int:
    with self._interface.lock:
      client_id = max(self._interface.outputs.keys()) + 1 if self._interface.outputs else 0
      self._interface.outputs[client_id] = None
    self._client_id = client_id
    self._executor.active_clients += 1
    return client_id

  def finish(self):
    if self._client_id is None:
      return
    with self._interface.lock:
      self._interface.outputs.pop(self._client_id, None)
    self._executor.active_clients -= 1
    self._client_id = None

  def predict(self, seed, image, fetches):
    if self._interface.exit_request.is_set():
      raise TerminationException()
    unknown = [f for f in fetches if f != 'logits']
    if unknown:
      raise KeyError('unsupported fetches: %r' % unknown)
    with timer_counter(self.counters, 'client-wait'):
      with self.engine_lock:
        logits = self.engine.predict(seed, image)
    return {'logits': logits[..., np.newaxis]}


class BatchExecutor:
  """Base class for FFN executors (executor.py:142-204)."""

  def __init__(self, interface, model, model_info, counters, batch_size):
    self._interface = interface
    self.model = model
    self.counters = counters
    self.batch_size = batch_size
    self.active_clients = 0
    self._input_seed_size = np.array(model_info.input_seed_size[::-1]).tolist()
    self._input_image_size = np.array(model_info.input_image_size[::-1]).tolist()
    self._pred_size = np.array(model_info.pred_mask_size[::-1]).tolist()

  def start_server(self):
    raise NotImplementedError()

  def stop_server(self):
    raise NotImplementedError()

  def get_client(self, subvol_counters):
    raise NotImplementedError()

  @property
  def num_devices(self):
    return 1


class B200Executor(BatchExecutor):
  """Owns the CUDA engine for one GPU.

  Args:
    interface: ExecutorInterface
    model: object with `.info` (ModelInfo, xyz), `.depth`, `.features` (e.g. ConvStack3DFFNModel)
    counters: Counters
    batch_size: kept for signature compatibility (requests are served as they arrive)
    checkpoint_path: TF1 bundle prefix (model.ckpt-N) or an .npz written by
      tf_checkpoint / tests; weights may instead be passed directly via `weights`/`biases`.
    device: CUDA device ordinal
    compute_mode: ffn_b200._lib.COMPUTE_FP16_TC (default) or COMPUTE_FP32
  """

  def __init__(self, interface, model, counters, batch_size=1, checkpoint_path=None, weights=None,
               biases=None, device=0, compute_mode=_lib.COMPUTE_FP16_TC):
    super().__init__(interface, model, model.info, counters, batch_size)
    if getattr(model, 'features', 32) != 32:
      raise ValueError('only 32 feature maps are supported')
    info = model.info
    if not (np.array_equal(info.pred_mask_size, info.input_seed_size) and
            np.array_equal(info.input_seed_size, info.input_image_size)):
      raise ValueError('pred / seed / image FoV sizes must be equal')
    if weights is None:
      if checkpoint_path is None:
        raise ValueError('a checkpoint path or explicit weights are required')
      with timer_counter(counters, 'restore-tf-checkpoint'):
        if checkpoint_path.endswith('.npz'):
          weights, biases = tf_checkpoint.load_convstack_npz(checkpoint_path)
        else:
          weights, biases = tf_checkpoint.load_convstack_weights(checkpoint_path, model.depth)
    # batch_size = number of canvases served concurrently (InferenceRequest.batch_size,
    # doc/manual.md:89-97).  The reference batches their FoVs into one session.run; here every
    # concurrent canvas gets its own engine with an equal share of the SMs, and the engines'
    # persistent kernels run side by side on the GPU (one host thread per canvas).
    n = max(1, int(batch_size))
    fov = tuple(int(v) for v in info.input_image_size[::-1])
    deltas = tuple(int(v) for v in info.deltas[::-1])
    self.engines, self.locks = [], []
    first = engine_lib.Engine(weights, biases, fov_zyx=fov, deltas_zyx=deltas, device=device, compute_mode=compute_mode)
    sms = first.info()['grid']
    self.engines.append(first)
    for _ in range(1, n):
      self.engines.append(engine_lib.Engine(weights, biases, fov_zyx=fov, deltas_zyx=deltas, device=device,
                                            compute_mode=compute_mode))
    if n > 1:
      for e in self.engines:
        e.set_grid(sms // n)
    self.locks = [threading.RLock() for _ in self.engines]
    self._next_slot = 0
    self._running = False

  @property
  def engine(self):
    return self.engines[0]

  @property
  def lock(self):
    return self.locks[0]

  def start_server(self):
    self._interface.exit_request.clear()
    self._running = True

  def stop_server(self):
    if self._running:
      self._interface.exit_request.set()
      self._running = False

  def get_client(self, subvol_counters):
    with self._interface.lock:
      slot = self._next_slot % len(self.engines)
      self._next_slot += 1
    return B200ExecutorClient(subvol_counters, self._interface, self, slot)

  @property
  def num_devices(self):
    return 1

  def close(self):
    self.stop_server()
    for e in self.engines:
      e.close()
    self.engines = []


# Name kept so code written against the reference's class keeps importing.
ThreadingBatchExecutor = B200Executor
