"""ctypes binding of libmpengine.so (include/mp_engine.h) + zero-copy torch views.

PyTorch is used only to wrap the engine's device buffers as tensors and to supply
streams; all compute is in the library's CUDA kernels. There is no CPU path: if
the library or a CUDA device is missing, construction raises.
"""

from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MP_ENGINE_LIB') or os.path.join(_HERE, 'libmpengine.so')

MP_FLAG_RENDER_WORLD = 1
MP_FLAG_RENDER_PLAYERS = 2
MP_FLAG_DEFAULT = 3

EXPORTED_SYMBOLS = (
    'mp_create', 'mp_destroy', 'mp_set_flags', 'mp_reset', 'mp_step',
    'mp_step_state', 'mp_render', 'mp_get_buffers', 'mp_step_host',
    'mp_reset_host', 'mp_launch_count', 'mp_algorithmic_bytes', 'mp_debug_render_tables', 'mp_debug_render_plan', 'mp_state_size', 'mp_state_save', 'mp_state_load',
    'mp_step_host_async', 'mp_wait', 'mp_exchange_create', 'mp_ipc_export', 'mp_ipc_open', 'mp_enable_peer_access',
    'mp_exchange_connect', 'mp_exchange_wait', 'mp_exchange_slot', 'mp_debug_lane_map', 'mp_debug_observations',
    'mp_gather_obs_create', 'mp_gather_obs_connect', 'mp_gather_obs_enable', 'mp_gather_obs_wait', 'mp_gather_obs_slot',
    'mp_last_error', 'mp_version',
)


class MpBuffers(ctypes.Structure):
  _fields_ = [
      ('num_envs', ctypes.c_int32), ('num_players', ctypes.c_int32),
      ('rgb_h', ctypes.c_int32), ('rgb_w', ctypes.c_int32),
      ('world_h', ctypes.c_int32), ('world_w', ctypes.c_int32),
      ('num_actions', ctypes.c_int32), ('num_scalar_obs', ctypes.c_int32),
      ('rgb', ctypes.c_void_p), ('world_rgb', ctypes.c_void_p),
      ('reward', ctypes.c_void_p), ('discount', ctypes.c_void_p),
      ('step_type', ctypes.c_void_p), ('scalar_obs', ctypes.c_void_p),
      ('avatar_state', ctypes.c_void_p), ('grid', ctypes.c_void_p),
      ('grid_layers', ctypes.c_int32), ('grid_cells', ctypes.c_int32),
      ('grid_cells_padded', ctypes.c_int32),
      ('timestep_packed', ctypes.c_void_p),
      ('events', ctypes.c_void_p), ('event_count', ctypes.c_void_p), ('max_events', ctypes.c_int32),
      ('scalar_block', ctypes.c_void_p), ('scalar_block_bytes', ctypes.c_uint64),
      ('gathered', ctypes.c_void_p), ('gathered_world', ctypes.c_int32),
      ('gathered_rgb', ctypes.c_void_p), ('gathered_world_rgb', ctypes.c_void_p), ('gathered_obs_slot_bytes', ctypes.c_uint64),
  ]


class MpHostOutputs(ctypes.Structure):
  _fields_ = [
      ('rgb', ctypes.c_void_p), ('world_rgb', ctypes.c_void_p),
      ('reward', ctypes.c_void_p), ('discount', ctypes.c_void_p),
      ('step_type', ctypes.c_void_p), ('scalar_obs', ctypes.c_void_p),
      ('scalar_block', ctypes.c_void_p), ('events', ctypes.c_void_p), ('event_count', ctypes.c_void_p),
  ]


_lib = None


def load_library() -> ctypes.CDLL:
  """Loads libmpengine.so; raises if it has not been built (no fallback)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        f'{LIB_PATH} is missing: build it with `python -m meltingpot_b200.build` '
        '(the engine has no CPU or PyTorch fallback)')
  lib = ctypes.CDLL(LIB_PATH)
  vp = ctypes.c_void_p
  lib.mp_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                            ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                            ctypes.c_uint32, ctypes.POINTER(vp)]
  lib.mp_destroy.argtypes = [vp]
  lib.mp_set_flags.argtypes = [vp, ctypes.c_uint32]
  lib.mp_reset.argtypes = [vp, vp, vp]
  lib.mp_step.argtypes = [vp, vp, vp]
  lib.mp_step_state.argtypes = [vp, vp, vp]
  lib.mp_render.argtypes = [vp, vp]
  lib.mp_get_buffers.argtypes = [vp, ctypes.POINTER(MpBuffers)]
  lib.mp_step_host.argtypes = [vp, vp, ctypes.POINTER(MpHostOutputs), vp]
  lib.mp_reset_host.argtypes = [vp, ctypes.POINTER(MpHostOutputs), vp]
  lib.mp_launch_count.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_algorithmic_bytes.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64),
                                       ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_state_size.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_state_save.argtypes = [vp, vp, vp]
  lib.mp_state_load.argtypes = [vp, vp, ctypes.c_uint64, vp]
  lib.mp_debug_render_plan.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
  lib.mp_debug_render_tables.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), vp, vp]
  lib.mp_step_host_async.argtypes = [vp, vp, ctypes.POINTER(MpHostOutputs), ctypes.c_int, vp]
  lib.mp_wait.argtypes = [vp, ctypes.c_int]
  lib.mp_exchange_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_ipc_export.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_ipc_open.argtypes = [ctypes.c_int, vp, ctypes.c_uint64, ctypes.POINTER(vp)]
  lib.mp_enable_peer_access.argtypes = [ctypes.c_int, ctypes.c_int]
  lib.mp_exchange_connect.argtypes = [vp, ctypes.POINTER(vp)]
  lib.mp_exchange_wait.argtypes = [vp, vp]
  lib.mp_exchange_slot.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_gather_obs_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_gather_obs_connect.argtypes = [vp, ctypes.POINTER(vp)]
  lib.mp_gather_obs_enable.argtypes = [vp, ctypes.c_int]
  lib.mp_gather_obs_wait.argtypes = [vp, vp]
  lib.mp_gather_obs_slot.argtypes = [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_uint64)]
  lib.mp_debug_observations.argtypes = [vp] * 6
  lib.mp_debug_lane_map.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_uint32)]
  lib.mp_last_error.restype = ctypes.c_char_p
  lib.mp_version.restype = ctypes.c_char_p
  _lib = lib
  return lib


EVENT_NAMES = {1: 'zap', 2: 'edible_consumed', 3: 'player_cleaned', 4: 'claimed_resource',
               5: 'destroyed_resource', 6: 'sanctioning', 7: 'removal_due_to_sanctioning', 8: 'coin_consumed',
               9: 'mining', 10: 'extraction', 11: 'extraction_pair'}
# argument names of each event's dict payload (second one unused for single-argument events)
EVENT_FIELDS = {1: ('source', 'target'), 2: ('player_index',), 3: ('player_index',), 4: ('player_index',),
                5: ('player_index',), 6: ('source', 'target'), 7: ('source', 'target'), 8: ('player_index', 'matched'),
                9: ('player', 'ore_type'), 10: ('player', 'ore_type'), 11: ('player_a', 'player_b_and_ore_type')}


class EngineError(RuntimeError):
  pass


def _check(rc: int) -> None:
  if rc != 0:
    msg = load_library().mp_last_error().decode('utf-8', 'replace')
    if rc in (-1, -2):
      raise ValueError(f'mp_engine error {rc}: {msg}')
    raise EngineError(f'mp_engine error {rc}: {msg}')


class _Handle:
  """Owns one mp_handle. Every tensor view of the engine's buffers holds a reference, so the device memory is
  released (mp_destroy) only once the Engine has been closed AND the last view is gone: a retained observation
  tensor can never dangle."""

  def __init__(self, lib, handle):
    self._lib = lib
    self.h = handle

  def __del__(self):
    try:
      if self.h:
        self._lib.mp_destroy(self.h)
        self.h = None
    except Exception:  # pylint: disable=broad-except
      pass


class _CudaView:
  """Exposes a raw device pointer through __cuda_array_interface__ (v2)."""

  def __init__(self, ptr: int, shape, typestr: str, owner):
    self._owner = owner  # the _Handle: keeps the device memory alive while tensors exist
    self.__cuda_array_interface__ = {
        'shape': tuple(int(s) for s in shape), 'typestr': typestr,
        'data': (int(ptr), False), 'version': 2, 'strides': None,
    }


class Engine:
  """One engine handle = `num_envs` env instances on one GPU."""

  def __init__(self, blob: bytes, num_envs: int, device: int = 0, seed: int = 1,
               env_index_base: int = 0, flags: int = MP_FLAG_DEFAULT):
    import torch  # pylint: disable=g-import-not-at-top
    if not torch.cuda.is_available():
      raise EngineError('CUDA is not available: the B200 engine has no CPU path')
    self._torch = torch
    self._lib = load_library()
    self._blob = bytes(blob)
    self.device = int(device)
    self.num_envs = int(num_envs)
    torch.cuda.init()
    with torch.cuda.device(self.device):
      torch.cuda.current_stream()  # make sure the primary context exists
    handle = ctypes.c_void_p()
    _check(self._lib.mp_create(self._blob, len(self._blob), self.num_envs,
                               self.device, ctypes.c_uint64(seed),
                               ctypes.c_uint64(env_index_base),
                               ctypes.c_uint32(flags), ctypes.byref(handle)))
    self._owner = _Handle(self._lib, handle)
    self._h = handle
    bufs = MpBuffers()
    _check(self._lib.mp_get_buffers(self._h, ctypes.byref(bufs)))
    self.buffers = bufs
    self.num_players = int(bufs.num_players)
    self.num_actions = int(bufs.num_actions)
    self.num_scalar_obs = int(bufs.num_scalar_obs)
    B, P = self.num_envs, self.num_players
    dev = torch.device('cuda', self.device)

    def view(ptr, shape, typestr, dtype):
      return torch.as_tensor(_CudaView(ptr, shape, typestr, self._owner), device=dev, dtype=dtype)

    self.rgb = view(bufs.rgb, (B, P, bufs.rgb_h, bufs.rgb_w, 3), '|u1', torch.uint8)
    self.world_rgb = view(bufs.world_rgb, (B, bufs.world_h, bufs.world_w, 3), '|u1', torch.uint8)
    self.reward = view(bufs.reward, (B, P), '<f8', torch.float64)
    self.discount = view(bufs.discount, (B,), '<f8', torch.float64)
    self.step_type = view(bufs.step_type, (B,), '<i8', torch.int64)
    self.scalar_obs = view(bufs.scalar_obs, (max(self.num_scalar_obs, 1), B, P), '<f8', torch.float64)
    self.avatar_state = view(bufs.avatar_state, (B, P, 4), '<i4', torch.int32)
    self.grid = view(bufs.grid, (B, bufs.grid_layers, bufs.grid_cells_padded), '<i2', torch.int16)
    self.timestep_packed = view(bufs.timestep_packed, (B, P + 2), '<f8', torch.float64)
    self.events = view(bufs.events, (B, bufs.max_events, 3), '<i4', torch.int32)
    self.event_count = view(bufs.event_count, (B,), '<i4', torch.int32)

  # -- lifecycle -----------------------------------------------------------------
  _VIEWS = ('rgb', 'world_rgb', 'reward', 'discount', 'step_type', 'scalar_obs', 'avatar_state', 'grid',
            'timestep_packed', 'events', 'event_count', 'gathered', 'gathered_rgb', 'gathered_world_rgb')

  def close(self) -> None:
    """Drops this object's references. mp_destroy runs when the last tensor view handed out has been released too
    (tensors a caller still holds stay readable; the engine itself can no longer be stepped)."""
    if getattr(self, '_h', None):
      self._h = None
      for name in self._VIEWS:
        self.__dict__.pop(name, None)
      self._owner = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  def _stream(self, stream) -> ctypes.c_void_p:
    if stream is None:
      stream = self._torch.cuda.current_stream(self.device)
    return ctypes.c_void_p(stream.cuda_stream)

  # -- device-resident API ---------------------------------------------------------
  def set_flags(self, flags: int) -> None:
    _check(self._lib.mp_set_flags(self._h, ctypes.c_uint32(flags)))

  def reset(self, mask=None, stream=None) -> None:
    ptr = None
    if mask is not None:
      assert mask.dtype == self._torch.uint8 and mask.is_cuda and mask.numel() == self.num_envs
      ptr = ctypes.c_void_p(mask.data_ptr())
    _check(self._lib.mp_reset(self._h, ptr, self._stream(stream)))

  def step(self, actions, stream=None) -> None:
    """actions: int32 CUDA tensor [B, P] of discrete action ids."""
    self._check_actions(actions)
    _check(self._lib.mp_step(self._h, ctypes.c_void_p(actions.data_ptr()), self._stream(stream)))

  def step_state(self, actions, stream=None) -> None:
    self._check_actions(actions)
    _check(self._lib.mp_step_state(self._h, ctypes.c_void_p(actions.data_ptr()), self._stream(stream)))

  def render(self, stream=None) -> None:
    _check(self._lib.mp_render(self._h, self._stream(stream)))

  def _check_actions(self, actions) -> None:
    torch = self._torch
    if (actions.dtype != torch.int32 or not actions.is_cuda or not actions.is_contiguous()
        or actions.shape != (self.num_envs, self.num_players)
        or actions.device.index != self.device):
      raise ValueError('actions must be a contiguous int32 CUDA tensor [B, P] on the engine device')

  # -- host-buffer API (end-to-end path) -----------------------------------------
  def make_host_outputs(self, rgb=True, world_rgb=True, events=False) -> Dict[str, 'np.ndarray']:
    """Pinned host tensors for mp_step_host / mp_step_host_async, as a dict of torch CPU tensors.

    reward / discount / step_type / scalar_obs are views of one pinned block laid out like mp_buffers.scalar_block,
    so the engine moves all scalar outputs of a step with a single device->host copy.
    """
    torch = self._torch
    b = self.buffers
    B, P = self.num_envs, self.num_players
    ns = max(self.num_scalar_obs, 1)
    n_words = int(b.scalar_block_bytes) // 8
    assert n_words == B * P + 2 * B + ns * B * P
    block = torch.empty((n_words,), dtype=torch.float64).pin_memory()
    o = 0
    out = {'scalar_block': block}
    out['reward'] = block[o:o + B * P].view(B, P); o += B * P
    out['discount'] = block[o:o + B]; o += B
    out['step_type'] = block[o:o + B].view(torch.int64); o += B
    out['scalar_obs'] = block[o:o + ns * B * P].view(ns, B, P)
    if rgb:
      out['rgb'] = torch.empty((B, P, b.rgb_h, b.rgb_w, 3), dtype=torch.uint8).pin_memory()
    if world_rgb:
      out['world_rgb'] = torch.empty((B, b.world_h, b.world_w, 3), dtype=torch.uint8).pin_memory()
    if events:
      out['events'] = torch.empty((B, b.max_events, 3), dtype=torch.int32).pin_memory()
      out['event_count'] = torch.empty((B,), dtype=torch.int32).pin_memory()
    return out

  def make_host_actions(self):
    """Pinned int32 [B, P] host tensor for the actions of mp_step_host / mp_step_host_async."""
    return self._torch.zeros((self.num_envs, self.num_players), dtype=self._torch.int32).pin_memory()

  @staticmethod
  def _host_struct(outputs) -> MpHostOutputs:
    s = MpHostOutputs()
    for name in ('rgb', 'world_rgb', 'reward', 'discount', 'step_type', 'scalar_obs', 'scalar_block', 'events', 'event_count'):
      t = outputs.get(name) if outputs else None
      setattr(s, name, ctypes.c_void_p(t.data_ptr()) if t is not None else None)
    return s

  def step_host(self, actions_host, outputs, stream=None) -> None:
    """actions_host: int32 CPU tensor [B, P] (pinned for full speed)."""
    assert actions_host.dtype == self._torch.int32 and not actions_host.is_cuda
    assert actions_host.is_contiguous() and actions_host.shape == (self.num_envs, self.num_players)
    s = self._host_struct(outputs)
    _check(self._lib.mp_step_host(self._h, ctypes.c_void_p(actions_host.data_ptr()),
                                  ctypes.byref(s), self._stream(stream)))

  def step_host_async(self, actions_host, outputs, slot: int, stream=None) -> None:
    """Pipelined step (mp_step_host_async): returns at once; `wait(slot)` before reading `outputs` or reusing
    `actions_host`. Alternate slot 0 / 1 between consecutive calls."""
    assert actions_host.dtype == self._torch.int32 and not actions_host.is_cuda
    assert actions_host.is_contiguous() and actions_host.shape == (self.num_envs, self.num_players)
    s = self._host_struct(outputs)
    _check(self._lib.mp_step_host_async(self._h, ctypes.c_void_p(actions_host.data_ptr()), ctypes.byref(s),
                                        int(slot), self._stream(stream)))

  def wait(self, slot: int) -> None:
    _check(self._lib.mp_wait(self._h, int(slot)))

  # -- stacked timestep across GPUs (mp_exchange_*) ------------------------------------
  def exchange_create(self, rank: int, world: int):
    """Allocates this rank's exchange block; returns (device pointer, bytes)."""
    ptr, n = ctypes.c_void_p(), ctypes.c_uint64(0)
    _check(self._lib.mp_exchange_create(self._h, int(rank), int(world), ctypes.byref(ptr), ctypes.byref(n)))
    self._x_world, self._x_rank = int(world), int(rank)
    bufs = MpBuffers()
    _check(self._lib.mp_get_buffers(self._h, ctypes.byref(bufs)))
    self.buffers = bufs
    torch = self._torch
    self.gathered = torch.as_tensor(
        _CudaView(bufs.gathered, (2, world * self.num_envs, self.num_players + 2), '<f8', self._owner),
        device=torch.device('cuda', self.device), dtype=torch.float64)
    return int(ptr.value), int(n.value)

  def exchange_connect(self, peer_blocks) -> None:
    """peer_blocks: every rank's block pointer (ints) as mapped into this process, in rank order."""
    arr = (ctypes.c_void_p * len(peer_blocks))(*[ctypes.c_void_p(int(p)) for p in peer_blocks])
    _check(self._lib.mp_exchange_connect(self._h, arr))

  def exchange_wait(self, stream=None) -> None:
    _check(self._lib.mp_exchange_wait(self._h, self._stream(stream)))

  def exchange_slot(self):
    slot, step = ctypes.c_int(0), ctypes.c_uint64(0)
    _check(self._lib.mp_exchange_slot(self._h, ctypes.byref(slot), ctypes.byref(step)))
    return int(slot.value), int(step.value)

  def gathered_timestep(self):
    """The stacked [world * B, P + 2] timestep rows of the most recent step (call exchange_wait first)."""
    return self.gathered[self.exchange_slot()[0]]

  # -- stacked observations across GPUs (mp_gather_obs_*) ----------------------------------
  def gather_obs_create(self, rank: int, world: int):
    """Allocates this rank's stacked-observation block; returns (device pointer, bytes)."""
    ptr, n = ctypes.c_void_p(), ctypes.c_uint64(0)
    _check(self._lib.mp_gather_obs_create(self._h, int(rank), int(world), ctypes.byref(ptr), ctypes.byref(n)))
    bufs = MpBuffers()
    _check(self._lib.mp_get_buffers(self._h, ctypes.byref(bufs)))
    self.buffers = bufs
    torch = self._torch
    dev = torch.device('cuda', self.device)
    B, P = self.num_envs, self.num_players
    slot = int(bufs.gathered_obs_slot_bytes)
    self.gathered_rgb = [torch.as_tensor(_CudaView(bufs.gathered_rgb + k * slot, (world * B, P, bufs.rgb_h, bufs.rgb_w, 3), '|u1', self._owner),
                                         device=dev, dtype=torch.uint8) for k in range(2)]
    self.gathered_world_rgb = [torch.as_tensor(_CudaView(bufs.gathered_world_rgb + k * slot, (world * B, bufs.world_h, bufs.world_w, 3), '|u1', self._owner),
                                               device=dev, dtype=torch.uint8) for k in range(2)]
    return int(ptr.value), int(n.value)

  def gather_obs_connect(self, peer_blocks) -> None:
    arr = (ctypes.c_void_p * len(peer_blocks))(*[ctypes.c_void_p(int(p)) for p in peer_blocks])
    _check(self._lib.mp_gather_obs_connect(self._h, arr))

  def gather_obs_enable(self, on: bool) -> None:
    _check(self._lib.mp_gather_obs_enable(self._h, int(bool(on))))

  def gather_obs_wait(self, stream=None) -> None:
    _check(self._lib.mp_gather_obs_wait(self._h, self._stream(stream)))

  def gathered_observations(self):
    """(rgb [world * B, P, h, w, 3], world_rgb [world * B, H, W, 3]) of the most recent render (gather_obs_wait first)."""
    slot = ctypes.c_int(0)
    _check(self._lib.mp_gather_obs_slot(self._h, ctypes.byref(slot), None))
    return self.gathered_rgb[slot.value], self.gathered_world_rgb[slot.value]

  def debug_observations(self, layer: bool = True, zap_matrix: bool = True, stream=None):
    """{'POSITION' [B,P,2], 'ORIENTATION' [B,P], 'LAYER' [B,P,vh,vw,L], 'ZAP_MATRIX' [B,P,P]} int32 CUDA tensors of the
    current timestep (mp_debug_observations)."""
    torch = self._torch
    dev = torch.device('cuda', self.device)
    b = self.buffers
    B, P = self.num_envs, self.num_players
    out = {'POSITION': torch.empty((B, P, 2), dtype=torch.int32, device=dev),
           'ORIENTATION': torch.empty((B, P), dtype=torch.int32, device=dev)}
    if layer:
      out['LAYER'] = torch.empty((B, P, b.rgb_h // 8, b.rgb_w // 8, b.grid_layers), dtype=torch.int32, device=dev)
    if zap_matrix:
      out['ZAP_MATRIX'] = torch.empty((B, P, P), dtype=torch.int32, device=dev)
    ptr = lambda k: ctypes.c_void_p(out[k].data_ptr()) if k in out else None
    _check(self._lib.mp_debug_observations(self._h, ptr('POSITION'), ptr('ORIENTATION'), ptr('LAYER'), ptr('ZAP_MATRIX'),
                                           self._stream(stream)))
    return out

  def reset_host(self, outputs, stream=None) -> None:
    s = self._host_struct(outputs)
    _check(self._lib.mp_reset_host(self._h, ctypes.byref(s), self._stream(stream)))

  # -- introspection -----------------------------------------------------------------
  def launch_count(self) -> int:
    n = ctypes.c_uint64(0)
    _check(self._lib.mp_launch_count(self._h, ctypes.byref(n)))
    return int(n.value)

  def save_state(self, stream=None) -> bytes:
    """Snapshot of every env instance (mp_state_save); restore with load_state on an identically built engine."""
    n = ctypes.c_uint64(0)
    _check(self._lib.mp_state_size(self._h, ctypes.byref(n)))
    buf = ctypes.create_string_buffer(n.value)
    _check(self._lib.mp_state_save(self._h, buf, self._stream(stream)))
    return buf.raw

  def load_state(self, snapshot: bytes, stream=None) -> None:
    snapshot = bytes(snapshot)
    buf = ctypes.create_string_buffer(snapshot, len(snapshot))
    _check(self._lib.mp_state_load(self._h, buf, ctypes.c_uint64(len(snapshot)), self._stream(stream)))

  def render_plan(self):
    """Layout the renderer chose for this substrate (diagnostic)."""
    out = (ctypes.c_int32 * 8)()
    _check(self._lib.mp_debug_render_plan(self._h, out))
    keys = ('teams', 'team_threads', 'wstrip_log2', 'smem_bytes', 'atlas_sprites', 'rec_stride', 'stage_bytes', 'grid_bytes')
    return dict(zip(keys, (int(v) for v in out)))

  def render_tables(self):
    """(pair[n, n], flags[n]) uint8 numpy arrays of the renderer's sprite tables (diagnostic)."""
    import numpy as np
    n = ctypes.c_int32(0)
    _check(self._lib.mp_debug_render_tables(self._h, ctypes.byref(n), None, None))
    pair = np.zeros((n.value, n.value), np.uint8)
    flags = np.zeros((n.value,), np.uint8)
    _check(self._lib.mp_debug_render_tables(self._h, ctypes.byref(n), pair.ctypes.data, flags.ctypes.data))
    return pair, flags

  def algorithmic_bytes(self):
    a, r = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _check(self._lib.mp_algorithmic_bytes(self._h, ctypes.byref(a), ctypes.byref(r)))
    return int(a.value), int(r.value)


def ipc_export(device_ptr: int):
  """(64-byte CUDA IPC handle, offset) naming `device_ptr` for another process (mp_ipc_export)."""
  handle = ctypes.create_string_buffer(64)
  off = ctypes.c_uint64(0)
  _check(load_library().mp_ipc_export(ctypes.c_void_p(int(device_ptr)), handle, ctypes.byref(off)))
  return handle.raw, int(off.value)


def ipc_open(device: int, handle: bytes, offset: int) -> int:
  ptr = ctypes.c_void_p()
  buf = ctypes.create_string_buffer(bytes(handle), 64)
  _check(load_library().mp_ipc_open(int(device), buf, ctypes.c_uint64(int(offset)), ctypes.byref(ptr)))
  return int(ptr.value)


def enable_peer_access(device: int, peer_device: int) -> None:
  _check(load_library().mp_enable_peer_access(int(device), int(peer_device)))
