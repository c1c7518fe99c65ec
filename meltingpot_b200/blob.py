"""Reader/writer for compiled-substrate blobs (layout: include/mpb_format.h)."""

from __future__ import annotations

import struct
from typing import Dict, Mapping

import numpy as np

MAGIC = b'MPB1'
VERSION = 4
NAME_LEN = 32

_DTYPES = {
    0: np.dtype(np.uint8),
    1: np.dtype(np.uint16),
    2: np.dtype(np.int32),
    3: np.dtype(np.float64),
    4: np.dtype(np.int64),
    5: np.dtype('S1'),
}
_CODES = {v: k for k, v in _DTYPES.items()}

_HEADER = struct.Struct('<4sIII')
_SECTION = struct.Struct('<32sII4IQQ')


def pack(sections: Mapping[str, np.ndarray]) -> bytes:
  """Serialises named arrays. Strings may be passed as `bytes`/`str`."""
  arrays = {}
  for name, value in sections.items():
    if isinstance(value, str):
      value = value.encode('utf-8')
    if isinstance(value, (bytes, bytearray)):
      value = np.frombuffer(bytes(value), dtype='S1')
    value = np.ascontiguousarray(value)
    if value.dtype not in _CODES:
      raise TypeError(f'section {name!r}: unsupported dtype {value.dtype}')
    if value.ndim > 4:
      raise ValueError(f'section {name!r}: ndim {value.ndim} > 4')
    if len(name.encode()) >= NAME_LEN:
      raise ValueError(f'section name too long: {name!r}')
    arrays[name] = value
  n = len(arrays)
  offset = _HEADER.size + n * _SECTION.size
  offset = (offset + 15) & ~15
  table = []
  payload = []
  for name, arr in arrays.items():
    shape = list(arr.shape) + [0] * (4 - arr.ndim)
    nbytes = arr.nbytes
    table.append(_SECTION.pack(name.encode(), _CODES[arr.dtype], arr.ndim,
                               *shape, offset, nbytes))
    pad = (-nbytes) % 16
    payload.append(arr.tobytes() + b'\0' * pad)
    offset += nbytes + pad
  head = _HEADER.pack(MAGIC, VERSION, n, 0) + b''.join(table)
  head += b'\0' * ((-len(head)) % 16)
  return head + b''.join(payload)


def unpack(blob: bytes) -> Dict[str, np.ndarray]:
  """Parses a blob back into named (read-only) arrays."""
  magic, version, n, _ = _HEADER.unpack_from(blob, 0)
  if magic != MAGIC or version != VERSION:
    raise ValueError(f'not an MPB{VERSION} blob (magic={magic!r} v={version})')
  out = {}
  for i in range(n):
    (name, code, ndim, s0, s1, s2, s3, offset, nbytes) = _SECTION.unpack_from(
        blob, _HEADER.size + i * _SECTION.size)
    name = name.split(b'\0', 1)[0].decode()
    shape = (s0, s1, s2, s3)[:ndim]
    dtype = _DTYPES[code]
    arr = np.frombuffer(blob, dtype=dtype, count=nbytes // dtype.itemsize,
                        offset=offset).reshape(shape)
    out[name] = arr
  return out


def section_text(sections: Mapping[str, np.ndarray], name: str) -> str:
  return sections[name].tobytes().decode('utf-8')
