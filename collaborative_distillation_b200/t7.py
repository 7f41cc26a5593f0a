"""Minimal Torch7 (.t7) binary reader -- replaces `torch.utils.serialization.load_lua`, which torch >= 1.0 no longer
ships, for the one thing the reference needs it for: reading the conv weights of the `nn.Sequential` VGG / decoder
files of `--mode original` (reference: model/model_original.py:26-29 `load_lua(model)`, utils.py:64-67
`load_param_from_t7(model, i, layer)` = `model.get(i).weight / .bias`).

Format (little endian, as written by torch7's File:writeObject in binary mode):
  object  := int32 type, payload
  type 0 nil | 1 number (float64) | 2 string (int32 n, n bytes) | 3 table | 4 torch object | 5 boolean (int32)
  table   := int32 ref-index, [int32 n, n x (key object, value object)]        (payload only the first time an index is seen)
  torch   := int32 ref-index, [version string "V <n>", class-name string, class payload]
  tensor  := int32 ndim, int64 size[ndim], int64 stride[ndim], int64 storage-offset (1-based), storage object
  storage := int64 n, n raw elements
  any other class (nn.*) := one object (normally a table of its fields)
"""
from __future__ import annotations

import struct

import numpy as np
import torch

_TENSOR = {"torch.FloatTensor": np.float32, "torch.DoubleTensor": np.float64, "torch.LongTensor": np.int64,
           "torch.IntTensor": np.int32, "torch.ByteTensor": np.uint8, "torch.CudaTensor": np.float32}
_STORAGE = {"torch.FloatStorage": np.float32, "torch.DoubleStorage": np.float64, "torch.LongStorage": np.int64,
            "torch.IntStorage": np.int32, "torch.ByteStorage": np.uint8, "torch.CudaStorage": np.float32}


class T7Object:
    """a deserialised torch class instance that is not a tensor: `.torch_typename` and its fields as attributes"""

    def __init__(self, typename, fields):
        self.torch_typename = typename
        self._fields = fields if isinstance(fields, dict) else {"value": fields}

    def __getattr__(self, k):
        f = self.__dict__.get("_fields", {})
        if k in f:
            return f[k]
        raise AttributeError(k)

    def get(self, i):
        """`model:get(i+1)` -- 0-based child of a container, like the legacy python nn wrapper used by the reference"""
        mods = self._fields["modules"]
        return mods[i] if isinstance(mods, list) else mods[i + 1]


class _Reader:
    def __init__(self, f):
        self.f, self.memo = f, {}

    def _read(self, fmt):
        n = struct.calcsize(fmt)
        b = self.f.read(n)
        if len(b) != n:
            raise EOFError("truncated .t7 file")
        return struct.unpack("<" + fmt, b)

    def int32(self):
        return self._read("i")[0]

    def int64(self):
        return self._read("q")[0]

    def string(self):
        n = self.int32()
        return self.f.read(n).decode("latin-1")

    def obj(self):
        t = self.int32()
        if t == 0:
            return None
        if t == 1:
            return self._read("d")[0]
        if t == 2:
            return self.string()
        if t == 5:
            return self.int32() == 1
        if t == 3:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            out = {}
            self.memo[idx] = out
            n = self.int32()
            for _ in range(n):
                k = self.obj()
                v = self.obj()
                if isinstance(k, float) and k == int(k):
                    k = int(k)
                out[k] = v
            # lua arrays: keys 1..n -> python list
            if out and all(isinstance(k, int) for k in out) and sorted(out) == list(range(1, len(out) + 1)):
                lst = [out[i] for i in range(1, len(out) + 1)]
                self.memo[idx] = lst
                return lst
            return out
        if t == 4:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            ver = self.string()
            cls = self.string() if ver.startswith("V ") else ver
            if cls in _TENSOR:
                nd = self.int32()
                size = [self.int64() for _ in range(nd)]
                stride = [self.int64() for _ in range(nd)]
                off = self.int64() - 1
                storage = self.obj()
                if storage is None or nd == 0:
                    out = torch.empty(0, dtype=torch.from_numpy(np.zeros(0, _TENSOR[cls])).dtype)
                else:
                    out = torch.as_strided(storage, size, stride, off).clone()
                self.memo[idx] = out
                return out
            if cls in _STORAGE:
                n = self.int64()
                dt = np.dtype(_STORAGE[cls])
                raw = self.f.read(n * dt.itemsize)
                if len(raw) != n * dt.itemsize:
                    raise EOFError("truncated .t7 file")
                data = np.frombuffer(raw, dtype=dt).copy()
                out = torch.from_numpy(data)
                self.memo[idx] = out
                return out
            placeholder = T7Object(cls, {})
            self.memo[idx] = placeholder
            fields = self.obj()
            placeholder._fields = fields if isinstance(fields, dict) else {"value": fields}
            return placeholder
        raise ValueError("unsupported .t7 object type %d" % t)


def load_t7(path: str):
    """read a binary-mode .t7 file; tensors come back as torch tensors, nn modules as T7Object (fields as attributes)"""
    with open(path, "rb") as f:
        return _Reader(f).obj()


def load_param_from_t7(model, in_layer_index: int, out_layer) -> None:
    """utils.py:64-67 of the reference: copy weight/bias of sequential child `in_layer_index` (0-based) into a Conv2d"""
    m = model.get(in_layer_index)
    with torch.no_grad():
        out_layer.weight.copy_(m.weight.float().view_as(out_layer.weight))
        out_layer.bias.copy_(m.bias.float().view_as(out_layer.bias))


# ---------------------------------------------------------------------------------------------------------------------
# a tiny writer, used by the tests to build synthetic nn.Sequential files (and handy for converting back)
class _Writer:
    def __init__(self, f):
        self.f, self.next = f, 1

    def w(self, fmt, *v):
        self.f.write(struct.pack("<" + fmt, *v))

    def string(self, s):
        b = s.encode("latin-1")
        self.w("i", len(b))
        self.f.write(b)

    def obj(self, o):
        if o is None:
            self.w("i", 0)
        elif isinstance(o, bool):
            self.w("ii", 5, int(o))
        elif isinstance(o, (int, float)):
            self.w("id", 1, float(o))
        elif isinstance(o, str):
            self.w("i", 2)
            self.string(o)
        elif torch.is_tensor(o):
            t = o.detach().contiguous()
            cls = {torch.float32: "Float", torch.float64: "Double", torch.int64: "Long"}[t.dtype]
            self.w("ii", 4, self.next); self.next += 1
            self.string("V 1"); self.string("torch.%sTensor" % cls)
            self.w("i", t.dim())
            for s in t.shape:
                self.w("q", s)
            for s in t.stride():
                self.w("q", s)
            self.w("q", 1)
            self.w("ii", 4, self.next); self.next += 1
            self.string("V 1"); self.string("torch.%sStorage" % cls)
            self.w("q", t.numel())
            self.f.write(t.numpy().tobytes())
        elif isinstance(o, T7Object):
            self.w("ii", 4, self.next); self.next += 1
            self.string("V 1"); self.string(o.torch_typename)
            self.obj(o._fields)
        elif isinstance(o, (list, tuple)):
            self.obj({i + 1: v for i, v in enumerate(o)})
        elif isinstance(o, dict):
            self.w("ii", 3, self.next); self.next += 1
            self.w("i", len(o))
            for k, v in o.items():
                self.obj(k)
                self.obj(v)
        else:
            raise TypeError(type(o))


def save_t7(path: str, obj) -> None:
    with open(path, "wb") as f:
        _Writer(f).obj(obj)
