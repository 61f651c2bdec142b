"""Torch7 `.t7` checkpoint reader: lets `checkpoint-*-video.t7` files of the reference load.

`torch.load(path).model` is the only way weights enter the reference (fast_artistic_video_core.lua:38-47); the checkpoint
table is `{opt, ..., iter, model}` with the model a float `nn.Sequential` (train_video.lua:507-541).

The serialisation format lives in the un-vendored torch7 rock (`File.lua`): PARITY UNPINNED.  It is restated here from
its published description (binary mode, little endian, 8-byte longs):
  object := int32 type, then
    0 nil | 1 number: float64 | 2 string: int32 len + bytes | 5 boolean: int32
    3 table : int32 index, [int32 count, count x (key object, value object)]   (index seen before => back-reference)
    4 torch : int32 index, [string "V <n>", string className, payload]          (same back-reference rule)
       tensor payload : int32 ndim, int64 size[ndim], int64 stride[ndim], int64 storageOffset (1-based), storage object
       storage payload: int64 n, n raw elements
       any other class: one table object holding its fields (nn modules have no custom :write)
`write_checkpoint` produces files in the same format from a state dict (used by the tests; there is no network to fetch
the released checkpoints, so the reader is validated on round trips only).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5

_TENSOR = {"torch.FloatTensor": np.float32, "torch.DoubleTensor": np.float64, "torch.LongTensor": np.int64,
           "torch.IntTensor": np.int32, "torch.ByteTensor": np.uint8, "torch.CudaTensor": np.float32}
_STORAGE = {k.replace("Tensor", "Storage"): v for k, v in _TENSOR.items()}


class T7Object(dict):
    """A deserialised torch class instance: dict of fields + .torch_type."""

    def __init__(self, torch_type, fields=None):
        super().__init__(fields or {})
        self.torch_type = torch_type


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p, self.objects = data, 0, {}

    def _unpack(self, fmt, n):
        v = struct.unpack_from("<" + fmt, self.d, self.p)
        self.p += n
        return v[0]

    def int(self):
        return self._unpack("i", 4)

    def long(self):
        return self._unpack("q", 8)

    def string(self):
        n = self.int()
        s = self.d[self.p:self.p + n].decode("latin1")
        self.p += n
        return s

    def obj(self):
        t = self.int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            return self._unpack("d", 8)
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.int() == 1
        if t == TYPE_TABLE:
            idx = self.int()
            if idx in self.objects:
                return self.objects[idx]
            tab = {}
            self.objects[idx] = tab
            for _ in range(self.int()):
                k = self.obj()
                v = self.obj()
                if isinstance(k, float) and k == int(k):
                    k = int(k)
                tab[k] = v
            return tab
        if t == TYPE_TORCH:
            idx = self.int()
            if idx in self.objects:
                return self.objects[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version  # legacy files have no version string
            if cls in _TENSOR:
                nd = self.int()
                size = [self.long() for _ in range(nd)]
                stride = [self.long() for _ in range(nd)]
                off = self.long() - 1
                self.objects[idx] = None
                storage = self.obj()
                if storage is None or nd == 0:
                    arr = np.zeros(size or (0,), _TENSOR[cls])
                else:
                    arr = np.lib.stride_tricks.as_strided(storage[off:], shape=size,
                                                          strides=[s * storage.itemsize for s in stride]).copy()
                self.objects[idx] = arr
                return arr
            if cls in _STORAGE:
                n = self.long()
                dt = np.dtype(_STORAGE[cls]).newbyteorder("<")
                arr = np.frombuffer(self.d, dt, count=n, offset=self.p).copy()
                self.p += n * dt.itemsize
                self.objects[idx] = arr
                return arr
            o = T7Object(cls)
            self.objects[idx] = o
            fields = self.obj()
            if isinstance(fields, dict):
                o.update(fields)
            return o
        raise ValueError(f".t7: unsupported object type {t} at byte {self.p}")


def load(path: str):
    with open(path, "rb") as f:
        return _Reader(f.read()).obj()


# ---- nn.Sequential -> (arch string, state dict in fav.h naming) ------------------------------------------------------
def _cls(m):
    return m.torch_type.split(".", 1)[1] if isinstance(m, T7Object) else None


def _mods(m) -> List:
    mods = m.get("modules", {})
    return [mods[k] for k in sorted(mods)]


class GeometryError(ValueError):
    """The checkpoint's module tree does not have the geometry build_model would give its arch string."""


def model_to_state(model, geometry: dict = None) -> Tuple[str, Dict[str, np.ndarray], float, int]:
    """Walk the checkpoint's nn.Sequential exactly as models_video.build_model lays it out (models_video.lua:55-140,
    plus the lazily inserted SpatialReflectionPadding, train_video.lua:319-324) and return
    (arch string, {name: array}, tanh_constant, reflect_pad).  `geometry`, if given, receives what the file says about
    padding: 'padding_type' ('reflect-start' | 'zero', inferred from the residual blocks: pad-0 convs + ShaveImage vs pad-1
    convs + Identity, models_video.lua:21-53), 'has_reflect_pad', and every conv's (name, padW, padW implied by its token);
    a layout build_model cannot produce for these two padding types raises GeometryError instead of being run with a
    different geometry."""
    toks, state, tanh_c, pad = [], {}, 150.0, 0
    geo = geometry if geometry is not None else {}
    geo.update(padding_type=None, has_reflect_pad=False, conv_pads=[])

    def check_pad(name, m, implied):
        pw, ph = int(m.get("padW", 0)), int(m.get("padH", m.get("padW", 0)))
        geo["conv_pads"].append((name, pw, implied))
        if pw != implied or ph != implied:
            raise GeometryError(f".t7 model: {name} has padding {pw}x{ph}, its arch token implies {implied}")
    mods = _mods(model)
    i = 0

    def conv_params(name, m, transposed):
        w = np.asarray(m["weight"], np.float32)
        cin, cout, k = int(m["nInputPlane"]), int(m["nOutputPlane"]), int(m["kW"])
        state[name + ".weight"] = w.reshape((cin, cout, k, k) if transposed else (cout, cin, k, k))
        state[name + ".bias"] = np.asarray(m["bias"], np.float32).reshape(cout)

    def in_params(name, m):
        state[name + ".weight"] = np.asarray(m["weight"], np.float32).reshape(-1)
        state[name + ".bias"] = np.asarray(m["bias"], np.float32).reshape(-1)

    def take_norm_relu(name):
        nonlocal i
        if i < len(mods) and _cls(mods[i]) == "InstanceNormalization":
            in_params(name + ".n", mods[i])
            i += 1
        if i < len(mods) and _cls(mods[i]) == "ReLU":
            i += 1

    while i < len(mods):
        m = mods[i]
        c = _cls(m)
        name = f"l{len(toks)}"
        i += 1
        if c == "SpatialReflectionPadding":
            pad = int(m["pad_l"])
            if not (i == 1 and pad == int(m["pad_r"]) == int(m["pad_t"]) == int(m["pad_b"])):
                raise GeometryError(".t7 model: only the leading symmetric SpatialReflectionPadding of padding_type "
                                    "'reflect-start' is supported (train_video.lua:319-324)")
            geo["has_reflect_pad"] = True
        elif c == "SpatialConvolution":
            k, s = int(m["kW"]), int(m["dW"])
            if k == 3 and s == 2 and int(m["padW"]) == 1:
                toks.append(f"d{int(m['nOutputPlane'])}")
            else:
                toks.append(f"c{k}s{s}-{int(m['nOutputPlane'])}")
            check_pad(name, m, 1 if toks[-1][0] == "d" else (k - 1) // 2)  # models_video.lua:65-80,90-93
            conv_params(name, m, False)
            take_norm_relu(name)
        elif c == "SpatialFullConvolution":
            k, s = int(m["kW"]), int(m["dW"])
            toks.append(f"u{int(m['nOutputPlane'])}" if (k == 3 and s == 2) else f"f{k}s{s}-{int(m['nOutputPlane'])}")
            check_pad(name, m, (k - 1) // 2)  # :81-89,99-102
            conv_params(name, m, True)
            take_norm_relu(name)
        elif c == "SpatialUpSamplingNearest":
            toks.append(f"U{int(m['scale_factor'])}")
            take_norm_relu(name)
        elif c == "Sequential":  # residual block: Sequential{ConcatTable{conv_block, ShaveImage}, CAddTable}
            concat = _mods(m)[0]
            block = _mods(_mods(concat)[0])
            convs = [b for b in block if _cls(b) == "SpatialConvolution"]
            norms = [b for b in block if _cls(b) == "InstanceNormalization"]
            assert len(convs) == 2 and len(norms) == 2, "unexpected residual block layout"
            extra = [_cls(b) for b in block if _cls(b) not in ("SpatialConvolution", "InstanceNormalization", "ReLU")]
            skip = _cls(_mods(concat)[1]) if len(_mods(concat)) > 1 else None
            rp = int(convs[0].get("padW", 0))
            ptype = "reflect-start" if (rp == 0 and skip == "ShaveImage") else ("zero" if (rp == 1 and skip == "Identity") else None)
            if extra or ptype is None or (geo["padding_type"] not in (None, ptype)):
                raise GeometryError(f".t7 model: residual block {name} (conv padding {rp}, skip {skip}, extra modules {extra}) is not "
                                    "a padding_type 'reflect-start' or 'zero' block (models_video.lua:21-53)")
            geo["padding_type"] = ptype
            for cn, cm in ((name + ".c1", convs[0]), (name + ".c2", convs[1])):
                check_pad(cn, cm, rp)
            toks.append(f"R{int(convs[0]['nOutputPlane'])}")
            conv_params(name + ".c1", convs[0], False); in_params(name + ".n1", norms[0])
            conv_params(name + ".c2", convs[1], False); in_params(name + ".n2", norms[1])
        elif c == "Tanh":
            pass
        elif c == "MulConstant":
            tanh_c = float(m["constant_scalar"])
        elif c == "TotalVariation":
            pass  # identity in forward (TotalVariation.lua:12-15)
        else:
            raise ValueError(f".t7 model: unsupported module {m.torch_type}")
    if geo["padding_type"] is None:
        geo["padding_type"] = "reflect-start" if geo["has_reflect_pad"] else "zero"
    if geo["has_reflect_pad"] != (geo["padding_type"] == "reflect-start" and pad > 0) and any(t[0] in "RC" for t in toks):
        raise GeometryError(".t7 model: the leading SpatialReflectionPadding and the residual blocks' padding disagree")
    return ",".join(toks), state, tanh_c, pad


def load_checkpoint(path: str, geometry: dict = None):
    ck = load(path)
    model = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
    return model_to_state(model, geometry)


# ---- writer (tests) ---------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.b, self.n = bytearray(), 0

    def int(self, v):
        self.b += struct.pack("<i", v)

    def long(self, v):
        self.b += struct.pack("<q", v)

    def string(self, s):
        self.int(len(s))
        self.b += s.encode("latin1")

    def obj(self, o):
        if o is None:
            self.int(TYPE_NIL)
        elif isinstance(o, bool):
            self.int(TYPE_BOOLEAN); self.int(1 if o else 0)
        elif isinstance(o, (int, float)):
            self.int(TYPE_NUMBER); self.b += struct.pack("<d", float(o))
        elif isinstance(o, str):
            self.int(TYPE_STRING); self.string(o)
        elif isinstance(o, np.ndarray):
            a = np.ascontiguousarray(o, np.float32)
            self.int(TYPE_TORCH); self.n += 1; self.int(self.n); self.string("V 1"); self.string("torch.FloatTensor")
            self.int(a.ndim)
            for s in a.shape:
                self.long(s)
            st = [int(x // a.itemsize) for x in a.strides]
            for s in st:
                self.long(s)
            self.long(1)
            self.int(TYPE_TORCH); self.n += 1; self.int(self.n); self.string("V 1"); self.string("torch.FloatStorage")
            self.long(a.size)
            self.b += a.tobytes()
        elif isinstance(o, T7Object):
            self.int(TYPE_TORCH); self.n += 1; self.int(self.n); self.string("V 1"); self.string(o.torch_type)
            self.obj(dict(o))
        elif isinstance(o, dict):
            self.int(TYPE_TABLE); self.n += 1; self.int(self.n); self.int(len(o))
            for k, v in o.items():
                self.obj(k); self.obj(v)
        else:
            raise TypeError(type(o))


def write_checkpoint(path: str, arch: str, state: Dict[str, np.ndarray], tanh_constant: float = 150.0,
                     reflect_pad: int = 40, padding_type: str = "reflect-start") -> None:
    """Serialise a state dict as the `{model = nn.Sequential{...}}` table train_video.lua:507-541 saves
    (padding_type 'reflect-start': leading reflection pad, pad-0 residual convs + ShaveImage; 'zero': pad-1 residual convs +
    Identity, models_video.lua:21-53)."""
    from . import synth

    def seq(mods, cls="nn.Sequential"):
        return T7Object(cls, {"modules": {i + 1: m for i, m in enumerate(mods)}})

    def conv(name, cin, cout, k, s, p, tr=False):
        return T7Object("nn.SpatialFullConvolution" if tr else "nn.SpatialConvolution",
                        {"nInputPlane": cin, "nOutputPlane": cout, "kW": k, "kH": k, "dW": s, "dH": s, "padW": p, "padH": p,
                         "weight": state[name + ".weight"], "bias": state[name + ".bias"]})

    def inorm(name):
        return T7Object("nn.InstanceNormalization", {"weight": state[name + ".weight"], "bias": state[name + ".bias"],
                                                      "eps": 1e-5, "nOutput": int(state[name + ".weight"].size)})

    zero = padding_type == "zero"
    mods = [] if zero else [T7Object("nn.SpatialReflectionPadding", {"pad_l": reflect_pad, "pad_r": reflect_pad, "pad_t": reflect_pad,
                                                                     "pad_b": reflect_pad})]
    for i, s in enumerate(synth.parse_arch(arch)):
        n = f"l{i}"
        if s["kind"] == "conv":
            mods.append(conv(n, s["cin"], s["cout"], s["k"], s["stride"], s["pad"]))
        elif s["kind"] == "fullconv":
            mods.append(conv(n, s["cin"], s["cout"], s["k"], s["stride"], s["pad"], True))
        elif s["kind"] == "up":
            mods.append(T7Object("nn.SpatialUpSamplingNearest", {"scale_factor": s["scale"]}))
        elif s["kind"] == "res":
            rp = 1 if zero else 0
            block = seq([conv(n + ".c1", s["cin"], s["cout"], 3, 1, rp), inorm(n + ".n1"), T7Object("nn.ReLU", {"inplace": True}),
                         conv(n + ".c2", s["cout"], s["cout"], 3, 1, rp), inorm(n + ".n2")])
            skip = T7Object("nn.Identity", {}) if zero else T7Object("nn.ShaveImage", {"size": 2})
            mods.append(seq([seq([block, skip], "nn.ConcatTable"), T7Object("nn.CAddTable", {})]))
        if s["in_norm"]:
            mods.append(inorm(n + ".n"))
        if s["relu"]:
            mods.append(T7Object("nn.ReLU", {"inplace": True}))
    mods += [T7Object("nn.Tanh", {}), T7Object("nn.MulConstant", {"constant_scalar": tanh_constant}),
             T7Object("nn.TotalVariation", {"strength": 1e-6})]
    w = _Writer()
    w.obj({"opt": {"arch": arch, "padding_type": padding_type}, "iter": 60000, "model": seq(mods)})
    with open(path, "wb") as f:
        f.write(bytes(w.b))
