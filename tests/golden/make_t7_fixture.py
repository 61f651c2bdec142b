"""Writes tests/golden/tiny_video_model.t7: a Torch7 binary checkpoint assembled BY HAND from the published description of
torch7's File.lua serialisation (little endian; int32 type tags 0 nil / 1 number / 2 string / 3 table / 4 torch object /
5 boolean; tables and torch objects carry an int32 reference index and are written once; torch objects: "V 1", class
name, payload).  It does NOT import fav_b200.t7 (whose writer the reader was so far only validated against), and it uses
what real checkpoints contain and that writer never emits: every parameter tensor is a VIEW into ONE flat FloatStorage at
its own storageOffset (what nn.Module:getParameters() leaves behind, train_video.lua:131), gradWeight / gradBias tensors,
an `output` DoubleTensor, number keys mixed with string keys, and a back-reference to a table written earlier.

Model: {SpatialReflectionPadding(4), c3s1-4 on 7 input channels, IN, ReLU, one R4 residual block, c3s1-3, Tanh,
MulConstant(150), TotalVariation} -- arch "c3s1-4,R4,c3s1-3", reflect pad 4.  Parameter values are closed-form
(value = 0.001 * flat index - 0.2) so the test recomputes them without this script.
"""
import os
import struct

import numpy as np

out = bytearray()
next_index = [0]


def i32(v): out.extend(struct.pack("<i", v))
def i64(v): out.extend(struct.pack("<q", v))
def raw_string(s): i32(len(s)); out.extend(s.encode("ascii"))
def number(v): i32(1); out.extend(struct.pack("<d", float(v)))
def string(s): i32(2); raw_string(s)
def boolean(b): i32(5); i32(1 if b else 0)
def nil(): i32(0)


def new_index():
    next_index[0] += 1
    return next_index[0]


def table(pairs, index=None):
    """pairs: list of (key writer, value writer) thunks"""
    i32(3)
    idx = index if index is not None else new_index()
    i32(idx)
    i32(len(pairs))
    for k, v in pairs:
        k(); v()
    return idx


def backref_table(idx): i32(3); i32(idx)


def torch_object(cls, payload):
    i32(4); i32(new_index()); raw_string("V 1"); raw_string(cls); payload()


# ---- one flat parameter storage, as getParameters() leaves it
shapes = [("l0.weight", (4, 7, 3, 3)), ("l0.bias", (4,)), ("l0.n.weight", (4,)), ("l0.n.bias", (4,)),
          ("l1.c1.weight", (4, 4, 3, 3)), ("l1.c1.bias", (4,)), ("l1.n1.weight", (4,)), ("l1.n1.bias", (4,)),
          ("l1.c2.weight", (4, 4, 3, 3)), ("l1.c2.bias", (4,)), ("l1.n2.weight", (4,)), ("l1.n2.bias", (4,)),
          ("l2.weight", (3, 4, 3, 3)), ("l2.bias", (3,))]
offsets, total = {}, 0
for name, shp in shapes:
    offsets[name] = total
    total += int(np.prod(shp))
flat = (0.001 * np.arange(total, dtype=np.float64) - 0.2).astype("<f4")
storage_index = [None]


def flat_storage():
    i32(4)
    if storage_index[0] is not None:  # written once, referenced by index afterwards
        i32(storage_index[0])
        return
    storage_index[0] = new_index()
    i32(storage_index[0]); raw_string("V 1"); raw_string("torch.FloatStorage"); i64(total); out.extend(flat.tobytes())


def param_tensor(name):
    shp = dict(shapes)[name]

    def payload():
        i32(len(shp))
        for s in shp: i64(s)
        stride = [int(np.prod(shp[k + 1:])) for k in range(len(shp))]
        for s in stride: i64(s)
        i64(offsets[name] + 1)  # 1-based storageOffset
        flat_storage()
    return lambda: torch_object("torch.FloatTensor", payload)


def small_tensor(cls, storage_cls, fmt, values):
    def payload():
        i32(1); i64(len(values)); i64(1); i64(1)
        torch_object(storage_cls, lambda: (i64(len(values)), out.extend(struct.pack("<%d%s" % (len(values), fmt), *values))))
    return lambda: torch_object(cls, payload)


def empty_tensor(cls):
    return lambda: torch_object(cls, lambda: (i32(0), i64(1), nil()))  # ndim 0, storageOffset 1, no storage


K = lambda s: (lambda: string(s))
N = lambda v: (lambda: number(v))


def module(cls, fields):
    return lambda: torch_object(cls, lambda: table(fields))


def conv(name, cin, cout, pad):
    return module("nn.SpatialConvolution", [
        (K("nInputPlane"), N(cin)), (K("nOutputPlane"), N(cout)), (K("kW"), N(3)), (K("kH"), N(3)), (K("dW"), N(1)), (K("dH"), N(1)),
        (K("padW"), N(pad)), (K("padH"), N(pad)), (K("weight"), param_tensor(name + ".weight")), (K("bias"), param_tensor(name + ".bias")),
        (K("gradWeight"), small_tensor("torch.FloatTensor", "torch.FloatStorage", "f", [0.0, 0.0])),
        (K("gradBias"), empty_tensor("torch.FloatTensor")), (K("output"), small_tensor("torch.DoubleTensor", "torch.DoubleStorage", "d", [1.0, 2.0, 3.0])),
        (K("train"), lambda: boolean(False))])


def inorm(name):
    return module("nn.InstanceNormalization", [(K("weight"), param_tensor(name + ".weight")), (K("bias"), param_tensor(name + ".bias")),
                                               (K("eps"), N(1e-5)), (K("nOutput"), N(4)), (K("prev_N"), N(-1))])


def sequential(cls, mods):
    return module(cls, [(K("modules"), lambda: table([(N(i + 1), m) for i, m in enumerate(mods)])), (K("train"), lambda: boolean(False))])


res_block = sequential("nn.Sequential", [
    sequential("nn.ConcatTable", [sequential("nn.Sequential", [conv("l1.c1", 4, 4, 0), inorm("l1.n1"), module("nn.ReLU", [(K("inplace"), lambda: boolean(True))]),
                                                              conv("l1.c2", 4, 4, 0), inorm("l1.n2")]),
                                  module("nn.ShaveImage", [(K("size"), N(2))])]),
    module("nn.CAddTable", [(K("inplace"), lambda: boolean(False))])])
model = sequential("nn.Sequential", [
    module("nn.SpatialReflectionPadding", [(K("pad_l"), N(4)), (K("pad_r"), N(4)), (K("pad_t"), N(4)), (K("pad_b"), N(4))]),
    conv("l0", 7, 4, 1), inorm("l0.n"), module("nn.ReLU", [(K("inplace"), lambda: boolean(True))]),
    res_block, conv("l2", 4, 3, 1), module("nn.Tanh", []), module("nn.MulConstant", [(K("constant_scalar"), N(150)), (K("inplace"), lambda: boolean(False))]),
    module("nn.TotalVariation", [(K("strength"), N(1e-6))])])

opt_index = [None]


def opt_table():
    opt_index[0] = table([(K("arch"), lambda: string("c3s1-4,R4,c3s1-3")), (K("padding_type"), lambda: string("reflect-start")),
                          (K("tanh_constant"), N(150)), (K("use_instance_norm"), N(1))])


# the checkpoint table of train_video.lua:507-541: {opt, train_loss_history (number keys), iter, model, opt again by reference}
table([(K("opt"), opt_table),
       (K("train_loss_history"), lambda: table([(N(1), N(12.5)), (N(2), N(11.25))])),
       (K("iter"), N(60000)),
       (K("model"), model),
       (K("opt_again"), lambda: backref_table(opt_index[0]))])

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_video_model.t7")
open(path, "wb").write(bytes(out))
print(path, len(out), "bytes")
