"""CPU emulation of the tcgen05 kernel's addressing (tests/emu/conv_emu.cu) against a direct convolution, for every
layer shape of the supported archs plus edge shapes: validates K-step tables, patch segments, parity-split input,
sub-pixel phases of the transposed convolution, weight packing and the fp16 hi/lo split (3-MMA scheme)."""
import ctypes as C
import os
import subprocess

import pytest

from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU, "libfav_emu.so")
    src = os.path.join(EMU, "conv_emu.cu")
    hdr = os.path.join(ROOT, "fast-artistic-videos_b200", "csrc", "conv_plan.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                               "-Wno-deprecated-gpu-targets", "-o", so, src])
    lib = C.CDLL(so)
    lib.emu_last_error.restype = C.c_char_p
    return lib


CASES = [  # cin, cout, k, stride, pad, transposed, adj, H, W
    (7, 32, 9, 1, 4, 0, 0, 10, 140),     # l0  c9s1-32 (Cin 7 -> 8, tap pairing)
    (32, 64, 3, 2, 1, 0, 0, 12, 268),    # l1  d64 (parity-split input)
    (64, 128, 3, 2, 1, 0, 0, 8, 260),    # l2  d128
    (128, 128, 3, 1, 0, 0, 0, 6, 134),   # residual convs
    (128, 64, 3, 2, 1, 1, 1, 5, 70),     # u64 (4 sub-pixel phases)
    (64, 32, 3, 2, 1, 1, 1, 5, 130),     # u32
    (32, 3, 9, 1, 4, 0, 0, 10, 132),     # last c9s1-3 (Cout 3 -> 16, row groups)
    (128, 64, 3, 1, 1, 0, 0, 5, 130),    # c3s1-64 of the paper arch
    (3, 32, 9, 1, 4, 0, 0, 9, 20),       # image model first layer, frame narrower than one tile
    (128, 128, 3, 1, 0, 0, 0, 3, 3),     # 1x1 output
]


@pytest.mark.parametrize("case", CASES)
def test_emulated_kernel_addressing_matches_direct_conv(emu, case):
    me, mr, sm, nm = C.c_double(), C.c_double(), C.c_int(), C.c_int()
    rc = emu.emu_conv_check(*case, 1, C.byref(me), C.byref(mr), C.byref(sm), C.byref(nm))
    assert rc == 0, emu.emu_last_error().decode()
    assert me.value <= 2e-6 * max(mr.value, 1.0), (me.value, mr.value)  # ~2^-21: hi*hi + lo*hi + hi*lo
    assert sm.value <= 227 * 1024
