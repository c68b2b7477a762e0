"""K1's host-side tables and the verdict of its arithmetic path (no GPU): the 5 x 256 fp32 table
must equal the reference's per-pixel arithmetic (vpd_dataset/common.py:52-69) bit for bit, and
whenever the library says the stem-layout kernel may COMPUTE its values, fmaf(u, scale, shift)
must round to the same bf16 as the table for every byte of every channel."""
import ctypes

import numpy as np
import pytest

from vpd_b200 import synth
from vpd_b200._lib import lib


def _bf16_bits(x):
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32)


def _reference_table(mean, std):
    u = np.arange(256)
    out = np.empty((5, 256), np.float32)
    for c in range(3):
        x = (u.astype(np.float32) / np.float32(255.0)).astype(np.float32)
        d = (x - np.float32(mean[c])).astype(np.float32)
        out[c] = (d / np.float32(std[c])).astype(np.float32)
    for c in (3, 4):
        out[c] = (u.astype(np.float64) / 255.0 - 0.5).astype(np.float32)
    return out


CASES = [synth.FS_MEAN_STD, ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
         ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)),
         ((0.3127, 0.7311, 0.1234), (0.0713, 0.3333, 0.9871))]


@pytest.mark.parametrize('mean,std', CASES)
def test_tables_and_arithmetic_constants(mean, std):
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    lut = np.zeros((5, 256), np.float32)
    sc = np.zeros(5, np.float32)
    sh = np.zeros(5, np.float32)
    ok = lib().call('vpd_assemble_tables', m, s, lut.ctypes.data, sc.ctypes.data, sh.ctypes.data)
    ref = _reference_table(mean, std)
    assert np.array_equal(lut.view(np.uint32), ref.view(np.uint32))
    if ok:
        u = np.arange(256, dtype=np.float64)
        for c in range(5):
            # exact in float64 (8-bit x 24-bit product + 24-bit addend), then ONE rounding = fmaf
            got = (u * float(sc[c]) + float(sh[c])).astype(np.float32)
            assert np.array_equal(_bf16_bits(got), _bf16_bits(ref[c])), c


def test_arithmetic_path_is_found_for_the_benchmarked_constants():
    mean, std = synth.FS_MEAN_STD
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    assert lib().call('vpd_assemble_tables', m, s, None, None, None) == 1
