"""GPU: the PTX bodies of the arithmetic primitives against Python big integers (audit hook kgv_debug_selftest)."""
import random

import pytest

import pyref

pytestmark = pytest.mark.gpu
P, N = pyref.P, pyref.N
LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72


def _operands(seed, n=512):
    rnd = random.Random(seed)
    edge = [0, 1, 2, P - 1, P, P + 1, N - 1, N, N + 1, 2**256 - 1, 2**255, 2**128, 2**128 - 1, 0xFFFFFFFF, 2**224 - 1]
    vals = edge + [rnd.randrange(2**256) for _ in range(n - len(edge))]
    a = [rnd.choice(vals) if rnd.random() < 0.3 else rnd.randrange(2**256) for _ in range(n)]
    b = [rnd.choice(vals) if rnd.random() < 0.3 else rnd.randrange(2**256) for _ in range(n)]
    return a, b


def test_wide_products(gpu_ctx):
    a, b = _operands(1)
    for got, x, y in zip(gpu_ctx.debug_selftest(0, a, b), a, b):
        assert got == x * y
    for got, x in zip(gpu_ctx.debug_selftest(1, a, b), a):
        assert got == x * x


def test_field_ops(gpu_ctx):
    a, b = _operands(2)
    M = 2**256
    for op, f in [(2, lambda x, y: x * y), (3, lambda x, y: x * x), (8, lambda x, y: x + y), (9, lambda x, y: x - y)]:
        for got, x, y in zip(gpu_ctx.debug_selftest(op, a, b), a, b):
            assert got < M and got % P == f(x, y) % P, (op, hex(x), hex(y))
    for got, x in zip(gpu_ctx.debug_selftest(7, a[:64], b[:64]), a[:64]):
        assert got % P == pow(x, P - 2, P)


def test_scalar_ops(gpu_ctx):
    a, b = _operands(3)
    a = [x % N for x in a]
    b = [y % N for y in b]
    for got, x, y in zip(gpu_ctx.debug_selftest(4, a, b), a, b):
        assert got == x * y % N
    for got, x in zip(gpu_ctx.debug_selftest(5, a, b), a):
        assert got == x * x % N
    for got, x, y in zip(gpu_ctx.debug_selftest(10, a, b), a, b):
        assert got == x * y % N
    nz = [x or 1 for x in a[:64]]
    for got, x in zip(gpu_ctx.debug_selftest(6, nz, nz), nz):
        assert got == pow(x, -1, N)


def test_glv_split(gpu_ctx):
    a, b = _operands(4)
    a = [x % N for x in a]
    for got, k in zip(gpu_ctx.debug_selftest(11, a, b), a):
        w = [(got >> (32 * i)) & 0xFFFFFFFF for i in range(16)]
        k1 = sum(w[i] << (32 * i) for i in range(5)) * (-1 if w[5] else 1)
        k2 = sum(w[8 + i] << (32 * i) for i in range(5)) * (-1 if w[13] else 1)
        assert (k1 + k2 * LAMBDA - k) % N == 0 and abs(k1) < 2**128 and abs(k2) < 2**128
