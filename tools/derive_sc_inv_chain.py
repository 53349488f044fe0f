"""Derives (and checks against pow(a, n-2, n)) the addition chain sc_inv uses (rusty_kaspa_b200/csrc/kgv_secp.cuh): the 127 leading one bits of
n - 2 through x_k = a^(2^k - 1), the remaining 129 bits by a sliding window (width <= 3) over the odd powers a, a^3, a^5, a^7.
Prints the (squarings, odd power) steps of the low part."""
import random
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
bits = bin(N - 2)[2:]
run = len(bits) - len(bits.lstrip("1"))
low = bits[run:]
steps, i, pending = [], 0, 0
while i < len(low):
    if low[i] == "0":
        pending += 1; i += 1
        continue
    w = min(3, len(low) - i)
    while low[i + w - 1] != "1":
        w -= 1
    steps.append((pending + w, int(low[i:i + w], 2))); pending = 0; i += w
if pending:
    steps.append((pending, 0))


def chain(a):
    mul = lambda x, y: (x * y) % N
    sq = lambda x, k: pow(x, 1 << k, N)
    x2 = mul(sq(a, 1), a); x3 = mul(sq(x2, 1), a); x6 = mul(sq(x3, 3), x3); x8 = mul(sq(x6, 2), x2); x14 = mul(sq(x8, 6), x6)
    x28 = mul(sq(x14, 14), x14); x56 = mul(sq(x28, 28), x28); x112 = mul(sq(x56, 56), x56); x126 = mul(sq(x112, 14), x14)
    t = mul(sq(x126, 1), a)
    tab = {1: a, 3: x2, 5: mul(x2, sq(a, 1)), 7: x3}
    for s, v in steps:
        t = sq(t, s)
        if v:
            t = mul(t, tab[v])
    return t


assert run == 127 and len(low) == 129
for _ in range(200):
    a = random.randrange(1, N)
    assert chain(a) == pow(a, N - 2, N)
print("leading ones:", run, "| low part:", sum(s for s, _ in steps), "squarings,", sum(1 for _, v in steps if v), "multiplications")
print(steps)
