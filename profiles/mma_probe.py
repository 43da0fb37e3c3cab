"""Probe of the tcgen05 shared-memory descriptor semantics (run on the GPU box).

The operand under test is a verbatim shared-memory image whose float i holds i (two runs: i % 2048 and i // 2048, both
exact in TF32); the other operand is an identity in the known-good K-major layout, so D[m][n] is the ADDRESS (in floats)
the tensor core fetched for the logical element of the probed operand:
  probing B: A = I  ->  D[m][n] = B(n, k=m)
  probing A: B = I  ->  D[m][n] = A(m, k=n)
Printed for several (LBO, SBO) settings so the address function can be read off."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simgan_b200 import _lib


def kmajor_image(X):
    E, K = X.shape
    img = torch.zeros(K // 4, E, 4)
    img[:] = X.view(E, K // 4, 4).permute(1, 0, 2)
    return img.reshape(-1).contiguous()


def run_raw(M, N, K, a_mn, b_mn, Aimg, Bimg, strides):
    dA, dB = Aimg.cuda(), Bimg.cuda()          # keep both alive: temporaries would alias in the caching allocator
    D = torch.zeros(M, N, device="cuda")
    rs = (C.c_int * 8)(*strides)
    _lib.check(_lib.lib().sg_selftest_mma(M, N, K, a_mn, b_mn, 1, _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(D), rs,
                                          _lib.current_stream()), "sg_selftest_mma")
    torch.cuda.synchronize()
    return D.cpu()


def decode_b(M, N, K, b_mn, blbo, bsbo, bstep, blt=0):
    A = kmajor_image(torch.eye(M, K))
    idx = torch.arange(N * K)
    a_str = [16 * M, 128, 32 * M]
    lo = run_raw(M, N, K, 0, b_mn, A, (idx % 2048).float(), a_str + [blbo, bsbo, bstep, 0, blt])
    hi = run_raw(M, N, K, 0, b_mn, A, (idx // 2048).float(), a_str + [blbo, bsbo, bstep, 0, blt])
    return (hi.round().long() * 2048 + lo.round().long())       # [k=m][n] -> float address


def decode_a(M, N, K, a_mn, albo, asbo, astep, alt=0):
    B = kmajor_image(torch.eye(N, K))
    idx = torch.arange(M * K)
    b_str = [16 * N, 128, 32 * N]
    lo = run_raw(M, N, K, a_mn, 0, (idx % 2048).float(), B, [albo, asbo, astep] + b_str + [alt, 0])
    hi = run_raw(M, N, K, a_mn, 0, (idx // 2048).float(), B, [albo, asbo, astep] + b_str + [alt, 0])
    return (hi.round().long() * 2048 + lo.round().long()).t()   # [k=n][m] -> float address


def show(tag, adr, es=(0, 1, 2, 3, 4, 5, 8, 16, 32, 63), ks=(0, 1, 2, 3, 4, 7, 8, 9, 16, 32)):
    print(tag)
    for k in ks:
        if k < adr.shape[0]:
            print("  k=%2d:" % k, " ".join("e%d->%d" % (e, int(adr[k, e])) for e in es if e < adr.shape[1]))


M, N, K = 128, 64, 128
print("== B operand, N=%d K=%d" % (N, K))
for (b_mn, blbo, bsbo, bstep, blt) in [(0, 16 * N, 128, 32 * N, 0), (1, 512, 16 * N, 32 * N, 1), (1, 16 * N, 512, 32 * N, 1),
                                       (1, 512, 2048, 4096, 1), (1, 128, 16 * K, 128, 2), (1, 128, 16 * K, 128, 6)]:
    show("b_mn=%d lbo=%d sbo=%d kstep=%d layout=%d" % (b_mn, blbo, bsbo, bstep, blt), decode_b(M, N, K, b_mn, blbo, bsbo, bstep, blt),
         es=(0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 31, 32, 33, 63))
M, N, K = 128, 64, 64
print("== A operand, M=%d K=%d" % (M, K))
for (a_mn, albo, asbo, astep, alt) in [(0, 16 * M, 128, 32 * M, 0), (1, 512, 16 * M, 32 * M, 1), (1, 16 * M, 512, 32 * M, 1)]:
    show("a_mn=%d lbo=%d sbo=%d kstep=%d layout=%d" % (a_mn, albo, asbo, astep, alt), decode_a(M, N, K, a_mn, albo, asbo, astep, alt),
         es=(0, 1, 2, 3, 4, 5, 6, 7, 8, 16, 31, 32, 33, 64, 127))
