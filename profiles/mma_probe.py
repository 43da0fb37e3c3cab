"""Probe of the tcgen05 shared-memory descriptor semantics (run on the GPU box): A = identity (K-major, known good),
B = a verbatim shared-memory image whose float i holds i (in two runs: i % 2048 and i // 2048, both exact in TF32), so
D[m][n] is the ADDRESS (in floats) the tensor core fetched for the logical B element (n, k = m)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simgan_b200 import _lib

def kmajor_image(X):
    E, K = X.shape
    img = torch.zeros(K // 4, E, 4)
    img[:] = X.view(E, K // 4, 4).permute(1, 0, 2)
    return img.reshape(-1).contiguous()

def run_raw(M, N, K, a_mn, b_mn, Aimg, Bimg, strides):
    D = torch.zeros(M, N, device="cuda")
    rs = (C.c_int * 6)(*strides)
    _lib.check(_lib.lib().sg_selftest_mma(M, N, K, a_mn, b_mn, 1, _lib.ptr(Aimg.cuda()), _lib.ptr(Bimg.cuda()), _lib.ptr(D), rs,
                                          _lib.current_stream()))
    torch.cuda.synchronize()
    return D.cpu()

def decode_b(M, N, K, b_mn, blbo, bsbo, bstep):
    A = kmajor_image(torch.eye(M, K))
    idx = torch.arange(N * K)
    a_str = [16 * M, 128, 32 * M]
    lo = run_raw(M, N, K, 0, b_mn, A, (idx % 2048).float(), a_str + [blbo, bsbo, bstep])
    hi = run_raw(M, N, K, 0, b_mn, A, (idx // 2048).float(), a_str + [blbo, bsbo, bstep])
    return (hi.long() * 2048 + lo.long())       # [k=m][n] -> float address

M, N, K = 128, 64, 128
for (b_mn, blbo, bsbo, bstep) in [(0, 16 * N, 128, 32 * N), (1, 128, 16 * K, 128), (1, 16 * K, 128, 128), (1, 256, 1024, 128), (1, 1024, 256, 128)]:
    adr = decode_b(M, N, K, b_mn, blbo, bsbo, bstep)
    print("b_mn=%d lbo=%d sbo=%d kstep=%d" % (b_mn, blbo, bsbo, bstep))
    for k in (0, 1, 2, 7, 8, 9, 16):
        print("  k=%2d:" % k, " ".join("n%d->%d" % (n, int(adr[k, n])) for n in (0, 1, 2, 3, 4, 5, 8, 16, 32, 63)))
