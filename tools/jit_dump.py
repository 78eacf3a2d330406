#!/usr/bin/env python
"""Dump the generated (JIT) source of every pass of a workload and, optionally, compile it offline with nvcc
to inspect registers / SASS:  python tools/jit_dump.py --n 30 --out /tmp/jit/p [--nvcc]"""
import argparse, ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rustqip_b200 import circuits
from rustqip_b200._abi import QipOp, marshal_ops, prec_of

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=30)
ap.add_argument("--depth", type=int, default=40)
ap.add_argument("--workload", default="random")
ap.add_argument("--dtype", default="f64")
ap.add_argument("--out", default="/tmp/jit/p")
ap.add_argument("--nvcc", action="store_true")
a = ap.parse_args()
so = os.environ.get("EMUL_SO") or os.path.join(ROOT, "tests", "native", "_build", "libplan_emul.so")
L = C.CDLL(so)
L.emul_dump_jit.restype = C.c_int
L.emul_dump_jit.argtypes = [C.c_int, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_char_p]
dtype = np.complex128 if a.dtype == "f64" else np.complex64
ops = circuits.qft(a.n) if a.workload == "qft" else circuits.random_circuit(a.n, a.depth, 0x5EED0002, "H,T,CNOT")
arr, keep = marshal_ops(ops, prec_of(dtype))
k = L.emul_dump_jit(prec_of(dtype), a.n, arr, len(ops), a.out.encode())
print("passes:", k)
if a.nvcc:
    for i in range(k):
        src = "%s_%03d.cu" % (a.out, i)
        r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-fmad=false", "-lineinfo", "-Xptxas", "-v",
                            "-cubin", "-o", src[:-3] + ".cubin", src], capture_output=True, text=True)
        print(src, [l for l in r.stderr.splitlines() if "registers" in l or "error" in l][:3])
