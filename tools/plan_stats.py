"""Print fusion-planner statistics (passes, super-ops, elementary-op histogram) for the bench workloads.

CPU only: uses the planner through tests/native/plan_emulator.cpp (built by __graft_entry__.build()).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustqip_b200 import circuits  # noqa: E402
from rustqip_b200._abi import QipOp, marshal_ops  # noqa: E402

CASES = {0: "END", 1: "D1R_FULL", 4: "D1C_FULL", 7: "D1R_MASK", 10: "D1C_MASK", 13: "PHASE", 14: "DENSE3",
         15: "X_FULL", 18: "X_MASK", 21: "PHASEN", 22: "PHASE_J", 25: "D1R_C1", 31: "D1R_C2", 34: "PHASE_2", 37: "HAD"}


def main():
    lib = C.CDLL(os.path.join(ROOT, "tests", "native", "_build", "libplan_emul.so"))
    lib.emul_plan_stats.restype = C.c_int
    lib.emul_plan_stats.argtypes = [C.c_int, C.c_uint32, C.POINTER(QipOp), C.c_size_t, C.c_uint32, C.c_uint32,
                                    C.c_int, C.c_uint32, C.c_void_p]
    work = [("n30 f64 d40 HTCNOT", 30, circuits.random_circuit(30, 40, 0x5EED0002), 1),
            ("cfg2 n28 f64", 28, circuits.config2(), 1),
            ("qft30 f32", 30, circuits.qft(30), 0),
            ("cfg5 n30 H,CZ,CNOT", 30, circuits.random_circuit(30, 30, 0x5EED0005, "H,CZ,CNOT"), 1)]
    for name, n, ops, prec in work:
        arr, keep = marshal_ops(ops, prec)
        st = np.zeros(64, dtype=np.uint64)
        rc = lib.emul_plan_stats(prec, n, arr, len(ops), 0, 0, 1, 0, st.ctypes.data)
        assert rc == 0, rc
        hist = {}
        for i in range(16, 64):
            if st[i]:
                base = max(k for k in CASES if k <= i - 16)
                hist[f"{CASES[base]}+{i - 16 - base}"] = int(st[i])
        print(f"{name}: gates={len(ops)} passes={st[0]} singles={st[1]} micro={st[2]} super={st[4]} elems={st[5]} "
              f"cond={st[15]} diag_terms={st[7]}\n   {hist}")


if __name__ == "__main__":
    main()
