#!/usr/bin/env python
"""Regenerates the fixtures in tests/golden/.

The reference is Rust and cannot run in this image, so there are no reference-generated
vectors.  What is committed here:
  * cswap_readme.json -- the known answer of BASELINE.json configs[0] (README CSWAP circuit),
    DERIVED BY HAND in SURVEY.md section 8 (Q-KA) from the reference's decomposition rules:
    192 pipeline entries, init index 4, amplitudes +0.5 @ {4,32,68}, -0.5 @ {96}.  This script
    only re-checks that the oracle reproduces it.
  * oracle_vectors.npz -- ORACLE-generated (not reference-generated) input/output pairs for
    small seeded circuits, so that the GPU path can be checked on a box without rebuilding
    anything: a regression net, not an independent pin.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import qip_oracle as qo  # noqa: E402
from rustqip_b200 import circuits  # noqa: E402
from rustqip_b200.builder import readme_cswap_circuit  # noqa: E402


def main():
    b, q, ra, rb, handle = readme_cswap_circuit()
    idx = b.initial_index([(ra, 0b000), (rb, 0b001)])
    st = qo.run_pipeline(b.n(), b.unitary_ops(), idx)
    ka = {"n": 7, "pipeline_depth": 192, "unitary_entries": 191, "init_index": 4,
          "amplitudes": {"4": [0.5, 0.0], "32": [0.5, 0.0], "68": [0.5, 0.0], "96": [-0.5, 0.0]},
          "p_q0": 0.5, "p_q1": 0.5,
          "post_measure_0": {"4": 0.7071067811865476, "32": 0.7071067811865476},
          "post_measure_1": {"68": 0.7071067811865476, "96": -0.7071067811865476},
          "source": "SURVEY.md section 8 Q-KA (derived from README.md:31-55, builder_traits.rs:408-476,505-538, builder.rs:409-421,754-764)"}
    assert b.pipeline_depth() == ka["pipeline_depth"] and idx == ka["init_index"]
    for i in range(128):
        want = complex(*ka["amplitudes"].get(str(i), [0.0, 0.0]))
        assert abs(st[i] - want) < 1e-12, (i, st[i])
    json.dump(ka, open(os.path.join(HERE, "cswap_readme.json"), "w"), indent=1)

    out = {}
    for name, n, ops, dtype in [("rand_htcnot_n10_f64", 10, circuits.random_circuit(10, 6, 0x5EED0002), np.complex128),
                                ("rand_hczcnot_n9_f32", 9, circuits.random_circuit(9, 5, 0x5EED0005, "H,CZ,CNOT"), np.complex64),
                                ("qft_n8_f32", 8, circuits.qft(8), np.complex64),
                                ("dense4_n8_f64", 8, circuits.config4(8, blocks=4), np.complex128)]:
        psi = circuits.random_state(n, 0x5EED0003, dtype)
        out[name + "_in"] = psi
        out[name + "_out"] = qo.run_pipeline(n, ops, state=psi, dtype=dtype)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
