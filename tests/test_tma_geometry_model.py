"""Tile geometry of the fused pass against a model of the TMA tiled copy (no GPU).

The pass kernels address the shared-memory image of a tile as `slot(t) = t ^ ((t >> 3) & 7)` (f64, 16-byte units;
f32: `t ^ (((t >> 4) & 7) << 1)`, 8-byte units) for the tile-local index t.  That image is produced by
`cp.async.bulk.tensor.5d` boxes whose tensor map (rustqip_b200/csrc/tile_kernel.cu: make_tile_map), box offsets
(planner.cpp: chunk_off) and coordinates (jit_codegen.cpp: box_coords) are restated here together with the documented
semantics of a tiled TMA copy with CU_TENSOR_MAP_SWIZZLE_128B (box elements row-major with dimension 0 innermost; the
16-byte chunk index within a 128-byte row XORed with the row index modulo 8).  For every geometry the planner can
choose -- L = 5 / 6 contiguous low bits is what runs on hardware today, L = 4 (f64) is the candidate of DESIGN.md
section 8 item 0 -- the model must put amplitude idx(t) of the state at slot(t) of the image."""
import numpy as np
import pytest


def tile_base(tile, L, hi_pos):
    base = tile << L
    for q in hi_pos:  # insert a zero bit at every high tile bit position, ascending (jit_codegen.cpp: tile_base)
        base = ((base >> q) << (q + 1)) | (base & ((1 << q) - 1))
    return base


def model_image(f64, n_local, T, L, hi_pos, tile, rng):
    """Returns (expected, got): arrays of global amplitude indices per shared-memory slot."""
    amp = 16 if f64 else 8
    esz = 8 if f64 else 4          # tensor-map element = one real
    low3 = 3 if f64 else 4         # log2(amplitudes per 128-byte row)
    m = T - L
    h1, h2, h3 = hi_pos[0], hi_pos[1], hi_pos[2]
    # make_tile_map
    dims = [128 // esz, 1 << (h1 - low3), 1 << (h2 - h1), 1 << (h3 - h2), 1 << (n_local - h3)]
    strides = [esz, 128, amp << h1, amp << h2, amp << h3]   # bytes; stride of dim 0 is the element size
    box = [128 // esz, 1 << (L - low3), 2, 2, 2]
    assert box[1] <= 256 and all(b <= d for b, d in zip(box, dims))
    box_bytes = amp << (L + 3)
    nbox = 1 << (m - 3)
    chunk_off = [sum(((ch >> i) & 1) << hi_pos[i] for i in range(m)) for ch in range(1 << m)]
    base = tile_base(tile, L, hi_pos)
    n_amp = 1 << T
    got = np.full(n_amp, -1, dtype=np.int64)
    for b in range(nbox):
        idx = base + chunk_off[b << 3]                                   # p.box_off[b]
        c = [0,                                                          # box_coords
             (idx >> low3) & ((1 << (h1 - low3)) - 1),
             (idx >> h1) & ((1 << (h2 - h1)) - 1),
             (idx >> h2) & ((1 << (h3 - h2)) - 1),
             idx >> h3]
        lin = 0
        for i4 in range(box[4]):
            for i3 in range(box[3]):
                for i2 in range(box[2]):
                    for i1 in range(box[1]):
                        for i0 in range(box[0]):
                            gbyte = sum((c[k] + i) * strides[k] for k, i in enumerate((i0, i1, i2, i3, i4)))
                            sbyte = b * box_bytes + lin * esz
                            sbyte ^= ((sbyte >> 7) & 7) << 4             # 128-byte swizzle
                            lin += 1
                            if gbyte % amp == 0:                          # first real of an amplitude
                                assert sbyte % amp == 0
                                got[sbyte // amp] = gbyte // amp
    expected = np.empty(n_amp, dtype=np.int64)
    for t in range(n_amp):
        idx = base + (t & ((1 << L) - 1)) + sum(((t >> (L + i)) & 1) << hi_pos[i] for i in range(m))
        slot = t ^ ((t >> 3) & 7) if f64 else t ^ (((t >> 4) & 7) << 1)
        expected[slot] = idx
    return expected, got


@pytest.mark.parametrize("f64,T,L", [(True, 12, 5), (False, 13, 6),      # what runs on hardware
                                     (True, 12, 4), (True, 12, 6), (False, 13, 5), (False, 13, 7),
                                     (True, 9, 4), (True, 8, 5), (False, 10, 6)])
def test_tma_boxes_produce_the_image_the_kernels_address(f64, T, L):
    rng = np.random.default_rng(100 * T + L)
    n_local = 24
    m = T - L
    for _ in range(3):
        hi_pos = sorted(int(x) for x in rng.choice(np.arange(L, n_local), m, replace=False))
        tile = int(rng.integers(0, 1 << (n_local - T)))
        expected, got = model_image(f64, n_local, T, L, hi_pos, tile, rng)
        assert np.array_equal(expected, got), (hi_pos, tile)
