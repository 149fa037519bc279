"""Generates tests/golden/*.npz.

The reference ships no golden vectors and cannot run in this image, so these
fixtures are NOT reference outputs: inputs come from the reference's own generator restated in
oracle/vkrs_oracle_host.cpp (mt19937 + uniform_int_distribution, MultiRadixSort.cpp:121-133) with
fixed seeds, and the expected outputs are numpy's sort / stable argsort of them -- the property the
reference's testSort checks (MultiRadixSort.cpp:148-161).  They pin the oracle and the CUDA path
against accidental change and record the stage intermediates (histogram matrix, pass outputs) the
oracle produced when it was validated.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, n, seed, max_value, nb
    ("c1_single_1000_u32", 1000, 0x5EED0001, 0xFFFFFFFF, 32),   # BASELINE.json config 1
    ("ref28_1000", 1000, 0x5EED0011, 0x0FFFFFFF, 32),           # reference distribution
    ("ragged_8193_nb3", 8193, 0x5EED0021, 0xFFFFFFFF, 3),
    ("dups_5000_nb1", 5000, 0x5EED0031, 7, 1),
]


def main():
    for name, n, seed, mx, nb in CASES:
        keys = O.generate_random(n, seed, mx)
        pc = O.push_constants(n, 8, nb)
        hist_shift8 = O.multi_histograms(keys, pc)
        pass_shift8 = O.multi_scatter(keys, hist_shift8, pc)
        buf0, buf1, hist = O.multi_sort(keys, nb)
        expect = np.sort(keys)
        assert np.array_equal(buf0, expect)
        order = np.argsort(keys, kind="stable").astype(np.uint32)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), keys=keys, sorted=expect, stable_order=order, nb=np.uint32(nb),
            hist_shift8=hist_shift8, pass_shift8=pass_shift8, final_buf1=buf1, final_hist=hist)
        print(name, n, "W=", pc.g_num_workgroups)


if __name__ == "__main__":
    main()
