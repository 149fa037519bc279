"""Generates tests/golden/*.npz FROM THE REFERENCE'S OWN SHADERS.

The reference ships no golden vectors and its Vulkan host program cannot run in this image,
but its three compute shaders compile as C++ behind oracle/ref_shim/ (`make -C oracle ref`, needs /root/reference).
Inputs come from the reference's generator restated in oracle/vkrs_oracle_host.cpp (mt19937 +
uniform_int_distribution, MultiRadixSort.cpp:121-133; std::mt19937 known-answer test in tests/test_oracle.py) with
fixed seeds; every stage recorded here -- histogram matrix and pass output of the shift-8 pass, the sorted result,
the scratch buffer and the histogram buffer after the last pass -- is what multi_radixsort_histograms.comp /
multi_radixsort.comp produce when run by MultiRadixSort::execute's loop (oracle/ref_shim/ref_driver.cpp).  The
stable order (key + payload extension) is numpy's stable argsort, which the shaders' rank order implies
(SURVEY.md section 4.3).  /root/reference does not travel: the vectors are committed, this script is how they were made.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (input generator only)
from oracle import ref as R  # noqa: E402  (the reference's shaders)

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, n, seed, max_value, nb
    ("c1_single_1000_u32", 1000, 0x5EED0001, 0xFFFFFFFF, 32),   # BASELINE.json config 1
    ("ref28_1000", 1000, 0x5EED0011, 0x0FFFFFFF, 32),           # reference distribution
    ("ragged_8193_nb3", 8193, 0x5EED0021, 0xFFFFFFFF, 3),
    ("dups_5000_nb1", 5000, 0x5EED0031, 7, 1),
    ("ref28_100003_nb32", 100003, 0x5EED0041, 0x0FFFFFFF, 32),  # the reference's shipped nb, several work groups
]


def main():
    assert R.available(), "oracle/_ref cannot be built here (/root/reference absent)"
    for name, n, seed, mx, nb in CASES:
        keys = O.generate_random(n, seed, mx)
        pc = R.push_constants(n, 8, nb)
        hist_shift8 = R.multi_histograms(keys, pc)
        pass_shift8 = R.multi_scatter(keys, hist_shift8, pc)
        buf0, buf1, hist = R.multi_sort(keys, nb)
        expect = np.sort(keys)
        assert np.array_equal(buf0, expect)
        if n <= 20000:
            assert np.array_equal(R.single_sort(keys), expect)
        order = np.argsort(keys, kind="stable").astype(np.uint32)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), keys=keys, sorted=expect, stable_order=order, nb=np.uint32(nb),
            hist_shift8=hist_shift8, pass_shift8=pass_shift8, final_buf1=buf1, final_hist=hist,
            generator=np.array("oracle/_ref"))
        print(name, n, "W=", pc.g_num_workgroups)


if __name__ == "__main__":
    main()
