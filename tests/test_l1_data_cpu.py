"""The reference-named L1 data symbols of libgmr1_b200.so (src/l1/conv.h:36-44, crc.h:36-38, punct.h:47-106) and
gmr1_puncturer_generate (punct.c:48-135) against the reference: the committed fixture made from the reference
build, and the reference build itself when it is present.  Host data only - no GPU."""
import os

import numpy as np
import pytest

import l1_data
import osmo_gmr_b200


@pytest.fixture(scope="module")
def ours():
    return l1_data.read_all(osmo_gmr_b200.LIB_PATH)


def _same(a, b):
    assert sorted(a) == sorted(b)
    for k in a:
        assert a[k].shape == b[k].shape and (a[k] == b[k]).all(), k


def test_symbols_match_committed_reference_fixture(ours):
    gold = dict(np.load(os.path.join(l1_data.ROOT, "tests", "golden", "l1_data.npz")))
    assert len(gold) == 9 * 3 + 3 + 51 + len(l1_data.GENERATE)
    _same(ours, gold)


def test_symbols_match_reference_build(ours):
    ref = os.path.join(l1_data.ROOT, "oracle", "_ref", "libgmr1_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    _same(ours, l1_data.read_all(ref))


def test_generate_rejects_rate_mismatch(ours):
    assert ours[f"generate_{len(l1_data.GENERATE) - 1}"].tolist() == [-22]
