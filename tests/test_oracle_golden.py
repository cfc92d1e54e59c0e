"""The CPU oracle against the reference's own outputs (tests/golden/, produced by the unmodified
reference binary): every text output file must be byte-identical, for every KAT and variant."""
import filecmp
import os

import pytest

import kats
import runners

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(k, v) for k in kats.KATS for v in kats.KATS[k]["variants"]]


@pytest.mark.parametrize("kat,variant", CASES)
def test_oracle_matches_reference_output(kat, variant, tmp_path):
    cmd, args = kats.KATS[kat]["variants"][variant]
    vdir = os.path.join(GOLD, kat, variant)
    runners.run_oracle(os.path.join(GOLD, kat, "input"), cmd, args, str(tmp_path))
    files = runners.expected_files(vdir, skip_bed=True)      # the oracle restates the counting path; bed lines are the product's host pass
    assert files, "golden directory is empty"
    for fn in files:
        got = os.path.join(str(tmp_path), fn)
        assert os.path.exists(got), fn
        assert filecmp.cmp(got, os.path.join(vdir, fn), shallow=False), "%s differs from the reference's" % fn
