"""The device logic (iteres_b200/csrc/itx_logic.cuh, compiled for the host by tests/emu/) against the
reference's own outputs in tests/golden/: same byte-for-byte check as for the oracle, on every KAT
variant the device path implements, at several chunk sizes (so that the speculative record-boundary
discovery and its repair path are exercised on every input)."""
import filecmp
import os

import pytest

import emu_lib
import kats
import oracle_lib
import runners

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(k, v) for k in kats.KATS for v in kats.KATS[k]["variants"]
         if not runners.needs_host_order(*kats.KATS[k]["variants"][v])]


@pytest.mark.parametrize("chunk", [256, 4096])
@pytest.mark.parametrize("kat,variant", CASES)
def test_device_logic_matches_reference_output(kat, variant, chunk, tmp_path):
    cmd, args = kats.KATS[kat]["variants"][variant]
    vdir = os.path.join(GOLD, kat, variant)
    mk = lambda *a: emu_lib.EmuIndex(*a, chunk=chunk)
    scan = lambda ix, bam, opts: ix.scan_stream(runners.sam_to_bam_bytes(bam) if bam.endswith(".sam") else oracle_lib.inflate_bam(bam), opts)
    runners.run_itx(mk, scan, os.path.join(GOLD, kat, "input"), cmd, args, str(tmp_path))
    files = runners.expected_files(vdir)
    assert files
    for fn in files:
        got = os.path.join(str(tmp_path), fn)
        assert os.path.exists(got), fn
        assert filecmp.cmp(got, os.path.join(vdir, fn), shallow=False), "%s differs from the reference's" % fn
