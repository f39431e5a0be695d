"""The kernels whose numbers are in profiles/ (the echo / broadcast instantiations `k_round<*, 0>`,
`k_commit`, `k_release`, `k_barrier`) must not change unnoticed: profiles/sass_fingerprint.json
holds a hash of their SASS (tools/sass_fingerprint.py).  A deliberate change of those kernels comes
with `python tools/sass_fingerprint.py --write` and a new measurement."""
import json
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit's cuobjdump")
def test_measured_kernels_are_the_fingerprinted_ones():
    import sass_fingerprint as F
    if not os.path.exists(F.SO):
        pytest.skip("library not built")
    want = json.load(open(F.OUT))
    got = F.fingerprint()
    import re
    # k_round<class, family 0, generic | shape-specialised>
    measured = [k for k in want if re.match(r"msd::k_round<\d, 0, (true|false)>$", k)] + \
               ["msd::k_commit", "msd::k_release", "msd::k_barrier"]
    assert len(measured) == 4 + 3 + 3
    for k in measured:
        assert got[k] == want[k], "SASS of %s changed: re-measure and run tools/sass_fingerprint.py --write" % k
