"""the JSON -> .bin converters reproduce the reference's shipped binaries byte for byte
(example/bn128/proving_key.json -> proving_key.bin, witness.json -> witness.bin).  Needs the reference checkout
(the 76 MB JSON is not staged for the GPU box), so it runs in the build container only."""
import hashlib
import json
import os

import pytest

REF = "/root/reference/example/bn128"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "proving_key.json")), reason="reference checkout not mounted")


def test_buildwitness_matches_reference_binary():
    from wasmsnark_b200.tools import buildwitness
    out = buildwitness.build(json.load(open(os.path.join(REF, "witness.json"))))
    assert out == open(os.path.join(REF, "witness.bin"), "rb").read()


@pytest.mark.slow
def test_buildpkey_matches_reference_binary():
    from wasmsnark_b200.tools import buildpkey
    out = buildpkey.build(json.load(open(os.path.join(REF, "proving_key.json"))))
    ref = open(os.path.join(REF, "proving_key.bin"), "rb").read()
    assert len(out) == len(ref) == 34339292
    assert hashlib.sha256(out).hexdigest() == hashlib.sha256(ref).hexdigest()
