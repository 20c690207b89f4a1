"""typical() parity with the reference sampler (NumCpp) on synthetic logits: the drawn token
sequence of a default-seeded process must be identical (tests/golden/make_golden.py)."""
import json
import os
import subprocess

import pytest

from util import ROOT, compile_cpp


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("samp") / "sampler_cli")
    return compile_cpp(os.path.join(ROOT, "tests", "helpers", "sampler_cli.cpp"), out)


def test_sequences_match_reference(cli):
    with open(os.path.join(ROOT, "tests", "golden", "sampler_golden.json")) as f:
        runs = json.load(f)["runs"]
    for run in runs:
        r = subprocess.run([cli, str(run["n"]), repr(run["temp"]), repr(run["tau"]), repr(run["scale"])],
                           capture_output=True, text=True, check=True)
        got = [int(x) for x in r.stdout.split()]
        assert got == run["tokens"], run
