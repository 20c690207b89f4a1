"""Tokenizer parity: this repo's GPT2Tokenizer vs golden vectors produced by the reference's
own tokenizer.h (tests/golden/make_golden.py). Token ids must be bit-exact."""
import json
import os
import subprocess

import pytest

from util import ROOT, VOCAB_DIR, compile_cpp


@pytest.fixture(scope="module")
def tok_cli(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tok") / "tok_cli")
    return compile_cpp(os.path.join(ROOT, "tests", "helpers", "tok_cli.cpp"), out)


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json")) as f:
        return json.load(f)


def run_cli(cli, hex_lines):
    r = subprocess.run([cli, VOCAB_DIR + "/vocab.json", VOCAB_DIR + "/merges.txt"], input="\n".join(hex_lines) + "\n",
                       capture_output=True, text=True, check=True)
    rows = []
    for line in r.stdout.split("\n")[:len(hex_lines)]:
        ids_s, back = line.split(" | ") if " | " in line else (line.replace(" |", ""), "")
        rows.append(([int(x) for x in ids_s.split()], back.strip()))
    return rows, int(r.stderr.split()[-1])


def test_ids_and_roundtrip_match_reference(tok_cli, golden):
    cases = golden["cases"]
    rows, vocab_size = run_cli(tok_cli, [c["hex"] for c in cases])
    assert vocab_size == golden["vocab_size"] == 50277
    bad = [(c["hex"], c["ids"], ids) for c, (ids, _) in zip(cases, rows) if ids != c["ids"]]
    assert not bad, "first mismatch: %r" % (bad[0],)
    for c, (_, back) in zip(cases, rows):
        assert back == c["decoded_hex"]


def test_survey_known_answers(tok_cli):
    """SURVEY.md section 4 table (produced by the reference tokenizer)."""
    kat = {
        "To see the world in a grain of": [1992, 923, 253, 1533, 275, 247, 13723, 273],
        "\n\n### Response:": [187, 187, 4118, 19371, 27],
        "    indented": [209, 209, 209, 801, 8006],  # the lost first merge (quirk Q1)
        "café 日本": [68, 2320, 860, 209, 49868],
        "I'll we've don't": [42, 1833, 359, 1849, 1053, 626],
    }
    rows, _ = run_cli(tok_cli, [k.encode().hex() for k in kat])
    for (text, want), (ids, back) in zip(kat.items(), rows):
        assert ids == want, text
        assert bytes.fromhex(back).decode() == text


def test_missing_files_return_nullopt(tok_cli, tmp_path):
    r = subprocess.run([tok_cli, str(tmp_path / "nope.json"), str(tmp_path / "nope.txt")], input="", capture_output=True, text=True)
    assert r.returncode == 2 and "could not open" in r.stderr
