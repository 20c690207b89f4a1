"""`.pth` -> `.bin` converter against golden hashes produced by the reference's own converter
class (tests/golden/make_converter_golden.py), the format table in include/rwkv/rwkv/format.h, and the oracle."""
import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from util import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden", "converter_golden.json")


@pytest.fixture(scope="module")
def conv():
    path = os.path.join(ROOT, "rwkv-cpp-accelerated_b200", "tools", "convert_model.py")
    spec = importlib.util.spec_from_file_location("convert_model_b200", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("case", json.load(open(GOLDEN))["cases"], ids=lambda c: "L%d_E%d_%s" % (c["n_layers"], c["n_embed"], c["dtype"]))
def test_bin_sections_match_reference_converter(conv, pkg, tmp_path, case):
    L, E = case["n_layers"], case["n_embed"]
    w = conv.synthetic_state_dict(L, E, case["seed"], getattr(torch, case["dtype"]))
    pth, out = str(tmp_path / "m.pth"), str(tmp_path / "m.bin")
    torch.save(w, pth)
    assert conv.convert(pth, out) == (L, E)
    blob = open(out, "rb").read()
    assert len(blob) == pkg.build.file_bytes(L, E)                       # format.h agrees on the total size
    assert np.frombuffer(blob[:16], "<i8").tolist() == [L, E]            # header (cpp_save_tensor.cpp:77-78)
    off = 16
    for i, (name, _) in enumerate(conv.ORDER):
        n = case["bytes"][i]
        assert hashlib.sha256(blob[off:off + n]).hexdigest() == case["sha256"][i], "tensor %d (%s) differs from the reference converter" % (i, name)
        off += n
    assert off == len(blob)


def test_quantiser_properties(conv):
    g = torch.Generator().manual_seed(5)
    w = torch.randn(48, 32, generator=g)
    q, ran, zp = conv.quantize_matrix(w)
    assert q.shape == (32, 48) and q.dtype == torch.uint8 and ran.dtype == torch.float32 and zp.dtype == torch.float32
    assert int(q.max()) >= 254 and int(q.min()) == 0                      # every input column spans the byte range
    deq = q.t().double() * ran.double() + zp.double()                     # w[out][in] ~ q*ran + zp
    err = (deq - w.double()).abs()
    assert float(err.max()) <= float(ran.max()) * 1.0001                  # truncation error below one step
    assert abs(float((deq - w.double()).mean())) < 1e-3 * float(ran.mean())   # and unbiased thanks to the zp correction


def test_converted_model_runs_in_the_oracle(conv, pkg, tmp_path):
    """The written file is a loadable model: the CPU oracle decodes it to finite logits."""
    from oracle.oracle import Oracle
    pkg.build.build_oracle()
    w = conv.synthetic_state_dict(2, 64, 3)
    pth, out = str(tmp_path / "m.pth"), str(tmp_path / "m.bin")
    torch.save(w, pth)
    conv.convert(pth, out)
    orc = Oracle(out, threads=2)
    tok = 4118
    for _ in range(3):
        logits = orc.forward(tok)
        assert np.all(np.isfinite(logits)) and logits.shape == (50277,)
        tok = int(logits.argmax())
    orc.close()
