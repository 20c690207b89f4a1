"""The pybind module `rwkv`: the reference's surface (bindings/pybind/c_binding.cpp:158-175,
binding.py:11-69). CPU part: module builds, exports the eleven functions, tokenizer works.
GPU part mirrors the reference's tests/test_pybind.py:1-37 — with assertions."""
import importlib
import os
import sys

import numpy as np
import pytest

from util import PKG_DIR, VOCAB_DIR

PYBIND_DIR = os.path.join(PKG_DIR, "bindings", "pybind")
NAMES = {"initRwkv", "modelForward", "loadModel", "initState", "getState", "initOutput", "getOutput", "initTokenizer",
         "tokenizerEncode", "tokenizerDecode", "typicalSample"}


@pytest.fixture(scope="module")
def binding(pkg):
    assert pkg.build.build_pybind()
    if PYBIND_DIR not in sys.path:
        sys.path.insert(0, PYBIND_DIR)
    os.environ["SO_LIB_PATH"] = "rwkv"
    return importlib.import_module("binding")


def test_surface(binding):
    exported = {n for n in dir(binding.CPP_LIB) if not n.startswith("_")}
    assert NAMES <= exported


def test_tokenizer_wrapper(binding):
    tok = binding.TokenizerWrapper(vocab_path=VOCAB_DIR + "/vocab.json", merges_path=VOCAB_DIR + "/merges.txt")
    ids = tok.encode("To see the world in a grain of")
    assert ids == [1992, 923, 253, 1533, 275, 247, 13723, 273]
    assert "".join(tok.decode(i) for i in ids) == "To see the world in a grain of"
    with pytest.raises(ValueError):
        binding.TokenizerWrapper(vocab_path="/nonexistent/vocab.json", merges_path="/nonexistent/merges.txt")


@pytest.mark.gpu
def test_model_wrapper_matches_oracle(binding, make_model):
    from oracle.oracle import Oracle
    path = make_model(3, 768)
    model = binding.ModelWrapper(model_path=path)
    tok = binding.TokenizerWrapper(vocab_path=VOCAB_DIR + "/vocab.json", merges_path=VOCAB_DIR + "/merges.txt")
    prompt = tok.encode("To see the world in a grain of")
    model.init_state()
    model.load_context(prompt)
    logits, state = model.forward(prompt[-1])
    orc = Oracle(path)
    for t in prompt:
        orc.forward(t)
    ref = orc.forward(prompt[-1])
    assert logits.shape == (50277,) and logits.dtype == np.float32
    assert np.abs(logits - ref).max() / np.abs(ref).max() < 1e-3
    assert len(state) == 5 and all(s.shape == (3 * 768,) for s in state)
    for got, key in zip(state, ("xy", "aa", "bb", "pp", "dd")):
        assert np.abs(got - orc.state[key]).max() / max(np.abs(orc.state[key]).max(), 1e-6) < 1e-3
    new_token = model.sample()
    assert 0 <= new_token < 50277 and isinstance(tok.decode(new_token), str)
    model.init_state()  # really resets (the reference's initState does not)
    again, _ = model.forward(prompt[0])
    orc.reset()
    assert np.abs(again - orc.forward(prompt[0])).max() / np.abs(again).max() < 1e-3
