"""Regenerate tests/golden/converter_golden.json (run in the build container, where /root/reference exists).

Builds a seeded synthetic RWKV-4 state dict, runs the REFERENCE's own converter class on it
(converter/convert_model.py: ConvertRWKV, imported from /root/reference, unmodified) and records the SHA-256 of
every tensor it would hand to its writer (cpp_save_tensor.cpp:75-95), in file order. tests/test_converter.py
checks that rwkv-cpp-accelerated_b200/tools/convert_model.py writes a file with exactly those sections."""
import hashlib
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/converter/convert_model.py"
CASES = [(2, 64, 11, "float32"), (1, 96, 12, "bfloat16")]


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ours = load(os.path.join(ROOT, "rwkv-cpp-accelerated_b200", "tools", "convert_model.py"), "ours_convert")
    ref = load(REF, "ref_convert")
    out = {"generator": "tests/golden/make_converter_golden.py", "reference": "converter/convert_model.py (ConvertRWKV)",
           "torch": torch.__version__, "cases": []}
    for L, E, seed, dt in CASES:
        w = ours.synthetic_state_dict(L, E, seed, getattr(torch, dt))
        m = ref.ConvertRWKV(dict(w), E, L)   # the class sets consumed weights to None: give it a copy
        seq = [m.rx, m.emb, m.cudalnin, m.emptyState[0], m.emptyState[1], m.emptyState[2], m.emptyState[3], m.emptyState[4],
               m.buffer0, m.buffer1, m.buffer2, m.buffer3, m.mixk, m.mixv, m.mixr,
               m.attkeyweights, m.attvalueweights, m.attreceptanceweights, m.attkeyranges, m.attvalueranges,
               m.attreceptanceranges, m.attkeyzp, m.attvaluezp, m.attreceptancezp,
               m.attoutputweights, m.attoutputranges, m.attoutputzp, m.mixffnk, m.mixffnr,
               m.ffnkeyweights, m.ffnvalueweights, m.ffnreceptanceweights, m.ffnkeyranges, m.ffnvalueranges,
               m.ffnreceptanceranges, m.ffnkeyzp, m.ffnvaluezp, m.ffnreceptancezp,
               m.ffnkbuf, m.ffnvbuf, m.ffkeybuffer, m.decay, m.bonus, m.cudahead, m.cudaheadr, m.cudaheadzp]
        assert len(seq) == 46
        sha = [hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest() for t in seq]
        nbytes = [int(t.numel() * t.element_size()) for t in seq]
        out["cases"].append({"n_layers": L, "n_embed": E, "seed": seed, "dtype": dt, "sha256": sha, "bytes": nbytes})
        print("case L=%d E=%d %s: %d tensors, %d bytes" % (L, E, dt, len(seq), sum(nbytes)))
    with open(os.path.join(ROOT, "tests", "golden", "converter_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    sys.exit(main())
