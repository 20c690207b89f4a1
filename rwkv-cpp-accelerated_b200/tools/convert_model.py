"""RWKV-4 `.pth` checkpoint -> reference-format uint8 `.bin`, written from scratch.

Produces byte-for-byte what the reference's converter produces (converter/convert_model.py:14-176 builds
the tensors, converter/cpp_save_tensor.cpp:75-95 writes them) without its torch C++ extension:
    python convert_model.py model.pth [model.bin]

File layout (include/rwkv/rwkv/format.h is the table this follows): two little-endian int64
{n_layers, n_embed}, then the 46 tensors in enum order, unpadded - including the scratch buffers and the
zero / -1e30 state sections the engines ignore (convert_model.py:19-25, 99-106).

Q8 scheme (convert_model.py:108-119) for a Linear weight W[out][in]: per INPUT column j
    lo_j = min_out W, ran_j = (max_out W - lo_j) / 255, q = trunc((W - lo_j) / ran_j) in 0..255,
    zp_j = lo_j + mean_out(frac((W - lo_j) / ran_j)) * ran_j          (bias correction of the truncation)
stored transposed [in][out] uint8 with ranges[in], zp[in] as f32; all arithmetic in f64.
Pinned bit-exact against the reference's own Python class: tests/golden/converter_golden.json.
"""
import struct
import sys

import numpy as np
import torch

VOCAB = 50277

# (name in the file table, dtype) in enum order - checked against format.h by tests/test_converter.py
ORDER = [
    ("xbuf", "f8"), ("embed", "f4"), ("layernorms", "f8"), ("state_xy", "f8"), ("state_aa", "f8"), ("state_bb", "f8"),
    ("state_pp", "f8"), ("state_dd", "f8"), ("buffer1", "f8"), ("buffer2", "f4"), ("buffer3", "f4"), ("buffer4", "f4"),
    ("mix_k", "f8"), ("mix_v", "f8"), ("mix_r", "f8"), ("km", "u1"), ("vm", "u1"), ("rm", "u1"), ("kr", "f4"),
    ("vr", "f4"), ("rr", "f4"), ("o1", "f4"), ("o2", "f4"), ("o3", "f4"), ("att_out", "u1"), ("att_out_r", "f4"),
    ("att_out_o", "f4"), ("ffn_mix_k", "f8"), ("ffn_mix_v", "f8"), ("ffn_k", "u1"), ("ffn_v", "u1"), ("ffn_r", "u1"),
    ("ffn_kr", "f4"), ("ffn_vr", "f4"), ("ffn_rr", "f4"), ("ffn_ko", "f4"), ("ffn_vo", "f4"), ("ffn_ro", "f4"),
    ("ffn_k_buffer", "f8"), ("ffn_v_buffer", "f8"), ("ffn_r_buffer", "f4"), ("decay", "f8"), ("bonus", "f8"),
    ("head", "u1"), ("head_r", "f4"), ("head_o", "f4"),
]


def quantize_matrix(w):
    """W[out][in] (any float dtype) -> (q [in][out] uint8, ranges [in] f32, zp [in] f32)."""
    x = w.to(torch.float64)
    lo = x.min(dim=0).values
    span = x - lo
    ran = span.max(dim=0).values / 255
    scaled = span / ran
    zp = lo + scaled.frac().mean(dim=0) * ran
    return scaled.t().to(torch.uint8).contiguous(), ran.to(torch.float32), zp.to(torch.float32)


def model_dims(w):
    n_embed = int(w["blocks.0.att.key.weight"].shape[0])
    n_layers = sum(1 for k in w if k.startswith("blocks.") and k.endswith("ln1.bias"))
    return n_layers, n_embed


def build_tensors(w):
    """state dict -> (n_layers, n_embed, dict name -> contiguous torch tensor of the file dtype)."""
    L, E = model_dims(w)
    out = {}

    def per_layer(fmt):
        return [w[fmt.format(i)] for i in range(L)]

    def vec64(fmt):
        return torch.stack([t.squeeze() for t in per_layer(fmt)]).to(torch.float64).contiguous()

    ln_names = ["blocks.0.ln0.weight", "blocks.0.ln0.bias"]
    for i in range(L):
        ln_names += ["blocks.%d.ln1.weight" % i, "blocks.%d.ln1.bias" % i, "blocks.%d.ln2.weight" % i, "blocks.%d.ln2.bias" % i]
    ln_names += ["ln_out.weight", "ln_out.bias"]

    out["xbuf"] = torch.arange(E, dtype=torch.float64)
    out["embed"] = w["emb.weight"].to(torch.float32).contiguous()
    out["layernorms"] = torch.stack([w[n] for n in ln_names]).to(torch.float64).contiguous()
    zeros = torch.zeros(L, E, dtype=torch.float64)
    out["state_xy"], out["state_aa"], out["state_bb"], out["state_dd"] = zeros, zeros, zeros, zeros
    # the reference builds its placeholder from a float32 tensor: the stored value is float32(-1e30) widened
    out["state_pp"] = torch.full((L, E), -1e30, dtype=torch.float32).to(torch.float64)
    out["buffer1"] = torch.arange(E, dtype=torch.float64)
    out["buffer2"] = torch.arange(VOCAB, dtype=torch.float32)
    out["buffer3"] = torch.arange(E, dtype=torch.float32)
    out["buffer4"] = torch.arange(E, dtype=torch.float32)
    out["mix_k"] = vec64("blocks.{}.att.time_mix_k")
    out["mix_v"] = vec64("blocks.{}.att.time_mix_v")
    out["mix_r"] = vec64("blocks.{}.att.time_mix_r")
    out["ffn_mix_k"] = vec64("blocks.{}.ffn.time_mix_k")
    out["ffn_mix_v"] = vec64("blocks.{}.ffn.time_mix_r")  # the file's "ffnmixv" slot holds time_mix_r (convert_model.py:55-56)
    out["decay"] = -torch.exp(vec64("blocks.{}.att.time_decay"))
    out["bonus"] = vec64("blocks.{}.att.time_first")

    def family(fmt, qn, rn, on):
        parts = [quantize_matrix(t) for t in per_layer(fmt)]
        out[qn] = torch.stack([p[0] for p in parts]).contiguous()
        out[rn] = torch.stack([p[1] for p in parts]).contiguous()
        out[on] = torch.stack([p[2] for p in parts]).contiguous()

    family("blocks.{}.att.key.weight", "km", "kr", "o1")
    family("blocks.{}.att.value.weight", "vm", "vr", "o2")
    family("blocks.{}.att.receptance.weight", "rm", "rr", "o3")
    family("blocks.{}.att.output.weight", "att_out", "att_out_r", "att_out_o")
    family("blocks.{}.ffn.key.weight", "ffn_k", "ffn_kr", "ffn_ko")
    family("blocks.{}.ffn.value.weight", "ffn_v", "ffn_vr", "ffn_vo")
    family("blocks.{}.ffn.receptance.weight", "ffn_r", "ffn_rr", "ffn_ro")
    out["ffn_k_buffer"] = torch.arange(E, dtype=torch.float64)
    out["ffn_v_buffer"] = torch.arange(E, dtype=torch.float64)
    out["ffn_r_buffer"] = torch.arange(4 * E, dtype=torch.float32)
    out["head"], out["head_r"], out["head_o"] = quantize_matrix(w["head.weight"])
    return L, E, out


def expected_elems(name, L, E):
    big = {"embed": VOCAB * E, "head": VOCAB * E, "layernorms": 4 * (L + 1) * E, "buffer2": VOCAB}
    if name in big:
        return big[name]
    if name in ("xbuf", "buffer1", "buffer3", "buffer4", "ffn_k_buffer", "ffn_v_buffer", "head_r", "head_o"):
        return E
    if name == "ffn_r_buffer":
        return 4 * E
    if name in ("km", "vm", "rm", "att_out", "ffn_r"):
        return L * E * E
    if name in ("ffn_k", "ffn_v"):
        return 4 * L * E * E
    if name in ("ffn_vr", "ffn_vo"):
        return 4 * L * E
    return L * E


def write_bin(path, L, E, tensors):
    np_dtype = {"f8": np.float64, "f4": np.float32, "u1": np.uint8}
    with open(path, "wb") as f:
        f.write(struct.pack("<qq", L, E))
        for name, dt in ORDER:
            a = tensors[name].contiguous().numpy()
            if a.dtype != np_dtype[dt] or a.size != expected_elems(name, L, E):
                raise ValueError("tensor %s: dtype %s / %d elements, expected %s / %d"
                                 % (name, a.dtype, a.size, dt, expected_elems(name, L, E)))
            f.write(a.tobytes())


def convert(pth_path, bin_path):
    w = torch.load(pth_path, map_location="cpu")
    for k in ("emb.weight", "ln_out.weight", "ln_out.bias", "blocks.0.ln0.weight", "blocks.0.ln0.bias", "head.weight"):
        if k not in w:
            raise ValueError("not an RWKV-4 checkpoint: missing %s" % k)
    L, E, tensors = build_tensors(w)
    print("n_layers %d  n_embed %d" % (L, E))
    write_bin(bin_path, L, E, tensors)
    return L, E


def synthetic_state_dict(L, E, seed, dtype=torch.float32):
    """A random RWKV-4 shaped state dict (tests and the golden generator share it)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dtype)

    w = {"emb.weight": rnd(VOCAB, E, scale=0.1), "head.weight": rnd(VOCAB, E, scale=E ** -0.5),
         "ln_out.weight": 1 + rnd(E, scale=0.1), "ln_out.bias": rnd(E, scale=0.1),
         "blocks.0.ln0.weight": 1 + rnd(E, scale=0.1), "blocks.0.ln0.bias": rnd(E, scale=0.1)}
    for i in range(L):
        b = "blocks.%d." % i
        for ln in ("ln1", "ln2"):
            w[b + ln + ".weight"] = 1 + rnd(E, scale=0.1)
            w[b + ln + ".bias"] = rnd(E, scale=0.1)
        for m in ("att.time_mix_k", "att.time_mix_v", "att.time_mix_r", "ffn.time_mix_k", "ffn.time_mix_r"):
            w[b + m] = torch.rand(1, 1, E, generator=g).to(dtype)
        w[b + "att.time_decay"] = rnd(E)
        w[b + "att.time_first"] = rnd(E, scale=0.5)
        for m in ("att.key", "att.value", "att.receptance", "att.output", "ffn.receptance"):
            w[b + m + ".weight"] = rnd(E, E, scale=E ** -0.5)
        w[b + "ffn.key.weight"] = rnd(4 * E, E, scale=E ** -0.5)
        w[b + "ffn.value.weight"] = rnd(E, 4 * E, scale=(4 * E) ** -0.5)
    return w


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    convert(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "model.bin")
