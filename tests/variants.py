"""Test-only checkpoint variants derived from the synthetic checkpoints (no real weights offline).

* `patched_model_dir` copies model.safetensors of a synthetic checkpoint and rewrites rows of the tied
  embedding / lm_head table in place (numpy memmap over the tensor's byte range), so tests can build
  - an EOS-capable model: row 151645 (<|im_end|>) scaled up until greedy decoding terminates by itself
    after a few dozen tokens (random-init weights otherwise never emit EOS), which exercises the stop
    conditions of qwen_asr.c:788-818 on both sides;
  - exact logit ties: a row duplicated at a lower / higher index (ties -> lowest index,
    qwen_asr_kernels.c:536-541).
* `write_full_vocab` writes a vocab.json in which every token id decodes to a distinct string, so the
  TEXT the reference's tokenizer produces is a one-to-one image of the ids (the default synthetic
  vocab.json has three entries: every id decodes to "").  Ids 0..255 are the byte-level alphabet (GPT-2
  byte -> unicode table), so prompts ("language English") encode without merges.txt.
"""
import json
import os
import shutil
import struct

import numpy as np

EMBED = "thinker.model.embed_tokens.weight"
TOKEN_IM_END = 151645


def tensor_span(path, name):
    """(absolute byte offset, nbytes, shape, dtype) of a tensor inside a .safetensors file."""
    with open(path, "rb") as f:
        hlen = struct.unpack("<Q", f.read(8))[0]
        header = json.loads(f.read(hlen))
    t = header[name]
    o0, o1 = t["data_offsets"]
    return 8 + hlen + o0, o1 - o0, t["shape"], t["dtype"]


def bytes_to_unicode():
    """The GPT-2 byte-level alphabet (what decode_gpt2_token inverts, qwen_asr_tokenizer.c)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


def write_full_vocab(path, vocab_size=151936):
    b2u = bytes_to_unicode()
    vocab = {b2u[b]: b for b in range(256)}
    for i in range(256, vocab_size):
        vocab[f"t{i}."] = i
    with open(path, "w", encoding="utf-8") as f:
        json.dump(vocab, f, ensure_ascii=False)


def patched_model_dir(base_dir, tag, patch_rows, full_vocab=True):
    """Copy of `base_dir` named <base>_<tag> whose embedding rows were rewritten by
    patch_rows(E) (E = uint16 memmap [vocab, hidden] of bf16 bits).  Cached on disk."""
    out = base_dir.rstrip("/") + "_" + tag
    done = os.path.join(out, ".done")
    if os.path.exists(done):
        return out
    os.makedirs(out, exist_ok=True)
    dst = os.path.join(out, "model.safetensors")
    shutil.copyfile(os.path.join(base_dir, "model.safetensors"), dst)
    off, nbytes, shape, dtype = tensor_span(dst, EMBED)
    assert dtype == "BF16" and nbytes == shape[0] * shape[1] * 2
    E = np.memmap(dst, dtype=np.uint16, mode="r+", offset=off, shape=tuple(shape))
    patch_rows(E)
    E.flush()
    del E
    if full_vocab:
        write_full_vocab(os.path.join(out, "vocab.json"))
    else:
        shutil.copyfile(os.path.join(base_dir, "vocab.json"), os.path.join(out, "vocab.json"))
    with open(done, "w") as f:
        f.write("ok\n")
    return out


def bf16_to_f32(u16):
    return (np.asarray(u16, np.uint32) << 16).view(np.float32)


def f32_to_bf16(x):
    u = np.asarray(x, np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def scale_row(E, row, gain):
    E[row] = f32_to_bf16(bf16_to_f32(E[row]) * np.float32(gain))


def eos_model_dir(base_dir, gain=3.0):
    """EOS-capable variant: the <|im_end|> row of the tied lm_head scaled by `gain`."""
    return patched_model_dir(base_dir, f"eos{gain:g}", lambda E: scale_row(E, TOKEN_IM_END, gain))
