#!/usr/bin/env python
"""Generate tests/golden/xxh64_chain_golden.json.

Pinned against python `xxhash` 3.7.0 (the reference C implementation behind a
Python binding) — an implementation independent of oracle/ and of the kernels:
  * XXH64 digests of seeded random byte strings of many lengths (all tail paths);
  * hash chains for 64-byte blocks (16 uint32 tokens), for
    the reference's own blockSize: 5 over ASCII text
    (/root/reference/pkg/router/strategy.go:57), and for 32/96/128-byte blocks,
    including truncation at maxPrefixBlocksToMatch, trailing partial blocks and
    prompts shorter than one block.
The reference holds no golden vectors for this path; these are the
strongest pins available.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import struct

import xxhash

HERE = os.path.dirname(os.path.abspath(__file__))


def chain(data: bytes, B: int, M: int, h0: int):
    out, prev = [], h0
    for i in range(min(len(data) // B, M)):
        prev = xxhash.xxh64_intdigest(data[i * B:(i + 1) * B] + struct.pack("<Q", prev))
        out.append(prev)
    return out


def main():
    rng = random.Random(0xF0510000)
    doc = {"generator": "tests/golden/make_golden.py", "xxhash_version": xxhash.VERSION,
           "known_answers": {"": "ef46db3751d8e999", "a": "d24ec4f1a98c6e5b", "abc": "44bc2cf5ad770999"},
           "xxh64": [], "chains": []}
    for n in list(range(0, 80)) + [95, 96, 97, 127, 128, 129, 255, 256, 1000, 4099]:
        data = bytes(rng.getrandbits(8) for _ in range(n))
        doc["xxh64"].append({"hex": data.hex(), "digest": f"{xxhash.xxh64_intdigest(data):016x}"})
    h0 = xxhash.xxh64_intdigest(b"synthetic/model")
    doc["h0_model"] = "synthetic/model"
    doc["h0"] = f"{h0:016x}"
    cases = [
        (64, 16, 64 * 16), (64, 16, 64 * 16 + 13), (64, 4, 64 * 9), (64, 8, 63), (64, 8, 0), (64, 8, 64),
        (32, 8, 32 * 5 + 1), (96, 8, 96 * 3 + 50), (128, 4, 128 * 6), (5, 256, 1280), (5, 256, 1283), (5, 16, 4),
        (5, 6, 77), (16, 8, 16 * 5), (40, 8, 40 * 3 + 7), (1, 12, 9),
    ]
    for B, M, n in cases:
        if B == 5:
            data = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz ,.") for _ in range(n))
        else:
            data = bytes(rng.getrandbits(8) for _ in range(n))
        doc["chains"].append({"block_bytes": B, "max_blocks": M, "hex": data.hex(),
                              "chain": [f"{h:016x}" for h in chain(data, B, M, h0)]})
    with open(os.path.join(HERE, "xxh64_chain_golden.json"), "w") as f:
        json.dump(doc, f, indent=0)
    print("wrote", len(doc["xxh64"]), "digests and", len(doc["chains"]), "chains")


if __name__ == "__main__":
    main()
