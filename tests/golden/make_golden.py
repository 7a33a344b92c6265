#!/usr/bin/env python3
"""Pack the consensus-spec KZG vectors of the reference into a small committed fixture.

Source (read-only, only in the build container): /root/reference/tests/<api>/kzg-mainnet/<case>/data.yaml
(SURVEY.md §4 / §8c).  Those files are 90 MB of hex because the same few blobs/cells are repeated in
hundreds of cases.  This script de-duplicates every byte string >= 1024 B into a content table and
writes

    tests/golden/vectors.json      {api: {case: {"input": ..., "output": ...}}}; big strings -> "@<id>"
    tests/golden/chunks.bin.xz     concatenation of the de-duplicated byte strings (lzma)
    tests/golden/chunks.idx.json   [[offset, length], ...] per id

The GPU box has no /root/reference, so tests read only these fixtures (tests/golden_vectors.py).
Run:  python tests/golden/make_golden.py   (idempotent)
"""
import hashlib, json, lzma, os, sys
import yaml

SRC = "/root/reference/tests"
OUT = os.path.dirname(os.path.abspath(__file__))
BIG = 1024

chunks, index, by_hash = bytearray(), [], {}

def intern(b: bytes) -> str:
    h = hashlib.sha256(b).digest()
    if h not in by_hash:
        by_hash[h] = len(index)
        index.append([len(chunks), len(b)])
        chunks.extend(b)
    return "@%d" % by_hash[h]

def conv(x):
    if isinstance(x, str) and x.startswith("0x"):
        raw = x[2:]
        # odd-length / non-hex strings are deliberate "invalid input" cases: keep verbatim
        try:
            b = bytes.fromhex(raw)
        except ValueError:
            return x
        return intern(b) if len(b) >= BIG else x
    if isinstance(x, list):
        return [conv(v) for v in x]
    if isinstance(x, dict):
        return {k: conv(v) for k, v in x.items()}
    return x

def main():
    loader = getattr(yaml, "CSafeLoader", yaml.SafeLoader)
    out = {}
    for api in sorted(os.listdir(SRC)):
        base = os.path.join(SRC, api, "kzg-mainnet")
        if not os.path.isdir(base):
            continue
        out[api] = {}
        for case in sorted(os.listdir(base)):
            with open(os.path.join(base, case, "data.yaml")) as f:
                d = yaml.load(f, Loader=loader)
            out[api][case] = conv(d)
        print(api, len(out[api]), file=sys.stderr)
    with open(os.path.join(OUT, "vectors.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"), sort_keys=True)
    with open(os.path.join(OUT, "chunks.idx.json"), "w") as f:
        json.dump(index, f, separators=(",", ":"))
    with open(os.path.join(OUT, "chunks.bin.xz"), "wb") as f:
        f.write(lzma.compress(bytes(chunks), preset=6))
    print("chunks:", len(index), "raw bytes:", len(chunks), file=sys.stderr)

if __name__ == "__main__":
    main()
