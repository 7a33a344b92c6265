"""Loader for the packed consensus-spec vectors (tests/golden, see tests/golden/make_golden.py)."""
import functools, json, lzma, os

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=None)
def _load():
    with open(os.path.join(_DIR, "vectors.json")) as f:
        vec = json.load(f)
    with open(os.path.join(_DIR, "chunks.idx.json")) as f:
        idx = json.load(f)
    with open(os.path.join(_DIR, "chunks.bin.xz"), "rb") as f:
        raw = lzma.decompress(f.read())
    return vec, idx, raw


def _resolve(x):
    """'@id' -> bytes; '0x..' -> bytes (or the original str if it is not valid hex: those are the
    deliberately malformed inputs); lists/dicts recursively; everything else unchanged."""
    _, idx, raw = _load()
    if isinstance(x, str):
        if x.startswith("@"):
            off, ln = idx[int(x[1:])]
            return raw[off : off + ln]
        if x.startswith("0x"):
            try:
                return bytes.fromhex(x[2:])
            except ValueError:
                return x
        return x
    if isinstance(x, list):
        return [_resolve(v) for v in x]
    if isinstance(x, dict):
        return {k: _resolve(v) for k, v in x.items()}
    return x


def apis():
    return sorted(_load()[0].keys())


def cases(api):
    """-> list of (case_name, input_dict, output) with bytes resolved."""
    vec = _load()[0][api]
    return [(name, _resolve(c["input"]), _resolve(c["output"])) for name, c in sorted(vec.items())]
