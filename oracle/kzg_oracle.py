"""CPU restatement of the c-kzg-4844 public API -- ORACLE / TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this module,
and only as the checker.  The product (`c-kzg-4844_b200/`) never does; it fails loudly without
its CUDA extension.

Each function restates the reference function cited in its docstring (paths relative to
/root/reference).  BADARGS is modelled by raising `BadArgs`.  Pure-Python loops: fine for a few
blobs; bulk differential checks use the compiled reference in oracle/_ref (oracle/build_ref.sh).

Parity pin: tests/test_oracle_golden.py runs this file against the consensus-spec vectors
(tests/golden, packed from /root/reference/tests by tests/golden/make_golden.py).
"""
import hashlib

from . import bls12_381 as B
from .bls12_381 import P, R

FIELD_ELEMENTS_PER_BLOB = 4096  # src/eip4844/blob.h:29
FIELD_ELEMENTS_PER_EXT_BLOB = 8192  # src/eip4844/blob.h:42
FIELD_ELEMENTS_PER_CELL = 64  # src/eip7594/cell.h:28
CELLS_PER_EXT_BLOB = 128  # src/eip7594/cell.h:37
CELLS_PER_BLOB = 64
BYTES_PER_BLOB = 131072
BYTES_PER_CELL = 2048
PRIMITIVE_ROOT = 7  # src/setup/setup.c:58
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"FSBLOBVERIFY_V1_"  # src/eip4844/eip4844.c:45
RANDOM_CHALLENGE_DOMAIN_VERIFY_BLOB_KZG_PROOF_BATCH = b"RCKZGBATCH___V1_"  # eip4844.c:48
RANDOM_CHALLENGE_DOMAIN_VERIFY_CELL_KZG_PROOF_BATCH = b"RCKZGCBATCH__V1_"  # eip7594.c:43


class BadArgs(ValueError):
    """C_KZG_BADARGS (src/common/ret.h:26)."""


# ---------------------------------------------------------------------------------------------
# utils (src/common/utils.c, src/common/bytes.c)
# ---------------------------------------------------------------------------------------------


def reverse_bits_limited(n, value):
    """src/common/utils.c:85"""
    bits = n.bit_length() - 1
    return int(format(value, "0%db" % bits)[::-1], 2) if bits else 0


def bit_reversal_permutation(seq):
    """src/common/utils.c:103"""
    n = len(seq)
    return [seq[reverse_bits_limited(n, i)] for i in range(n)]


def bytes_to_bls_field(b):
    """src/common/bytes.c:64 -- big-endian, must be canonical (< r)."""
    v = int.from_bytes(b, "big")
    if len(b) != 32 or v >= R:
        raise BadArgs("non-canonical field element")
    return v


def bytes_from_bls_field(v):
    """src/common/bytes.c:52"""
    return int(v % R).to_bytes(32, "big")


def hash_to_bls_field(b):
    """src/common/bytes.c:123 -- big-endian integer reduced mod r."""
    return int.from_bytes(b, "big") % R


def validate_kzg_g1(b):
    """src/common/bytes.c:81 -- uncompress, accept infinity, else require the G1 subgroup."""
    p = B.g1_uncompress(bytes(b))
    if p is None:
        raise BadArgs("bad G1 encoding")
    if not B.g1_is_inf(p) and not B.g1_in_subgroup(p):
        raise BadArgs("G1 point not in subgroup")
    return p


bytes_to_kzg_commitment = validate_kzg_g1  # src/common/bytes.c:101
bytes_to_kzg_proof = validate_kzg_g1  # src/common/bytes.c:112
bytes_from_g1 = B.g1_compress  # src/common/bytes.c:42


def blob_to_polynomial(blob):
    """src/eip4844/blob.c:31"""
    if len(blob) != BYTES_PER_BLOB:
        raise BadArgs("blob length")
    return [bytes_to_bls_field(blob[32 * i : 32 * i + 32]) for i in range(FIELD_ELEMENTS_PER_BLOB)]


def compute_powers(x, n):
    """src/common/utils.c:151"""
    out, cur = [], 1
    for _ in range(n):
        out.append(cur)
        cur = cur * x % R
    return out


# ---------------------------------------------------------------------------------------------
# MSM (src/common/lincomb.c)
# ---------------------------------------------------------------------------------------------


def g1_lincomb_naive(points, scalars):
    """src/common/lincomb.c:34"""
    acc = B.G1_INF
    for p, s in zip(points, scalars):
        acc = B.g1_add(acc, B.g1_mul(p, s % R))
    return acc


def g1_lincomb_fast(points, scalars, c=8):
    """src/common/lincomb.c:65 -> blst_p1s_mult_pippenger (blst/src/multi_scalar.c:370-434).
    Bucket method with unsigned c-bit windows; points at infinity are skipped as the reference
    does (lincomb.c:92-99).  Result is the same group element whatever the window size."""
    pts = [(p, s % R) for p, s in zip(points, scalars) if not B.g1_is_inf(p) and s % R]
    if not pts:
        return B.G1_INF
    nwin = (255 + c - 1) // c
    total = B.G1_INF
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = B.g1_dbl(total)
        buckets = [B.G1_INF] * (1 << c)
        for p, s in pts:
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = B.g1_add(buckets[d], p)
        run, acc = B.G1_INF, B.G1_INF
        for d in range((1 << c) - 1, 0, -1):
            run = B.g1_add(run, buckets[d])
            acc = B.g1_add(acc, run)
        total = B.g1_add(total, acc)
    return total


# ---------------------------------------------------------------------------------------------
# Trusted setup (src/setup/setup.c)
# ---------------------------------------------------------------------------------------------


class Settings:
    """The parts of KZGSettings (src/setup/settings.h:27-79) the oracle needs."""

    def __init__(self):
        self.roots_of_unity = None  # [8193]
        self.brp_roots_of_unity = None  # [8192]
        self.g1_monomial = None  # [4096] Jacobian
        self.g1_lagrange_brp = None  # [4096] Jacobian
        self.g2_monomial = None  # [65] Jacobian


def compute_roots_of_unity():
    """src/setup/setup.c:130 (ROOT_OF_UNITY :81 = 7^((r-1)/8192))."""
    w = pow(PRIMITIVE_ROOT, (R - 1) // FIELD_ELEMENTS_PER_EXT_BLOB, R)
    roots = compute_powers(w, FIELD_ELEMENTS_PER_EXT_BLOB + 1)
    assert roots[-1] == 1
    return roots, bit_reversal_permutation(roots[:-1])


def load_trusted_setup(g1_monomial_bytes, g1_lagrange_bytes, g2_monomial_bytes, check=True):
    """src/setup/setup.c:392.  Points are decompressed WITHOUT a subgroup check (:447-477)."""
    if len(g1_monomial_bytes) != 48 * 4096 or len(g1_lagrange_bytes) != 48 * 4096:
        raise BadArgs("g1 byte count")
    if len(g2_monomial_bytes) != 96 * 65:
        raise BadArgs("g2 byte count")
    s = Settings()

    def g1s(buf):
        out = []
        for i in range(4096):
            p = B.g1_uncompress(buf[48 * i : 48 * i + 48])
            if p is None:
                raise BadArgs("setup g1")
            out.append(p)
        return out

    s.g1_monomial = g1s(g1_monomial_bytes)
    lag = g1s(g1_lagrange_bytes)
    s.g2_monomial = []
    for i in range(65):
        q = B.g2_uncompress(g2_monomial_bytes[96 * i : 96 * i + 96])
        if q is None:
            raise BadArgs("setup g2")
        s.g2_monomial.append(q)
    if check:  # is_trusted_setup_in_lagrange_form, setup.c:339
        if B.pairings_verify(lag[1], s.g2_monomial[0], lag[0], s.g2_monomial[1]):
            raise BadArgs("setup is in monomial form")
    s.roots_of_unity, s.brp_roots_of_unity = compute_roots_of_unity()
    s.g1_lagrange_brp = bit_reversal_permutation(lag)
    return s


def parse_trusted_setup_text(text):
    """src/setup/setup.c:519-582: counts, then G1 Lagrange, G2 monomial, G1 monomial hex."""
    tok = text.split()
    n1, n2 = int(tok[0]), int(tok[1])
    if n1 != 4096 or n2 != 65:
        raise BadArgs("setup counts")
    body = tok[2:]
    lag = b"".join(bytes.fromhex(t) for t in body[:n1])
    g2 = b"".join(bytes.fromhex(t) for t in body[n1 : n1 + n2])
    mono = b"".join(bytes.fromhex(t) for t in body[n1 + n2 : n1 + n2 + n1])
    return mono, lag, g2


def load_trusted_setup_file(path, check=True):
    with open(path) as f:
        mono, lag, g2 = parse_trusted_setup_text(f.read())
    return load_trusted_setup(mono, lag, g2, check)


# ---------------------------------------------------------------------------------------------
# EIP-4844 (src/eip4844/eip4844.c)
# ---------------------------------------------------------------------------------------------


def compute_challenge(blob, commitment_bytes):
    """src/eip4844/eip4844.c:147.  `commitment_bytes` = the canonical 48-byte compression."""
    h = hashlib.sha256()
    h.update(FIAT_SHAMIR_PROTOCOL_DOMAIN)
    h.update((0).to_bytes(8, "big") + FIELD_ELEMENTS_PER_BLOB.to_bytes(8, "big"))
    h.update(bytes(blob))
    h.update(bytes(commitment_bytes))
    return hash_to_bls_field(h.digest())


def evaluate_polynomial_in_evaluation_form(poly, x, s):
    """src/eip4844/eip4844.c:192 (barycentric formula, in-domain shortcut :213)."""
    roots = s.brp_roots_of_unity
    acc = 0
    for i in range(FIELD_ELEMENTS_PER_BLOB):
        if x == roots[i]:
            return poly[i]
    for i in range(FIELD_ELEMENTS_PER_BLOB):
        acc = (acc + poly[i] * roots[i] % R * pow(x - roots[i], -1, R)) % R
    acc = acc * pow(FIELD_ELEMENTS_PER_BLOB, -1, R) % R
    return acc * (pow(x, FIELD_ELEMENTS_PER_BLOB, R) - 1) % R


def blob_to_kzg_commitment(blob, s):
    """src/eip4844/eip4844.c:264"""
    poly = blob_to_polynomial(blob)
    return bytes_from_g1(g1_lincomb_fast(s.g1_lagrange_brp, poly))


def compute_kzg_proof_impl(poly, z, s):
    """src/eip4844/eip4844.c:417-494 (quotient in evaluation form, in-domain case :460-481)."""
    roots = s.brp_roots_of_unity
    y = evaluate_polynomial_in_evaluation_form(poly, z, s)
    q = [0] * FIELD_ELEMENTS_PER_BLOB
    m = None
    for i in range(FIELD_ELEMENTS_PER_BLOB):
        if z == roots[i]:
            m = i
            continue
        q[i] = (poly[i] - y) * pow(roots[i] - z, -1, R) % R
    if m is not None:
        acc = 0
        for i in range(FIELD_ELEMENTS_PER_BLOB):
            if i == m:
                continue
            num = (poly[i] - y) * roots[i] % R
            den = z * (z - roots[i]) % R
            acc = (acc + num * pow(den, -1, R)) % R
        q[m] = acc
    return bytes_from_g1(g1_lincomb_fast(s.g1_lagrange_brp, q)), y


def compute_kzg_proof(blob, z_bytes, s):
    """src/eip4844/eip4844.c:382 -> (proof48, y32)"""
    poly = blob_to_polynomial(blob)
    z = bytes_to_bls_field(z_bytes)
    proof, y = compute_kzg_proof_impl(poly, z, s)
    return proof, bytes_from_bls_field(y)


def compute_blob_kzg_proof(blob, commitment_bytes, s):
    """src/eip4844/eip4844.c:506"""
    c = bytes_to_kzg_commitment(commitment_bytes)
    poly = blob_to_polynomial(blob)
    z = compute_challenge(blob, bytes_from_g1(c))
    return compute_kzg_proof_impl(poly, z, s)[0]


def verify_kzg_proof_impl(c, z, y, proof, s):
    """src/eip4844/eip4844.c:343:  e(C - [y]G1, G2) == e(proof, [tau]G2 - [z]G2)"""
    x_minus_z = B.g2_add(s.g2_monomial[1], B.g2_neg(B.g2_mul(B.G2_GEN_J, z)))
    p_minus_y = B.g1_sub(c, B.g1_mul(B.G1_GEN_J, y))
    return B.pairings_verify(p_minus_y, B.G2_GEN_J, proof, x_minus_z)


def verify_kzg_proof(commitment_bytes, z_bytes, y_bytes, proof_bytes, s):
    """src/eip4844/eip4844.c:302"""
    c = bytes_to_kzg_commitment(commitment_bytes)
    z = bytes_to_bls_field(z_bytes)
    y = bytes_to_bls_field(y_bytes)
    pr = bytes_to_kzg_proof(proof_bytes)
    return verify_kzg_proof_impl(c, z, y, pr, s)


def verify_blob_kzg_proof(blob, commitment_bytes, proof_bytes, s):
    """src/eip4844/eip4844.c:546"""
    c = bytes_to_kzg_commitment(commitment_bytes)
    poly = blob_to_polynomial(blob)
    pr = bytes_to_kzg_proof(proof_bytes)
    z = compute_challenge(blob, bytes_from_g1(c))
    y = evaluate_polynomial_in_evaluation_form(poly, z, s)
    return verify_kzg_proof_impl(c, z, y, pr, s)


def compute_r_powers_for_verify_kzg_proof_batch(cs, zs, ys, prs):
    """src/eip4844/eip4844.c:597"""
    n = len(cs)
    h = hashlib.sha256()
    h.update(RANDOM_CHALLENGE_DOMAIN_VERIFY_BLOB_KZG_PROOF_BATCH)
    h.update(FIELD_ELEMENTS_PER_BLOB.to_bytes(8, "big") + n.to_bytes(8, "big"))
    for c, z, y, p in zip(cs, zs, ys, prs):
        h.update(bytes_from_g1(c) + bytes_from_bls_field(z) + bytes_from_bls_field(y) + bytes_from_g1(p))
    return compute_powers(hash_to_bls_field(h.digest()), n)


def verify_kzg_proof_batch(cs, zs, ys, prs, s):
    """src/eip4844/eip4844.c:697"""
    n = len(cs)
    rp = compute_r_powers_for_verify_kzg_proof_batch(cs, zs, ys, prs)
    proof_lincomb = g1_lincomb_naive(prs, rp)
    c_minus_y = [B.g1_sub(cs[i], B.g1_mul(B.G1_GEN_J, ys[i])) for i in range(n)]
    r_times_z = [rp[i] * zs[i] % R for i in range(n)]
    proof_z_lincomb = g1_lincomb_naive(prs, r_times_z)
    c_minus_y_lincomb = g1_lincomb_naive(c_minus_y, rp)
    rhs = B.g1_add(c_minus_y_lincomb, proof_z_lincomb)
    return B.pairings_verify(proof_lincomb, s.g2_monomial[1], rhs, B.G2_GEN_J)


def verify_blob_kzg_proof_batch(blobs, commitments_bytes, proofs_bytes, s):
    """src/eip4844/eip4844.c:775 (n==0 -> true :791, n==1 -> single verify :798)."""
    n = len(blobs)
    if n == 0:
        return True
    if n == 1:
        return verify_blob_kzg_proof(blobs[0], commitments_bytes[0], proofs_bytes[0], s)
    cs, zs, ys, prs = [], [], [], []
    for i in range(n):
        c = bytes_to_kzg_commitment(commitments_bytes[i])
        poly = blob_to_polynomial(blobs[i])
        z = compute_challenge(blobs[i], bytes_from_g1(c))
        y = evaluate_polynomial_in_evaluation_form(poly, z, s)
        pr = bytes_to_kzg_proof(proofs_bytes[i])
        cs.append(c), zs.append(z), ys.append(y), prs.append(pr)
    return verify_kzg_proof_batch(cs, zs, ys, prs, s)
