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


# ---------------------------------------------------------------------------------------------
# EIP-7594 (src/eip7594/*.c) -- restated with Python integers; slow (seconds per blob), used on a
# handful of cases, the bulk differential checks use oracle/_ref.
# ---------------------------------------------------------------------------------------------


def fr_fft(vals, s, inverse=False):
    """src/eip7594/fft.c:100-146: radix-2 transform over the 8192-th root domain with stride 8192/n;
    the inverse scales by 1/n."""
    n = len(vals)
    assert n & (n - 1) == 0 and n <= FIELD_ELEMENTS_PER_EXT_BLOB
    stride = FIELD_ELEMENTS_PER_EXT_BLOB // n
    roots = s.roots_of_unity

    def rec(v, st):
        m = len(v)
        if m == 1:
            return v
        ev, od = rec(v[0::2], st * 2), rec(v[1::2], st * 2)
        out = [0] * m
        for i in range(m // 2):
            w = roots[(FIELD_ELEMENTS_PER_EXT_BLOB - i * st) % FIELD_ELEMENTS_PER_EXT_BLOB] if inverse else roots[i * st]
            t = od[i] * w % R
            out[i] = (ev[i] + t) % R
            out[i + m // 2] = (ev[i] - t) % R
        return out

    out = rec([v % R for v in vals], stride)
    if inverse:
        inv = pow(n, -1, R)
        out = [v * inv % R for v in out]
    return out


def g1_fft(points, s, inverse=False):
    """src/eip7594/fft.c:164-240 (g1_fft / g1_ifft_unscaled: no 1/n)."""
    n = len(points)
    stride = FIELD_ELEMENTS_PER_EXT_BLOB // n
    roots = s.roots_of_unity

    def rec(v, st):
        m = len(v)
        if m == 1:
            return v
        ev, od = rec(v[0::2], st * 2), rec(v[1::2], st * 2)
        out = [None] * m
        for i in range(m // 2):
            w = roots[(FIELD_ELEMENTS_PER_EXT_BLOB - i * st) % FIELD_ELEMENTS_PER_EXT_BLOB] if inverse else roots[i * st]
            t = od[i] if w == 1 else B.g1_mul(od[i], w)
            out[i] = B.g1_add(ev[i], t)
            out[i + m // 2] = B.g1_sub(ev[i], t)
        return out

    return rec(list(points), stride)


def poly_lagrange_to_monomial(lagrange, s):
    """src/eip7594/poly.c:58: bit-reversal then inverse FFT."""
    return fr_fft(bit_reversal_permutation(list(lagrange)), s, inverse=True)


def compute_cells(blob, s):
    """cells half of compute_cells_and_kzg_proofs (src/eip7594/eip7594.c:88-121) -> (cells bytes, monomial)"""
    mono = poly_lagrange_to_monomial(blob_to_polynomial(blob), s) + [0] * FIELD_ELEMENTS_PER_BLOB
    data = bit_reversal_permutation(fr_fft(mono, s))
    return b"".join(bytes_from_bls_field(v) for v in data), mono


def fk20_x_ext_fft_columns(s):
    """init_fk20_multi_settings (src/setup/setup.c:238-330): x_ext_fft[offset] = G1-FFT of 64 setup points
    padded with 64 identities; cached on the settings object."""
    if getattr(s, "_fk20", None) is None:
        cols = []
        for offset in range(FIELD_ELEMENTS_PER_CELL):
            start = FIELD_ELEMENTS_PER_BLOB - FIELD_ELEMENTS_PER_CELL - 1 - offset
            x = [s.g1_monomial[start - i * FIELD_ELEMENTS_PER_CELL] for i in range(CELLS_PER_BLOB - 1)] + [B.G1_INF]
            cols.append(g1_fft(x + [B.G1_INF] * CELLS_PER_BLOB, s))
        s._fk20 = cols
    return s._fk20


def compute_fk20_cell_proofs(mono, s):
    """src/eip7594/fk20.c:139-286 -> 128 Jacobian points in natural order (caller bit-reverses)."""
    n2 = 2 * CELLS_PER_BLOB
    cols = fk20_x_ext_fft_columns(s)
    inv = pow(n2, -1, R)
    coeffs = []
    for i in range(FIELD_ELEMENTS_PER_CELL):
        c = [0] * n2  # circulant_coeffs_stride, fk20.c:55-78
        dmi = FIELD_ELEMENTS_PER_BLOB - 1 - i
        c[0] = mono[dmi]
        for j in range(1, CELLS_PER_BLOB - 1):
            c[n2 - j] = mono[dmi - j * FIELD_ELEMENTS_PER_CELL]
        coeffs.append([v * inv % R for v in fr_fft(c, s)])
    u = [g1_lincomb_fast([cols[i][j] for i in range(FIELD_ELEMENTS_PER_CELL)], [coeffs[i][j] for i in range(FIELD_ELEMENTS_PER_CELL)], c=5) for j in range(n2)]
    v = g1_fft(u, s, inverse=True)
    v = v[:CELLS_PER_BLOB] + [B.G1_INF] * CELLS_PER_BLOB
    return g1_fft(v, s)


def compute_cells_and_kzg_proofs(blob, s, want_proofs=True):
    """src/eip7594/eip7594.c:61 -> (cells 262144 B, proofs 6144 B or None)"""
    cells, mono = compute_cells(blob, s)
    if not want_proofs:
        return cells, None
    proofs = bit_reversal_permutation(compute_fk20_cell_proofs(mono, s))
    return cells, b"".join(bytes_from_g1(p) for p in proofs)


def recover_cells(cell_indices, cells_fr, s):
    """src/eip7594/recovery.c:200-365 (cells_fr: 8192 values in cell order, zeros where missing)."""
    n = FIELD_ELEMENTS_PER_EXT_BLOB
    cells_brp = bit_reversal_permutation(list(cells_fr))
    missing = [reverse_bits_limited(CELLS_PER_EXT_BLOB, i) for i in range(CELLS_PER_EXT_BLOB) if i not in cell_indices]
    # vanishing polynomial (recovery.c:46-162)
    short = [1]
    for m in missing:
        r = s.roots_of_unity[m * (n // CELLS_PER_EXT_BLOB)]
        nxt = [0] * (len(short) + 1)
        for j, cj in enumerate(short):
            nxt[j] = (nxt[j] - cj * r) % R
            nxt[j + 1] = (nxt[j + 1] + cj) % R
        short = nxt
    z = [0] * n
    for i, cj in enumerate(short):
        z[i * FIELD_ELEMENTS_PER_CELL] = cj
    z_eval = fr_fft(z, s)
    ez = [a * b % R for a, b in zip(cells_brp, z_eval)]
    q = fr_fft(ez, s, inverse=True)
    shift = lambda p, f: [c * pow(f, k, R) % R for k, c in enumerate(p)]
    q_coset = fr_fft(shift(q, 7), s)
    z_coset = fr_fft(shift(z, 7), s)
    p_coset = [a * pow(b, -1, R) % R for a, b in zip(q_coset, z_coset)]
    p = shift(fr_fft(p_coset, s, inverse=True), pow(7, -1, R))
    return bit_reversal_permutation(fr_fft(p, s))


def recover_cells_and_kzg_proofs(cell_indices, cells, s, want_proofs=True):
    """src/eip7594/eip7594.c:177"""
    nc = len(cell_indices)
    if nc > CELLS_PER_EXT_BLOB or nc < CELLS_PER_BLOB or len(cells) != nc:
        raise BadArgs("cell count")
    for i, ci in enumerate(cell_indices):
        if ci >= CELLS_PER_EXT_BLOB or (i and ci <= cell_indices[i - 1]):
            raise BadArgs("cell index")
    data = [0] * FIELD_ELEMENTS_PER_EXT_BLOB
    for ci, cell in zip(cell_indices, cells):
        for j in range(FIELD_ELEMENTS_PER_CELL):
            data[ci * FIELD_ELEMENTS_PER_CELL + j] = bytes_to_bls_field(cell[32 * j : 32 * j + 32])
    rec = data if nc == CELLS_PER_EXT_BLOB else recover_cells(list(cell_indices), data, s)
    out_cells = b"".join(bytes_from_bls_field(v) for v in rec)
    if not want_proofs:
        return out_cells, None
    mono = poly_lagrange_to_monomial(rec, s)
    proofs = bit_reversal_permutation(compute_fk20_cell_proofs(mono, s))
    return out_cells, b"".join(bytes_from_g1(p) for p in proofs)


def compute_verify_cell_kzg_proof_batch_challenge(unique_commitments, commitment_indices, cell_indices, cells, proofs):
    """src/eip7594/eip7594.c:390-482"""
    h = hashlib.sha256()
    h.update(RANDOM_CHALLENGE_DOMAIN_VERIFY_CELL_KZG_PROOF_BATCH)
    for v in (FIELD_ELEMENTS_PER_BLOB, FIELD_ELEMENTS_PER_CELL, len(unique_commitments), len(cell_indices)):
        h.update(v.to_bytes(8, "big"))
    for c in unique_commitments:
        h.update(bytes(c))
    for ci, col, cell, pr in zip(commitment_indices, cell_indices, cells, proofs):
        h.update(ci.to_bytes(8, "big") + col.to_bytes(8, "big") + bytes(cell) + bytes(pr))
    return hash_to_bls_field(h.digest())


def verify_cell_kzg_proof_batch(commitments, cell_indices, cells, proofs, s):
    """src/eip7594/eip7594.c:825-974"""
    n = len(cell_indices)
    if n == 0:
        return True
    if any(c >= CELLS_PER_EXT_BLOB for c in cell_indices):
        raise BadArgs("cell index")
    uniq, cidx = [], []
    for c in commitments:  # deduplicate_commitments, :345-376
        c = bytes(c)
        if c not in uniq:
            uniq.append(c)
        cidx.append(uniq.index(c))
    r = compute_verify_cell_kzg_proof_batch_challenge(uniq, cidx, cell_indices, cells, proofs)
    rp = compute_powers(r, n)
    proofs_g1 = [bytes_to_kzg_proof(p) for p in proofs]
    proof_lincomb = g1_lincomb_fast(proofs_g1, rp, c=5)
    cm_g1 = [bytes_to_kzg_commitment(c) for c in uniq]
    weights = [0] * len(uniq)
    for k in range(n):
        weights[cidx[k]] = (weights[cidx[k]] + rp[k]) % R
    final = g1_lincomb_fast(cm_g1, weights, c=5)
    # aggregated interpolation polynomial (:615-770)
    agg = [[0] * FIELD_ELEMENTS_PER_CELL for _ in range(CELLS_PER_EXT_BLOB)]
    used = set()
    for k in range(n):
        col = cell_indices[k]
        used.add(col)
        for j in range(FIELD_ELEMENTS_PER_CELL):
            agg[col][j] = (agg[col][j] + bytes_to_bls_field(cells[k][32 * j : 32 * j + 32]) * rp[k]) % R
    poly = [0] * FIELD_ELEMENTS_PER_CELL
    for col in sorted(used):
        coeffs = fr_fft(bit_reversal_permutation(agg[col]), s, inverse=True)
        hinv = s.roots_of_unity[FIELD_ELEMENTS_PER_EXT_BLOB - reverse_bits_limited(CELLS_PER_EXT_BLOB, col)]
        for k in range(FIELD_ELEMENTS_PER_CELL):
            poly[k] = (poly[k] + coeffs[k] * pow(hinv, k, R)) % R
    interp = g1_lincomb_fast(s.g1_monomial[:FIELD_ELEMENTS_PER_CELL], poly, c=5)
    final = B.g1_sub(final, interp)
    wp = [rp[k] * s.roots_of_unity[reverse_bits_limited(CELLS_PER_EXT_BLOB, cell_indices[k]) * FIELD_ELEMENTS_PER_CELL] % R for k in range(n)]
    final = B.g1_add(final, g1_lincomb_fast(proofs_g1, wp, c=5))
    return B.pairings_verify(final, B.G2_GEN_J, proof_lincomb, s.g2_monomial[FIELD_ELEMENTS_PER_CELL])
