"""The signed c-bit window recoding used by the fixed-base table kernels (fk20.cu fk20_msm_kernel, msm_direct.cu,
msm_affine.cu bam_digits): a Python restatement of the exact expression the kernels evaluate -- funnel shift over
nine 32-bit limbs, carry into the next window, magnitude <= 2^(c-1) -- checked for every window width the memory
plan can pick (api.cu plan_commit_window / plan_fk_window) on edge and random scalars: the digits must rebuild the
scalar, stay inside the table (1 <= magnitude <= 2^(c-1)) and leave no carry after the top window.  CPU only; the
kernels themselves are pinned by the GPU parity tests."""
import random

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def recode(k, c):
    s = [(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)] + [0]
    w_count = (256 + c - 1) // c
    m = 1 << (c - 1)
    dmask, dfull = (1 << c) - 1, 1 << c
    carry, out = 0, []
    for w in range(w_count):
        o = w * c
        lo, hi, sh = s[o >> 5], s[(o >> 5) + 1], o & 31
        funnel = ((lo | (hi << 32)) >> sh) & 0xFFFFFFFF  # __funnelshift_r(lo, hi, sh)
        d = (funnel & dmask) + carry
        neg = d > m
        carry = 1 if neg else 0
        mag = (dfull - d) if neg else d
        out.append((mag, neg))
    return out, carry, w_count, m


def test_digits_rebuild_the_scalar_for_every_planned_width():
    rnd = random.Random(3)
    scalars = [0, 1, 2, R - 1, R - 2, (1 << 254) - 1, 1 << 254, (1 << 255) - 1, 0x5555555555555555555555555555555555555555555555555555555555555555,
               0x2AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA]
    scalars += [rnd.randrange(R) for _ in range(300)]
    scalars += [((1 << 255) - 1) ^ (1 << rnd.randrange(255)) for _ in range(50)]  # long carry chains
    for c in (8, 10, 12, 13, 14):
        for k in scalars:
            assert k < (1 << 255)
            digits, carry, w_count, m = recode(k, c)
            assert carry == 0, (c, hex(k))  # the top window absorbs the carry: the scalar is below 2^255
            total = 0
            for w, (mag, neg) in enumerate(digits):
                assert 0 <= mag <= m
                total += (-mag if neg else mag) << (c * w)
            assert total == k, (c, hex(k))


def test_window_counts_and_table_sizes():
    # geometry quoted in DESIGN.md / cells.h / msm_direct.cu
    geo = {c: ((256 + c - 1) // c, 1 << (c - 1)) for c in (8, 10, 12, 13, 14)}
    assert geo == {8: (32, 128), 10: (26, 512), 12: (22, 2048), 13: (20, 4096), 14: (19, 8192)}
    gb = lambda pts, c: pts * geo[c][0] * geo[c][1] * 96 / 2**30
    assert round(gb(8192, 12), 1) == 33.0 and round(gb(8192, 8), 1) == 3.0 and round(gb(8192, 10), 1) == 9.8
    assert round(gb(4096, 14), 1) == 57.0 and round(gb(4096, 13), 1) == 30.0 and round(gb(4096, 12), 1) == 16.5
