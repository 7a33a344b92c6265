"""Run the packed consensus-spec vectors against any backend exposing the CKZG method set
(oracle.ref_lib.CKZG on either library, or tests.py_oracle_backend.PyOracle).

Mirrors the reference's own vector runner, bindings/python/tests.py:37-252: `output: null` means the
call must fail (C_KZG_BADARGS, or a binding-level length check -- those live in the bindings, not
in C: bindings/go/main.go:379-386, so they are applied here before the call).
"""
import golden_vectors as gv


class Invalid(Exception):
    pass


def _need(b, n):
    if not isinstance(b, (bytes, bytearray)) or len(b) != n:
        raise Invalid("length")
    return bytes(b)


def _needs(lst, n):
    return [_need(b, n) for b in lst]


def _call(api, be, i):
    if api == "blob_to_kzg_commitment":
        return be.blob_to_kzg_commitment(_need(i["blob"], 131072))
    if api == "compute_kzg_proof":
        p, y = be.compute_kzg_proof(_need(i["blob"], 131072), _need(i["z"], 32))
        return [p, y]
    if api == "compute_blob_kzg_proof":
        return be.compute_blob_kzg_proof(_need(i["blob"], 131072), _need(i["commitment"], 48))
    if api == "verify_kzg_proof":
        return be.verify_kzg_proof(_need(i["commitment"], 48), _need(i["z"], 32), _need(i["y"], 32), _need(i["proof"], 48))
    if api == "verify_blob_kzg_proof":
        return be.verify_blob_kzg_proof(_need(i["blob"], 131072), _need(i["commitment"], 48), _need(i["proof"], 48))
    if api == "verify_blob_kzg_proof_batch":
        blobs, cs, ps = _needs(i["blobs"], 131072), _needs(i["commitments"], 48), _needs(i["proofs"], 48)
        if not (len(blobs) == len(cs) == len(ps)):
            raise Invalid("count")
        return be.verify_blob_kzg_proof_batch(b"".join(blobs), b"".join(cs), b"".join(ps))
    if api == "compute_cells":
        cells, _ = be.compute_cells_and_kzg_proofs(_need(i["blob"], 131072), True, False)
        return [cells[2048 * k : 2048 * k + 2048] for k in range(128)]
    if api == "compute_cells_and_kzg_proofs":
        cells, proofs = be.compute_cells_and_kzg_proofs(_need(i["blob"], 131072), True, True)
        return [[cells[2048 * k : 2048 * k + 2048] for k in range(128)], [proofs[48 * k : 48 * k + 48] for k in range(128)]]
    if api == "recover_cells_and_kzg_proofs":
        cells = _needs(i["cells"], 2048)
        idx = [int(x) for x in i["cell_indices"]]
        if len(idx) != len(cells):
            raise Invalid("count")
        oc, op = be.recover_cells_and_kzg_proofs(idx, b"".join(cells), True)
        return [[oc[2048 * k : 2048 * k + 2048] for k in range(128)], [op[48 * k : 48 * k + 48] for k in range(128)]]
    if api == "verify_cell_kzg_proof_batch":
        cs, cells, ps = _needs(i["commitments"], 48), _needs(i["cells"], 2048), _needs(i["proofs"], 48)
        idx = [int(x) for x in i["cell_indices"]]
        if not (len(cs) == len(cells) == len(ps) == len(idx)):
            raise Invalid("count")
        return be.verify_cell_kzg_proof_batch(b"".join(cs), idx, b"".join(cells), b"".join(ps))
    raise KeyError(api)


def run_case(api, be, inp):
    """-> the API result, or None if the call is rejected (BADARGS / malformed input)."""
    try:
        return _call(api, be, inp)
    except Invalid:
        return None
    except Exception as e:  # backend BadArgs types differ; anything flagged BADARGS counts
        if type(e).__name__ == "BadArgs":
            return None
        raise


def run_api(api, be, only=None, limit=None):
    """Returns list of (case, got, expected) mismatches; runs `limit` cases at most."""
    bad, n = [], 0
    for name, inp, want in gv.cases(api):
        if only is not None and not only(name):
            continue
        if limit is not None and n >= limit:
            break
        n += 1
        got = run_case(api, be, inp)
        if isinstance(got, tuple):
            got = list(got)
        if got != want:
            bad.append((name, got, want))
    return bad, n
