"""Helper of tests/test_gpu_scale.py::test_pinned_host_batch_with_and_without_column_pieces (run in a subprocess so that
CKZG_B200_TAIL_PIECES, which the library reads once, can differ from the parent's): pinned host batches of 1024 and 700
blobs must verify, one flipped byte in the last / in the fourth 16 KiB piece of a tail blob must not, and the pageable
route must agree."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

mod = entry.load_package()
ts = mod.load_trusted_setup()
n = 1024
host = torch.from_numpy(bench.synth_blobs(n, 31337))
dev = host.cuda()
cms = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
prs = torch.empty(48 * n, dtype=torch.uint8, device="cuda")
mod.blob_to_kzg_commitment_device(cms.data_ptr(), dev.data_ptr(), n, ts)
mod.compute_blob_kzg_proof_device(prs.data_ptr(), dev.data_ptr(), cms.data_ptr(), n, ts)
hc, hp = cms.cpu().pin_memory(), prs.cpu().pin_memory()
for m in (1024, 700):
    pin = host[: 131072 * m].clone().pin_memory()
    assert mod.verify_blob_kzg_proof_batch_host(pin.data_ptr(), hc.data_ptr(), hp.data_ptr(), m, ts) is True
    pin[131072 * (m - 3) + 131071] ^= 1  # low byte of the last field element: still canonical
    assert mod.verify_blob_kzg_proof_batch_host(pin.data_ptr(), hc.data_ptr(), hp.data_ptr(), m, ts) is False
    pin[131072 * (m - 3) + 131071] ^= 1
    pin[131072 * (m - 100) + 16384 * 3 + 31] ^= 1  # a byte in the fourth piece
    assert mod.verify_blob_kzg_proof_batch_host(pin.data_ptr(), hc.data_ptr(), hp.data_ptr(), m, ts) is False
pageable = host[: 131072 * 700].numpy().copy()
assert mod.verify_blob_kzg_proof_batch_host(pageable.ctypes.data, hc.data_ptr(), hp.data_ptr(), 700, ts) is True
print("pieces_check ok, CKZG_B200_TAIL_PIECES =", os.environ.get("CKZG_B200_TAIL_PIECES"))
