"""JPEG decode on the GPU through the C ABI (spm_jpeg_decode) against PIL, bit for bit; then the whole input path from
file bytes (decode -> Resize / CenterCrop / ToTensor -> encoder) against the same path fed with PIL-decoded frames."""
import numpy as np
import pytest
import torch

from tests import helpers as H
from tests import jpeg_cases as J

pytestmark = pytest.mark.gpu


def test_jpeg_decode_matches_pil_bit_for_bit():
    from clip_spm_b200 import ops
    cs = J.cases()
    assert len(cs) >= 60
    for name, data in cs:
        got = ops.decode_jpegs([data])[0].cpu().numpy()
        assert np.array_equal(got, J.pil_decode(data)), name


def test_jpeg_batches_and_restart_intervals():
    """many files per call (one thread per restart interval on the device), mixed content, same geometry"""
    from clip_spm_b200 import ops
    files = [J.encode(256, 340, kind, 70 + (i % 4) * 8, 2, seed=i, **({"restart_marker_rows": 2} if i % 3 == 0 else {}))
             for i, kind in enumerate(["noise", "smooth", "blocks"] * 16)]
    out = ops.decode_jpegs(files).cpu().numpy()
    assert out.shape == (48, 256, 340, 3)
    for i, f in enumerate(files):
        assert np.array_equal(out[i], J.pil_decode(f)), i
    assert ops.jpeg_info(files[0])[:2] == (256, 340)
    assert ops.decode_jpegs([]).shape[0] == 0


def test_jpeg_errors_are_loud():
    from clip_spm_b200 import ops
    a, b = J.encode(64, 64, "smooth", 80, 2), J.encode(64, 80, "smooth", 80, 2)
    with pytest.raises(RuntimeError, match="differs"):
        ops.decode_jpegs([a, b])
    with pytest.raises(RuntimeError, match="SOI"):
        ops.decode_jpegs([b"garbage"])
    import io
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(np.zeros((32, 32, 3), np.uint8)).save(buf, "JPEG", progressive=True)
    with pytest.raises(RuntimeError, match="progressive"):
        ops.decode_jpegs([buf.getvalue()])


def test_input_path_from_jpeg_files_equals_path_from_pil_frames():
    """files -> GPU decode -> GPU transform -> encoder == PIL decode -> the same GPU transform -> encoder (bit-equal
    frames give bit-equal features), and evaluate_jpeg == evaluate_frames_u8 on PIL-decoded frames"""
    from clip_spm_b200 import ops
    from oracle import clipspm_oracle as O
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci)
    ep = O.make_episode(7000, 2, 1, 1, 2, 24, "P0", images=False)
    su_files = [J.encode(256, 340, "smooth", 85, 2, seed=i) for i in range(4)]     # S*T = 2*2
    qu_files = [J.encode(256, 340, "noise", 85, 2, seed=10 + i) for i in range(4)]  # Q*T = 2*2
    su_pil = torch.from_numpy(np.stack([J.pil_decode(f) for f in su_files])).cuda()
    qu_pil = torch.from_numpy(np.stack([J.pil_decode(f) for f in qu_files])).cuda()
    assert torch.equal(ops.decode_jpegs(su_files), su_pil)
    assert torch.equal(net.encode_frames_u8(ops.decode_jpegs(su_files)), net.encode_frames_u8(su_pil))
    a = net.evaluate_jpeg(su_files, ep["context_labels"], qu_files, ep["real_support_labels"], ep["real_target_labels"],
                          ep["target_labels"])
    b = net.evaluate_frames_u8(su_pil, ep["context_labels"], qu_pil, ep["real_support_labels"], ep["real_target_labels"],
                               ep["target_labels"])
    for k in ("logits", "loss", "acc", "pred"):
        assert torch.equal(a[k], b[k]), k
    assert torch.isfinite(a["logits"]).all()


def test_listing_sweep_from_jpeg_files_equals_sweep_from_pil_decoded_frames():
    """the reference's test loop from the FILES: sampler plan -> JPEG bytes -> GPU decode / transform / forward
    (sweep.run_listing_sweep(jpeg=True)) == the same sweep fed with the PIL decode of the same files"""
    from clip_spm_b200 import frames as F, sweep
    ci = H.case_inputs("vit_2w1s_t2_p0")
    net = H.build_cuda_model(ci, max_episodes=2)
    files, decoded = F.Split(), F.Split()
    k = 0
    for vid in range(3):
        for cls in range(4):
            n = 4 + (cls + vid) % 3
            fs = [J.encode(120, 160, ("smooth", "noise", "blocks")[(k + i) % 3], 80, 2, seed=100 * k + i) for i in range(n)]
            files.add_vid(fs, cls)
            decoded.add_vid([torch.from_numpy(J.pil_decode(f)) for f in fs], cls)
            k += 1
    a = sweep.run_listing_sweep(net, files, lambda fr: fr, 5, 2, 1, 1, seed=50, episodes_per_call=2, jpeg=True)
    b = sweep.run_listing_sweep(net, decoded, lambda fr: fr, 5, 2, 1, 1, seed=50, episodes_per_call=2)
    assert a["n"] == b["n"] == 5
    assert abs(a["accuracy"] - b["accuracy"]) < 1e-9 and abs(a["loss"] - b["loss"]) < 1e-6
