"""The batched loader (seqikpy_b200.loader) against the reference's own raw-format converters: tests/golden/loader.npz holds
small inputs in the three formats and what seqikpy.alignment.convert_from_{anipose,df3d,df3dpp}_to_dict (imported from
/root/reference by oracle/make_golden.py) make of them.  Host logic only -- no GPU, no arithmetic."""
import pickle

import numpy as np
import pytest

from seqikpy_b200 import alignment as A
from seqikpy_b200.loader import BatchedPoseLoader, to_pose_dicts


@pytest.fixture(scope="module")
def gold():
    import conftest
    return dict(np.load(conftest.GOLDEN / "loader.npz", allow_pickle=False))


def anipose_inputs(gold):
    table = {str(k): gold["anipose_table"][i] for i, k in enumerate(gold["anipose_columns"])}
    pts = {str(seg): [str(k) for k in gold[f"anipose_kps_{seg}"]] for seg in gold["anipose_segments"]}
    return table, pts


def test_anipose_key_order_and_layout(gold):
    table, pts = anipose_inputs(gold)
    shifted = {k: v + 1.0 for k, v in table.items()}                          # a second, different recording
    ld = BatchedPoseLoader(["RF", "LF"], fmt="anipose", pts2align=pts, extra_segments=("R_head", "L_head", "Thorax"), pin_memory=False)
    out = ld.load([table, shifted])
    assert tuple(out["legs"].shape) == (2, 2, 40, 5, 3) and out["legs"].dtype.is_floating_point and out["legs"].is_contiguous()
    for li, leg in enumerate(("RF", "LF")):
        ref = gold[f"anipose_ref_{leg}_leg"]
        assert np.array_equal(out["legs"][0, li].numpy(), ref.astype(np.float32))
        assert np.array_equal(out["legs"][1, li].numpy(), (ref + 1.0).astype(np.float32))
    for seg in ("R_head", "L_head", "Thorax"):
        assert np.array_equal(out[seg][0].numpy(), gold[f"anipose_ref_{seg}"].astype(np.float32))
    # our per-recording converter (the reference's name) gives the reference's float64 arrays exactly
    ours = A.convert_from_anipose_to_dict(table, pts)
    assert list(ours.keys()) == list(pts.keys())
    for seg in pts:
        assert np.array_equal(ours[seg], gold[f"anipose_ref_{seg}"])


def test_df3d_and_df3dpp(gold):
    segs = ["RF_leg", "RM_leg", "RH_leg", "LF_leg", "LM_leg", "LH_leg"]
    legs = [s[:2] for s in segs]
    idx = {s: gold[f"df3d_idx_{s}"] for s in segs}
    out = BatchedPoseLoader(legs, fmt="df3d", pts2align=idx, pin_memory=False).load([gold["df3d_array"]] * 3)
    assert tuple(out["legs"].shape) == (3, 6, 30, 5, 3)
    for li, s in enumerate(segs):
        assert np.array_equal(out["legs"][2, li].numpy(), gold[f"df3d_ref_{s}"].astype(np.float32))
        assert np.array_equal(A.convert_from_df3d_to_dict(gold["df3d_array"], idx)[s], gold[f"df3d_ref_{s}"])
    pp = {s: {kp: {"raw_pos_aligned": gold[f"df3dpp_raw_{s}"][i]} for i, kp in enumerate(("Coxa", "Femur", "Tibia", "Tarsus", "Claw"))} for s in segs}
    out = BatchedPoseLoader(legs, fmt="df3dpp", pin_memory=False).load([pp])
    for li, s in enumerate(segs):
        assert np.array_equal(out["legs"][0, li].numpy(), gold[f"df3dpp_ref_{s}"].astype(np.float32))
        assert np.array_equal(A.convert_from_df3dpp_to_dict(pp, segs)[s], gold[f"df3dpp_ref_{s}"])
    dicts = to_pose_dicts(out, legs)
    assert list(dicts[0].keys()) == segs and dicts[0]["RM_leg"].shape == (30, 5, 3) and dicts[0]["RM_leg"].dtype == np.float64


def test_lengths_paths_and_errors(gold, tmp_path):
    table, pts = anipose_inputs(gold)
    short = {k: v[:25] for k, v in table.items()}
    ld = BatchedPoseLoader(["RF", "LF"], fmt="anipose", pts2align=pts, pin_memory=False)
    assert ld.load([table, short])["legs"].shape[2] == 25                     # default: the shortest recording
    with pytest.raises(ValueError):
        BatchedPoseLoader(["RF", "LF"], fmt="anipose", pts2align=pts, n_frame=30, pin_memory=False).load([table, short])
    with pytest.raises(ValueError):
        BatchedPoseLoader(["RF"], fmt="df3d")
    with pytest.raises(ValueError):
        BatchedPoseLoader(["RF"], fmt="hdf5")
    for i, rec in enumerate((table, short)):
        d = tmp_path / f"rec{i}" / "pose-3d"
        d.mkdir(parents=True)
        with open(d / "pose3d.pkl", "wb") as f:
            pickle.dump(rec, f)
    out = ld.load_paths([tmp_path / "rec0", tmp_path / "rec1"])
    assert tuple(out["legs"].shape) == (2, 2, 25, 5, 3)
    assert np.array_equal(out["legs"][0, 0].numpy(), gold["anipose_ref_RF_leg"][:25].astype(np.float32))
    with pytest.raises(FileNotFoundError):                                    # like AlignPose.from_file_path (reference tests/test_alignment.py:22-27)
        ld.load_paths([tmp_path / "nothing_here"])
