"""CPU: the label-synthesis oracle (oracle/labels_ref.py) against the fixture generated from the
reference's own dataset/target_generation.py (tests/golden/make_golden_labels.py) and, in the build container, against
the reference functions themselves on fresh random inputs."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
G = np.load(os.path.join(HERE, "golden", "labels_golden.npz"))


def test_pose_targets_match_fixture():
    from oracle import labels_ref as R
    for b in range(G["joints"].shape[0]):
        m, a = R.gen_pose_target(G["joints"][b], G["vis"][b], 4, 96, 96, 7, aux=True)
        assert np.array_equal(m, G["pose"][b]) and np.array_equal(a, G["pose_aux"][b])


def test_edge_flip_match_fixture():
    from oracle import labels_ref as R
    for b in range(G["label"].shape[0]):
        assert np.array_equal(R.generate_edge(G["label"][b]), G["edge"][b])
        assert np.array_equal(R.flip_parsing(G["label"][b]), G["flip"][b])
        assert np.array_equal(R.flip_joints(G["joints"][b], 384), G["flip_joints"][b])


def test_oracle_matches_reference_on_fresh_inputs():
    if not os.path.isdir("/root/reference/dataset"):
        pytest.skip("reference not mounted")
    import make_golden_labels as M
    from oracle import labels_ref as R
    T = M.reference_module()
    joints, vis, label = M.inputs(seed=7, b=2)
    for b in range(2):
        m, a = T.gen_pose_target(joints[b], vis[b], 4, 96, 96, 7, aux=True)
        mo, ao = R.gen_pose_target(joints[b], vis[b], 4, 96, 96, 7, aux=True)
        assert np.array_equal(m, mo) and np.array_equal(a, ao)
        assert np.array_equal(T.generate_edge(label[b]), R.generate_edge(label[b], mark_ignore=False))
        assert np.array_equal(T.gen_parsing_target(label[b], flip_param=True, stride=1), R.flip_parsing(label[b]))
        assert np.array_equal(T.flip_joints(joints[b], 384), R.flip_joints(joints[b], 384))
    # empty / degenerate cases: no visible joint, constant label, all-ignore label
    m, _ = R.gen_pose_target(joints[0], np.zeros(16, dtype=np.int32), 4, 96, 96, 7)
    assert np.array_equal(m, T.gen_pose_target(joints[0], np.zeros(16, dtype=np.int32), 4, 96, 96, 7)[0])
    for lab in (np.full((16, 24), 3, np.uint8), np.full((16, 24), 255, np.uint8)):
        assert np.array_equal(T.generate_edge(lab), R.generate_edge(lab, mark_ignore=False))
