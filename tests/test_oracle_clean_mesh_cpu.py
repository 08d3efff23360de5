"""The CPU restatement of utils/clean_mesh.py (oracle/clean_mesh_oracle.py) against hand-checkable cases."""
import numpy as np
import torch

import clean_mesh_oracle as CO


def test_disk_matches_skimage_definition():
    d = CO.disk(2)
    assert d.shape == (5, 5) and d.sum() == 13 and d[0, 2] and not d[0, 1]
    m = np.zeros((1, 9, 9), dtype=bool)
    m[0, 4, 4] = True
    assert np.array_equal(CO.dilate(m, 2)[0][2:7, 2:7], d)


def test_first_hit_picks_the_nearest_triangle():
    v = np.array([[-1, -1, 1], [1, -1, 1], [0, 1, 1], [-1, -1, 2], [1, -1, 2], [0, 1, 2]], dtype=np.float64)
    f = np.array([[3, 4, 5], [0, 1, 2]])
    o = np.zeros((3, 3))
    d = np.array([[0, 0, 1.0], [0, 0, -1.0], [0.9, 0.9, 1.0]])
    idx, margin = CO.first_hits(v, f, o, d)
    assert idx.tolist() == [1, -1, -1]
    assert margin[0] > 0.2


def test_face_adjacency_and_components():
    # two triangles sharing an edge, one isolated triangle, one edge shared by three faces
    f = np.array([[0, 1, 2], [2, 1, 3], [4, 5, 6], [7, 8, 9], [8, 7, 10], [7, 8, 11]])
    adj = CO.face_adjacency(f)
    assert sorted(map(tuple, np.sort(adj, axis=1).tolist())) == [(0, 1)]
    keep, _ = CO.components_keep(f, 2)
    assert keep.tolist() == [True, True, False, False, False, False]


def test_vertex_valid_counts_views():
    intr = torch.eye(4)[None].repeat(2, 1, 1)
    intr[:, 0, 0] = intr[:, 1, 1] = 10.0
    intr[:, 0, 2] = intr[:, 1, 2] = 4.5
    c2w = torch.eye(4)[None].repeat(2, 1, 1)
    masks = torch.zeros(2, 10, 10, dtype=torch.bool)
    masks[:, 4:6, 4:6] = True
    v = np.array([[0, 0, 1.0], [0.4, 0.4, 1.0], [0, 0, -1.0]])
    assert CO.vertex_valid(v, masks, intr, c2w, 1).tolist() == [True, False, False]
    assert CO.vertex_valid(v, masks, intr, c2w, 2).tolist() == [False, False, False]
