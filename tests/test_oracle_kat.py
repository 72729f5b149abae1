"""Known-answer tests for the TF-half restatement (hand-checked)."""
import numpy as np
import pytest
import torch

from oracle import losses_torch as L

E = torch.tensor([[0, 0], [1, 0], [0, 2], [3, 0]], dtype=torch.float64)

KATS = [
    ("batch_hard", [1, 1, 2, 2], 0.2, 4.6, [[0, 1], [1, 0], [-3, 1], [2, -2]]),
    ("batch_hard", [1, 1, 2, 2], "soft", 4.524355377,
     [[-0.0474258732, 1.0473024786], [1.0947283518, 0], [-2.9996298163, 0.9524507322], [1.9523273377, -1.9997532108]]),
    ("lifted", [1, 1, 2, 2], 1.0, 5.079997649,
     [[0.0100392764, 0.7310585786], [0.8588364384, 0.2689414214], [-2.8655089465, 0.9999864381], [1.9966332317, -1.9999864381]]),
    ("batch_hard", [0, 1, 1, 2], "soft", 1.531039002,
     [[0.5611507372, 0.8354955184], [0.4485812620, -1.9577969928], [-0.9788984964, 1.1223014743], [-0.0308335028, 0]]),
]


@pytest.mark.parametrize("kind,pids,margin,loss,grad", KATS)
def test_kat(kind, pids, margin, loss, grad):
    out, g, d = L.loss_and_grad(kind, E, torch.tensor(pids, dtype=torch.float64), margin)
    assert float(out[0].detach()) == pytest.approx(loss, abs=1e-8)
    assert np.allclose(g.numpy(), np.array(grad), atol=1e-8)


def test_kat_details():
    pids = torch.tensor([1., 1., 2., 2.], dtype=torch.float64)
    out, _, d = L.loss_and_grad("batch_hard", E, pids, 0.2)
    assert d.tolist()[0] == [0, 1, 4, 9]
    assert out[4].tolist() == [1, 1, 13, 13] and out[5].tolist() == [4, 4, 4, 4]
    assert out[2].tolist() == pytest.approx([0, 0, 9.2, 9.2])
    p, n = L.mined_indices(d, pids)
    assert p.tolist() == [1, 0, 3, 2] and n.tolist() == [2, 3, 0, 1]
    out, _, _ = L.loss_and_grad("batch_hard", E, torch.tensor([0., 1., 1., 2.], dtype=torch.float64), "soft")
    assert out[3].tolist() == pytest.approx([0, 2 / 7, 2 / 7, 3 / 7])
    assert out[4].tolist() == [0, 5, 5, 0] and out[5].tolist() == [1, 1, 4, 4]


def test_edge_semantics():
    # single class: no negatives -> loss terms 0, weights 0/0 (undefined input, App. A.2) ; weighted=False is finite
    pids = torch.ones(4, dtype=torch.float64)
    out, g, _ = L.loss_and_grad("batch_hard", E, pids, 0.2, weighted=False)
    assert float(out[0].detach()) == 0.0 and torch.isinf(out[5]).all() and (g == 0).all()
    out, g, _ = L.loss_and_grad("lifted", E, pids, 1.0, weighted=False)
    assert float(out[0].detach()) == 0.0 and (g == 0).all()
    # exact ties split the gradient evenly (TF _MinOrMaxGrad)
    E2 = torch.tensor([[0., 0.], [1., 0.], [-1., 0.], [0., 5.]], dtype=torch.float64)
    out, g, _ = L.loss_and_grad("batch_hard", E2, torch.tensor([1., 1., 1., 2.], dtype=torch.float64), 0.2, weighted=False)
    # row 0 has two furthest positives at distance 1 -> each gets half
    assert g[0, 0] == pytest.approx(0.0)


def test_contrib_triplet_semihard_known_answer():
    """tf.contrib triplet_semihard_loss restatement on SURVEY App. B's four points, by hand:
    E = [[0,0],[1,0],[0,2],[3,0]], labels [1,1,2,2], squared distances d01=1 d02=4 d03=9 d12=5 d13=4 d23=13, margin 0.5.
    (a=0,p=1): d_ap=1, negatives {4, 9} beyond it -> n*=4 -> max(0.5+1-4, 0) = 0
    (a=1,p=0): d_ap=1, negatives {5, 4} -> n*=4 -> 0
    (a=2,p=3): d_ap=13, negatives {4, 5}, none beyond -> largest = 5 -> 0.5+13-5 = 8.5
    (a=3,p=2): d_ap=13, negatives {9, 4}, none beyond -> largest = 9 -> 0.5+13-9 = 4.5
    loss = (0 + 0 + 8.5 + 4.5) / 4 = 3.25"""
    import torch
    from oracle import losses_torch as L
    e = torch.tensor([[0., 0.], [1., 0.], [0., 2.], [3., 0.]], dtype=torch.float64)
    lab = torch.tensor([1, 1, 2, 2])
    assert float(L.contrib_triplet_semihard_loss(lab, e, 0.5)) == pytest.approx(3.25, abs=1e-12)
    # anchors without any negative: masked_maximum degenerates to the row minimum (0): l = margin + d_ap
    assert float(L.contrib_triplet_semihard_loss(torch.tensor([7, 7]), e[:2], 0.5)) == pytest.approx(1.5, abs=1e-12)


def test_contrib_lifted_struct_known_answer():
    """tf.contrib lifted_struct_loss restatement on the same four points, margin 1, Euclidean distances
    d01=1 d02=2 d03=3 d12=sqrt5 d13=2 d23=sqrt13:  S_0 = e^(1-2)+e^(1-3), S_1 = e^(1-sqrt5)+e^(1-2),
    S_2 = e^(1-2)+e^(1-sqrt5), S_3 = e^(1-3)+e^(1-2);  J_01 = log(S_0+S_1)+1, J_23 = log(S_2+S_3)+sqrt13;
    4 ordered positive pairs: loss = 0.25 * 2 (J_01^2 + J_23^2) / 2."""
    import math
    import torch
    from oracle import losses_torch as L
    e = torch.tensor([[0., 0.], [1., 0.], [0., 2.], [3., 0.]], dtype=torch.float64)
    s0 = math.exp(-1) + math.exp(-2); s1 = math.exp(1 - math.sqrt(5)) + math.exp(-1)
    s2 = math.exp(-1) + math.exp(1 - math.sqrt(5)); s3 = math.exp(-2) + math.exp(-1)
    j01 = max(math.log(s0 + s1) + 1, 0.0); j23 = max(math.log(s2 + s3) + math.sqrt(13), 0.0)
    want = 0.25 * 2 * (j01 ** 2 + j23 ** 2) / 2
    assert float(L.contrib_lifted_struct_loss(torch.tensor([1, 1, 2, 2]), e, 1.0)) == pytest.approx(want, rel=1e-12)
