"""TwoStreamDenoiser (reference models/model.py:437-547, models/modules.py): the oracle restatement against
golden outputs of the unmodified reference."""
import pytest
import torch

from conftest import load_golden
from oracle import twostream as OT
from oracle.make_golden_twostream import CASES, fill, inputs


def golden_state(name):
    g = load_golden("twostream_" + name)
    shapes = {k: tuple(int(x) for x in v.split(",")) if v else () for k, v in zip(g["shapes_keys"], g["shapes_vals"])}
    sd = fill(shapes, CASES[name]["seed"])
    return g, sd


@pytest.mark.parametrize("name", ["small", "config"])
def test_oracle_matches_reference(name):
    c = CASES[name]
    g, sd = golden_state(name)
    x, t, labels, views, prev = inputs(c)
    with torch.no_grad():
        y0, z0 = OT.twostream_forward(sd, c, x, t, labels, views)
        y1, z1 = OT.twostream_forward(sd, c, x, t, labels, views, prev_latent=prev)
        y2, z2 = OT.twostream_forward(sd, c, x, t, torch.zeros_like(labels), None, prev_latent=torch.from_numpy(g["z0"]))
    rel = lambda a, b: float((a - torch.from_numpy(b)).norm() / torch.from_numpy(b).norm())
    assert rel(y0, g["y0"]) < 2e-6 and rel(z0, g["z0"]) < 2e-6
    assert rel(y1, g["y1"]) < 2e-6 and rel(z1[:, ::8], g["z1"]) < 2e-6
    assert rel(y2, g["y2"]) < 2e-6 and rel(z2[:, ::8], g["z2"]) < 2e-6
