"""TEST INFRASTRUCTURE ONLY.  Pins the oracle's MODEL STRUCTURE against the reference's own model.py, IN THE BUILD CONTAINER.

The reference's `GeneratorBE / GeneratorBE3 / EncoderBE / EncoderBE3 / AE / AE3` (model.py:5-216) are imported unchanged
from /root/reference and executed under oracle/tf_shim.install_structural: `tf.variable_scope`, `slim.conv2d/conv3d/
fully_connected` and `get_variables` are provided by the shim (variable names, shapes and creation order are recorded), the
layers' arithmetic is the oracle's restated `conv_nd` / `linear` (TensorFlow's kernels are not installable).  While
generating, this script asserts that
  * the variables the reference's code asks for == the oracle's layout tables (names, shapes, ORDER),
  * the returned `variables` lists == those tables,
  * the reference-structure outputs == oracle/ref_model.py's forward functions bit for bit,
and writes tests/golden/model_structure.npz (inputs + outputs of the reference-structure run at small sizes), which travels
to the GPU box where /root/reference does not exist.

    python -m oracle.make_golden_model
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import tf_shim  # noqa: E402
from oracle import ref_model as M  # noqa: E402
from oracle import ref_ops as R  # noqa: E402

FILTERS, Z_NUM, SEED = 8, 5, 20261017
CASES = {  # name: (builder, spatial+[C], kwargs)
    "gen2d": ("GeneratorBE", [16, 12, 1], dict(num_conv=2)),
    "gen3d": ("GeneratorBE3", [16, 16, 8, 3], dict(num_conv=2)),
    "gen2d_repeat": ("GeneratorBE", [32, 24, 2], dict(num_conv=1, repeat=3)),
    "enc2d": ("EncoderBE", [16, 12, 2], dict(num_conv=2)),
    "enc3d": ("EncoderBE3", [16, 16, 8, 3], dict(num_conv=1)),
    "ae2d": ("AE", [16, 12, 2], dict(num_conv=3)),
    "ae3d": ("AE3", [16, 16, 8, 3], dict(num_conv=2)),
    "ae2d_sparse": ("AE", [16, 12, 2], dict(num_conv=2, use_sparse=True)),
    # arch=dg: the patch discriminator sees concat(velocity, vorticity): 2+1 channels in 2D, 3+3 in 3D (trainer.py:153)
    "disc2d": ("DiscriminatorPatch", [16, 24, 3], dict()),
    "disc3d": ("DiscriminatorPatch3", [16, 8, 16, 6], dict()),
}


def run_case(model, name):
    builder, shape, kw = CASES[name]
    g = torch.Generator().manual_seed(SEED + sum(map(ord, name)))
    B = 2
    if builder.startswith("Generator"):
        tab, _, _ = M.generator_layout(shape, FILTERS, kw.get("num_conv", 4), kw.get("repeat", 0), z_dim=3, name="G")
        inp = torch.rand(B, 3, generator=g) * 2 - 1
    elif builder.startswith("Discriminator"):
        tab = M.discriminator_layout(shape[-1], FILTERS, len(shape) - 1, name="D")
        inp = torch.randn(B, *shape, generator=g)
    elif builder.startswith("Encoder"):
        tab, _ = M.encoder_layout(shape, FILTERS, Z_NUM, kw.get("num_conv", 3), kw.get("repeat", 0), name="enc")
        inp = torch.randn(B, *shape, generator=g)
    else:
        tab = M.ae_layout(shape, FILTERS, Z_NUM, kw.get("num_conv", 4), kw.get("repeat", 0), name="AE")
        inp = torch.randn(B, *shape, generator=g)
    var = M.init_variables(tab, SEED)
    for k in var:                                    # non-zero biases, so a dropped / misplaced bias shows
        if k.endswith("biases"):
            var[k] = torch.randn(var[k].shape, generator=g) * 0.1
    store = tf_shim.VariableStore(var)
    tf_shim.install_structural(store, R.conv_nd, R.linear)
    fn = getattr(model, builder)
    if builder.startswith("Generator"):
        out, variables = fn(inp, FILTERS, shape, **kw)
        mine = M.generator_forward(inp, var, shape, FILTERS, kw.get("num_conv", 4), kw.get("repeat", 0), "G")
        outs = {"out": out}
        assert torch.equal(out, mine), name
    elif builder.startswith("Discriminator"):
        out, variables = fn(inp, FILTERS)
        mine = M.discriminator_forward(inp, var, "D")
        outs = {"out": out}
        assert torch.equal(out, mine), name
        # reuse=True (trainer.py:156: the generated field goes through the SAME discriminator): no new variable is requested
        n_req = len(store.requested)
        out2, variables2 = fn(inp * 0.5, FILTERS, reuse=True)
        assert len(store.requested) == n_req and list(variables2) == list(variables), name
        assert torch.equal(out2, M.discriminator_forward(inp * 0.5, var, "D")), name
    elif builder.startswith("Encoder"):
        out, variables = fn(inp, FILTERS, Z_NUM, **kw)
        mine = M.encoder_forward(inp, var, FILTERS, kw.get("num_conv", 3), kw.get("repeat", 0), "enc")
        outs = {"out": out}
        assert torch.equal(out, mine), name
    else:
        out, z, variables = fn(inp, FILTERS, Z_NUM, **kw)
        mo, mz = M.ae_forward(inp, var, FILTERS, Z_NUM, kw.get("num_conv", 4), kw.get("repeat", 0), "AE", kw.get("use_sparse", False))
        outs = {"out": out, "z": z}
        assert torch.equal(out, mo) and torch.equal(z, mz), name
    # names, shapes and creation order: reference code == oracle layout table
    assert [n for n, _ in store.requested] == list(tab.keys()), (name, [n for n, _ in store.requested][:4], list(tab)[:4])
    assert [tuple(s) for _, s in store.requested] == [tuple(s) for s in tab.values()], name
    assert list(variables) == list(tab.keys()), name
    return inp, outs, list(tab.keys())


def main(reference_root="/root/reference", out_dir=os.path.join(ROOT, "tests", "golden")):
    model = tf_shim.import_reference_model(reference_root)
    blob = {}
    for name in CASES:
        inp, outs, names = run_case(model, name)
        blob[name + "/in"] = inp.numpy()
        for k, v in outs.items():
            blob[name + "/" + k] = v.detach().numpy()
        blob[name + "/variables"] = np.array(names)
        print("%-14s %2d variables, out %s: reference structure == oracle" % (name, len(names), tuple(outs["out"].shape)))
    np.savez_compressed(os.path.join(out_dir, "model_structure.npz"), **blob)
    print("written", os.path.join(out_dir, "model_structure.npz"))


if __name__ == "__main__":
    main()
