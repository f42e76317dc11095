"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/reference_model_signatures.json IN THE BUILD CONTAINER: the parameter
names and literal defaults of the reference's model builders (model.py:5,48,118,154,190,204), read from the source with
`ast` (model.py cannot be imported: TensorFlow 1.15 is not installable), and tests/golden/reference_trainer_methods.json: the
method names of Trainer (trainer.py) and Trainer3 (trainer3.py).  The fixture travels to the GPU box, where
/root/reference does not exist.

    python -m oracle.make_signatures
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NAMES = ("GeneratorBE", "GeneratorBE3", "EncoderBE", "EncoderBE3", "AE", "AE3", "DiscriminatorPatch", "DiscriminatorPatch3", "NN")
# ops.py functions of the drop-in boundary (SURVEY.md 8b "Ops"): layer wrappers, resampling, stencils, numpy twins
OPS_NAMES = ("lrelu", "conv2d", "conv3d", "linear", "batch_norm", "upscale", "upscale3", "jacobian", "jacobian3", "curl",
             "divergence", "divergence3", "vort_np", "curl_np", "grad_np", "jacobian_np3")


def _signatures(path, names):
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            args = [a.arg for a in node.args.args]
            defaults = [ast.unparse(d) for d in node.args.defaults]
            pad = [None] * (len(args) - len(defaults))
            out[node.name] = {"lineno": node.lineno, "params": [[a, d] for a, d in zip(args, pad + defaults)]}
    assert set(out) == set(names), sorted(set(names) - set(out))
    return out


def main(reference_root="/root/reference", out_dir=os.path.join(ROOT, "tests", "golden")):
    out = _signatures(os.path.join(reference_root, "model.py"), NAMES)
    ops = _signatures(os.path.join(reference_root, "ops.py"), OPS_NAMES)
    with open(os.path.join(out_dir, "reference_ops_signatures.json"), "w") as f:
        json.dump(ops, f, indent=1, sort_keys=True)
    print("written", os.path.join(out_dir, "reference_ops_signatures.json"))
    with open(os.path.join(out_dir, "reference_model_signatures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("written", os.path.join(out_dir, "reference_model_signatures.json"))
    methods = {}
    for fname, cname in (("trainer.py", "Trainer"), ("trainer3.py", "Trainer3")):
        t = ast.parse(open(os.path.join(reference_root, fname)).read())
        cls = [n for n in t.body if isinstance(n, ast.ClassDef) and n.name == cname][0]
        methods[cname] = [n.name for n in cls.body if isinstance(n, ast.FunctionDef)]
    with open(os.path.join(out_dir, "reference_trainer_methods.json"), "w") as f:
        json.dump(methods, f, indent=1, sort_keys=True)
    print("written", os.path.join(out_dir, "reference_trainer_methods.json"))


if __name__ == "__main__":
    main()
