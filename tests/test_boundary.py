"""CPU tests of the drop-in boundary's host logic: the flag surface equals the reference's (golden generated from
/root/reference/config.py by oracle/make_golden.py), unknown flags are tolerated (config.py:73 parse_known_args), the
error behaviour of main()/Trainer mirrors main.py:29-30 / trainer.py:80,167."""
import json
import os

from deepfluids_b200 import config as C


def test_flag_names_and_defaults_match_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "reference_config_defaults.json")))
    cfg, unparsed = C.get_config([])
    mine = vars(cfg)
    for k, v in ref.items():
        assert k in mine, "missing reference flag --%s" % k
        assert mine[k] == v, (k, mine[k], v)
    extra = set(mine) - set(ref)
    assert extra == {"synthetic", "synthetic_samples", "max_step", "precision", "grad_accum"}   # new knobs are optional, defaults inert
    assert cfg.synthetic is False and cfg.max_step == 0


def test_unknown_flags_are_tolerated_like_the_reference():
    # run.bat:56 passes `--filter=64`; argparse prefix matching + parse_known_args make that legal in the reference
    cfg, unparsed = C.get_config(["--filter=64", "--not_a_flag=1", "--is_3d=True"])
    assert cfg.filters == 64 and cfg.is_3d is True and "--not_a_flag=1" in unparsed


def test_str2bool():
    assert C.str2bool("True") and C.str2bool("1") and not C.str2bool("no")


def test_model_builders_keep_the_reference_signatures(golden_dir):
    """GeneratorBE(3) / EncoderBE(3) / AE(3): same parameter names, order and literal defaults as reference model.py
    (fixture written by oracle/make_signatures.py from the reference source)."""
    import inspect
    from deepfluids_b200 import model as M
    ref = json.load(open(os.path.join(golden_dir, "reference_model_signatures.json")))
    for name, spec in ref.items():
        fn = getattr(M, name)
        mine = list(inspect.signature(fn).parameters.values())
        assert [p.name for p in mine] == [a for a, _ in spec["params"]], name
        for p, (a, d) in zip(mine, spec["params"]):
            if d is None:
                assert p.default is inspect.Parameter.empty, (name, a)
            elif d == "lrelu":
                assert p.default is M.lrelu, (name, a)
            elif d == "tf.nn.elu":
                assert p.default is M.elu, (name, a)
            else:
                assert p.default == eval(d), (name, a, p.default, d)    # literals only: 'G', 4, 3, 0, False


def test_trainer_method_surface_matches_reference(golden_dir):
    """every method of the reference's Trainer / Trainer3 exists on the mirror (de / ae / dg / nn implemented, the rendering
    helpers raise NotImplementedError naming the reason); names recorded from the reference source by oracle/make_signatures.py"""
    from deepfluids_b200.trainer import Trainer
    from deepfluids_b200.trainer3 import Trainer3
    ref = json.load(open(os.path.join(golden_dir, "reference_trainer_methods.json")))
    for m in ref["Trainer"]:
        assert callable(getattr(Trainer, m, None)), "Trainer.%s missing" % m
    for m in ref["Trainer3"]:
        assert callable(getattr(Trainer3, m, None)), "Trainer3.%s missing" % m
    t = Trainer.__new__(Trainer)
    for m in ("generate", "get_vort_image"):                  # rendering helpers: out of scope, say so loudly
        try:
            getattr(t, m)(None)
            raise AssertionError("%s should raise" % m)
        except NotImplementedError as e:
            assert "outside the B200 hot path" in str(e)


def test_trainer_graph_tensor_names_exist():
    """names other code reads off the reference's Trainer: the ones the fused step never materialises are
    on-demand properties"""
    from deepfluids_b200.trainer import Trainer
    for a in ("x_jaco", "x_vort", "G_jaco_", "G_vort_", "s", "z", "x_"):
        assert isinstance(getattr(Trainer, a), property), a
    src = open(os.path.join(os.path.dirname(__file__), "..", "deep-fluids_b200", "trainer.py")).read()
    for a in ("G_s", "G_", "G_var", "g_loss", "g_loss_l1", "g_loss_j_l1", "g_optim", "g_lr", "step", "loss", "loss_l1",
              "loss_j_l1", "loss_p", "optim"):
        assert ("self.%s " % a) in src or ("self.%s=" % a) in src or ("self.%s," % a) in src, a


def test_parameter_sweep_dump_matches_reference_test_(tmp_path):
    """Trainer.test_ (trainer.py:314-354): the reference's own method (imported unchanged through the shim, build container
    only) and the mirror's, both driven by the same stand-in generator, must write the same `<model_dir>/10_2/<i>.npz`
    files: z_c construction (p1, p2 fixed, last parameter swept over linspace(-1, 1, y3)), batching by test_b_num, denorm."""
    import numpy as np
    import pytest
    import torch
    if not os.path.exists("/root/reference/trainer.py"):
        pytest.skip("reference source not present on this machine")
    from oracle import tf_shim
    from deepfluids_b200.trainer import Trainer
    ref_trainer, _, _ = tf_shim.import_reference_trainers("/root/reference")

    def fake_generator(z):                       # any deterministic map z [n,3] -> fields [n,4,3,2]
        z = np.asarray(z, dtype=np.float32)
        base = np.arange(24, dtype=np.float32).reshape(4, 3, 2) / 24
        return (z[:, 0, None, None, None] + 2 * z[:, 1, None, None, None] + 3 * z[:, 2, None, None, None] * base).astype(np.float32)

    class BM(object):
        y_num = [11, 5, 8]
        x_range = 2.5

        def denorm(self, x=None, y=None):
            return (x * self.x_range if x is not None else None), y

    class Sess(object):
        def run(self, fetch, feed):
            assert fetch == "G_-tensor" and list(feed) == ["z-placeholder"]
            assert feed["z-placeholder"].shape == (4, 3)           # exactly test_b_num rows per run
            return fake_generator(feed["z-placeholder"])

    r = object.__new__(ref_trainer.Trainer)
    r.build_test_model = lambda: None
    r.batch_manager, r.test_b_num, r.c_num, r.model_dir, r.sess, r.G_, r.z = BM(), 4, 3, str(tmp_path / "ref"), Sess(), "G_-tensor", "z-placeholder"
    ref_trainer.Trainer.test_(r)

    m = Trainer.__new__(Trainer)
    m.build_test_model = lambda: None
    m.batch_manager, m.test_b_num, m.c_num, m.model_dir = BM(), 4, 3, str(tmp_path / "mine")
    m.generate_velocity = lambda z: torch.from_numpy(fake_generator(z))
    out_dir = m.test_()
    assert out_dir == os.path.join(str(tmp_path / "mine"), "10_2")
    ref_files = sorted(os.listdir(os.path.join(str(tmp_path / "ref"), "10_2")))
    assert sorted(os.listdir(out_dir)) == ref_files == sorted("%d.npz" % i for i in range(8))
    for f in ref_files:
        a = np.load(os.path.join(str(tmp_path / "ref"), "10_2", f))["x"]
        b = np.load(os.path.join(out_dir, f))["x"]
        np.testing.assert_array_equal(b, a)


def test_prepare_dirs_matches_reference_util(tmp_path, monkeypatch):
    """util.prepare_dirs_and_logger / save_config (util.py:17-59): the reference's own functions (imported with empty
    stand-ins for its plotting imports; build container only) and the mirror's produce the same data_path, the same
    model_dir rule (load_path wins; a preset model_dir is kept; else log_dir/<dataset>/<MMDD_HHMMSS>_<arch>_<tag>) and the
    same params.json."""
    import argparse
    import importlib.util
    import re
    import sys
    import types
    import pytest
    if not os.path.exists("/root/reference/util.py"):
        pytest.skip("reference source not present on this machine")
    from deepfluids_b200 import util as U
    fakes = {}
    for name in ("PIL", "PIL.Image", "PIL.ImageOps", "imageio", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            fakes[name] = types.ModuleType(name)
    if "PIL" in fakes:
        fakes["PIL"].Image = fakes.get("PIL.Image")
    for k, v in fakes.items():
        monkeypatch.setitem(sys.modules, k, v)
    spec = importlib.util.spec_from_file_location("_dfl_reference_util", "/root/reference/util.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cwd = os.getcwd()

    def cfg(**kw):
        base = dict(data_dir="data", dataset="smoke_pos21_size5_f200", log_dir=str(tmp_path / "log"), arch="de", tag="test", load_path="")
        base.update(kw)
        return argparse.Namespace(**base)

    try:
        for kw in ({}, {"load_path": str(tmp_path / "resume")}, {"model_dir": str(tmp_path / "preset")}):
            a, b = cfg(**kw), cfg(**kw)
            ref.prepare_dirs_and_logger(a)
            os.chdir(cwd)                                     # (the reference chdir()s into its source directory)
            U.prepare_dirs_and_logger(b)
            assert a.data_path == b.data_path == os.path.join("data", "smoke_pos21_size5_f200")
            if kw:
                assert a.model_dir == b.model_dir == list(kw.values())[0]
            else:
                pat = re.escape(os.path.join(str(tmp_path / "log"), "smoke_pos21_size5_f200", "")) + r"\d{4}_\d{6}_de_test$"
                assert re.match(pat, a.model_dir) and re.match(pat, b.model_dir)
            assert os.path.isdir(a.model_dir) and os.path.isdir(b.model_dir)
            ref.save_config(a)
            U.save_config(b)
            ja = json.load(open(os.path.join(a.model_dir, "params.json")))
            jb = json.load(open(os.path.join(b.model_dir, "params.json")))
            ja.pop("model_dir"), jb.pop("model_dir")
            assert ja == jb
    finally:
        os.chdir(cwd)


def test_ops_keep_the_reference_signatures(golden_dir):
    """ops.py of the boundary (SURVEY 8b "Ops"): same parameter names, order and literal defaults as reference ops.py
    (fixture written by oracle/make_signatures.py from the reference source); DiscriminatorPatch(3) likewise."""
    import inspect
    from deepfluids_b200 import model as M, ops as O
    ref = json.load(open(os.path.join(golden_dir, "reference_ops_signatures.json")))
    assert {"conv2d", "conv3d", "linear", "batch_norm", "curl", "jacobian", "jacobian3", "lrelu", "upscale", "upscale3", "divergence",
            "divergence3", "curl_np", "vort_np", "grad_np", "jacobian_np3"} <= set(ref)
    for name, spec in ref.items():
        mine = list(inspect.signature(getattr(O, name)).parameters.values())
        assert [p.name for p in mine] == [a for a, _ in spec["params"]], name
        for p, (a, d) in zip(mine, spec["params"]):
            if d is None:
                assert p.default is inspect.Parameter.empty, (name, a)
            elif d == "lrelu":
                assert p.default is O.lrelu, (name, a)
            else:
                assert p.default == eval(d), (name, a, p.default, d)
    refm = json.load(open(os.path.join(golden_dir, "reference_model_signatures.json")))
    for name in ("DiscriminatorPatch", "DiscriminatorPatch3"):
        mine = list(inspect.signature(getattr(M, name)).parameters.values())
        assert [(p.name, None if p.default is inspect.Parameter.empty else repr(p.default)) for p in mine] == \
               [(a, d) for a, d in refm[name]["params"]], name
    nn = list(inspect.signature(M.NN).parameters.values())
    assert [p.name for p in nn] == [a for a, _ in refm["NN"]["params"]]
    assert nn[4].default is O.elu and nn[5].default == 0.1 and nn[3].default == 'NN'        # act=tf.nn.elu, dropout=0.1


def test_variable_scope_naming_follows_slim():
    """tf.variable_scope / slim default layer names as the reference's models rely on them: unnamed layers are `Conv`,
    `Conv_1`, ... within their scope, re-entering the scope with reuse=True finds the same variables, creating an existing
    variable without reuse raises, asking for a missing one under reuse raises."""
    import pytest
    from deepfluids_b200 import ops as O
    st = O.reset_variables(7)
    with O.variable_scope("D") as vs:
        names = [O._STORE.unique_default("Conv") for _ in range(3)]
        assert names == ["Conv", "Conv_1", "Conv_2"] and vs.name == "D"
        with O.variable_scope("Conv"):
            w = st.get("weights", (3, 3, 4, 8), "cpu")
            st.get("biases", (8,), "cpu", zeros=True)
    assert O.get_variables("D") == ["D/Conv/weights", "D/Conv/biases"] and float(O.get_variable("D/Conv/biases").abs().sum()) == 0.0
    lim = (6.0 / (9 * 4 + 9 * 8)) ** 0.5
    assert float(w.abs().max()) <= lim and float(w.abs().max()) > 0.5 * lim            # xavier-uniform (slim default)
    with O.variable_scope("D", reuse=True):
        assert O._STORE.unique_default("Conv") == "Conv"                                # counters restart on re-entry
        with O.variable_scope("Conv"):
            assert st.get("weights", (3, 3, 4, 8), "cpu") is w
            with pytest.raises(ValueError, match="shape"):
                st.get("weights", (3, 3, 4, 9), "cpu")
        with O.variable_scope("Conv_9"):
            with pytest.raises(ValueError, match="does not exist"):
                st.get("weights", (3, 3, 4, 8), "cpu")
    with O.variable_scope("D"):
        with O.variable_scope("Conv"):
            with pytest.raises(ValueError, match="already exists"):
                st.get("weights", (3, 3, 4, 8), "cpu")
    O.reset_variables()


def test_numpy_twins_match_the_oracle():
    """ops.vort_np / curl_np / grad_np / jacobian_np3 (reference ops.py:305-374) == the pinned oracle stencils, bit for bit"""
    import numpy as np
    import torch
    from deepfluids_b200 import ops as O
    from oracle import ref_ops as R
    g = np.random.default_rng(3)
    x = g.standard_normal((2, 6, 7, 2)).astype(np.float32)
    assert np.array_equal(O.curl_np(x), R.curl(torch.from_numpy(x)).numpy())
    assert np.array_equal(O.vort_np(x), R.jacobian(torch.from_numpy(x))[1].numpy())
    j2 = R.jacobian(torch.from_numpy(np.repeat(x[..., :1], 2, -1)))[0].numpy()
    assert np.array_equal(O.grad_np(x), np.stack([j2[..., 0], j2[..., 1]], -1))
    x3 = g.standard_normal((2, 5, 6, 7, 3)).astype(np.float32)
    j, c = O.jacobian_np3(x3)
    jr, cr = R.jacobian3(torch.from_numpy(x3))
    assert np.array_equal(j, jr.numpy()) and np.array_equal(c, cr.numpy())


def test_fused_engine_scopes_raise_on_reuse_before_creation():
    """tf.variable_scope(reuse=True) on variables that do not exist raises in TF; the fused-engine scopes do the same instead
    of silently building fresh seed-123 weights (no CUDA call is reached)"""
    import pytest
    import torch
    from deepfluids_b200 import model as M
    M.reset()
    with pytest.raises(ValueError):
        M.GeneratorBE(torch.zeros(2, 3), 128, [32, 48, 1], reuse=True)
    with pytest.raises(ValueError):
        M.EncoderBE3(torch.zeros(2, 16, 16, 16, 3), 128, 16, reuse=True)
    with pytest.raises(ValueError):
        M.AE(torch.zeros(2, 32, 48, 2), 128, 16, reuse=True)


def test_scope_engine_reuse_rules_on_fake_engines():
    """host logic of model._scope_engine (no CUDA): reuse=False replaces the scope's variables and drops their siblings;
    reuse=True at the scope's batch size returns the scope's engine, at another batch size ONE sibling built over the same
    parameter object; every reuse call re-packs the operands"""
    from deepfluids_b200 import model as M
    M.reset()
    built = []

    class Fake(object):
        def __init__(self, B, params):
            self.B, self.params, self.repacks = B, params if params is not None else object(), 0
            built.append(self)

        def repack(self):
            self.repacks += 1

    key = ("G", 2)
    e4 = M._scope_engine(key, False, 4, lambda shared: Fake(4, shared))
    assert M._scope_engine(key, True, 4, lambda shared: Fake(4, shared)) is e4 and e4.repacks == 1 and len(built) == 1
    e2 = M._scope_engine(key, True, 2, lambda shared: Fake(2, shared))
    assert e2 is not e4 and e2.params is e4.params and e2.repacks == 1
    assert M._scope_engine(key, True, 2, lambda shared: Fake(2, shared)) is e2 and len(built) == 2 and e2.repacks == 2
    # an encoder scope of the same name is a different scope, not a sibling
    enc = M._scope_engine(("G", 2, "enc"), False, 4, lambda shared: Fake(4, shared))
    assert M._ENGINES[("G", 2, ("batch", 2))] is e2
    # re-creating the scope drops the siblings of the old variables, keeps the other scope
    n4 = M._scope_engine(key, False, 4, lambda shared: Fake(4, shared))
    assert n4 is not e4 and ("G", 2, ("batch", 2)) not in M._ENGINES and M._ENGINES[("G", 2, "enc")] is enc
    M.reset()
