"""TEST INFRASTRUCTURE ONLY.  Pins the oracle's LOSS / OPTIMIZER WIRING against the reference's own trainers, IN THE BUILD
CONTAINER.

The reference's `Trainer.build_model`, `Trainer3.build_model`, `Trainer.build_model_ae`, `Trainer3.build_model_ae`
(trainer.py:136-184, :357-396; trainer3.py:14-63, :240-279) are imported unchanged and run on a stub `self` under the
structural shim (oracle/tf_shim.py): the generator / AE graph, curl, Jacobians, the loss expression, the optimizer
construction and the `minimize(loss, global_step, var_list)` call are the reference's code; layer arithmetic is the
oracle's restated primitives; the build stops at the first placeholder (TensorBoard summaries follow).  Asserted while
generating:
  * loss, loss_l1, loss_j_l1 (and loss_p / the KL term for the AE) == oracle/ref_train.py bit for bit,
  * d loss / d variables through torch autograd of the reference-built expression == the oracle's gradients,
  * `var_list` == the oracle's variable table (names, order), `global_step` is the trainer's step variable,
  * the optimizer is AdamOptimizer(g_lr, beta1=, beta2=) with no epsilon argument (TF default 1e-8).
Writes tests/golden/trainer_wiring.npz (inputs, losses, gradient checksums).

    python -m oracle.make_golden_trainer
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
from oracle import ref_train as T  # noqa: E402

FILTERS, Z_NUM, SEED, P_NUM = 8, 6, 20261018, 2
CASES = {  # name: (is_3d, arch, spatial, num_conv, use_sparse)
    "de2d": (False, "de", [16, 12], 2, False),
    "de3d": (True, "de", [16, 16, 8], 2, False),
    "ae2d": (False, "ae", [16, 12], 3, False),
    "ae3d": (True, "ae", [16, 16, 8], 2, False),
    "ae2d_sparse": (False, "ae", [16, 12], 2, True),
    # arch=dg (trainer.py:149-156,174-182 / trainer3.py:27-34,53-61): generator + patch discriminator, LSGAN terms
    "dg2d": (False, "dg", [16, 16], 2, False),
    "dg3d": (True, "dg", [16, 16, 16], 1, False),
}
W1, W2, W4, W5, SPARSITY = 0.7, 1.3, 0.9, 0.5, 0.05
W3 = 0.6


class _Q(object):
    def size(self):
        return 0


class _BM(object):
    q = _Q()


def make_inputs(name):
    is3d, arch, spatial, num_conv, use_sparse = CASES[name]
    g = torch.Generator().manual_seed(SEED + sum(map(ord, name)))
    B, C = 2, (3 if is3d else 2)
    x = torch.randn(B, *spatial, C, generator=g).clamp_(-1, 1)
    if arch in ("de", "dg"):
        y = torch.rand(B, 3, generator=g) * 2 - 1
        tab, _, _ = M.generator_layout(spatial + [3 if is3d else 1], FILTERS, num_conv, 0, z_dim=3, name="G")
        if arch == "dg":
            tab = type(tab)(tab)
            tab.update(M.discriminator_layout(6 if is3d else 3, FILTERS, 3 if is3d else 2, "D"))
    else:
        y = torch.rand(B, P_NUM, 4, generator=g) * 2 - 1              # [B, dof, frames]; the loss uses y[:, :, -1]
        tab = M.ae_layout(spatial + [C], FILTERS, Z_NUM, num_conv, 0, name="AE")
    var = M.init_variables(tab, SEED)
    for k in var:
        if k.endswith("biases"):
            var[k] = torch.randn(var[k].shape, generator=g) * 0.1
    return x, y, tab, var


def run_case(trainer_mod, trainer3_mod, ref_ops, name):
    is3d, arch, spatial, num_conv, use_sparse = CASES[name]
    x, y, tab, var = make_inputs(name)
    leaves = {k: v.clone().requires_grad_(True) for k, v in var.items()}
    store = tf_shim.VariableStore(leaves)
    tf_shim.install_structural(store, R.conv_nd, R.linear)
    record = {}
    tf_shim.install_training(record)
    cls = trainer3_mod.Trainer3 if is3d else trainer_mod.Trainer
    t = object.__new__(cls)
    t.batch_manager = _BM()
    t.x, t.y = x, y
    t.x_jaco, t.x_vort = (ref_ops.jacobian3 if is3d else ref_ops.jacobian)(x)      # trainer.py:28-32
    t.arch, t.use_c, t.filters, t.num_conv, t.repeat = arch, True, FILTERS, num_conv, 0
    t.output_shape = spatial + [3 if is3d else 1]                                  # trainer.py:48-53
    t.optimizer, t.g_lr, t.beta1, t.beta2, t.step = "adam", "g_lr-variable", 0.5, 0.999, "step-variable"
    t.w1, t.w2 = W1, W2
    if arch == "dg":
        t.w3 = W3
    if arch == "ae":
        t.z_num, t.use_sparse, t.sparsity, t.w4, t.w5, t.p_num = Z_NUM, use_sparse, SPARSITY, W4, W5, P_NUM
    try:
        (cls.build_model_ae if arch == "ae" else cls.build_model)(t)
        raise AssertionError("the build did not reach its first placeholder")
    except tf_shim.BuildDone:
        pass
    opt, mini = record["optimizer"], record["minimize"]
    assert opt["kind"] == "adam" and opt["args"] == ("g_lr-variable",) and opt["kwargs"] == {"beta1": 0.5, "beta2": 0.999}, opt
    if arch == "dg":
        return run_dg_case(name, t, record, x, y, tab, var, leaves, num_conv)
    assert mini["global_step"] == "step-variable" and mini["var_list"] == list(tab.keys()), name
    loss = mini["loss"]
    grads = torch.autograd.grad(loss, [leaves[k] for k in tab])
    out = {"x": x.numpy(), "y": y.numpy()}
    if arch == "de":
        o_loss, o_l1, o_jl1, _, _, o_grads = T.generator_loss_and_grads(y, x, var, FILTERS, num_conv, 0, W1, W2, True, "G")
        assert loss is t.g_loss
        assert torch.equal(t.g_loss.detach(), o_loss) and torch.equal(t.g_loss_l1.detach(), o_l1) and torch.equal(t.g_loss_j_l1.detach(), o_jl1), name
        out.update(loss=o_loss.numpy(), l1=o_l1.numpy(), jl1=o_jl1.numpy())
    else:
        o_loss, o_l1, o_jl1, o_lp, _, o_z, o_grads = T.ae_loss_and_grads(x, y[:, :, -1], var, P_NUM, FILTERS, Z_NUM, num_conv, 0, W1, W2, W4, True,
                                                                         "AE", use_sparse, SPARSITY, W5)
        assert loss is t.loss
        assert torch.equal(t.loss_l1.detach(), o_l1) and torch.equal(t.loss_j_l1.detach(), o_jl1) and torch.equal(t.loss_p.detach(), o_lp), name
        assert torch.allclose(t.loss.detach(), o_loss, rtol=1e-6, atol=0), (name, float(t.loss), float(o_loss))
        assert torch.equal(t.z.detach(), o_z), name
        out.update(loss=o_loss.numpy(), l1=o_l1.numpy(), jl1=o_jl1.numpy(), loss_p=o_lp.numpy())
    # relative to the largest gradient entry of the model: the output conv's bias gradient is mathematically zero (the curl
    # of a constant vanishes), so a per-variable relative measure would compare rounding noise with rounding noise
    scale = max(float(o_grads[k].abs().max()) for k in tab)
    worst = max(float((g - o_grads[k]).abs().max()) for k, g in zip(tab, grads)) / scale
    assert worst <= 1e-5, (name, worst)
    out["grad_abs_sums"] = np.array([float(o_grads[k].abs().sum()) for k in tab])
    return out, worst


def run_dg_case(name, t, record, x, y, tab, var, leaves, num_conv):
    """arch=dg: TWO minimize calls on ONE optimizer object (trainer.py:181,184): d_optim = minimize(d_loss, var_list=D_var)
    without global_step, then g_optim = minimize(g_loss, global_step=step, var_list=G_var)."""
    g_names = [k for k in tab if k.startswith("G/")]
    d_names = [k for k in tab if k.startswith("D/")]
    m_d, m_g = record["minimize_all"]
    assert m_d["optimizer_id"] == m_g["optimizer_id"], "the reference uses one AdamOptimizer for both (shared beta powers)"
    assert m_d["var_list"] == d_names and m_d["global_step"] is None and m_d["loss"] is t.d_loss
    assert m_g["var_list"] == g_names and m_g["global_step"] == "step-variable" and m_g["loss"] is t.g_loss
    assert list(t.G_var) == g_names and list(t.D_var) == d_names
    losses, o_gg, o_dg = T.dg_losses_and_grads(y, x, {k: var[k] for k in g_names}, {k: var[k] for k in d_names}, FILTERS,
                                               num_conv, 0, W1, W2, W3)
    for k in ("g_loss", "g_loss_l1", "g_loss_j_l1", "g_loss_real", "d_loss_fake", "d_loss_real", "d_loss", "D_x", "D_G"):
        assert torch.equal(getattr(t, k).detach(), losses[k]), (name, k)
    gg = torch.autograd.grad(t.g_loss, [leaves[k] for k in g_names], retain_graph=True)
    dg = torch.autograd.grad(t.d_loss, [leaves[k] for k in d_names])
    worst = 0.0
    for names, mine, ora in ((g_names, gg, o_gg), (d_names, dg, o_dg)):
        scale = max(float(ora[k].abs().max()) for k in names)
        worst = max(worst, max(float((g - ora[k]).abs().max()) for k, g in zip(names, mine)) / scale)
    assert worst <= 1e-5, (name, worst)
    out = {"x": x.numpy(), "y": y.numpy(), "loss": losses["g_loss"].numpy(), "d_loss": losses["d_loss"].numpy(),
           "g_loss_real": losses["g_loss_real"].numpy(), "D_x": losses["D_x"].numpy(), "D_G": losses["D_G"].numpy(),
           "g_grad_abs_sums": np.array([float(o_gg[k].abs().sum()) for k in g_names]),
           "d_grad_abs_sums": np.array([float(o_dg[k].abs().sum()) for k in d_names])}
    return out, worst


def run_init_case(trainer_mod, trainer3_mod, is3d, lr_update, start_step):
    """Run the reference's Trainer.__init__ itself (trainer.py:13-105) on a stub config / batch manager: input wiring,
    target Jacobian, output_shape rule, max_step, the learning-rate variable and its update expression; __init__ then
    enters build_model, which the shim stops at its first placeholder.  Returns the half-constructed trainer + record."""
    import argparse
    spatial = [16, 16, 8] if is3d else [16, 12]
    g = torch.Generator().manual_seed(SEED + start_step)
    B, C = 2, (3 if is3d else 2)
    x = torch.randn(B, *spatial, C, generator=g).clamp_(-1, 1)
    y = torch.rand(B, 3, generator=g) * 2 - 1
    tab, _, _ = M.generator_layout(spatial + [3 if is3d else 1], FILTERS, 2, 0, z_dim=3, name="G")
    leaves = {k: v.clone().requires_grad_(True) for k, v in M.init_variables(tab, SEED).items()}
    store = tf_shim.VariableStore(leaves)
    tf_shim.install_structural(store, R.conv_nd, R.linear)
    record = {}
    tf_shim.install_training(record)

    class BM(_BM):
        c_num, dof = 3, 0
        epochs_per_step = 8.0 / 21000.0                       # batch_size / num_samples (data.py:52)

        def batch(self):
            return x, y

    cfg = argparse.Namespace(is_3d=is3d, dataset="smoke", data_type="velocity", arch="de", res_x=spatial[-1], res_y=spatial[-2],
                             res_z=spatial[0] if is3d else 0, batch_size=B, test_batch_size=100, repeat=0, filters=FILTERS, num_conv=2,
                             w1=W1, w2=W2, use_curl=True, optimizer="adam", beta1=0.5, beta2=0.999, model_dir="unused", load_path="",
                             start_step=start_step, max_epoch=100, lr_update=lr_update, lr_min=2.5e-6, lr_max=1e-4, lr_update_step=120000,
                             log_step=500, test_step=1000, save_sec=3600, is_train=True)
    cls = trainer3_mod.Trainer3 if is3d else trainer_mod.Trainer
    t = object.__new__(cls)
    try:
        cls.__init__(t, cfg, BM())
        raise AssertionError("__init__ did not reach the first placeholder of build_model")
    except tf_shim.BuildDone:
        pass
    return t, record, x, cfg


def check_init(trainer_mod, trainer3_mod, ref_ops):
    """learning-rate schedules (trainer.py:69-80) and the input wiring of __init__ vs the oracle"""
    out = {}
    for is3d in (False, True):
        max_step = int(100 // (8.0 / 21000.0))                 # = 262499: float floor division, as the reference evaluates it
        steps = (0, 1, 7, 1000, max_step // 2, max_step)
        vals = []
        for s in steps:
            t, rec, x, cfg = run_init_case(trainer_mod, trainer3_mod, is3d, "decay", s)
            assert t.max_step == max_step == 262499
            assert t.output_shape == list(x.shape[1:-1]) + [3 if is3d else 1]                     # trainer.py:48-53
            ref_j, ref_w = (ref_ops.jacobian3 if is3d else ref_ops.jacobian)(x)
            assert torch.equal(t.x_jaco, ref_j) and torch.equal(t.x_vort, ref_w) and t.c_num == 3
            a = [r for r in rec["assign"] if r["name"] == "g_lr_update"]
            assert len(a) == 1 and a[0]["ref"] is t.g_lr and float(t.g_lr.value) == float(torch.tensor(cfg.lr_max, dtype=torch.float32))
            assert rec["optimizer"]["args"] == (t.g_lr,) and rec["minimize"]["global_step"] is t.step
            v = float(a[0]["value"])
            want = T.lr_decay(s, t.max_step, cfg.lr_max, cfg.lr_min)
            assert abs(v - want) <= 2e-7 * cfg.lr_max, (s, v, want)                                 # fp32 vs double evaluation
            vals.append(v)
        assert vals[0] == float(torch.tensor(1e-4, dtype=torch.float32)) and abs(vals[-1] - 2.5e-6) < 1e-11    # lr_max at step 0, lr_min at max_step
        out["lr_decay_3d" if is3d else "lr_decay_2d"] = np.array(vals)
        t, rec, x, cfg = run_init_case(trainer_mod, trainer3_mod, is3d, "step", 0)
        a = [r for r in rec["assign"] if r["name"] == "g_lr_update"]
        assert abs(float(a[0]["value"]) - T.lr_step(cfg.lr_max, cfg.lr_min)) <= 1e-11
    out["lr_steps"] = np.array(steps)
    try:
        run_init_case(trainer_mod, trainer3_mod, False, "bogus", 0)
        raise AssertionError("invalid lr_update accepted")
    except Exception as e:                                                                        # trainer.py:80
        assert "Invalid lr update method" in str(e)
    return out


def main(reference_root="/root/reference", out_dir=os.path.join(ROOT, "tests", "golden")):
    trainer_mod, trainer3_mod, ref_ops = tf_shim.import_reference_trainers(reference_root)
    blob = {}
    for name in CASES:
        out, worst = run_case(trainer_mod, trainer3_mod, ref_ops, name)
        for k, v in out.items():
            blob[name + "/" + k] = v
        print("%-12s loss %.6f: reference wiring == oracle (max relative gradient difference %.1e)" % (name, float(out["loss"]), worst))
    for k, v in check_init(trainer_mod, trainer3_mod, ref_ops).items():
        blob["init/" + k] = v
    print("Trainer.__init__ / Trainer3.__init__: input wiring, output_shape, max_step, cosine and step LR schedules == oracle")
    np.savez_compressed(os.path.join(out_dir, "trainer_wiring.npz"), **blob)
    print("written", os.path.join(out_dir, "trainer_wiring.npz"))


if __name__ == "__main__":
    main()
