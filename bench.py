#!/usr/bin/env python
"""bench.py -- deep-fluids generator train step (fwd + curl/Jacobian loss + bwd + Adam) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2sq|c3|c4|c5|tiny] [--scaling weak|strong] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path over one batch of synthetic input = one `sess.run(g_optim)` of the reference
(trainer.py:269 / trainer3.py:156).  Prints ONE JSON line (rank 0).  Fields beyond the base contract:
  roofline      dominant kernel (tcgen05 conv, fwd+dgrad launches) achieved TFLOP/s vs the measured bf16 peak
  cpu_baseline  the CPU oracle (torch-CPU restatement of the reference path; TF1 cannot run here) on the host cores
  e2e           same metric through Trainer.train_step() with HOST (pinned) inputs, H2D + D2H inside the timed region
`--impl reference` times the oracle on the host cores (the reference's own CPU path is TensorFlow 1.15: not runnable).
"""
import argparse
import json
import math
import numpy as np
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "velocity_fields_per_sec_fwd_bwd"
UNIT = "fields/s"

# BASELINE.json configs (per-GPU batch; multi-GPU = weak scaling, batch sharded by replication of the per-GPU work)
WORKLOADS = {
    # name: (is_3d, res_z, res_y, res_x, per_gpu_batch, description)
    "c2": (False, 1, 128, 96, 64, "2D smoke_pos_size 128x96 generator+curl, batch 64/GPU (BASELINE configs[1])"),
    "c2sq": (False, 1, 128, 128, 64, "2D 128x128 generator+curl, batch 64/GPU (the metric's '128^2' grid; the reference's own 2D recipe is 128x96 = c2)"),
    "c3": (True, 64, 64, 64, 16, "3D smoke3_vel_buo 64^3 generator+curl, batch 16/GPU (BASELINE configs[2])"),
    "c4": (True, 128, 128, 128, 4, "3D smoke3_vel_buo 128^3 generator+curl+grad-loss, batch 4/GPU (BASELINE configs[3])"),
    "c5": (True, 128, 128, 128, 4, "3D AE 128^3 encoder-decoder (arch=ae), batch 4/GPU (BASELINE configs[4])"),
    "tiny": (False, 1, 32, 24, 4, "2D 32x24 plumbing check"),
}
ARCH = {"c5": "ae"}
# BASELINE configs[1] is quoted in fp32 (the reference's TF graphs are fp32): it runs the fp32-grade split-operand path
DEFAULT_PRECISION = {"c2": "fp32x3", "c2sq": "fp32x3"}
PRECISION_NOTE = {
    "bf16": "bf16 operands/activations, fp32 accumulate (TMEM), fp32 master weights + Adam",
    "fp32x3": "fp32-grade: activations/gradients/weights as (hi, lo) bf16 pairs (16-bit mantissa), x*w = 3 tcgen05 MMA "
              "terms (hi*hi + lo*hi + hi*lo), fp32 accumulate (TMEM), fp32 master weights + Adam",
}
# algorithmic conv FLOPs per field, fwd+bwd
FLOPS_PER_FIELD = {"c2": 58.0e9, "c2sq": 77.3e9, "c3": 3196.0e9, "c4": 25575.0e9, "c5": 49471.0e9, "tiny": None}
# strong scaling: the GLOBAL batch is fixed and split over the ranks; a rank whose share
# exceeds its per-GPU micro-batch accumulates gradients over share / micro-batch micro-steps (ONE all-reduce + ONE Adam
# update per optimizer step).  A "step" is then one optimizer step over the global batch.
STRONG_GLOBAL_BATCH = {"c2": 512, "c2sq": 512, "c3": 128, "c4": 32, "c5": 32, "tiny": 32}


# stdout carries exactly ONE line (the JSON): everything else any library prints to fd 1 (e.g. NCCL's version banner)
# is routed to stderr for the lifetime of the process
_REAL_STDOUT = None


def _guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_config(workload, extra=()):
    from deepfluids_b200 import config as C
    is3d, rz, ry, rx, b, _ = WORKLOADS[workload]
    argv = ["--synthetic=true", "--is_3d=%s" % ("true" if is3d else "false"), "--res_x=%d" % rx, "--res_y=%d" % ry,
            "--res_z=%d" % rz, "--batch_size=%d" % b, "--max_step=1000000", "--arch=%s" % ARCH.get(workload, "de")] + list(extra)
    cfg, _ = C.get_config(argv)
    return cfg


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (faithful torch-CPU restatement; kind = "port") on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_oracle_fields_per_sec(workload, steps, warmup, budget_s=25.0):
    import torch
    from oracle import ref_model as M          # allowed importer: bench.py cpu_baseline / --impl reference
    from oracle import ref_train as T
    is3d, rz, ry, rx, _, _ = WORKLOADS[workload]
    spatial = [rz, ry, rx] if is3d else [ry, rx]
    ncpu = os.cpu_count() or 1
    # measured on the GPU box (profiles/r01_cpu_threads_sweep.txt): oneDNN is fastest at 16 threads; at 128 threads
    # the same step is 200x slower (oversubscription), so "all the threads it can use" = min(cores, 16)
    cores = min(ncpu, 16)
    torch.set_num_threads(cores)
    flops = FLOPS_PER_FIELD[workload] or 1e9
    # bounded sample: batch sized for ~2 s per step at ~0.5 TFLOP/s, at least 1 field
    b = max(1, min(8, int(2.0 * 0.5e12 / flops)))
    cout = 3 if is3d else 1
    is_ae = ARCH.get(workload) == "ae"
    if is_ae:
        tab = M.ae_layout(spatial + [cout])
        warmup = 0                      # one 128^3 AE step is ~50 TFLOP: a single timed step is the bounded sample
    else:
        tab, _, _ = M.generator_layout(spatial + [cout])
    var = M.init_variables(tab, 123)
    opt = T.TFAdam(var, 0.5, 0.999)
    x, y = T.synthetic_batch(b, spatial, seed=123)
    ylast = torch.rand(b, 2) * 2 - 1
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        if is_ae:
            grads = T.ae_loss_and_grads(x, ylast, var, 2)[-1]
        else:
            loss, _, _, _, _, grads = T.generator_loss_and_grads(y, x, var)
        opt.step(var, grads, 1e-4)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_begin > budget_s and len(times) >= 1:
            break
    per_step = sum(times) / len(times)
    return {"value": b / per_step, "unit": UNIT, "cores": cores, "kind": "port", "steps_timed": len(times),
            "sample": "%d timed step(s) of batch %d (fwd+bwd+TF-Adam) of workload %s, fp32 torch-CPU/oneDNN oracle"
                      % (len(times), b, workload) + " (%d of %d host cores: fastest thread count measured)" % (cores, ncpu),
            "ms_per_step": per_step * 1e3, "batch": b}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_oracle_fields_per_sec(args.workload, max(1, args.steps), min(args.warmup, 1), budget_s=150.0)
    # `steps` is the number of steps actually TIMED (the bounded sample stops at its wall-clock budget); the request is
    # kept beside it so the line never claims more work than it measured
    out = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": r["steps_timed"], "steps_requested": args.steps, "warmup": min(args.warmup, 1),
           "ms_per_step": r["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOADS[args.workload][5], "batch_per_step": r["batch"],
                      "note": "reference CPU path = TensorFlow 1.15 (not installable); timed = oracle/ torch-CPU port"},
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from deepfluids_b200 import kernels as K
    from deepfluids_b200.data import BatchManager
    from deepfluids_b200.trainer import Trainer
    from deepfluids_b200.trainer3 import Trainer3

    precision = args.precision or DEFAULT_PRECISION.get(args.workload, "bf16")
    extra = ["--precision=%s" % precision]
    accum = 1
    if args.scaling == "strong":
        gb, micro = STRONG_GLOBAL_BATCH[args.workload], WORKLOADS[args.workload][4]
        assert gb % (micro * world) == 0, "global batch %d is not a multiple of %d ranks x micro-batch %d" % (gb, world, micro)
        accum = gb // (micro * world)
        extra.append("--grad_accum=%d" % accum)
    cfg = make_config(args.workload, extra)
    terms = 3 if precision == "fp32x3" else 1      # MMA terms executed per algorithmic multiply-add
    bm = BatchManager(cfg, device=dev, pool=2, rank=rank)
    tr = (Trainer3 if cfg.is_3d else Trainer)(cfg, bm)
    B = cfg.batch_size * accum          # fields per rank per (optimizer) step
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---------------- device-resident arm (`value`) ----------------
    def step_resident():
        tr.train_step()
        tr.update_lr(tr.step)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = K.PROF.launches
    ms = timed_region(step_resident, args.steps)
    launches = K.PROF.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---------------- end-to-end arm (`e2e`): host pinned inputs -> H2D -> train_step -> D2H loss ----------------
    xh = [x.cpu().pin_memory() for x, _ in bm._pool]
    yh = [y.cpu().pin_memory() for _, y in bm._pool]
    # Every step: one H2D copy of a pinned host batch (issued one step ahead on the copy stream, data.HostPrefetcher -- the
    # role of the reference's loader threads + FIFOQueue), the train step, one D2H read of the loss.  The loss of step i is
    # read (and checked) right after step i+1 has been enqueued, the way a training loop logs asynchronously; the last
    # step's loss is read before the timed region closes.
    from deepfluids_b200.data import HostPrefetcher
    pf = HostPrefetcher(bm._pool[0][0], bm._pool[0][1], dev)
    loss_h = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    cnt = [0]
    pending = [None]

    def read_loss(k):
        loss_ev[k].synchronize()
        assert loss_h[k][0] == loss_h[k][0], "NaN loss"

    class _HostSource(object):
        """micro-batches 2..accum of an optimizer step (strong scaling): pulled from the pinned host pool through the same
        one-step-ahead prefetcher, so every field of the step crosses PCIe inside the timed region"""
        def batch(self_inner):
            i = cnt[0]
            cnt[0] += 1
            pf.release()
            xd, yd = pf.get()
            pf.put(xh[(i + 1) % len(xh)], yh[(i + 1) % len(yh)])
            return xd, yd

    def step_e2e():
        i = cnt[0]
        cnt[0] += 1
        xd, yd = pf.get()                                # batch i (its copy was started during step i-1)
        pf.put(xh[(i + 1) % len(xh)], yh[(i + 1) % len(yh)])   # batch i+1 -> the other slot, under this step's kernels
        l3 = tr.train_step(xd, yd)
        pf.release()
        tr.update_lr(tr.step)
        loss_h[i & 1].copy_(l3, non_blocking=True)
        loss_ev[i & 1].record()
        if pending[0] is not None:
            read_loss(pending[0])                        # loss of step i-1
        pending[0] = i & 1

    def e2e_region(steps):
        for _ in range(steps):
            step_e2e()
        read_loss(pending[0])                            # the last step's loss, inside the timed region
        pending[0] = None

    if accum > 1:
        tr.batch_manager = _HostSource()
    pf.put(xh[0], yh[0])
    e2e_region(max(1, args.warmup // 2))
    ms_e = timed_region(lambda: e2e_region(args.steps), 1)
    e2e_value = B * world / (ms_e / args.steps * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in pf.slots[0]) * accum
    tr.batch_manager = bm

    # ---------------- per-kernel timing with CUDA events (2 extra instrumented steps, same stream) ----------------
    K.PROF.events = []
    tr.use_graph = False          # eager launches so every kernel can be bracketed by events
    tr.accum = 1                  # (per-kernel times are per micro-batch; with accumulation a step holds `accum` of them)
    for _ in range(2):
        tr.train_step()
    torch.cuda.synchronize()
    agg = {}
    for name, s, e, work, xwork in K.PROF.events:
        a = agg.setdefault(name, [0.0, 0.0, 0, 0.0])
        a[0] += s.elapsed_time(e) * 1e-3
        a[1] += work
        a[2] += 1
        a[3] += xwork
    K.PROF.events = None
    # every kernel of the step: ms per step and launches per step (CUDA events, eager replay of the same step)
    kernel_ms = {k: {"ms_per_step": round(v[0] / 2 * 1e3, 4), "launches_per_step": v[2] // 2}
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    # dominant kernel = the tap-window conv (its executed FLOPs == its algorithmic FLOPs); the per-tap kernel's launches
    # (stride-2 layers of the AE, phase-decomposed upsample-convs) are reported as their own entry below
    conv_t, conv_f, conv_n, conv_x = agg.get("conv_tc", [1e-9, 0.0, 1, 0.0])
    tap_t, tap_f, tap_n, tap_x = agg.get("conv_tap", [0.0, 0.0, 0, 0.0])
    wg_t, wg_f, wg_n, wg_x = agg.get("wgrad_tc", [1e-9, 0.0, 1, 0.0])
    st_t, st_b, st_n = agg.get("stencil_fused", [0.0, 0.0, 0, 0.0])[:3]
    fb_t, fb_b, fb_n = agg.get("lastconv_bwd_fused", [0.0, 0.0, 0, 0.0])[:3]
    stencil_note = "inside the step"
    if st_n == 0 and cfg.is_3d:
        # 3D: the loss stencil runs in the prologue of the fused first-backward kernel, so the step has no stencil launch.
        # The standalone kernel (dfl_stencil_loss_fwdbwd, the API entry point) is timed here on the step's own tensors,
        # outside the timed region, L2 flushed between launches, so the line still carries its roofline fraction.
        stencil_note = "standalone launches outside the timed region (in the step the stencil is fused into lastconv_bwd_fused_kernel)"
        xs, _ = bm._pool[0]
        pot_b = tr.engine.dec.pot if hasattr(tr.engine, "dec") else tr.engine.pot
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        dscr = torch.empty_like(pot_b)
        K.PROF.events = []
        for _ in range(4):
            flush.zero_()
            K.stencil_loss_fwdbwd(pot_b, xs, dpot=dscr)
        torch.cuda.synchronize()
        ev = [(s_.elapsed_time(e_) * 1e-3, w_) for n_, s_, e_, w_, _ in K.PROF.events if n_ == "stencil_fused"][1:]
        K.PROF.events = None
        st_t, st_b, st_n = sum(t for t, _ in ev), sum(w_ for _, w_ in ev), len(ev)
        del flush, dscr
    st_t = max(st_t, 1e-12)
    conv_f /= terms                     # PROF counts executed MMA flops; the roofline numerator is algorithmic flops
    wg_f /= terms
    # phase-decomposed upsample-conv launches are booked with the DENSE layer's algorithmic FLOPs (the roofline numerator
    # stays algorithmic); what the tensor cores executed is reported beside it
    eng_ = tr.engine.dec if hasattr(tr.engine, "dec") else tr.engine
    phase_on = bool(getattr(eng_, "phase", False))
    nd_ = 3 if cfg.is_3d else 2
    phase_note = None
    if phase_on:
        phase_note = ("first conv of every block after the first = conv3(upscale(s)): forward and data gradient run 2^nd taps "
                      "per output phase on the coarse tensor (%.3f of the dense FLOPs)%s" % (
                          (2.0 / 3.0) ** nd_, "; weight gradient = 4^nd-tap stride-2 correlation on the coarse grid (same ratio)"
                          if getattr(eng_, "phase_wgrad", False) else ""))
    achieved = conv_f / conv_t / 1e12
    roofline = {"kernel": "conv_tc2_kernel (tcgen05 tap-window implicit-GEMM conv, fwd + dgrad launches)", "bound": "tensor",
                "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_sustained"], "peak_source": "%s (sustained bf16, kernel timed inside a long step)" % peaks["src"],
                "mma_terms_per_flop": terms, "executed_tflops": conv_x / conv_t / 1e12,
                "frac_executed": conv_x / conv_t / 1e12 / peaks["tf_sustained"],
                "frac_note": "achieved / frac count the reference formulation's dense FLOPs (SURVEY 8d); with the phase "
                             "decomposition the tensor cores execute fewer (executed_tflops / frac_executed), so frac may exceed 1",
                "launches_timed": conv_n, "avg_launch_ms": conv_t / conv_n * 1e3, "traffic": load_traffic("conv_tc_kernel"),
                "traffic_source": "STATIC: dram bytes per launch of a full-resolution launch from the committed ncu --set full "
                                  "capture (profiles/ncu_summary.json), not measured in this run",
                "share_of_step": conv_t / 2 * accum / (ms_per_step * 1e-3),
                "others": {
                    "wgrad_tc_kernel": {"bound": "tensor", "achieved": wg_f / wg_t / 1e12, "unit": "TFLOP/s",
                                        "executed_tflops": wg_x / wg_t / 1e12,
                                        "frac_executed": wg_x / wg_t / 1e12 / peaks["tf_sustained"],
                                        "frac": wg_f / wg_t / 1e12 / peaks["tf_sustained"],
                                        "share_of_step": wg_t / 2 * accum / (ms_per_step * 1e-3)},
                    "stencil_fused_kernel": {"bound": "hbm", "achieved": st_b / st_t / 1e9, "unit": "GB/s",
                                             "frac": st_b / st_t / 1e9 / peaks["hbm_gbs"], "measured": stencil_note,
                                             "avg_launch_ms": st_t / max(st_n, 1) * 1e3}}}
    if tap_n:
        roofline["others"]["conv_tap_kernel"] = {
            "what": "per-tap tcgen05 conv: stride-2 layers" + ("; " + phase_note if phase_note else ""),
            "bound": "tensor", "achieved": tap_f / terms / tap_t / 1e12, "unit": "TFLOP/s",
            "achieved_is": "ALGORITHMIC FLOPs of the dense layer / time (can exceed the peak: the decomposition skips work)",
            "executed_tflops": tap_x / tap_t / 1e12, "frac_executed": tap_x / tap_t / 1e12 / peaks["tf_sustained"],
            "share_of_step": tap_t / 2 * accum / (ms_per_step * 1e-3)}
    if fb_n:
        roofline["others"]["lastconv_bwd_fused_kernel"] = {
            "what": "loss stencil (curl + Jacobian-L1 + adjoints) in the prologue of the output conv's backward (dgrad + wgrad + bias-grad)",
            "bound": "hbm", "achieved": fb_b / fb_t / 1e9, "unit": "GB/s", "frac": fb_b / fb_t / 1e9 / peaks["hbm_gbs"],
            "algorithmic_bytes_per_voxel": 1048 if cfg.is_3d else 1036, "avg_launch_ms": fb_t / fb_n * 1e3,
            "share_of_step": fb_t / 2 * accum / (ms_per_step * 1e-3)}
        if cfg.is_3d:
            # The same kernel timed ALONE on the step's own tensors (outside the timed region, L2 flushed between launches): inside
            # the power-capped step the SM clock is ~1.5 GHz and this kernel's time scales with it, so the in-step figure above and
            # this one bracket it (tools/lastconv_bwd_bench.py measures the same thing on synthetic tensors).
            try:
                fa = tr._fused_args(bm._pool[0][0])
                top, nc = eng_.rep - 1, eng_.num_conv
                P_ = eng_.params
                flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
                torch.cuda.synchronize()
                time.sleep(1.5)                  # leave the power-capped state of the step: "alone" means at the boost clock
                K.PROF.events = []
                sa_mhz = []
                for _ in range(6):
                    flush.zero_()
                    K.lastconv_curl_loss_bwd(eng_.s, eng_.pot, fa["x"], P_.p(eng_.last_name + "/weights"), eng_.y[top][nc - 1],
                                             eng_._gview(0, top), eng_._gview(1, top), P_.g(eng_.last_name + "/weights"),
                                             P_.g(eng_.last_name + "/biases"), fa["loss3"], fa["workspace"], fa["w1"], fa["w2"], 1.0)
                    torch.cuda.synchronize()
                    try:
                        sa_mhz.append(int(torch.cuda.clock_rate(dev)))
                    except Exception:
                        pass
                    time.sleep(0.05)
                ev = [(s_.elapsed_time(e_) * 1e-3, w_) for n_, s_, e_, w_, _ in K.PROF.events if n_ == "lastconv_bwd_fused"][1:]
                K.PROF.events = None
                del flush
                sa_t, sa_b = sum(t for t, _ in ev), sum(w_ for _, w_ in ev)
                roofline["others"]["lastconv_bwd_fused_kernel"]["standalone"] = {
                    "measured": "the same kernel on the step's own tensors, timed alone after the timed region (1.5 s idle first, "
                                "one launch at a time, L2 flushed between launches): SM clock not held down by the step's power cap",
                    "sm_mhz_after_launch": sa_mhz[1:],
                    "avg_launch_ms": sa_t / len(ev) * 1e3, "achieved": sa_b / sa_t / 1e9, "unit": "GB/s",
                    "frac": sa_b / sa_t / 1e9 / peaks["hbm_gbs"]}
            except Exception as ex:      # a diagnostic extra must never cost the bench line
                K.PROF.events = None
                roofline["others"]["lastconv_bwd_fused_kernel"]["standalone"] = {"error": repr(ex)[:200]}
    roofline["kernel_ms"] = kernel_ms
    fl = FLOPS_PER_FIELD[args.workload]
    if fl:
        roofline["step_conv_tflops_per_gpu"] = value / world * fl / 1e12      # algorithmic (dense-layer) FLOPs
        roofline["step_frac_of_peak"] = value / world * fl / 1e12 / peaks["tf_sustained"]
        if phase_on:
            exec_per_step = (conv_x + wg_x + tap_x) / 2 * accum
            roofline["step_executed_tflops_per_gpu"] = exec_per_step / (ms_per_step * 1e-3) / 1e12
            roofline["step_executed_frac_of_peak"] = roofline["step_executed_tflops_per_gpu"] / peaks["tf_sustained"]
            roofline["step_note"] = ("step_conv_tflops_per_gpu / step_frac_of_peak count the DENSE layers' algorithmic FLOPs; the "
                                     "phase-decomposed upsample-convs execute 8/27 (3D) / 4/9 (2D) of theirs: " + phase_note)

    phase_on_cfg = phase_on
    if rank == 0:
        cpu = cpu_oracle_fields_per_sec(args.workload, 3, 1) if world == 1 and not args.no_cpu else None
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
               "vs_baseline": None, "dtype": "bf16" if terms == 1 else "f32 (bf16x3 split operands)", "data": "synthetic",
               "config": {"workload": WORKLOADS[args.workload][5], "global_batch": B * world, "per_gpu_batch": B,
                          "micro_batch": cfg.batch_size, "grad_accum_micro_steps": accum,
                          "parallelism": "dp%d" % world, "filters": 128, "num_conv": 4, "phase_upconv": phase_on_cfg,
                          "l2": "per-step working set (>= %.0f MB of activations) exceeds the 126 MB L2; no flush needed"
                                % (B * float(np.prod(bm._pool[0][0].shape[1:-1])) * 128 * 2 * 6 / 1e6),
                          "precision": PRECISION_NOTE[precision]},
               "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                         "d2h_bytes_per_step": 12, "ms_per_step": ms_e / args.steps},
               "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def load_traffic(kernel):
    """dram bytes per launch from the committed ncu summary (profiles/ncu_summary.json), else null."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        return json.load(open(p))[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def main():
    # a stuck run (cold-box import, a hung kernel) leaves its Python stacks on stderr instead of a silent timeout
    import faulthandler
    faulthandler.dump_traceback_later(170, repeat=True, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", type=str, default=None, choices=["bf16", "fp32x3"],
                    help="default: fp32x3 for c2 (BASELINE quotes it in fp32), bf16 otherwise")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", type=str, default="weak", choices=["weak", "strong"],
                    help="weak: per-GPU batch fixed (default).  strong: global batch fixed (c4/c5: 32), gradient accumulation "
                         "over global/(micro*N) micro-steps per optimizer step")
    args = ap.parse_args()
    _guard_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
