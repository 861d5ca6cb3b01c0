"""Benchmark of the Mix-StAGE generator hot path on B200 (contract: see the task statement).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU

Headline workload (BASELINE.json configs[1]): full GAN train step of JointLateClusterSoftStyle4_G + pose discriminator
(gan=1, L1Loss), batch 16 per GPU, 64-frame windows, 4 speakers, num_clusters=8, fp64 master parameters and inputs as the
reference's trainer keeps them (trainer.py:138), synthetic N(0,1) inputs and seeded synthetic weights.  A "step" is one
training iteration: zero_grad, GAN.forward (generator step on even iterations, discriminator step on odd ones -- the
reference flips a fair coin, gan.py:105), backward, gradient all-reduce when N>1, clip_grad_norm_(1) and Adam(1e-4)
(trainer.py:1138-1146).  Metric: pose sequences (64-frame windows) per second, whole job.

The metric also names the forward (inference) half and the 8-GPU configurations, so the SAME JSON line carries, under
"configs", one object per further BASELINE config measured in the same process with the same rules (device-timed value,
e2e through the public API from pinned host buffers, roofline of the dominant kernel family, CPU baseline at N=1):
  configs[2]  style-sweep inference, B=1024 windows per GPU x 4 target styles (no collective)
  configs[3]  data-parallel training slice: B=128 per GPU (global 1024 at N=8), 8 speakers
  configs[4]  stress shape: 25 speakers, 16 clusters, soft style, 256-frame windows, bf16 (run when N=8 or --workload all)
`--workload {train,infer,config3,config4}` runs one of them alone as the headline line.
"""
import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

UNIT = "sequences/s"
MOD = ["audio/log_mel_400"]
DTYPE_NAME = {"fp32": "f32", "bf16x3": "bf16x3 (split-bf16 operands, f32 accumulate)", "bf16": "bf16 (f32 accumulate)"}

# name -> (config label, B per GPU, speakers, clusters, argmax, T, default precision)
TRAIN_WORKLOADS = {
    "train": ("configs[1]", 16, 4, 8, 1, 64, "bf16x3"),
    "config3": ("configs[3] per-GPU slice (global batch 1024 at N=8)", 128, 8, 8, 1, 64, "bf16x3"),
    "config4": ("configs[4] stress", 16, 25, 16, 0, 256, "bf16"),
}


def train_workload_name(label, B, T, S, K, am):
    """Canonical name of a train workload: the SAME string on the CUDA arm and on the reference arm (the driver compares
    the two `config` objects); implementation details (precision, graphs, exchange) go to `config_detail`."""
    return ("%s: GAN train step (G on even, D on odd iterations), B=%d per GPU, T=%d, S=%d, K=%d, %s style, gan=1, L1Loss, "
            "fp64 master params/inputs" % (label, B, T, S, K, "argmax" if am else "soft"))


def infer_workload_name(B, S):
    return ("configs[2]: inference sweep, B=%d windows per GPU x S=%d target styles per step (sample_all_styles), T=64, K=8, "
            "eval mode, fp64 parameters/inputs/outputs" % (B, S))


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel family from the committed ncu --set full capture (profiles/), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))[workload]
            return d["traffic_bytes_per_launch"], d["source"]
        except Exception:
            continue
    return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own modules when its sources are on this machine
# (/root/reference in the build container, baseline/_ref/**/src when a pod ships them), else the
# oracle port of the same algorithm -- always on the host cores
# --------------------------------------------------------------------------------------------
def find_reference_src():
    cands = [os.environ.get("MIXSTAGE_REFERENCE_SRC"), "/root/reference/src"]
    cands += sorted(glob.glob(os.path.join(ROOT, "baseline", "_ref", "**", "src"), recursive=True))
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "model", "joint_late_cluster_soft_style.py")):
            return c
    return None


def _spec(S, K=8, argmax=1, T=64):
    import mixstage_oracle as O
    return O.Spec(num_speakers=S, num_clusters=K, argmax=argmax, time_steps=T)


def _reference_train_steps(src, spec, B, T, steps, warmup):
    """The reference's own classes (unmodified files under `src`), its trainer's loop body restated around them:
    zero_grad, GAN.forward, sum of losses, backward, clip_grad_norm_(1), Adam(1e-4) (trainer.py:1138-1146,1268-1285)."""
    os.environ["MIXSTAGE_REFERENCE_SRC"] = src
    import importlib
    import ref_loader
    importlib.reload(ref_loader)
    import make_golden
    import mixstage_oracle as O
    ns = ref_loader.load()
    G, D, gan = make_golden.build(ns, spec, T)
    G.thresh.value, G.thresh.iters = 1.0, 1000
    audio, pose, labels, style = O.synth_inputs(B, T, spec)
    optG = torch.optim.Adam(G.parameters(), lr=1e-4)
    optD = torch.optim.Adam(D.parameters(), lr=1e-4)
    gan.train()
    orig = torch.rand

    def one(kind):
        G.zero_grad()
        D.zero_grad()
        torch.rand = make_golden.FixedRand([0.9 if kind == "G" else 0.1, 0.5])
        try:
            fake, il, _ = gan([audio.clone(), labels], pose, input_modalities=MOD, style=style, desc="train", sample_flag=0,
                              description="train")
        finally:
            torch.rand = orig
        sum(il).backward()
        net, opt = (G, optG) if kind == "G" else (D, optD)
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1)
        opt.step()

    for i in range(warmup):
        one("G" if i % 2 == 0 else "D")
    t0 = time.perf_counter()
    for i in range(steps):
        one("G" if i % 2 == 0 else "D")
    return time.perf_counter() - t0


def _port_train_steps(spec, B, T, steps, warmup, dtype=torch.float64):
    """The oracle's restatement of one reference train step (fwd + bwd + clip + Adam)."""
    import mixstage_oracle as O
    from oracle_cases import D_SEED, G_SEED, leafify
    sd = leafify(O.synth_state(O.g_state_shapes(spec), G_SEED, dtype))
    sdd = leafify(O.synth_state(O.d_state_shapes(spec.out_feats), D_SEED, dtype))
    audio, pose, labels, style = O.synth_inputs(B, T, spec, dtype=dtype)
    gleaf = [v for k, v in sd.items() if v.requires_grad and not k.startswith("style_dec_gr.")]
    dleaf = [v for v in sdd.values() if v.requires_grad]
    state = {id(v): (torch.zeros_like(v), torch.zeros_like(v)) for v in gleaf + dleaf}
    nstep = {"G": 0, "D": 0}

    def one(kind):
        for v in gleaf + dleaf:
            v.grad = None
        lg, ld = O.BNLog(), O.BNLog()
        fake, losses, _ = O.gan_forward(sd, sdd, spec, audio, labels, pose, style, step=kind, log_g=lg, log_d=ld)
        sum(losses).backward()
        leaves = gleaf if kind == "G" else dleaf
        used = [v for v in leaves if v.grad is not None]
        nstep[kind] += 1
        with torch.no_grad():
            O.clip_and_adam(used, [v.grad for v in used], [state[id(v)][0] for v in used],
                            [state[id(v)][1] for v in used], nstep[kind])
            for k, v in lg.updates.items():
                sd[k] = v
            for k, v in ld.updates.items():
                sdd[k] = v

    for i in range(warmup):
        one("G" if i % 2 == 0 else "D")
    t0 = time.perf_counter()
    for i in range(steps):
        one("G" if i % 2 == 0 else "D")
    return time.perf_counter() - t0


def cpu_train_steps(spec, B, T, steps, warmup):
    """(seconds for `steps` train steps on all host cores, kind, description of what ran)."""
    torch.set_num_threads(os.cpu_count())
    src = find_reference_src()
    if src is not None:
        try:
            return _reference_train_steps(src, spec, B, T, steps, warmup), "reference", "unmodified reference modules from %s" % src
        except Exception as e:            # missing dependency of the reference on this box: say so, fall back to the port
            sys.stderr.write("bench.py: reference modules under %s did not run (%r); timing the oracle port\n" % (src, e))
    return _port_train_steps(spec, B, T, steps, warmup), "port", "oracle port of the reference algorithm (torch CPU fp64)"


def cpu_infer_sweep(spec, B, T, steps, warmup, dtype=torch.float64):
    """The reference sampling loop on the host: S full forwards per batch (trainer.py:791-794, 1367-1386)."""
    import mixstage_oracle as O
    from oracle_cases import G_SEED
    torch.set_num_threads(os.cpu_count())
    S = spec.num_speakers
    audio, pose, labels, style = O.synth_inputs(B, T, spec, dtype=dtype)
    src = find_reference_src()
    kind, what = "port", "oracle port (torch CPU fp64)"
    one = None
    if src is not None:
        try:
            os.environ["MIXSTAGE_REFERENCE_SRC"] = src
            import importlib
            import ref_loader
            importlib.reload(ref_loader)
            import make_golden
            G, _, _ = make_golden.build(ref_loader.load(), spec, T)
            G.eval()

            def one():
                with torch.no_grad():
                    for shift in range(S):
                        G([audio.clone(), labels], pose, input_modalities=MOD, style=(style + shift) % S, sample_flag=1,
                          description="test")
            one()
            kind, what = "reference", "unmodified reference modules from %s" % src
        except Exception as e:
            sys.stderr.write("bench.py: reference modules under %s did not run (%r); timing the oracle port\n" % (src, e))
            one = None
    if one is None:
        sd = O.synth_state(O.g_state_shapes(spec), G_SEED, dtype)

        def one():
            with torch.no_grad():
                for shift in range(S):
                    O.g_forward(sd, spec, audio, labels, pose, (style + shift) % S, training=False, sample_flag=1,
                                description="test")
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    return time.perf_counter() - t0, kind, what


def reference_line(args):
    """--impl reference: the reference's CPU implementation of the workload on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload if args.workload != "all" else "train"
    steps, warmup = args.steps, args.warmup
    cores = os.cpu_count()
    if wl == "infer":
        B, S = args.batch or 1024, 4
        cb = 16                                  # bounded sample: the CPU sweeps B=16 windows per step, same shapes otherwise
        spec = _spec(S)
        dt, kind, what = cpu_infer_sweep(spec, cb, 64, steps, warmup)
        val = steps * cb * S / dt
        metric = "pose sequences/sec (64-frame windows), inference style sweep"
        workload = infer_workload_name(B, S)
        sample = "%d sweeps of B=%d x S=%d styles after %d warm-up, %s, %d threads" % (steps, cb, S, warmup, what, cores)
        gb = B * args.gpus
    else:
        label, B0, S0, K, am, T, _ = TRAIN_WORKLOADS[wl]
        B = args.batch or B0
        cb = min(B, 16)                          # bounded sample of the same workload: at most 16 sequences per CPU step
        S = args.speakers or S0
        spec = _spec(S, K, am, T)
        dt, kind, what = cpu_train_steps(spec, cb, T, steps, warmup)
        val = steps * cb / dt
        metric = "pose sequences/sec (%d-frame windows), GAN train step" % T
        workload = train_workload_name(label if (B, S) == (B0, S0) else "variant of " + label, B, T, S, K, am)
        sample = "%d train steps (G/D alternating) B=%d after %d warm-up, %s, %d threads" % (steps, cb, warmup, what, cores)
        gb = B * args.gpus
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / max(1, steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "global_batch": gb, "parallelism": "dp%d" % args.gpus},
        "config_detail": {"arm": "reference algorithm on the host cores (rank 0 only), fp64, %d threads" % cores},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.pk, self.pk_kind = peaks()

    def init_dist(self):
        if self.world > 1 and not self.dist.is_initialized():
            # the overlapped gradient exchange runs beside persistent chain launches on the SMs TrainStep leaves free
            os.environ.setdefault("NCCL_MAX_CTAS", "16")
            self.dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, nsteps, body, host_in_loop=False, finish=None):
        """EXACTLY nsteps steps between barrier + synchronize on both sides, CUDA events on the launch stream, max over
        ranks.  host_in_loop: the e2e leg, whose host read-backs belong to the step (wall clock bounds it from below)."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(nsteps):
            body(i)
        if finish is not None:
            finish()
        e1.record()
        self.barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if host_in_loop:
            ms = max(ms, wall * 1e3)
        return self.max_over_ranks(ms)


def _instrument(ops, names, run, busy_ms=60):
    """CUDA events around every launch of the named C-ABI entry points during run().  The stream is first kept busy by a
    spin kernel so that the host queues the whole instrumented step ahead of the GPU: every (event, kernel, event) triple
    then executes back to back and the interval holds the kernel alone, not the host's launch latency."""
    records = []
    orig_call = ops.call

    def timed_call(name, *a):
        if name in names or any(name.startswith(n) for n in names if n.endswith("_")):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_call(name, *a)
            e1.record()
            records.append((name, ops.last_gemm_flops, e0, e1))
        else:
            orig_call(name, *a)

    ops.call = timed_call
    try:
        torch.cuda.synchronize()
        torch.cuda._sleep(int(busy_ms * 1.9e6))
        run()
    finally:
        ops.call = orig_call
    torch.cuda.synchronize()
    return [(n, f, e0.elapsed_time(e1)) for n, f, e0, e1 in records]


GEMM_ENTRY_POINTS = ("ms_igemm_bf16", "ms_igemm_bf16_fused", "ms_igemm_bf16_mix", "ms_wgrad_bf16", "ms_conv_block_train_fwd",
                     "ms_conv_block_train_bwd", "ms_wgrad_bf16_acc", "ms_conv_chain_fwd", "ms_conv_chain_bwd",
                     "ms_wgrad_bf16_acc_multi")


def _roofline(ctx, recs, kernel_desc, traffic_key):
    ms = sum(r[2] for r in recs)
    fl = sum(r[1] for r in recs)
    peak_tf = ctx.pk.get("bf16_tflops_sustained", ctx.pk["bf16_tflops"])
    achieved = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
    tr, src = ncu_traffic(traffic_key)
    return {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "traffic": tr, "traffic_source": src, "kernel": kernel_desc % len(recs), "launches_timed": len(recs),
            "kernel_ms_total": ms, "algorithmic_gflop": fl / 1e9, "peak_source": "%s bf16 sustained" % ctx.pk_kind}


def bench_train(ctx, args, wl, steps, warmup, headline):
    import mixstage_b200 as M
    from mixstage_b200 import _lib, ops, parallel
    import mixstage_oracle as O
    from model_cases import build as build_model
    label, B0, S0, K, am, T, prec0 = TRAIN_WORKLOADS[wl]
    B = (args.batch if headline and args.batch else B0)
    S = (args.speakers if headline and args.speakers else S0)
    prec = (args.precision if headline and args.precision else prec0)
    ops.set_precision(prec)
    spec = _spec(S, K, am, T)
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    G, D, gan = build_model(spec, T, dev, torch.float64)
    gan.train()
    G.thresh.value, G.thresh.iters = 1.0, 1000            # past the curriculum: audio branch
    parallel.sync_host_rng(11212)
    ts = M.TrainStep(gan, lr=1e-4, max_norm=1.0, use_graphs=not args.no_graphs, overlap_allreduce=not args.no_overlap,
                     exchange_dtype=args.exchange_dtype)

    # per-rank synthetic shard (weak scaling: B sequences per GPU), pinned host copies for the e2e leg
    audio, pose, labels, style = O.synth_inputs(B, T, spec, seed=11212 + rank)
    host = [t.pin_memory() for t in (audio, labels, pose, style)]
    resident = [t.to(dev) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)
    # e2e leg: every step's five losses AND the generated poses are read on the host (the reference moves y_cap to the
    # host every iteration, trainer.py:654), through pinned buffers and one step behind the launch, so the host enqueues
    # step i+1 (H2D of its batch into the static buffers, graph replay) while the GPU still runs step i
    loss_pin = [torch.empty(5, dtype=torch.float64).pin_memory() for _ in range(2)]
    fake_pin = [torch.empty((B, T, spec.out_feats), dtype=torch.float64).pin_memory() for _ in range(2)]
    d2h = loss_pin[0].numel() * 8 + fake_pin[0].numel() * 8
    ev = [torch.cuda.Event() for _ in range(2)]
    log = []

    def step(i, batch, read_back):
        fake, losses = ts.step(*batch, kind="G" if i % 2 == 0 else "D")
        if read_back:
            b = i & 1
            loss_pin[b].copy_(losses, non_blocking=True)
            fake_pin[b].copy_(fake, non_blocking=True)
            ev[b].record()
            if i > 0:
                ev[b ^ 1].synchronize()
                log.append(float(loss_pin[b ^ 1][0]) + float(fake_pin[b ^ 1][0, 0, 0]))

    def drain():
        b = (steps - 1) & 1
        ev[b].synchronize()
        log.append(float(loss_pin[b][0]))

    nwarm = max(warmup, 4)                                 # at least one replay of each captured graph before timing
    for i in range(nwarm):
        step(i, resident, False)
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES + ts.launched
    ms = ctx.timed(steps, lambda i: step(i, resident, False))
    launches = _lib.LAUNCHES + ts.launched - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = ctx.timed(steps, lambda i: step(i, host, True), host_in_loop=True, finish=drain)
    per_graph = dict(("%s%s" % (k[0], "+pose" if k[1] else ""), v) for k, v in ts.kernels_per_graph.items())

    # ---- roofline of the dominant kernel family (tcgen05 implicit GEMMs, incl. the fused train-block kernels): one
    # instrumented G+D pair launched eagerly behind a spin kernel
    graphs_on = ts.use_graphs
    ts.use_graphs = False

    def pair():
        step(0, resident, False)
        step(1, resident, False)
    try:
        recs = _instrument(ops, GEMM_ENTRY_POINTS, pair, busy_ms=150)
    finally:
        ts.use_graphs = graphs_on
    total = steps * B * world
    out = {
        "metric": "pose sequences/sec (%d-frame windows), GAN train step" % T, "value": total / (ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": steps, "warmup": nwarm, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NAME[prec], "data": "synthetic",
        "config": {"workload": train_workload_name(label if (B, S) == (B0, S0) else "variant of " + label, B, T, S, K, am),
                   "global_batch": B * world, "parallelism": "dp%d" % world},
        "config_detail": {"precision": prec, "launch": "CUDA-graph replay" if ts.use_graphs else "eager launches",
                          "exchange": None if world == 1 else (
                              "fp32 weight-gradient accumulators all-reduced (NCCL AVG) %s; the ~1 %% of gradients outside them "
                              "gathered into one %s vector" % (
                                  "in buckets sent from backward hooks on a communication stream while the chain launches "
                                  "leave 16 SMs free" if not args.no_overlap else "after backward",
                                  "fp32" if args.exchange_dtype == "fp32" else "fp64")),
                          "l2": "no explicit flush: per-step working set (master params + packed weights + grads + Adam state "
                                "~0.9 GB) exceeds the 126 MB L2"},
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "mixstage_b200.TrainStep.step(pinned host batch); every step's losses and generated poses copied to "
                       "pinned host memory and read one step behind the launch (%d reads)" % len(log)},
        "gpu_launches": launches, "kernels_per_graph": per_graph,
        "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clocks else None,
        "roofline": _roofline(ctx, recs, "tcgen05 chain kernels (conv_chain_fwd / conv_chain_bwd: implicit GEMM + BatchNorm "
                              "statistics / normalise / backward of up to 12 blocks per launch) + chain weight-gradient launches, "
                              "%d launches per G+D step pair; algorithmic GEMM FLOPs over the WHOLE launch time (element-wise "
                              "phases and device barriers included)" + (", split-bf16 issues 3x the MMAs" if prec == "bf16x3" else ""),
                              "train"),
    }
    # whole-step fraction: algorithmic FLOPs of the alternating G/D loop / device-timed step
    out["roofline"]["whole_step_frac"] = (_step_gflop(spec, T) * B / (ms / steps * 1e-3) / 1e3) / out["roofline"]["peak"]
    if world > 1:
        # data-parallel replicas must hold bit-identical parameters after the timed steps (same reduced gradients, same update)
        chk = torch.stack([ts.fG.p.sum(), ts.fG.p.abs().sum(), ts.fD.p.sum(), ts.fD.p.abs().sum()]).to(torch.float64)
        allc = [torch.empty_like(chk) for _ in range(world)]
        ctx.dist.all_gather(allc, chk)
        out["replicas_in_sync"] = bool(all(torch.equal(allc[0], c) for c in allc))
    out["cpu_baseline"] = _cpu_leg_train(ctx, spec, B, T)
    del ts, G, D, gan
    torch.cuda.empty_cache()
    return out


def _step_gflop(spec, T):
    """Algorithmic GFLOP per sequence and per step of the alternating G/D loop ((G-step + D-step) / 2), SURVEY.md section
    8(d): G-step ~ 3 x train forward, D-step ~ generator forward + 2 x 3 x D; MMAC per sequence from Appendix A (T=64,
    K=8), scaled linearly in T and (sub-decoders) in K."""
    k, t = spec.num_clusters / 8.0, T / 64.0
    audio, unet, cls = 480.51, 65.80, 76.12
    dec, logit = 406.59 * k, 12.58 * k
    pse, d = 3.80, 3.29
    fwd_eval = (audio + unet + cls + dec + logit) * t
    fwd_train = fwd_eval + (2 * pse + d) * t
    g_step = 3 * fwd_train
    d_step = fwd_eval + pse * t + 2 * 3 * d * t
    return (g_step + d_step) / 2 * 2 / 1e3


def _cpu_leg_train(ctx, spec, B, T):
    if ctx.world > 1 or ctx.rank != 0:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": None,
                "sample": "not run: the CPU leg is measured on rank 0 at N=1 only"}
    cb = min(B, 16)
    n = 4 if T <= 64 else 2
    dt, kind, what = cpu_train_steps(spec, cb, T, n, 1)
    return {"value": n * cb / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
            "sample": "%d train steps (G/D alternating) B=%d after 1 warm-up, %s" % (n, cb, what)}


def bench_infer(ctx, args, steps, warmup, headline):
    """One step = the reference's sampling inner loop for one batch (trainer.py:791-794 with update_kwargs :1367-1386):
    B windows pushed through G.forward once per target style ((style + shift) % S), eval mode, no_grad.  Sequences per
    step = B * S.  Sharded over ranks with no collective.  The sweep of one batch is captured into a CUDA graph per
    input buffer (device-timed leg and e2e leg alike)."""
    from mixstage_b200 import _lib, ops
    import mixstage_oracle as O
    from model_cases import build as build_model
    prec = (args.precision if headline and args.precision else "bf16")
    ops.set_precision(prec)
    B, S, T = (args.batch if headline and args.batch else 1024), 4, 64
    spec = _spec(S)
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    G, D, gan = build_model(spec, T, dev, torch.float64)
    G.eval()
    G.thresh.value, G.thresh.iters = 1.0, 1000

    # two alternating per-rank batches (static device buffers): consecutive steps never see the same audio, so the
    # style-sweep cache (encoder output reused across the S styles of ONE batch) cannot carry work between steps
    hosts, static, graphs, outs = [], [], [], []
    for j in range(2):
        audio, pose, labels, style = O.synth_inputs(B, T, spec, seed=11212 + 2 * rank + j)
        h = [t.pin_memory() for t in (audio, labels, pose, style)]
        hosts.append(h)
        static.append([t.to(dev) for t in h])
    h2d = sum(t.numel() * t.element_size() for t in hosts[0])
    out_host = torch.empty((S, B, T, spec.out_feats), dtype=torch.float64).pin_memory()
    d2h = out_host.numel() * out_host.element_size()

    def sweep(j):
        audio, labels, pose, style = static[j]
        res = []
        with torch.no_grad(), G.style_sweep():
            for shift in range(S):
                st = (style + shift) % S
                out, _ = G([audio, labels], pose, input_modalities=MOD, style=st, sample_flag=1, description="test")
                res.append(out)
        return res

    use_graphs = not args.no_graphs
    launches_per_sweep = None
    if use_graphs:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for j in range(2):
                sweep(j)                      # warm-up: allocator, packed weights, folded BatchNorm
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for j in range(2):
            g = torch.cuda.CUDAGraph()
            l0 = _lib.LAUNCHES
            with torch.cuda.graph(g):
                res = sweep(j)
            launches_per_sweep = _lib.LAUNCHES - l0
            graphs.append(g)
            outs.append(res)

    def step(i, e2e):
        j = i % 2
        if e2e:
            for s_, h_ in zip(static[j], hosts[j]):
                s_.copy_(h_, non_blocking=True)
        if use_graphs:
            graphs[j].replay()
            res = outs[j]
        else:
            res = sweep(j)
        if e2e:
            for shift in range(S):
                out_host[shift].copy_(res[shift], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    nwarm = max(warmup, 3)
    for i in range(nwarm):
        step(i, False)
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES
    ms = ctx.timed(steps, lambda i: step(i, False))
    launches = (launches_per_sweep * steps) if use_graphs else (_lib.LAUNCHES - l0)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = ctx.timed(steps, lambda i: step(i, True), host_in_loop=True)
    recs = _instrument(ops, GEMM_ENTRY_POINTS, lambda: sweep(0), busy_ms=60)
    total = steps * B * S * world
    line = {
        "metric": "pose sequences/sec (64-frame windows), inference style sweep", "value": total / (ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": steps, "warmup": nwarm, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NAME[prec], "data": "synthetic",
        "config": {"workload": infer_workload_name(B, S), "global_batch": B * world, "parallelism": "dp%d" % world},
        "config_detail": {"precision": prec, "collective": "none", "cache": "encoder+UNet computed once per batch and reused across styles",
                          "launch": "one CUDA graph per input buffer" if use_graphs else "eager launches",
                          "l2": "two alternating input batches; per-step activations (>1 GB at B=1024) exceed the 126 MB L2"},
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "G.forward per style on a batch copied from pinned host memory + fp64 pose outputs copied back to pinned host memory"},
        "gpu_launches": launches,
        "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clocks else None,
        "roofline": _roofline(ctx, recs, "igemm_tc_persist_kernel / igemm_tc_pair_kernel incl. the mixture GEMMs, fused inference "
                              "epilogue (%d launches per step; algorithmic FLOPs" + (", split-bf16 issues 3x the MMAs)" if prec == "bf16x3" else ")"),
                              "infer"),
    }
    # whole-step fraction: UNPRUNED algorithmic FLOPs of the sweep (encoder once + S x the style-dependent half) / step time
    enc, rest = (480.51 + 65.80) * 2e-3, (76.12 + 406.59 + 12.58) * 2e-3
    line["roofline"]["whole_step_frac"] = ((enc + S * rest) * B / (ms / steps * 1e-3) / 1e3) / line["roofline"]["peak"]
    if world > 1 or rank != 0:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": None,
                                "sample": "not run: the CPU leg is measured on rank 0 at N=1 only"}
    else:
        cb = 16
        dt, kind, what = cpu_infer_sweep(spec, cb, T, 2, 1)
        line["cpu_baseline"] = {"value": 2 * cb * S / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                                "sample": "2 sweeps of B=%d x S=%d styles after 1 warm-up, %s, encoder recomputed per style as "
                                          "the reference does" % (cb, S, what)}
    del graphs, outs, G, D, gan
    torch.cuda.empty_cache()
    return line


def run_cuda(args):
    from mixstage_b200 import _lib
    ctx = Ctx()
    ctx.init_dist()
    _lib.load()
    wl = args.workload
    head = "train" if wl == "all" else wl
    if head == "infer":
        line = bench_infer(ctx, args, args.steps, args.warmup, True)
    else:
        line = bench_train(ctx, args, head, args.steps, args.warmup, True)
    if wl == "all":
        # the other BASELINE configs ride on the same line (fewer timed steps each, the same timing rules)
        extra = {}
        sub = max(6, min(args.steps, 10))
        for name in ("infer", "config3") + (("config4",) if (ctx.world == 8 or args.all_configs) else ()):
            try:
                if name == "infer":
                    extra["configs[2]"] = bench_infer(ctx, args, sub, 3, False)
                else:
                    extra["configs[3]" if name == "config3" else "configs[4]"] = bench_train(ctx, args, name, sub, 4, False)
            except Exception as e:            # a failing extra must not take the headline down with it
                extra[name] = {"error": repr(e)[:400]}
        line["configs"] = extra
    if ctx.rank == 0:
        print(json.dumps(line))
    if ctx.world > 1:
        _leave(ctx.dist)


def _leave(dist):
    """End of a multi-rank run.  Tearing NCCL down while captured CUDA graphs still hold all-reduce nodes of that
    communicator hung on the 2-GPU box (both ranks inside destroy_process_group), so: everything measured is already
    printed, meet once more, then leave without running the communicator's destructor."""
    sys.stdout.flush()
    sys.stderr.flush()
    try:
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--batch", type=int, default=None, help="sequences per GPU of the headline workload")
    ap.add_argument("--speakers", type=int, default=None, help="train workloads: number of speakers S")
    ap.add_argument("--workload", default="all", choices=["all", "train", "infer", "config3", "config4"],
                    help="all (default): configs[1] train step as the headline + configs[2]/[3] (and [4] at N=8) under 'configs'; "
                         "train / infer / config3 / config4: that workload alone")
    ap.add_argument("--all-configs", action="store_true", help="with --workload all: also run configs[4] when N != 8")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16x3", "bf16"], help="headline workload's arithmetic")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: one gradient all-reduce after backward instead of buckets during it")
    ap.add_argument("--exchange-dtype", default="fp32", choices=["fp32", "native"], help="N > 1: dtype of the gradient all-reduce")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_line(args)
    run_cuda(args)


if __name__ == "__main__":
    main()
