"""Benchmark of the Mix-StAGE generator hot path on B200 (contract: see the task statement).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU

Workload (BASELINE.json configs[1]): full GAN train step of JointLateClusterSoftStyle4_G +
pose discriminator (gan=1, L1Loss), batch 16 per GPU, 64-frame windows, 4 speakers,
num_clusters=8, fp64 master parameters and inputs as the reference's trainer keeps them
(trainer.py:138), synthetic N(0,1) inputs and seeded synthetic weights.  A "step" is one
training iteration: zero_grad, GAN.forward (generator step on even iterations, discriminator
step on odd ones -- the reference flips a fair coin, gan.py:105), backward, gradient
all-reduce when N>1, clip_grad_norm_(1) and Adam(1e-4) (trainer.py:1138-1146).
Metric: pose sequences (64-frame windows) per second, whole job.
"""
import argparse
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

METRIC = "pose sequences/sec (64-frame windows), GAN train step"
UNIT = "sequences/s"
MOD = ["audio/log_mel_400"]
T = 64


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel family from the committed ncu --set full capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        d = json.load(open(path))[workload]
        return d["traffic_bytes_per_launch"], d["source"]
    except Exception:
        return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
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
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# --------------------------------------------------------------------------------------------
def cpu_train_steps(B, S, steps, warmup, dtype=torch.float64):
    """Times the oracle's restatement of one reference train step (fwd + bwd + clip + Adam)."""
    import mixstage_oracle as O
    from oracle_cases import D_SEED, G_SEED, leafify
    torch.set_num_threads(os.cpu_count())
    spec = O.Spec(num_speakers=S)
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
        return float(sum(losses))

    for i in range(warmup):
        one("G" if i % 2 == 0 else "D")
    t0 = time.perf_counter()
    for i in range(steps):
        one("G" if i % 2 == 0 else "D")
    dt = time.perf_counter() - t0
    return dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, S = args.batch, args.speakers
    steps, warmup = args.steps, args.warmup
    dt = cpu_train_steps(B, S, steps, warmup)
    val = steps * B / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: GAN train step (alternating G/D), B=%d, T=64, S=%d, K=8, fp64" % (B, S)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": "%d train steps (G/D alternating) after %d warm-up, oracle port of the reference "
                                   "algorithm (torch CPU fp64, %d threads); the reference is Python and "
                                   "/root/reference does not exist on the GPU box" % (steps, warmup, os.cpu_count())},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------------
def conv_flops(name, desc):
    g = desc.groups
    if name == "ms_conv_fwd_f32" or name == "ms_conv_wgrad_f32":
        return 2.0 * desc.B * desc.Ho * desc.Wo * desc.Cout * (desc.Cin // g) * desc.kh * desc.kw
    if name == "ms_conv_dgrad_f32":
        return 2.0 * desc.B * desc.Ho * desc.Wo * desc.Cout * (desc.Cin // g) * desc.kh * desc.kw
    return 0.0


def run_cuda(args):
    import torch.distributed as dist
    import mixstage_b200 as M
    import mixstage_oracle as O
    from mixstage_b200 import _lib, ops, parallel
    from model_cases import build as build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    ops.set_precision(args.precision)

    B, S = args.batch, args.speakers
    spec = O.Spec(num_speakers=S)
    G, D, gan = build_model(spec, T, dev, torch.float64)
    gan.train()
    G.thresh.value, G.thresh.iters = 1.0, 1000            # past the curriculum: audio branch
    parallel.sync_host_rng(11212)
    # the repo's public train-step API: zero_grad + GAN.forward + backward + (all-reduce) + clip + Adam,
    # replayed from CUDA graphs (mixstage_b200/train_step.py)
    ts = M.TrainStep(gan, lr=1e-4, max_norm=1.0, use_graphs=not args.no_graphs, overlap_allreduce=args.overlap_allreduce)

    # per-rank synthetic shard (weak scaling: B sequences per GPU), pinned host copies for the e2e leg
    audio, pose, labels, style = O.synth_inputs(B, T, spec, seed=11212 + rank)
    host = [t.pin_memory() for t in (audio, labels, pose, style)]
    resident = [t.to(dev) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = 5 * 8

    # e2e leg: every step's five losses are read on the host, through pinned buffers and one step behind the launch, so the
    # host enqueues step i+1 (H2D of its batch into the static buffers, graph replay) while the GPU still runs step i
    loss_pin = [torch.empty(5, dtype=torch.float64).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    loss_log = []

    def step(i, batch, read_loss):
        kind = "G" if i % 2 == 0 else "D"
        fake, losses = ts.step(*batch, kind=kind)           # batch: host (pinned) or device tensors -> static buffers
        if read_loss:
            b = i & 1
            loss_pin[b].copy_(losses, non_blocking=True)    # D2H read of the step's five losses
            loss_ev[b].record()
            if i > 0:
                loss_ev[b ^ 1].synchronize()
                loss_log.append(float(loss_pin[b ^ 1][0]))  # the previous step's losses are on the host now
        return None

    def drain(nsteps):
        b = (nsteps - 1) & 1
        loss_ev[b].synchronize()
        loss_log.append(float(loss_pin[b][0]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(nsteps):
            step(i, host if e2e else resident, e2e)
        if e2e:
            drain(nsteps)                   # the last step's losses
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if e2e:
            ms = max(ms, wall * 1e3)        # the host read-back is part of the step
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(args.warmup, 4)):
        step(i, resident, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES + ts.launched
    ms = timed(args.steps, False)
    launches = _lib.LAUNCHES + ts.launched - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(args.steps, True)

    # ---- roofline of the dominant kernel family (implicit-GEMM convolutions): one instrumented
    # step pair with CUDA events around every conv launch on the launch stream
    records = []
    orig_call = ops.call

    def timed_call(name, *a):
        if name.startswith("ms_conv_") or name in ("ms_igemm_bf16", "ms_wgrad_bf16"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_call(name, *a)
            e1.record()
            if name.startswith("ms_conv_"):
                desc = [x for x in a if isinstance(x, _lib.ConvDesc)][0]
                fl = conv_flops(name, desc)
            else:
                fl = ops.last_gemm_flops
            records.append((name, fl, e0, e1))
        else:
            orig_call(name, *a)

    ops.call = timed_call
    graphs_on = ts.use_graphs
    ts.use_graphs = False                    # the instrumented pair runs the same body eagerly
    try:
        step(0, resident, False)
        step(1, resident, False)
    finally:
        ops.call = orig_call
        ts.use_graphs = graphs_on
    torch.cuda.synchronize()
    tc = [r for r in records if not r[0].startswith("ms_conv_")]
    dom = tc if tc else records                         # dominant kernel family: the tcgen05 implicit GEMMs when in use
    conv_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in dom)
    conv_fl = sum(f for _, f, _, _ in dom)
    all_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in records)
    pk, pk_kind = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    achieved_tf = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

    if rank == 0:
        cpu_steps = 4
        cpu_dt = cpu_train_steps(B, S, cpu_steps, 1)
        total = args.steps * B * world
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (split-bf16 operands, f32 accumulate)",
                                           "bf16": "bf16 (f32 accumulate)"}[args.precision], "data": "synthetic",
            "config": {"workload": "%s: GAN train step (G on even, D on odd iterations), B=%d per GPU, T=64, "
                                   "S=%d, K=8, gan=1, L1Loss, fp64 master params/inputs, precision=%s, %s" % (
                                       "configs[1]" if (B, S) == (16, 4) else "configs[3] per-GPU slice" if (B, S) == (128, 8) else "variant",
                                       B, S, args.precision, "CUDA-graph replay" if ts.use_graphs else "eager launches"),
                       "global_batch": B * world, "parallelism": "dp%d" % world,
                       "l2": "no explicit flush: per-step working set (fp64 params + packed fp32 weights + grads + "
                             "Adam state ~0.9 GB) exceeds the 126 MB L2"},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "mixstage_b200.TrainStep.step(pinned host batch); every step's losses copied to pinned host memory and "
                           "read one step behind the launch (%d reads)" % len(loss_log)},
            "gpu_launches": launches,
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clocks else None,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": ncu_traffic("train")[0],
                         "traffic_source": ncu_traffic("train")[1],
                         "kernel": ("igemm_tc_kernel + igemm_tc_persist_kernel + wgrad_tc_kernel (tcgen05 implicit GEMM: fwd+dgrad+wgrad, %d launches "
                                    "per G+D step pair; algorithmic FLOPs, split-bf16 issues 3x the MMAs)" % len(dom)) if tc else
                                   "conv_gemm_simt (fwd+dgrad+wgrad, %d launches per G+D step pair, fp32 CUDA cores)" % len(dom),
                         "gemm_ms_per_step_pair": conv_ms, "all_conv_ms_per_step_pair": all_ms,
                         "peak_source": "%s bf16 sustained" % pk_kind},
            "cpu_baseline": {"value": cpu_steps * B / cpu_dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d train steps (G/D alternating) B=%d after 1 warm-up, oracle port, torch CPU fp64" % (cpu_steps, B)},
        }
        print(json.dumps(line))
    if world > 1:
        _leave(dist)


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


# --------------------------------------------------------------------------------------------
# CUDA arm, BASELINE configs[2]: batched inference sweep (sample_all_styles over S speakers)
# --------------------------------------------------------------------------------------------
def run_infer(args):
    """One step = the reference's sampling inner loop for one batch (trainer.py:791-794 with update_kwargs :1367-1386):
    B windows pushed through G.forward once per target style ((style + shift) % S), eval mode, no_grad.  Sequences per
    step = B * S.  Sharded over ranks with no collective."""
    import torch.distributed as dist
    import mixstage_b200 as M
    import mixstage_oracle as O
    from mixstage_b200 import _lib, ops
    from model_cases import build as build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    prec = args.precision or "bf16"
    ops.set_precision(prec)
    B, S = args.batch or 1024, 4
    spec = O.Spec(num_speakers=S)
    G, D, gan = build_model(spec, T, dev, torch.float64)
    G.eval()
    G.thresh.value, G.thresh.iters = 1.0, 1000

    # two alternating per-rank batches: consecutive steps never see the same audio tensor, so the style-sweep cache
    # (encoder output reused across the S styles of ONE batch) cannot carry work from one step to the next
    hosts, residents = [], []
    for j in range(2):
        audio, pose, labels, style = O.synth_inputs(B, T, spec, seed=11212 + 2 * rank + j)
        h = [t.pin_memory() for t in (audio, labels, pose, style)]
        hosts.append(h)
        residents.append([t.to(dev) for t in h])
    h2d = sum(t.numel() * t.element_size() for t in hosts[0])
    out_host = torch.empty((S, B, T, spec.out_feats), dtype=torch.float64).pin_memory()
    d2h = out_host.numel() * out_host.element_size()

    def step(i, e2e):
        if e2e:
            audio, labels, pose, style = [t.to(dev, non_blocking=True) for t in hosts[i % 2]]
        else:
            audio, labels, pose, style = residents[i % 2]
        with torch.no_grad():
            for shift in range(S):
                st = (style + shift) % S
                out, _ = G([audio, labels], pose, input_modalities=MOD, style=st, sample_flag=1, description="test")
                if e2e:
                    out_host[shift].copy_(out, non_blocking=True)
        if e2e:
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(nsteps):
            step(i, e2e)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if e2e:
            ms = max(ms, wall * 1e3)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(args.warmup, 3)):
        step(i, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0, hits0 = _lib.LAUNCHES, G.encoder_cache_hits
    ms = timed(args.steps, False)
    launches = _lib.LAUNCHES - l0
    hits = G.encoder_cache_hits - hits0
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(args.steps, True)

    # roofline of the dominant kernel family: CUDA events around every tcgen05 launch of one instrumented step
    records = []
    orig_call = ops.call

    def timed_call(name, *a):
        if name in ("ms_igemm_bf16_fused", "ms_igemm_bf16"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_call(name, *a)
            e1.record()
            records.append((name, ops.last_gemm_flops, e0, e1))
        else:
            orig_call(name, *a)

    ops.call = timed_call
    try:
        step(args.steps, False)
    finally:
        ops.call = orig_call
    torch.cuda.synchronize()
    gemm_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in records)
    gemm_fl = sum(f for _, f, _, _ in records)
    pk, pk_kind = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    achieved_tf = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    if rank == 0:
        total = args.steps * B * S * world
        cpu_b = 16
        cpu_dt = cpu_infer_sweep(cpu_b, S, 2, 1)
        line = {
            "metric": "pose sequences/sec (64-frame windows), inference style sweep", "value": total / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (split-bf16 operands, f32 accumulate)", "bf16": "bf16 (f32 accumulate)"}[prec],
            "data": "synthetic",
            "config": {"workload": "configs[2]: inference sweep, B=%d windows per GPU x S=%d target styles per step "
                                   "(sample_all_styles), T=64, K=8, eval mode, fp64 parameters/inputs/outputs, precision=%s, "
                                   "encoder+UNet computed once per batch and reused across styles" % (B, S, prec),
                       "global_batch": B * world, "parallelism": "dp%d (no collective)" % world,
                       "l2": "two alternating input batches; per-step activations (>1 GB at B=1024) exceed the 126 MB L2"},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "G.forward per style on a pinned host batch + pose outputs copied back to pinned host memory"},
            "gpu_launches": launches, "encoder_cache_hits": hits,
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")} if clocks else None,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": ncu_traffic("infer")[0],
                         "traffic_source": ncu_traffic("infer")[1],
                         "kernel": "igemm_tc_persist_kernel / igemm_tc_pair_kernel, fused inference epilogue (%d launches per step; algorithmic FLOPs%s)" % (
                             len(records), ", split-bf16 issues 3x the MMAs" if prec == "bf16x3" else ""),
                         "gemm_ms_per_step": gemm_ms, "peak_source": "%s bf16 sustained" % pk_kind},
            "cpu_baseline": {"value": 2 * cpu_b * S / cpu_dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": "2 sweeps of B=%d x S=%d styles after 1 warm-up, oracle port (torch CPU fp64), encoder "
                                       "recomputed per style as the reference does" % (cpu_b, S)},
        }
        print(json.dumps(line))
    if world > 1:
        _leave(dist)


def cpu_infer_sweep(B, S, steps, warmup, dtype=torch.float64):
    """Times the oracle's restatement of the reference sampling loop: S full forwards per batch."""
    import mixstage_oracle as O
    from oracle_cases import G_SEED
    torch.set_num_threads(os.cpu_count())
    spec = O.Spec(num_speakers=S)
    sd = O.synth_state(O.g_state_shapes(spec), G_SEED, dtype)
    audio, pose, labels, style = O.synth_inputs(B, T, spec, dtype=dtype)

    def one():
        with torch.no_grad():
            for shift in range(S):
                O.g_forward(sd, spec, audio, labels, pose, (style + shift) % S, training=False, sample_flag=1, description="test")

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--batch", type=int, default=None, help="sequences per GPU (default 16 train, 1024 infer)")
    ap.add_argument("--speakers", type=int, default=4, help="train workload: number of speakers S (configs[1]: 4, configs[3]: 8)")
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train: BASELINE configs[1] GAN train step (default); infer: configs[2] style-sweep inference")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16x3", "bf16"], help="default bf16x3 train, bf16 infer")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="train, N > 1: exchange the generator's gradients segment by segment during backward (experimental)")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.workload == "infer" and args.impl == "cuda":
        return run_infer(args)
    if args.workload == "infer":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        B, S = args.batch or 16, 4
        dt = cpu_infer_sweep(B, S, args.steps, args.warmup)
        val = args.steps * B * S / dt
        print(json.dumps({
            "impl": "reference", "metric": "pose sequences/sec (64-frame windows), inference style sweep", "value": val,
            "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[2]: inference sweep, B=%d x S=%d styles per step, T=64, K=8, fp64" % (B, S)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d sweeps after %d warm-up, oracle port (torch CPU fp64, %d threads)" % (
                                 args.steps, args.warmup, os.cpu_count())},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    args.batch = args.batch or 16
    args.precision = args.precision or "bf16x3"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
