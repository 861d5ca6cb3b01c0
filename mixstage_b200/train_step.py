"""The hot loop body of the reference's trainer as ONE call, replayed from CUDA graphs.

Reference: `TrainerLateClusterStyleGAN` runs, per batch (src/model/trainer.py:590-675),
    zero_grad (:1104-1107) -> forward_pass = GAN.forward (:1158-1165, gan.py:86-164) -> calculate_loss = sum of the five
    internal losses (:1268-1285) -> optimize: backward, clip_grad_norm_(params, 1), G_optim.step() or D_optim.step()
    (:1138-1146; Adam lr 1e-4, :262-287).
`TrainStep.step(audio, labels, pose, style)` is that body.  At batch 16 the step is ~600 dependent microsecond-scale
kernels, so the host (Python, autograd bookkeeping, launches) -- not the GPU -- bounds the eager version.  Here the whole
body (packing of the updated weights, forward, backward, gradient all-reduce when data-parallel, fused clip + Adam over
flat buffers) is captured once per (step kind, curriculum branch) into a CUDA graph and replayed; the host only copies the
batch into static buffers, draws the same two coin flips the reference draws (gan.py:105, jlcss.py:127) and launches one
graph.

Parameters of G and of D are re-homed into one flat buffer each (the nn.Parameter objects become views, `state_dict` is
unchanged), with flat gradient / Adam-moment buffers beside them: the optimiser is two kernels (csrc/optimizer.cu) and the
data-parallel exchange one all-reduce per sub-network with no pack/unpack copies.  Gradients of parameters a step does not
touch stay zero, which reproduces the reference's torch-1.5 `zero_grad()` (zero-fill, not None) + Adam behaviour.
"""
from __future__ import annotations

import os as _os

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import MixStageError, call, dt_code, ptr, stream

_DP_DEBUG = _os.environ.get("MS_DP_DEBUG", "")     # timing experiments: "skip_acc", "skip_small" (results are then wrong), "serial_tail"
ALIGN = 4       # parameter offsets in elements: 16-byte aligned for fp32, 32-byte for fp64


class FlatState:
    """Flat parameter / gradient / Adam-moment buffers of one sub-network."""

    def __init__(self, module, skip=(), moment_dtype=None):
        """moment_dtype: storage type of Adam's m / v (None: the parameter dtype, as torch.optim.Adam keeps them).
        skip: name prefixes of parameters that no forward of `module` ever touches (parameter holders kept for
        state_dict compatibility).  Their gradient is identically zero, so Adam never moves them (m = v = 0 gives a zero
        update): leaving them out of the flat buffers changes nothing but the bytes zeroed, normed, all-reduced and
        stepped over."""
        skip = tuple(skip)
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad and not (skip and n.startswith(skip))]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        p0 = self.params[0]
        self.dtype, self.device = p0.dtype, p0.device
        if self.dtype not in (torch.float32, torch.float64):
            raise MixStageError("TrainStep supports fp32 / fp64 master parameters")
        off, self.offsets = 0, []
        for p in self.params:
            if p.dtype != self.dtype or p.device != self.device:
                raise ValueError("all parameters must share dtype/device")
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        # contiguous range of every top-level sub-module (parameters are registered module by module)
        self.segments = {}
        for i, n in enumerate(self.names):
            top = n.split(".")[0]
            beg = self.offsets[i]
            end = self.offsets[i + 1] if i + 1 < len(self.offsets) else self.numel
            b0, e0 = self.segments.get(top, (beg, beg))
            if e0 != beg and top in self.segments:
                raise ValueError("parameters of %s are not contiguous in registration order" % top)
            self.segments[top] = (b0, end)
        mk = lambda dt=self.dtype: torch.zeros(self.numel, dtype=dt, device=self.device)       # noqa: E731
        self.moment_dtype = moment_dtype or self.dtype
        if self.moment_dtype not in (torch.float32, self.dtype):
            raise MixStageError("Adam moments are stored in fp32 or in the parameter dtype")
        self.p, self.g, self.m, self.v = mk(), mk(), mk(self.moment_dtype), mk(self.moment_dtype)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.p[o:o + n].view(p.shape)
            p.grad = self.g[o:o + n].view(p.shape)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.sqnorm = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.x32 = None             # fp32 staging of the gradient exchange (allocated on the first multi-rank step)

    def zero_grad(self):
        self.g.zero_()
        for p, o in zip(self.params, self.offsets):      # re-attach if someone dropped .grad (zero_grad(set_to_none))
            if p.grad is None or p.grad.data_ptr() != self.g.data_ptr() + o * self.g.element_size():
                p.grad = self.g[o:o + p.numel()].view(p.shape)

    def reduce_range(self, beg, end, group=None, fp32=True, async_op=False):
        """Mean of g[beg:end] over the ranks, in place.  With fp64 master gradients and fp32=True the exchange itself is
        fp32 (half the NVLink bytes): one kernel scales by 1/world and narrows into a persistent staging buffer, the
        all-reduce (SUM) runs on the staging range, one kernel widens back.  Every rank performs the same element-wise
        operations, so replicas stay bit-identical.  Returns the async work handle (or None)."""
        if end <= beg:
            return None
        ws = dist.get_world_size(group)
        chunk = self.g[beg:end]
        n = end - beg
        st = stream()
        if fp32 and self.dtype == torch.float64:
            if self.x32 is None:
                self.x32 = torch.empty(self.numel, dtype=torch.float32, device=self.device)
            stage = self.x32[beg:end]
            call("ms_scale_cast", ptr(chunk), _lib.MS_F64, ptr(stage), _lib.MS_F32, n, 1.0 / ws, st)
            w = dist.all_reduce(stage, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
            if w is not None:
                # the widening below must see the REDUCED staging range.  NCCL runs the collective on its own stream: wait()
                # makes the launching (communication) stream wait for it -- a stream-level dependency, the host does not
                # block and the compute stream is not involved.  (gloo: a host wait; the widening runs on the host.)
                w.wait()
                w = None
            call("ms_scale_cast", ptr(stage), _lib.MS_F32, ptr(chunk), _lib.MS_F64, n, 1.0, st)
            return w
        call("ms_scale_cast", ptr(chunk), dt_code(self.dtype), ptr(chunk), dt_code(self.dtype), n, 1.0 / ws, st)
        return dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def allreduce_mean(self, group=None, fp32=True):
        if not (dist.is_available() and dist.is_initialized()):
            return
        if dist.get_world_size(group) == 1:
            return
        self.reduce_range(0, self.numel, group, fp32)

    def clip_adam(self, lr, lr_dev, betas, eps, max_norm):
        st = stream()
        dt = dt_code(self.dtype)
        call("ms_grad_sqnorm", ptr(self.g), dt, self.numel, ptr(self.sqnorm), ptr(self.step_count), st)
        call("ms_clip_adam_mixed", ptr(self.p), ptr(self.g), ptr(self.m), ptr(self.v), dt, dt_code(self.moment_dtype), self.numel,
             ptr(self.sqnorm), ptr(self.step_count), float(lr), float(betas[0]), float(betas[1]), float(eps), float(max_norm),
             ptr(lr_dev), st)


class TrainStep:
    """gan: mixstage_b200.GAN already on its device/dtype (`.to(device).double()` as the reference's trainer does,
    trainer.py:138).  Do not call `.to()/.double()` on the model afterwards: parameters are views of flat buffers."""

    # activation whose gradient has arrived -> generator sub-modules whose parameter gradients are final (see G._mark)
    READY = {"hc": ("logits", "decoder", "classify_cluster"), "unet_out": ("style_emb",), "encoder_out": ("unet",)}

    def __init__(self, gan, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, use_graphs=True, group=None,
                 input_modalities=("audio/log_mel_400",), description="train", overlap_allreduce=True,
                 exchange_dtype="fp32", rng_seed=None, check_agreement=False, comm_sms=16, side_sms=None,
                 moment_dtype="fp32"):
        """One TrainStep (and one CUDA device) per process: the scratch arena, the direct-gradient switch and the
        side stream are module-level state of mixstage_b200.ops, and the kernels' lazily set function attributes are
        per process.

        overlap_allreduce / exchange_dtype: data-parallel runs exchange the generator's gradients in four buckets
        launched from backward hooks on a communication stream (decoder + logits + classifier first, ... audio encoder
        last) while backward is still running, as fp32 ("fp32", default: half the bytes of the fp64 master gradients)
        or in the master dtype ("native").
        moment_dtype: storage of Adam's m / v: "fp32" (default: 40 instead of 56 bytes per fp64 parameter and step; the
        update itself is computed in fp64 from them, relative rounding 6e-8 per step) or "native" (the parameter dtype, as
        torch.optim.Adam keeps them in the reference, trainer.py:262-287).
        comm_sms: SMs left free beside the chain launches for the NCCL kernels of the overlapped exchange (multi-rank only).
        side_sms: SMs left free beside the chain launches for the side stream's weight-gradient launches (default: the
        MS_SIDE_SMS environment variable, else 0).
        rng_seed: seed of the generator the D/G coin and the curriculum draw come from.  None = the process-global CPU
        generator in a single-process run (the reference's own RNG consumption, gan.py:105 / jlcss.py:127) and a
        dedicated generator seeded with the reference's seed 11212 on every rank of a data-parallel run: ranks then
        choose the same (kind, branch) graphs whatever else consumes the global generator.
        check_agreement: debug aid, all-gathers (kind, branch) every step and raises on disagreement (host sync)."""
        self.gan, self.G, self.D = gan, gan.G, gan.D
        if moment_dtype not in ("fp32", "native"):
            raise MixStageError("moment_dtype must be 'fp32' or 'native'")
        mdt = torch.float32 if moment_dtype == "fp32" else None
        self.fG = FlatState(self.G, skip=getattr(self.G, "UNUSED_PARAMETER_PREFIXES", ()), moment_dtype=mdt)
        self.fD = FlatState(self.D, moment_dtype=mdt)
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        dev = self.fG.device
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float64, device=dev)
        self.use_graphs = bool(use_graphs) and dev.type == "cuda"
        self.group = group
        self.mod = list(input_modalities)
        self.description = description
        self.graphs = {}
        self.kernels_per_graph = {}
        self.launched = 0            # C-ABI kernel launches executed through graph replays
        self.static = None
        self.losses = None
        self.fake = None
        self.last_kind = None
        self.replays = 0
        self.eager_steps = 0         # steps whose batch shape differed from the captured one (run eagerly)
        self.warmup_iters = 2
        self.side = ops.SideWork(dev) if dev.type == "cuda" else None      # weight-gradient GEMMs beside the dgrad chain
        self._tables = {}            # (kind, tag) -> (signature, device table, n): entries of the batched weight re-tiling
        self.defer_dgrad_pack = _os.environ.get("MS_DEFER_DGRAD_PACK", "1") != "0"
        for pw in self._packed_of(self.G) + self._packed_of(self.D):
            pw._src.clear()          # recipes recorded before the parameters moved into the flat buffers are stale
        self._pver = None            # flat-parameter versions seen by the last refresh
        # opt-in: all-reduce the generator's gradients segment by segment while backward is still running (NOT yet
        # validated on multi-GPU hardware; numerics covered by tests/test_parallel_cpu.py with gloo)
        self.overlap = bool(overlap_allreduce)
        if exchange_dtype not in ("fp32", "native"):
            raise MixStageError("exchange_dtype must be 'fp32' or 'native'")
        self.exchange_fp32 = exchange_dtype == "fp32"
        self.comm = torch.cuda.Stream(device=dev) if (self.overlap and dev.type == "cuda") else None
        # the chain launches are persistent and would hold every SM: leave a few to the NCCL kernels of the overlapped
        # exchange (set NCCL_MAX_CTAS accordingly before the process group is created; bench.py does)
        self.comm_sms = int(_os.environ.get("MS_COMM_SMS", comm_sms))
        if side_sms is None:
            side_sms = int(_os.environ.get("MS_SIDE_SMS", "0"))
        free = max(self.comm_sms if (self.comm is not None and self._world() > 1) else 0, int(side_sms))
        if dev.type == "cuda" and free > 0:
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            call("ms_set_chain_sm_budget", max(16, n_sm - free))
        self._cap_stream = None
        self._reduced, self._works = [], []
        self._flushed = set()
        # tensor-core modes: the conv weight gradients are exchanged as their fp32 accumulators (see _exchange_acc)
        self._acc_plan = {}          # (kind, curriculum branch) -> [(stage, accumulator keys), ...] learned in the first body
        self._acc_marks = None       # learning pass: [(stage, accumulators touched so far)]
        self._small_idx = {}         # kind -> (entries seen, indices of the flat gradient elements outside the accumulators)
        if rng_seed is None and self._world() > 1:
            rng_seed = 11212
        self.rng = None if rng_seed is None else torch.Generator().manual_seed(int(rng_seed))
        self.check_agreement = bool(check_agreement)
        # [lambda_D, lambda_gan] in device memory: the GAN forward multiplies by these (captured graphs read the current
        # values at replay); the injectable scheduler (gan.lambda_scheduler) is stepped once per real iteration below
        self.lambda_dev = torch.tensor([float(gan.lambda_D), float(gan.lambda_gan)], dtype=torch.float64, device=dev)
        self._lam_host = [float(gan.lambda_D), float(gan.lambda_gan)]
        # persistent weight-gradient accumulators, one set per kind of step (a G-step also produces the discriminator's
        # -- unused -- weight gradients, a D-step only the discriminator's), and the device tables of their conversion
        self.wacc = {"G": ops.WgradAccum(), "D": ops.WgradAccum()}
        self._wtables = {}
        self._old_tables = []        # superseded packed-weight tables stay alive as long as this object does

    # ------------------------------------------------------------------ host-side decisions
    def set_lr(self, lr):
        """ExponentialLR etc. (trainer.py:311-313): the captured graphs read the rate from device memory."""
        self.lr = lr
        self.lr_dev.fill_(float(lr))

    def _decide(self, kind):
        """Same draws, in the same order, as the reference: D/G coin (gan.py:105), then the curriculum draw inside
        G.forward (jlcss.py:127; consumed in eval mode too)."""
        gan, G = self.gan, self.G
        coin = torch.rand(1, generator=self.rng).item()
        if kind is None:
            kind = "D" if coin < gan.D_prob else "G"
        u = torch.rand(1, generator=self.rng).item()
        if kind == "G":
            use_pose = u > G.thresh.step(True)
        else:
            G.thresh.step(False)
            use_pose = False
        if self.check_agreement and self._world() > 1:
            mine = (kind, bool(use_pose))
            seen = [None] * self._world()
            dist.all_gather_object(seen, mine, group=self.group)
            if any(o != mine for o in seen):
                raise MixStageError("TrainStep: ranks disagree on (step kind, curriculum branch): %s" % (seen,))
        # the injectable lambda schedule advances once per real training iteration (reference: gan.py:103)
        lam = [float(v) for v in gan.lambda_scheduler.step()]
        gan.lambda_D, gan.lambda_gan = lam
        if lam != self._lam_host:
            self._lam_host = lam
            self.lambda_dev.copy_(torch.tensor(lam, dtype=self.lambda_dev.dtype))
        return kind, bool(use_pose)

    # ------------------------------------------------------------------ the step body (eager, and what gets captured)
    def _body(self, kind, use_pose, audio, labels, pose, style):
        gan, G = self.gan, self.G
        wacc = self.wacc[kind]
        # Work the FORWARD does not depend on runs beside it on the side stream: zero-filling the flat gradient buffers and
        # the weight-gradient accumulators (first touched by backward) and, in a graphed generator step, the re-tiling of the
        # input-gradient copies of the weights (only this step's backward reads them).  Joined right before backward.
        pre = self.side is not None
        if pre:
            with self.side.fork():
                self.fG.zero_grad()
                self.fD.zero_grad()
                wacc.zero()
                if self.use_graphs and kind == "G" and self.defer_dgrad_pack:
                    self._refresh("G", "dgrad")
        else:
            self.fG.zero_grad()
            self.fD.zero_grad()
            wacc.zero()
        old_force, old_lam = gan.force_step, gan.lambda_dev
        gan.force_step = kind
        gan.lambda_dev = self.lambda_dev
        G.force_branch = "pose" if use_pose else "audio"
        # kernels accumulate parameter gradients straight into the flat buffers and draw their fp64 accumulators from
        # one arena cleared by a single memset (ops.py)
        ops.arena.begin(self.fG.device)
        ops.DIRECT_GRADS = True
        ops.CAST_CACHE = {}
        ops.SIDE = self.side
        ops.WACC = wacc
        ops.RAW_LOSSES = True
        overlap = self.overlap and kind == "G" and self._world() > 1
        self._reduced, self._works = [], []
        self._flushed = set()
        self._step_key = (kind, bool(use_pose))
        self._acc_marks = []
        self._acc_mode = None        # decided at the first hook / at the end: accumulators exist <=> tensor-core mode
        G.grad_ready_hook = self._on_ready if overlap else None
        try:
            fake, losses, _ = gan([audio, labels], pose, input_modalities=self.mod, style=style, sample_flag=0,
                                  description=self.description, desc=self.description)
            # one launch scales the terms (lambda_id, the device-resident GAN lambdas), widens them for the report and sums
            # them; its backward hands every loss its fp32 seed (instead of ~20 scalar casts, multiplies and adds)
            loss, report = ops.combine_losses(losses, self.lambda_dev)
            if pre:
                torch.cuda.current_stream().wait_stream(self.side.stream)
            loss.backward()
        finally:
            G.force_branch = None
            G.grad_ready_hook = None
            gan.force_step, gan.lambda_dev = old_force, old_lam       # a direct gan(...) call draws its own coin again
            ops.DIRECT_GRADS = False
            ops.RAW_LOSSES = False
            ops.CAST_CACHE = None
            ops.SIDE = None
            ops.WACC = None
            ops.arena.end()
            if self.side is not None:
                self.side.join()
        f = self.fG if kind == "G" else self.fD
        if self._world() > 1 and self.wacc[kind].touched:
            # tensor-core mode: exchange the fp32 accumulators (+ the few gradients that live only in the flat buffer), THEN
            # convert everything into the flat buffer with the whole machine
            # The tail after backward: the accumulators already reduced from hooks are converted while the last bucket is
            # still travelling, that bucket's accumulators right after it, and the gathered small gradients (disjoint
            # elements of the flat buffer) travel beside both conversion launches.
            late, late_works = self._exchange_acc_finish(kind, overlap)
            if late_works:
                self._flush_wgrads(kind, None, ("early",) + self._step_key, exclude=late)
                for w in late_works:
                    w.wait()
            self._flush_wgrads(kind, None, ("late",) + self._step_key if late_works else "all")
            if self.comm is not None:
                torch.cuda.current_stream().wait_stream(self.comm)
        elif overlap:
            self._finish_overlapped()
            self._flush_wgrads(kind, None, "tail")      # nothing left unless a gap-free layout skipped the "rest" bucket
        else:
            self._flush_wgrads(kind)
            f.allreduce_mean(self.group, self.exchange_fp32)
        f.clip_adam(self.lr, self.lr_dev, self.betas, self.eps, self.max_norm)
        if self.use_graphs:
            # keep every packed copy (bf16 re-tilings, fp32 biases, folded eval BatchNorm) of the sub-network that just
            # stepped in sync with its parameters, so that no forward has to re-pack anything
            self._refresh(kind, "fwd" if (kind == "G" and self.defer_dgrad_pack) else "all")
        return fake.detach(), report if report.dtype == fake.dtype else report.to(fake.dtype)

    # ------------------------------------------------------------------ overlapped gradient exchange (opt-in)
    def _world(self):
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def _reduce_range(self, beg, end, tops=None, tag=None):
        """Mean of the generator's flat gradient [beg, end) over the ranks on the communication stream.  The conv weight
        gradients of the range still sit in their fp32 accumulators (the weight-gradient launches add into them on the side
        stream): they are converted into the flat buffer FIRST, on the communication stream, behind both other streams."""
        if end <= beg:
            return
        if self.comm is None:                       # CPU (gloo): no streams
            self._flush_wgrads("G", tops, tag)
            self.fG.reduce_range(beg, end, self.group, self.exchange_fp32)
            return
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)                  # BatchNorm / bias gradients are written on the compute stream
        if self.side is not None:
            self.comm.wait_stream(self.side.stream) # weight gradients on the side stream
        with torch.cuda.stream(self.comm):
            self._flush_wgrads("G", tops, tag)
            w = self.fG.reduce_range(beg, end, self.group, self.exchange_fp32, async_op=True)
            if w is not None:
                self._works.append(w)

    # ------------------------------------------------------------------ exchange through the weight-gradient accumulators
    # In the tensor-core modes 99 % of the gradient bytes sit in the fp32 accumulators of the chain weight-gradient launches
    # until the conversion launch at the end of backward.  They are all-reduced THERE (mean, NCCL AVG): no fp64 -> fp32
    # staging pass, no widening pass, no per-bucket conversion squeezed onto the SMs the chain launches leave free.  The
    # buckets are "the accumulators touched since the previous hook", learned in the first body of every (kind, branch).
    def _reduce_tensors(self, tensors):
        if "skip_acc" in _DP_DEBUG:                # timing experiments only: replicas diverge
            return
        ws = self._world()
        nccl = self.fG.device.type == "cuda"
        for t in tensors:
            if nccl:
                self._works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.mul_(1.0 / ws)

    def _acc_bucket(self, keys):
        """All-reduce the accumulators `keys` on the communication stream, behind the side stream's weight-gradient launches."""
        wacc = self.wacc[self._step_key[0]]
        tensors = wacc.ranges(keys)
        if not tensors:
            return
        if self.comm is None:
            self._reduce_tensors(tensors)
            return
        if self.side is not None:
            self.comm.wait_stream(self.side.stream)
        with torch.cuda.stream(self.comm):
            self._reduce_tensors(tensors)

    def _small_indices(self, kind):
        """Indices of the flat gradient elements that no accumulator feeds (BatchNorm affine parameters, biases, the style
        embedding, the few convolutions on the CUDA-core kernels): ~1 % of the buffer, exchanged as one gathered vector."""
        f = self.fG if kind == "G" else self.fD
        ents = self.wacc[kind].entries
        cur = self._small_idx.get(kind)
        if cur is not None and cur[0] == len(ents):
            return cur[1]
        base, isz = f.g.data_ptr(), f.g.element_size()
        mask = torch.ones(f.numel, dtype=torch.bool)
        for e in ents.values():
            off = (e[1] - base) // isz
            if 0 <= off < f.numel:
                mask[off:off + e[3] * e[4] * e[5]] = False        # Cout * Cin_g * taps
        idx = mask.nonzero().flatten().to(f.device)
        if torch.cuda.is_available() and f.device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            raise MixStageError("internal: exchange index table changed during graph capture")
        if cur is not None:
            self._old_tables.append(cur)
        self._small_idx[kind] = (len(ents), idx)
        return idx

    def _exchange_acc_finish(self, kind, overlap):
        """End of backward: the buckets not sent from hooks (all of them when not overlapping or still learning), the
        gathered small gradients, then join.  Runs after the side stream has been joined."""
        wacc = self.wacc[kind]
        f = self.fG if kind == "G" else self.fD
        plan = self._acc_plan.get(self._step_key) if overlap else None
        sent = set()
        if plan is not None and self._acc_mode == "plan":
            for _, keys in plan:
                sent.update(keys)
        rest = [k for k in wacc.touched if k not in sent]
        if overlap and self._acc_mode != "plan":
            # learning pass: remember which accumulators every hook could have sent
            marks, prev, plan = self._acc_marks, 0, []
            for stage, m in marks:
                plan.append((stage, tuple(wacc.touched[prev:m])))
                prev = m
            self._acc_plan[self._step_key] = [p_ for p_ in plan if p_[1]]
        idx = self._small_indices(kind)
        ws = self._world()
        ctxm = torch.cuda.stream(self.comm) if self.comm is not None else None
        if self.comm is not None:
            self.comm.wait_stream(torch.cuda.current_stream())
        if ctxm is not None:
            ctxm.__enter__()
        try:
            acc_works, self._works = self._works, []
            self._reduce_tensors(wacc.ranges(rest))
            late_works, self._works = self._works, []
            if idx.numel() and "skip_small" not in _DP_DEBUG:
                small = f.g.index_select(0, idx)
                if self.exchange_fp32 and small.dtype == torch.float64:
                    small = small.float()
                small.mul_(1.0 / ws)
                w = dist.all_reduce(small, op=dist.ReduceOp.SUM, group=self.group, async_op=self.comm is not None)
                if w is not None:
                    w.wait()
                f.g.index_copy_(0, idx, small.to(f.g.dtype))
        finally:
            if ctxm is not None:
                ctxm.__exit__(None, None, None)
        # the compute stream continues as soon as the buckets sent from hooks have arrived; the caller waits for the last
        # bucket between its two conversion launches and joins the communication stream (small gradients) after them
        for w in acc_works:
            w.wait()
        if self.comm is not None and "serial_tail" in _DP_DEBUG:
            for w in late_works:
                w.wait()
            torch.cuda.current_stream().wait_stream(self.comm)
            return set(), []
        if not acc_works or not late_works:
            # nothing was sent early (learning pass, no overlap) or nothing is left: one conversion launch
            for w in late_works:
                w.wait()
            return set(), []
        return {wacc.slots[k].data_ptr() for k in rest}, late_works

    def _on_ready(self, stage):
        wacc = self.wacc[self._step_key[0]]
        if wacc.touched or self._acc_mode is not None:
            # tensor-core mode
            plan = self._acc_plan.get(self._step_key)
            if self._acc_mode is None:
                self._acc_mode = "plan" if plan is not None else "learn"
            if self._acc_mode == "learn":
                self._acc_marks.append((stage, len(wacc.touched)))
            else:
                for st_, keys in plan:
                    if st_ == stage:
                        self._acc_bucket(keys)
            return
        for top in self.READY.get(stage, ()):
            seg = self.fG.segments.get(top)
            if seg is not None and seg not in self._reduced:
                self._reduced.append(seg)
                self._reduce_range(*seg, tops=(top,), tag=top)

    def _finish_overlapped(self):
        """Everything not exchanged from a hook (audio / pose encoder, pose-style encoder, gaps), then join."""
        done = sorted(self._reduced)
        pos = 0
        first = True
        for beg, end in done + [(self.fG.numel, self.fG.numel)]:
            if beg > pos:
                # the remaining accumulators (all of them, whichever gap they fall in) go out with the first gap
                self._reduce_range(pos, beg, tops=None, tag="rest") if first else self._reduce_range(pos, beg, tops=(), tag=None)
                first = False
            pos = max(pos, end)
        for w in self._works:
            w.wait()
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
        self._reduced, self._works = [], []

    # ------------------------------------------------------------------ packed copies follow the parameters
    @staticmethod
    def _packed_of(module):
        from .layers import ConvNormRelu, PlainConv
        out = []
        for m in module.modules():
            if isinstance(m, ConvNormRelu):
                out.append(m._packed)
            for v in vars(m).values():
                if isinstance(v, PlainConv):
                    out.append(v.packed)
                elif isinstance(v, ops.MixedLogits):
                    out.append(v.packed)
            for cfg_pw in getattr(m, "_alt", {}).values():       # AudioEncoder: the pruned inference form of its last block
                out.append(cfg_pw[1])
        return out

    def _refresh(self, kind, which="all"):
        """Re-derive the packed copies of sub-network `kind` from its parameters.  which: "all", or for the generator of a
        graphed run "fwd" (everything but the input-gradient tilings: what the NEXT forward of either step kind reads) /
        "dgrad" (the input-gradient tilings: only a generator step's backward reads them, so a generator step re-tiles them
        at its START on the side stream, beside its forward, instead of at the end of the previous one)."""
        mod = self.G if kind == "G" else self.D
        entries = []
        if which == "dgrad":
            cur = self._tables.get((kind, "dgrad"))
            if cur is not None:
                call("ms_pack_igemm_weight_multi", ptr(cur[1]), cur[2], 0, stream())
            return
        for m in mod.modules():
            for v in vars(m).values():
                if isinstance(v, ops.MixedLogits):
                    v.refresh_dense()            # dense (P, K*256) copy of the grouped logits weight, before its re-tiling
        for pw in self._packed_of(mod):
            pw.refresh(entries)
        if not entries:
            return
        groups = [("all", entries)]
        if kind == "G" and self.defer_dgrad_pack:
            groups = [("fwd", [e for e in entries if e.mode == 0]), ("dgrad", [e for e in entries if e.mode != 0])]
        for tag, ents in groups:
            self._refresh_table(kind, tag, ents, launch=(which == "all" or tag == which or tag == "all"))

    def _refresh_table(self, kind, tag, entries, launch):
        if not entries:
            return
        sig = tuple((e.w, e.wp, e.wp_lo, e.mode, e.num_classes, e.class_n, e.ntaps, e.kpad) for e in entries)
        cur = self._tables.get((kind, tag))
        if cur is None or cur[0] != sig:
            if torch.cuda.is_current_stream_capturing():
                raise MixStageError("internal: packed-weight table changed during graph capture")
            if cur is not None:
                # graphs captured so far launch the batched re-tiling with the OLD table (pointer and entry count): keep
                # that table alive, and drop the graphs -- they would skip the new entries; the next step re-captures
                self._old_tables.append(cur)
                self.graphs.clear()
                self.kernels_per_graph.clear()
            import ctypes
            arr = (_lib.PackEntry * len(entries))(*entries)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            cur = (sig, host.to(self.fG.device), len(entries))
            self._tables[(kind, tag)] = cur
        if launch:
            call("ms_pack_igemm_weight_multi", ptr(cur[1]), cur[2], 0, stream())

    def _flush_wgrads(self, kind, tops=None, tag="all", exclude=None):
        """Weight-gradient accumulators of this step -> the flat gradient buffers, one launch.  tops: only the accumulators
        of these top-level sub-modules (the overlapped exchange converts a bucket's accumulators right before it reduces
        the bucket); None: everything not converted yet in this step.  exclude: accumulator addresses to leave for a later
        call (their all-reduce is still in flight)."""
        f = self.fG if kind == "G" else self.fD
        base, isz = f.g.data_ptr(), f.g.element_size()
        ents = []
        for key, e in self.wacc[kind].entries.items():
            if key in self._flushed:
                continue
            if exclude is not None and key[0] in exclude:
                continue
            if tops is not None:
                off = (e[1] - base) // isz
                if not any(f.segments[t][0] <= off < f.segments[t][1] for t in tops if t in f.segments):
                    continue
            ents.append(e)
            self._flushed.add(key)
        if not ents:
            return
        sig = tuple(ents)
        cur = self._wtables.get((kind, tag))
        if cur is None or cur[0] != sig:
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                raise MixStageError("internal: weight-gradient table changed during graph capture")
            if cur is not None:
                self._old_tables.append(cur)
                self.graphs.clear()
                self.kernels_per_graph.clear()
            arr = (_lib.WgradEntry * len(ents))()
            for i, e in enumerate(ents):
                (arr[i].acc, arr[i].dw, arr[i].pdt, arr[i].Cout, arr[i].Cin_g, arr[i].taps, arr[i].kpad, arr[i].accumulate) = e
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            cur = (sig, host.to(self.fG.device), len(ents))
            self._wtables[(kind, tag)] = cur
        call("ms_unpack_wgrad_multi", ptr(cur[1]), cur[2], 0, stream())

    def refresh(self):
        """Call after changing parameters or BatchNorm buffers from outside (load_state_dict, manual edits)."""
        self._refresh("G")
        self._refresh("D")
        self._pver = (self.fG.p._version, self.fD.p._version)

    def _capture(self, key, batch):
        kind, use_pose = key
        dev = self.fG.device
        if self.static is None:
            self.static = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in batch]
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)
        # eager warm-up on a side stream (allocator + lazy kernel attributes), with parameters / moments / BN buffers
        # restored afterwards so that capture does not advance the training state
        snap = self._snapshot()
        # warm-up and capture share ONE stream for the life of the TrainStep: autograd's AccumulateGrad nodes (the few
        # parameters whose gradients are not written into the flat buffers by the kernels) remember the stream they were
        # created on, and a different capture stream would make every captured backward synchronise with it
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(device=dev)
        side = self._cap_stream
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            if not self.graphs:
                # before the FIRST capture run every (kind, branch) body once: each one registers the packed copies it
                # reads (train-mode tilings, the eval-mode generator's folded BatchNorm / dense logits of a D-step, the
                # pose encoder of the curriculum branch), so that the refresh captured at the end of every graph covers
                # ALL copies of the sub-network it updates -- not only those its own body happened to touch
                for k2, p2 in (("G", False), ("D", False), ("G", True)):
                    if (k2, p2) != (kind, use_pose):
                        self._body(k2, p2, *self.static)
            for _ in range(self.warmup_iters):
                self._body(kind, use_pose, *self.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        # the warm-up bodies ended with _refresh(): every packed copy is valid and stays at its address, so the captured
        # forward launches no re-packing; the captured step ends with the same refresh for the sub-network it updates
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(g, stream=side):
            fake, losses = self._body(kind, use_pose, *self.static)
        n_launch = _lib.LAUNCHES - l0                         # C-ABI launches recorded in this graph
        torch.cuda.synchronize(dev)
        self._restore(snap)
        self.refresh()               # the restored parameters' packed copies (may drop OTHER graphs: see _refresh)
        self.graphs[key] = (g, fake, losses)
        self.kernels_per_graph[key] = n_launch

    def _snapshot(self):
        bufs = [b for m in (self.G, self.D) for b in m.buffers()]
        st = [self.fG, self.fD]
        return ([b.clone() for b in bufs], bufs,
                [(f.p.clone(), f.m.clone(), f.v.clone(), f.step_count.clone()) for f in st], st)

    def _restore(self, snap):
        saved, bufs, fs, st = snap
        for b, s in zip(bufs, saved):
            b.copy_(s)
        for f, (p, m, v, c) in zip(st, fs):
            f.p.copy_(p)
            f.m.copy_(m)
            f.v.copy_(v)
            f.step_count.copy_(c)
        ops.bump_weight_epoch()

    def step(self, audio, labels, pose, style, kind=None):
        """One training iteration.  Returns (fake_pose, losses (5,)) as device tensors that stay valid until the next
        call: [pose L1, G_gan, cluster CE, id_in, id_out] for a G-step, [real_D, fake_D, ...] for a D-step (gan.py)."""
        kind, use_pose = self._decide(kind)
        self.last_kind = kind
        self.gan.train()
        batch = (audio, labels, pose, style)
        if not self.use_graphs:
            dev = self.fG.device
            batch = [t.to(dev, non_blocking=True) for t in batch]
            self.fake, self.losses = self._body(kind, use_pose, *batch)
            ops.bump_weight_epoch()
            return self.fake, self.losses
        key = (kind, use_pose)
        if self.static is not None and any(s.shape != t.shape or s.dtype != t.dtype for s, t in zip(self.static, batch)):
            # e.g. the trailing partial batch of a DataLoader without drop_last: the same body, launched eagerly
            dev = self.fG.device
            self.fake, self.losses = self._body(kind, use_pose, *[t.to(dev, non_blocking=True) for t in batch])
            ops.bump_weight_epoch()
            self._pver = (self.fG.p._version, self.fD.p._version)
            self.eager_steps += 1
            return self.fake, self.losses
        if self._pver != (self.fG.p._version, self.fD.p._version):
            self.refresh()           # someone wrote the parameters between steps (load_state_dict, ...); may drop graphs
        if key not in self.graphs:
            self._capture(key, batch)
        for s, t in zip(self.static, batch):
            s.copy_(t, non_blocking=True)
        g, self.fake, self.losses = self.graphs[key]
        g.replay()
        self.replays += 1
        self.launched += self.kernels_per_graph.get(key, 0)
        ops.bump_weight_epoch()
        return self.fake, self.losses
