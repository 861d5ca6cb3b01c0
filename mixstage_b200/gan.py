"""GAN wrapper, mirror of reference src/model/gan.py (GAN.forward :86-164) with the velocity,
discriminator and L1 reductions running in the CUDA kernels.  Same constructor and
``forward(x_audio, y_pose, **kwargs) -> (fake_pose, internal_losses, args)`` contract, same
host RNG consumption (one ``torch.rand(1)`` per training forward, gan.py:105)."""
import torch
import torch.nn as nn

from . import ops


class LambdaScheduler:
    """Stand-in for pycasper.torchUtils.LambdaScheduler (un-vendored and un-pinned upstream, so parity is unpinned here;
    semantics defined in oracle/ref_loader.py: constant lambdas).  The schedule is an INJECTABLE object: pass
    ``GAN(..., lambda_scheduler=obj)`` (or assign ``gan.lambda_scheduler``) with any object whose ``step()`` returns
    ``[lambda_D, lambda_gan]`` for the coming training forward (reference call site gan.py:30-33,103) -- e.g. pycasper's
    own class when it is installed next to the reference."""

    def __init__(self, lambdas, **kwargs):
        self.lambdas = list(lambdas)
        self.kwargs = dict(kwargs)

    def step(self):
        return list(self.lambdas)


class RampLambdaScheduler(LambdaScheduler):
    """A schedule in the shape the reference's arguments describe (``kind='incremental', max_interval=300,
    max_lambda=2``): every ``max_interval`` training forwards each lambda grows by its initial value until it reaches
    ``max_lambda``.  NOT pinned against pycasper (source absent); provided so that a varying schedule can be exercised
    through the device-resident lambdas of TrainStep."""

    def __init__(self, lambdas, max_interval=300, max_lambda=2, **kwargs):
        super().__init__(lambdas, max_interval=max_interval, max_lambda=max_lambda, **kwargs)
        self.base = list(lambdas)
        self.max_interval, self.max_lambda = int(max_interval), float(max_lambda)
        self.calls = 0

    def step(self):
        k = self.calls // max(1, self.max_interval)
        self.calls += 1
        return [min(self.max_lambda, b * (1 + k)) for b in self.base]


class GAN(nn.Module):
    def __init__(self, G, D, dg_iter_ratio=1, lambda_D=1, lambda_gan=1, lr=0.0001, criterion='MSELoss', optim='Adam',
                 joint=False, update_D_prob_flag=True, no_grad=True, lambda_scheduler=None, **kwargs):
        super().__init__()
        self.G = G
        self.D = D
        self.D_prob = dg_iter_ratio / (dg_iter_ratio + 1)
        self.lambda_D = lambda_D
        self.lambda_gan = lambda_gan
        self.lambda_scheduler = lambda_scheduler if lambda_scheduler is not None else LambdaScheduler(
            [self.lambda_D, self.lambda_gan], kind='incremental', max_interval=300, max_lambda=2)
        # TrainStep keeps [lambda_D, lambda_gan] in device memory (captured CUDA graphs read them there) and steps the
        # scheduler itself, once per real iteration; None = this forward steps the scheduler and multiplies host floats
        self.lambda_dev = None
        self.G_flag = True
        self.lr = lr
        if criterion != 'L1Loss':
            raise NotImplementedError("mixstage_b200.GAN accelerates criterion='L1Loss' (the reference jobs' -loss)")
        if joint:
            raise NotImplementedError("mixstage_b200.GAN: joint=True is outside the accelerated path")
        self.joint = joint
        self.input_modalities = kwargs['input_modalities']
        self.update_D_prob_flag = update_D_prob_flag
        self.no_grad = no_grad
        self.force_step = None        # 'G' / 'D' overrides the coin flip (tests, bench)

    def get_velocity(self, x, x_audio=None):
        return ops.cast(ops.velocity(ops.cast(x, torch.float32).contiguous()), x.dtype)

    def estimate_weights(self, x_audio, y_pose, **kwargs):
        return torch.ones(y_pose.shape[0], device=y_pose.device), None

    @staticmethod
    def _l1(a, b=None, const=0.0, dtype=None, lam=None, lam_host=None, lam_dev=None):
        """mean |a - b| (or |a - const|) as a loss entry, optionally times lambda number `lam` (0: lambda_D, 1: lambda_gan):
        the host value lam_host, or the device-resident lam_dev[lam] under TrainStep."""
        a32 = ops.f32_of(a).contiguous()
        b32 = None if b is None else ops.f32_of(b).contiguous()
        l = ops.l1_mean(a32, b32, const)
        if lam is None:
            return ops.loss_term(l, dtype or a.dtype)
        if lam_dev is not None:
            return ops.loss_term(l, dtype or a.dtype, lam_dev=lam_dev, lam=lam)
        return ops.loss_term(l, dtype or a.dtype, weight=lam_host)

    def _score(self, pose):
        """D(velocity(pose)) without leaving fp32."""
        v = ops.velocity(ops.f32_of(pose).contiguous())
        s, _ = self.D(v)
        return s

    def forward(self, x_audio, y_pose, **kwargs):
        internal_losses = []
        W, _ = self.estimate_weights(x_audio, y_pose, **kwargs)
        dt = y_pose.dtype
        if 'input_modalities' not in kwargs:
            kwargs['input_modalities'] = self.input_modalities
        if self.training:
            ld = None
            lam_D = lam_gan = None
            if self.lambda_dev is None:
                self.lambda_D, self.lambda_gan = self.lambda_scheduler.step()
                lam_D, lam_gan = self.lambda_D, self.lambda_gan
            elif ops.RAW_LOSSES:
                ld = self.lambda_dev                   # fp64, read by the loss-combining kernel at replay
            else:
                ld = self.lambda_dev if self.lambda_dev.dtype == dt else self.lambda_dev.to(dt)     # views: read at replay
            if self.force_step is None:
                d_step = torch.rand(1).item() < self.D_prob          # gan.py:105
            else:
                d_step = self.force_step == 'D'                      # the caller (tests, TrainStep) drew the coin
            if d_step:
                self.G.eval()
                with torch.no_grad():
                    fake_pose, partial_i_loss, *args = self.G(x_audio, y_pose, **kwargs)
                    args = args[0] if len(args) > 0 else {}
                self.G.train(self.training)
                self.fake_flag = True
                f32 = getattr(fake_pose, "_ms_f32", None)
                fake_d = fake_pose.detach()
                if f32 is not None:
                    fake_d._ms_f32 = f32.detach()
                fake_score = self._score(fake_d)
                fake_D_loss = self._l1(fake_score, None, 0.0, dt, lam=0, lam_host=lam_D, lam_dev=ld)
                real_score = self._score(y_pose)
                real_D_loss = self._l1(real_score, None, 1.0, dt)
                internal_losses.append(real_D_loss)
                internal_losses.append(fake_D_loss)
                internal_losses += partial_i_loss
                self.G_flag = False
            else:
                fake_pose, partial_i_loss, *args = self.G(x_audio, y_pose, **kwargs)
                args = args[0] if len(args) > 0 else {}
                if self.no_grad:
                    with torch.no_grad():
                        fake_score = self._score(fake_pose)
                else:
                    fake_score = self._score(fake_pose)
                G_gan_loss = self._l1(fake_score, None, 1.0, dt, lam=1, lam_host=lam_gan, lam_dev=ld)
                pose_loss = self._l1(fake_pose, y_pose, 0.0, dt)
                internal_losses.append(pose_loss)
                internal_losses.append(G_gan_loss)
                internal_losses += partial_i_loss
                self.G_flag = True
        else:
            fake_pose, partial_i_loss, *args = self.G(x_audio, y_pose, **kwargs)
            args = args[0] if len(args) > 0 else {}
            internal_losses.append(self._l1(fake_pose, y_pose, 0.0, dt))
            internal_losses.append(torch.tensor(0))
            internal_losses += partial_i_loss
            self.G_flag = True
        args.update(dict(W=W))
        return fake_pose, internal_losses, args
