"""Pose preprocessing on the device -- the step immediately before the generator hot path.

Per batch the reference's trainer computes, on the host in fp64 (src/model/trainer.py:1290-1308):

    labels = self.cluster(self.transform_cluster(batch[pose]))      # KMeans.predict(RemoveJoints(raw pose))
    batch  = self.pre(batch)                                        # ZNorm
    y      = self.transform(batch[pose].to(device))                 # RemoveJoints

`PosePreprocessor` mirrors the three reference objects involved (`KMeans`, `ZNorm`, `RemoveJoints`,
src/data/transform.py:150-245, 247-415, 463-510) for the pose modality and runs them as ONE CUDA kernel
(csrc/preprocess.cu, `ms_pose_prepare`) over the raw pose batch already on the device: no per-batch CPU work and one H2D copy
of the raw fp64 pose instead of three tensors.  Arithmetic is fp64 as in the reference; cluster labels are the first index
of the minimum distance (torch.min's tie rule).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import MixStageError, call, ptr, stream

FEATS = {"pose": 1, "velocity": 2, "speed": 3, "acceleration": 4}


def _need_cuda(t, what):
    if not t.is_cuda:
        raise MixStageError("%s: CUDA tensor required (mixstage_b200 has no CPU fallback)" % what)


class PosePreprocessor:
    """mask: joints removed by RemoveJoints (default [0, 7, 8, 9], src/argsUtils.py:21); muvar: (mean, var) of the raw pose
    as ZNorm stores them (shape (..., Pr)); centers: (K, D) k-means centres (KMeans.centers); feats: the feature list the
    centres were fitted on (KMeans.feats; 'spatial' is not supported)."""

    def __init__(self, num_joints=52, mask=(0, 7, 8, 9), muvar=None, centers=None, feats=("pose", "velocity", "speed"),
                 eps=1e-8, device="cuda"):
        self.device = torch.device(device)
        self.J = int(num_joints)
        self.mask = sorted(int(m) for m in mask)
        keep = [j for j in range(self.J) if j not in set(self.mask)]
        # (B,T,2,J) view: x coordinates first, y coordinates second (transform.py:495, 501)
        cols = keep + [self.J + j for j in keep]
        self.Pr, self.P = 2 * self.J, len(cols)
        self.cols = torch.tensor(cols, dtype=torch.int32, device=self.device)
        self.eps = float(eps)
        self.mean = self.var = None
        if muvar is not None:
            self.mean = torch.as_tensor(muvar[0], dtype=torch.float64).reshape(-1).contiguous().to(self.device)
            self.var = torch.as_tensor(muvar[1], dtype=torch.float64).reshape(-1).contiguous().to(self.device)
            if self.mean.numel() != self.Pr or self.var.numel() != self.Pr:
                raise MixStageError("muvar must have %d entries" % self.Pr)
        for f in feats:
            if f not in FEATS:
                raise MixStageError("KMeans feature %r is not supported on the device" % (f,))
        self.feats = list(feats)
        self._feats_c = (ctypes.c_int32 * len(self.feats))(*[FEATS[f] for f in self.feats])
        self.D = sum(self.P // 2 if f == "speed" else self.P for f in self.feats)
        self.centers = None
        if centers is not None:
            c = torch.as_tensor(centers, dtype=torch.float64).contiguous().to(self.device)
            if c.dim() != 2 or c.shape[1] != self.D:
                raise MixStageError("centers must be (K, %d) for feats %s" % (self.D, self.feats))
            self.centers = c

    # ------------------------------------------------------------------
    def _check(self, x):
        _need_cuda(x, "PosePreprocessor")
        if x.dim() != 3 or x.shape[-1] != self.Pr:
            raise MixStageError("raw pose must be (B, T, %d)" % self.Pr)
        return x.to(torch.float64).contiguous()          # KMeans.predict: x.double() (transform.py:393)

    def __call__(self, pose_raw, soft_labels=False):
        """raw pose (B,T,2J) -> (y (B,T,P) = RemoveJoints(ZNorm(pose)), labels (B,T) int64, or (B,T,K) soft labels)."""
        x = self._check(pose_raw)
        if self.mean is None or self.centers is None:
            raise MixStageError("PosePreprocessor needs muvar and centers for the fused call")
        B, T, _ = x.shape
        K = self.centers.shape[0]
        y = torch.empty(B, T, self.P, dtype=torch.float64, device=x.device)
        lab = torch.empty(B, T, dtype=torch.int64, device=x.device) if not soft_labels else None
        soft = torch.empty(B, T, K, dtype=torch.float64, device=x.device) if soft_labels else None
        call("ms_pose_prepare", ptr(x), ptr(self.mean), ptr(self.var), ptr(self.cols), ptr(self.centers), B, T, self.Pr, self.P,
             K, self._feats_c, len(self.feats), self.eps, ptr(y), ptr(lab), ptr(soft), stream())
        return y, (soft if soft_labels else lab)

    def predict(self, pose_raw, soft_labels=False):
        """KMeans.predict(RemoveJoints(pose_raw)) (transform.py:392-407)."""
        x = self._check(pose_raw)
        if self.centers is None:
            raise MixStageError("PosePreprocessor.predict needs centers")
        B, T, _ = x.shape
        K = self.centers.shape[0]
        lab = torch.empty(B, T, dtype=torch.int64, device=x.device) if not soft_labels else None
        soft = torch.empty(B, T, K, dtype=torch.float64, device=x.device) if soft_labels else None
        call("ms_pose_prepare", ptr(x), None, None, ptr(self.cols), ptr(self.centers), B, T, self.Pr, self.P, K, self._feats_c,
             len(self.feats), self.eps, None, ptr(lab), ptr(soft), stream())
        return soft if soft_labels else lab

    def normalize(self, pose_raw):
        """RemoveJoints(ZNorm.znorm(pose_raw)) (transform.py:221-226, 499-508)."""
        x = self._check(pose_raw)
        if self.mean is None:
            raise MixStageError("PosePreprocessor.normalize needs muvar")
        B, T, _ = x.shape
        y = torch.empty(B, T, self.P, dtype=torch.float64, device=x.device)
        call("ms_pose_prepare", ptr(x), ptr(self.mean), ptr(self.var), ptr(self.cols), None, B, T, self.Pr, self.P, 0, None, 0,
             self.eps, ptr(y), None, None, stream())
        return y

    def inv_znorm(self, x):
        """ZNorm.inv_znorm (transform.py:228-229) on a full-width (.., 2J) pose."""
        _need_cuda(x, "inv_znorm")
        if x.shape[-1] != self.Pr or self.mean is None:
            raise MixStageError("inv_znorm: tensor (..., %d) and muvar required" % self.Pr)
        x = x.to(torch.float64).contiguous()
        out = torch.empty_like(x)
        call("ms_inv_znorm", ptr(x), ptr(self.mean), ptr(self.var), x.numel() // self.Pr, self.Pr, ptr(out), stream())
        return out


class PoseMetrics:
    """L1, VelL1 and PCK of the reference's evaluation loop (src/evaluation/metrics.py:94-131, 247-303, driven by
    TrainerBase.calculate_metrics, src/model/trainer.py:865-907) accumulated from ONE kernel per batch on the device
    (`ms_pose_metrics`); the running averages follow AverageMeter's weighting (metrics.py:36-62) so `get_averages(desc)`
    returns the reference's keys and values.  Inputs are the full-width normalised poses (masked joints re-inserted)."""

    def __init__(self, muvar, num_joints=52, mask=(0, 7, 8, 9), alphas=(0.1, 0.2), device="cuda"):
        self.device = torch.device(device)
        self.J = int(num_joints)
        self.mask = sorted(int(m) for m in mask)
        self.alphas = [float(a) for a in alphas]
        keep = torch.ones(self.J, dtype=torch.uint8)
        keep[self.mask] = 0
        self.keep = keep.to(self.device)
        self.nkeep = int(keep.sum())
        self.mean = torch.as_tensor(muvar[0], dtype=torch.float64).reshape(-1).contiguous().to(self.device)
        self.var = torch.as_tensor(muvar[1], dtype=torch.float64).reshape(-1).contiguous().to(self.device)
        self._alphas_c = (ctypes.c_double * len(self.alphas))(*self.alphas)
        self.reset()

    def reset(self):
        z = lambda: torch.zeros((), dtype=torch.float64)                                # noqa: E731
        self.l1_sum, self.vel_sum, self.n = z(), z(), 0
        self.pck_joint_sum = torch.zeros(len(self.alphas), self.J, dtype=torch.float64)   # sum of (per-call mean) * n
        self.pck_alpha_sum = torch.zeros(len(self.alphas), dtype=torch.float64)
        self.pck_n, self.pck_alpha_n = 0, 0
        self.pck_sum, self.pck_cnt = z(), 0

    def __call__(self, y_cap, y_gt):
        """y_cap, y_gt: (B, T, 2J) CUDA tensors (normalised).  One kernel, one 8*(2 + nalpha*J)-byte read-back."""
        _need_cuda(y_cap, "PoseMetrics")
        _need_cuda(y_gt, "PoseMetrics")
        if y_cap.shape != y_gt.shape or y_cap.dim() != 3 or y_cap.shape[-1] != 2 * self.J:
            raise MixStageError("PoseMetrics: tensors of shape (B, T, %d) required" % (2 * self.J))
        y = y_cap.to(torch.float64).contiguous()
        g = y_gt.to(torch.float64).contiguous()
        B, T, _ = y.shape
        acc = torch.empty(2, dtype=torch.float64, device=y.device)
        cnt = torch.empty(len(self.alphas) * self.J, dtype=torch.int64, device=y.device)
        call("ms_pose_metrics", ptr(y), ptr(g), ptr(self.mean), ptr(self.var), ptr(self.keep), B, T, self.J, self._alphas_c,
             len(self.alphas), ptr(acc), ptr(cnt), stream())
        acc, cnt = acc.cpu(), cnt.cpu().view(len(self.alphas), self.J).double()
        # L1 / VelL1: l1_loss means, AverageMeter.update(val, n=B)
        self.l1_sum += acc[0] / (B * T * 2 * self.nkeep) * B
        if T > 1:
            self.vel_sum += acc[1] / (B * (T - 1) * 2 * self.nkeep) * B
        self.n += B
        # PCK on (B*T) frames: per joint update(pck.mean(0)[j], n=frames); per alpha update(pck[:, kept].mean(), n=frames*|kept|)
        F_ = B * T
        self.pck_joint_sum += cnt / F_ * F_
        self.pck_n += F_
        kept = self.keep.cpu().bool()
        for a in range(len(self.alphas)):
            self.pck_alpha_sum[a] += cnt[a][kept].sum() / (F_ * self.nkeep) * (F_ * self.nkeep)
        self.pck_alpha_n += F_ * self.nkeep
        for a in range(len(self.alphas)):                   # metrics.py:271-272: running average of the running averages
            self.pck_sum += (self.pck_alpha_sum[a] / self.pck_alpha_n) * (F_ * self.nkeep)
            self.pck_cnt += F_ * self.nkeep

    def get_averages(self, desc):
        out = {"%s_L1" % desc: float(self.l1_sum / max(self.n, 1)), "%s_VelL1" % desc: float(self.vel_sum / max(self.n, 1))}
        for a, al in enumerate(self.alphas):
            for j in range(self.J):
                out["%s_pck_%s_%d" % (desc, al, j)] = float(self.pck_joint_sum[a, j] / max(self.pck_n, 1))
            out["%s_pck_%s" % (desc, al)] = float(self.pck_alpha_sum[a] / max(self.pck_alpha_n, 1))
        out["%s_pck" % desc] = float(self.pck_sum / max(self.pck_cnt, 1))
        return out
