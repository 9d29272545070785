"""Reference-shaped orchestration of the predict path (no Lightning dependency).

``ContourUncertaintyPipeline`` does what ``AleatoricUncertaintyTask._predict_step`` +
``ContourUncertaintyTask.predict_step`` do after the network forward
(cu/task/regression/aleatoric.py:80-135, contour_uncertainty.py:71-82) for T_e epistemic passes
given as a list of logit tensors, and returns a ``BatchResult``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .. import engine as E
from ..data.config import BatchResult


def sample_entropy(samples) -> np.ndarray:
    """UncertaintyTask.sample_entropy (cu/task/uncertainty.py:107-133).

    samples (T,C,H,W) or (T_e,T_a,C,H,W).  0/1 masks with C == 1 (the contour path): binary entropy (bits) of the sample
    mean in fp64 -- the count map is formed by an exact integer sum, the entropy table runs on the GPU (K4).  Probabilities
    or C > 1: fp32 like the reference's tensors (base 2 for C == 1, base C otherwise).
    """
    s = np.asarray(samples)
    if s.ndim == 5:
        s = s.reshape((-1,) + s.shape[2:])
    if s.shape[1] == 3 and tuple(s.shape[2:]) == (E.HW, E.HW) and np.isin(s, (0, 1)).all() and (s.sum(1) == 1).all():
        # one-hot BG / LV / MYO stack (labels with Label.MYO): exact class counts, fp64 entropy in base 3
        dev = torch.device("cuda", torch.cuda.current_device())
        pack = lambda m: torch.from_numpy(np.ascontiguousarray(np.packbits(m.astype(bool), axis=-1, bitorder="little")).view(np.int32)
                                          .reshape(1, s.shape[0], E.MASK_WORDS)).to(dev)
        return E.mask_class_entropy(pack(s[:, 1]), pack(s[:, 2]))[0].cpu().numpy()
    if s.shape[1] != 1 or not np.isin(s, (0, 1)).all():
        # probabilities / several classes: the general kernel (fp32 like the reference's tensors)
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32))[None].to(dev)
        return E.sample_entropy(t)[0].cpu().numpy()
    T = s.shape[0]
    count = torch.from_numpy(s[:, 0].astype(np.int32).sum(0, dtype=np.int32))
    dev = torch.device("cuda", torch.cuda.current_device())
    count = count.to(dev)
    diff = torch.zeros((1,) + tuple(count.shape), dtype=torch.int32, device=dev)
    diff[0, :, 0] = count[:, 0]
    diff[0, :, 1:] = count[:, 1:] - count[:, :-1]
    _, _, ent, _ = E.finalize_maps(diff, T, want_count=False)
    return ent[0].cpu().numpy()


def batch_result_from(res: E.UncertaintyResult, img, uncertainty_map=None, gt=None, id="", labels=("lv",),
                      return_pred_samples: bool = False, **kw) -> BatchResult:
    n = res.mu.shape[0]
    T = res.area.shape[1]
    ent = res.entropy.cpu().numpy()
    pred_samples = None
    if return_pred_samples:
        if res.masks is None:
            raise ValueError("run the engine with return_masks=True to materialise pred_samples")
        pred_samples = E.unpack_masks(res.masks).cpu().numpy().astype(np.int64).reshape(n, 1, T, 1, E.HW, E.HW)
    return BatchResult(
        id=id, labels=labels, img=np.asarray(img), gt=gt, pred=res.pred.cpu().numpy().astype(np.int64),
        uncertainty_map=ent if uncertainty_map is None else uncertainty_map, mu=res.mu.cpu().numpy(),
        mode=res.mu.cpu().numpy(), cov=res.cov.cpu().numpy(),
        contour_samples=None if res.contour_samples is None else res.contour_samples.cpu().numpy().reshape(n, 1, T, E.K, 2),
        pred_samples=pred_samples, entropy_map=ent, sample_area=res.area.cpu().numpy().reshape(n, 1, T), **kw)


def add_point_and_instant_uncertainty(res: BatchResult, mu_t: torch.Tensor, cov_t: torch.Tensor) -> BatchResult:
    """ContourUncertaintyTask.predict_step, lines 85-128 of cu/task/regression/contour_uncertainty.py: per-point and
    per-instant scalar uncertainties.  (N,K) algebra in numpy as in the reference; the spline-normal projection
    (projected_uncertainty_value) runs on the GPU."""
    LA = np.linalg
    cov = res.cov
    cov_det = LA.det(cov) ** 0.25
    cov_eigval = np.sqrt(LA.eig(cov)[0])
    res.point_uncertainty = {"cov_xx": cov[:, :, 0, 0] ** 0.5, "cov_yy": cov[:, :, 1, 1] ** 0.5, "cov_det": cov_det,
                             "cov_eigval_sum": cov_eigval.sum(-1)}
    if res.post_cov is not None:
        pe = np.sqrt(LA.eig(res.post_cov)[0])
        res.point_uncertainty.update({"post_cov_xx": res.post_cov[:, :, 0, 0] ** 0.5, "post_cov_yy": res.post_cov[:, :, 1, 1] ** 0.5,
                                      "post_cov_det": LA.det(res.post_cov) ** 0.25, "post_cov_eigval_sum": pe.sum(-1)})
    unc, _ = E.projected_uncertainty(mu_t.to(torch.float32).contiguous(), cov_t.to(torch.float32).contiguous())
    mask = res.pred != 0                                               # Label.BG == 0
    with np.errstate(divide="ignore", invalid="ignore"):
        res.instant_uncertainty = {"cov_det_mean": np.mean(cov_det, axis=-1),
                                   "cov_eigenvalue_mean": np.mean(cov_eigval, axis=(-1, -2)),
                                   "cov_projection": unc.sum(1).cpu().numpy(),
                                   "umap_mean": np.sum(res.uncertainty_map, axis=(-2, -1)) / np.sum(mask, axis=(-2, -1))}
        if res.entropy_map is not None:
            res.instant_uncertainty["entropy_mean"] = np.sum(res.entropy_map, axis=(-2, -1)) / np.sum(mask, axis=(-2, -1))
    return res


class ContourUncertaintyPipeline:
    """dsnt-al / dsnt-al2 predict path on logits.  ``covar`` = hparams.covar (config/task/dsnt-al.yaml)."""

    def __init__(self, psm_path: Optional[str] = None, t_a: int = 25, covar: bool = True, sampler: str = "psm",
                 device="cuda", sequence_sampler: bool = False, seq_psm_path: Optional[str] = None):
        """``sequence_sampler`` = hparams.sequence_sampler (aleatoric.py:43-49): consecutive images are
        (ED, ES) pairs of one view and are sampled jointly through the 84-dim sequence PSM."""
        self.t_a = t_a
        self.engine = E.UncertaintyEngine(psm_path or E.DEFAULT_PSM, device=device, sampler=sampler, covar=covar)
        self.seq = E.SequenceModelConstants(seq_psm_path or E.DEFAULT_SEQ_PSM, device=self.engine.device) if sequence_sampler else None

    def predict(self, logits_per_model: Sequence[torch.Tensor]):
        """DSNTAleatoric.predict (dsnt_al.py:133-151): stack T_e passes -> (N,T_e,K,2), (N,T_e,K,2,2)."""
        S, cov = [], []
        for lg in logits_per_model:
            _, mu, sig = E.dsnt_moments(lg, is_logit=True, covar=self.engine.covar)
            S.append(mu)
            cov.append(sig)
        return torch.stack(S, 1), torch.stack(cov, 1)

    def predict_step(self, logits, img=None, seed: int = 0, z=None, return_pred_samples: bool = False, **kw) -> BatchResult:
        """logits (N,21,256,256) CUDA, or a list of T_e such tensors (ensemble / MC-dropout passes) -> BatchResult.

        Follows AleatoricUncertaintyTask._predict_step (aleatoric.py:80-135) + predict_step
        (contour_uncertainty.py:71-82): T_a samples per (image, pass), masks of all T_e*T_a samples
        aggregated per image, epistemic + aleatoric covariance, posterior moments of the samples.
        """
        passes = list(logits) if isinstance(logits, (list, tuple)) else [logits]
        mu, cov = self.predict(passes)                                  # (N,T_e,K,2), (N,T_e,K,2,2)
        return self.from_moments(mu, cov, img=img, seed=seed, z=z, return_pred_samples=return_pred_samples, **kw)

    def from_moments(self, mu: torch.Tensor, cov: torch.Tensor, **kw) -> BatchResult:
        """Everything of _predict_step / predict_step after ``predict`` (aleatoric.py:88-135, contour_uncertainty.py:71-130):
        per-pass moments mu (N,T_e,K,2), cov (N,T_e,K,2,2) fp32 CUDA -> BatchResult."""
        return add_point_and_instant_uncertainty(*self.moments_to_result(mu, cov, **kw))

    def moments_to_result(self, mu: torch.Tensor, cov: torch.Tensor, img=None, seed: int = 0, z=None,
                          return_pred_samples: bool = False, first_instant=None, umap_fn=None, **kw):
        """AleatoricUncertaintyTask._predict_step after ``predict`` (aleatoric.py:88-135) with the entropy map of
        predict_step (contour_uncertainty.py:76-77) already filled from the on-device count map.  Returns
        (BatchResult, mu (N,K,2), cov (N,K,2,2) device tensors).  ``umap_fn``: a per-image functor (mu, cov, labels) ->
        (H,W) to use instead of the batched analytic u-map kernel."""
        mu = E.nat.require_cuda(mu, torch.float32, "mu")
        cov = E.nat.require_cuda(cov, torch.float32, "cov")
        if mu.dim() != 4 or mu.shape[2:] != (E.K, 2) or tuple(cov.shape) != tuple(mu.shape) + (2,):
            raise ValueError("mu (N,T_e,21,2) / cov (N,T_e,21,2,2) expected")
        n, te, T = mu.shape[0], mu.shape[1], self.t_a
        m2, c2 = mu.reshape(n * te, E.K, 2).contiguous(), cov.reshape(n * te, E.K, 2, 2).contiguous()
        if self.seq is not None:
            # aleatoric.py:66-69: per pass, the (ED, ES) pairs are sampled jointly; (n,2,K,2) per pair -> (N,T_e,T_a,K,2)
            if n % 2:
                raise ValueError("sequence_sampler needs (ED, ES) pairs: an even number of images")
            per_pass = []
            for e in range(te):
                zt = None if z is None else torch.as_tensor(z, dtype=torch.float32, device=mu.device).reshape(te, n // 2, T, 2, E.Z_PSM)[e].contiguous()
                fi = None if first_instant is None else torch.as_tensor(first_instant, device=mu.device).to(torch.uint8).reshape(te, n // 2, T)[e].contiguous()
                sp = E.sequence_sample(self.engine.model, self.seq, mu[:, e].reshape(n // 2, 2, E.K, 2).contiguous(),
                                       cov[:, e].reshape(n // 2, 2, E.K, 2, 2).contiguous(), T, z=zt, first_instant=fi, seed=seed + 7919 * e)
                per_pass.append(sp.permute(0, 2, 1, 3, 4).reshape(n, T, E.K, 2))         # (Np,T,2,..) -> (N,T,..)
            samples = torch.stack(per_pass, 1).contiguous()                                # (N,T_e,T_a,K,2)
            diff = torch.zeros((n, E.HW, E.HW), dtype=torch.int32, device=mu.device)
            masks, area = E.rasterize_contours(samples.reshape(-1, E.K, 2), return_masks=return_pred_samples, diff=diff,
                                               per_image=te * T)
            r = dict(diff=diff, area=area, masks=masks)
        else:
            a_t = self.engine.factor(m2, c2)
            zt = None if z is None else torch.as_tensor(z, dtype=torch.float32, device=mu.device).reshape(n * te, T, a_t.shape[1])
            r = E.sample_raster_aggregate(m2, a_t, T, z=zt, seed=seed, return_masks=return_pred_samples, images_per_map=te)
            samples = r["contours"].reshape(n, te, T, E.K, 2)
        count, mean, ent, pred, var = E.finalize_maps(r["diff"], te * T, want_variance=True)
        # aleatoric.py:90-94 (tiny (N,T_e,K) tensors: plain tensor algebra)
        mu_mean = mu.mean(1, keepdim=True)
        d = (mu - mu_mean)[..., None]
        cov_out = cov.mean(1) + (d * d.transpose(-1, -2)).mean(1)
        # aleatoric.py:96-108: per-pass sample mean / unbiased covariance (kernel), then the same split
        pm, pc = E.sample_moments(samples)                              # (N,T_e,K,2), (N,T_e,K,2,2) fp64
        d = (pm - pm.mean(1, keepdim=True))[..., None]
        post_cov = pc.mean(1) + (d * d.transpose(-1, -2)).mean(1)
        if img is None:
            img = np.zeros((n, E.HW, E.HW), dtype=np.float32)
        ent_np = ent.cpu().numpy()
        pred_samples = None
        if return_pred_samples:
            pred_samples = E.unpack_masks(r["masks"]).cpu().numpy().astype(np.int64).reshape(n, te, T, 1, E.HW, E.HW)
        mu_mean_t = mu.mean(1).contiguous()
        mu_np = mu_mean_t.cpu().numpy()
        umap = kw.pop("uncertainty_map", None)
        labels = kw.pop("labels", ("lv",))
        cov_np = cov_out.cpu().numpy()
        if umap is None and umap_fn is not None:                        # aleatoric.py:114, a caller-supplied functor
            umap = np.array([umap_fn(mu_np[i], cov_np[i], labels) for i in range(n)])
        if umap is None:                                                # aleatoric.py:114: umap_fn = USUMap (analytic map / max)
            umap = E.umap_analytic(mu_mean_t, cov_out.contiguous(), normalize=True).cpu().numpy()
        res = BatchResult(
            id=kw.pop("id", ""), labels=labels, img=np.asarray(img), gt=kw.pop("gt", None),
            pred=pred.cpu().numpy().astype(np.int64), uncertainty_map=umap,
            mu=mu_np, mode=mu_np, cov=cov_np, post_mu=pm.mean(1).cpu().numpy(), post_cov=post_cov.cpu().numpy(),
            contour_samples=samples.cpu().numpy(), pred_samples=pred_samples, entropy_map=ent_np,
            sample_area=r["area"].reshape(n, te, T).cpu().numpy(), variance_map=var.cpu().numpy(),
            packed_samples=None if r["masks"] is None else r["masks"].reshape(n, te, T, E.MASK_WORDS).cpu().numpy(), **kw)
        return res, mu_mean_t, cov_out.contiguous()



class SkewUncertaintyPipeline:
    """dsnt-skew predict path: ``SkewUncertaintyTask._predict_step`` (cu/task/regression/aleatoric_skew.py:55-127) +
    ``predict_step`` (contour_uncertainty.py:71-130) after the network forward.

    The network gives heatmap logits and, from its bottleneck, the raw skewness vectors (``skew_block``); both stay
    torch/cuDNN.  Everything after is here: DSNT moments, the numerical skew-PSM sampler (25 samples per pass in the
    reference, aleatoric_skew.py:63), posterior moments, projected mode + skew u-map, mask of the mode, sample masks
    -> entropy map."""

    def __init__(self, psm_path: Optional[str] = None, t_a: int = 25, covar: bool = True, device="cuda"):
        self.t_a = t_a
        self.covar = covar
        self.device = torch.device(device)
        self.model = E.ShapeModelConstants(psm_path or E.DEFAULT_PSM, device=self.device)

    def predict(self, logits_per_model: Sequence[torch.Tensor], alpha_per_model: Sequence[torch.Tensor], negate_alpha_y: bool = True):
        """DSNTSkew.predict (dsnt_skew.py:153-199): -> S (N,T_e,K,2), cov (N,T_e,K,2,2), alpha (N,T_e,K,2); the raw
        skew_block output has its y component negated (:164)."""
        S, cov, al = [], [], []
        for lg, a in zip(logits_per_model, alpha_per_model):
            _, mu, sig = E.dsnt_moments(lg, is_logit=True, covar=self.covar)
            a = a.to(mu.device, torch.float32)
            if negate_alpha_y:
                a = a * torch.tensor([1.0, -1.0], device=a.device)
            S.append(mu); cov.append(sig); al.append(a)
        return torch.stack(S, 1), torch.stack(cov, 1), torch.stack(al, 1)

    def predict_step(self, logits, alpha, img=None, seed: int = 0, z=None, expo=None, return_pred_samples: bool = False,
                     negate_alpha_y: bool = True, **kw) -> BatchResult:
        passes = list(logits) if isinstance(logits, (list, tuple)) else [logits]
        alphas = list(alpha) if isinstance(alpha, (list, tuple)) else [alpha]
        mu, cov, al = self.predict(passes, alphas, negate_alpha_y)
        return self.from_moments(mu, cov, al, img=img, seed=seed, z=z, expo=expo, return_pred_samples=return_pred_samples, **kw)

    def from_moments(self, mu: torch.Tensor, cov: torch.Tensor, al: torch.Tensor, **kw) -> BatchResult:
        """Everything of SkewUncertaintyTask._predict_step / predict_step after ``predict`` (aleatoric_skew.py:63-127,
        contour_uncertainty.py:71-130): mu (N,T_e,K,2), cov (N,T_e,K,2,2), alpha (N,T_e,K,2) fp32 CUDA -> BatchResult."""
        return add_point_and_instant_uncertainty(*self.moments_to_result(mu, cov, al, **kw))

    def moments_to_result(self, mu: torch.Tensor, cov: torch.Tensor, al: torch.Tensor, img=None, seed: int = 0, z=None, expo=None,
                          return_pred_samples: bool = False, skew_umap_fn=None, **kw):
        """SkewUncertaintyTask._predict_step after ``predict`` (aleatoric_skew.py:63-127) + the entropy map of predict_step.
        Returns (BatchResult, mu (N,K,2), cov (N,K,2,2) device tensors)."""
        mu = E.nat.require_cuda(mu, torch.float32, "mu")
        cov = E.nat.require_cuda(cov, torch.float32, "cov")
        al = E.nat.require_cuda(al, torch.float32, "alpha")
        if mu.dim() != 4 or mu.shape[2:] != (E.K, 2) or tuple(cov.shape) != tuple(mu.shape) + (2,) or al.shape != mu.shape:
            raise ValueError("mu (N,T_e,21,2) / cov (N,T_e,21,2,2) / alpha (N,T_e,21,2) expected")
        n, te, T = mu.shape[0], mu.shape[1], self.t_a
        m2, c2, a2 = (t.reshape((n * te,) + t.shape[2:]).contiguous() for t in (mu, cov, al))
        dev = mu.device
        zt = None if z is None else torch.as_tensor(z, dtype=torch.float32, device=dev).reshape(n * te, T, 3, 3).contiguous()
        et = None if expo is None else torch.as_tensor(expo, dtype=torch.float32, device=dev).reshape(n * te, T, E.N_DRAW_SKEW, E.HW * E.HW).contiguous()
        samples = E.skew_sample(self.model, m2, c2, a2, T, z=zt, expo=et, seed=seed).reshape(n, te, T, E.K, 2)
        # aleatoric_skew.py:65-73
        mu_mean = mu.mean(1, keepdim=True)
        d = (mu - mu_mean)[..., None]
        cov_out = (cov.mean(1) + (d * d.transpose(-1, -2)).mean(1)).contiguous()
        mu_m, al_m = mu.mean(1).contiguous(), al.mean(1).contiguous()
        # :75-79  post_mu over all passes and samples, post_cov = np.cov of the pooled samples
        pm, pc = E.sample_moments(samples.reshape(n, te * T, E.K, 2).contiguous())
        # :82-91  projected mode + skew u-map (USSkewUmap), mask of the mode, masks of the samples
        labels = kw.pop("labels", ("lv",))
        if skew_umap_fn is not None:                                    # a caller-supplied per-image functor (:85-88)
            mu_np, cov_np, al_np = mu_m.cpu().numpy(), cov_out.cpu().numpy(), al_m.cpu().numpy()
            mode_l, umap_l = zip(*[skew_umap_fn(mu_np[i], cov_np[i], al_np[i], labels) for i in range(n)])
            mode = torch.as_tensor(np.array(mode_l), dtype=torch.float32, device=dev)
            umap = torch.as_tensor(np.array(umap_l), device=dev)
        else:
            mode, umap = E.skew_umap(mu_m, cov_out, al_m, linear_close=True, normalize=True)
        pmask, _ = E.rasterize_contours(mode.contiguous(), round_landmarks=True, return_masks=True)
        pred = E.unpack_masks(pmask).to(torch.int64)
        diff = torch.zeros((n, E.HW, E.HW), dtype=torch.int32, device=dev)
        masks, area = E.rasterize_contours(samples.reshape(-1, E.K, 2), return_masks=return_pred_samples, diff=diff, per_image=te * T)
        _, _, ent, _, var = E.finalize_maps(diff, te * T, want_count=False, want_variance=True)
        if img is None:
            img = np.zeros((n, E.HW, E.HW), dtype=np.float32)
        pred_samples = None
        if return_pred_samples:
            pred_samples = E.unpack_masks(masks).cpu().numpy().astype(np.int64).reshape(n, te, T, E.HW, E.HW)
        res = BatchResult(
            id=kw.pop("id", ""), labels=labels, img=np.asarray(img), gt=kw.pop("gt", None),
            pred=pred.cpu().numpy(), uncertainty_map=umap.cpu().numpy(), mu=mu_m.cpu().numpy(), mode=mode.cpu().numpy(),
            cov=cov_out.cpu().numpy(), alpha=al_m.cpu().numpy(), post_mu=pm.cpu().numpy(), post_cov=pc.cpu().numpy(),
            contour_samples=samples.cpu().numpy(), pred_samples=pred_samples, entropy_map=ent.cpu().numpy(),
            sample_area=area.reshape(n, te, T).cpu().numpy(), variance_map=var.cpu().numpy(),
            packed_samples=None if masks is None else masks.reshape(n, te, T, E.MASK_WORDS).cpu().numpy(), **kw)
        return res, mu_m, cov_out
