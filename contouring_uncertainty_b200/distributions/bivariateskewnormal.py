"""Mirror of contour_uncertainty/distributions/bivariateskewnormal.py::BivariateSkewNormal (classmethods, torch in / out).

``logpdf`` / ``pdf`` / ``rvs_fast`` run sm_100a kernels; ``nll`` (training loss) stays differentiable torch as in the
reference.  ``pdf`` on the sampler's 256 x 256 integer grid for whole batches is ``engine.skew_pdf_grid``."""
import math

import torch

from .. import engine as E
from .bivariatenormal import BivariateNormal, _dev


class BivariateSkewNormal:
    log2 = torch.log(torch.tensor(2))

    @classmethod
    def logpdf(cls, x, loc, cov, alpha):
        """bivariateskewnormal.py:20-34: log 2 + log N(x; loc, cov) + log(Phi(alpha^T cov^-1/2 (x - loc)) + 1e-7)."""
        d = _dev(x)
        f = lambda t: t.to(d, torch.float32)
        return E.bivariate_pdf(f(x), f(loc), f(cov), f(alpha), log=True).to(x.device)

    @classmethod
    def pdf(cls, x, loc, cov, alpha):
        d = _dev(x)
        f = lambda t: t.to(d, torch.float32)
        return E.bivariate_pdf(f(x), f(loc), f(cov), f(alpha), log=False).to(x.device)

    @classmethod
    def affine(cls, x, loc, cov, alpha):
        """bivariateskewnormal.py:36-44 (torch; cov^-1/2 in closed form for symmetric positive-definite 2x2 blocks)."""
        s = torch.sqrt(torch.det(cov))[..., None, None]
        t = torch.sqrt(cov[..., 0, 0] + cov[..., 1, 1] + 2 * s[..., 0, 0])[..., None, None]
        eye = torch.eye(2, dtype=cov.dtype, device=cov.device)
        inv_sqrt = torch.inverse((cov + s * eye) / t)
        return torch.bmm(alpha.transpose(-1, -2), inv_sqrt) @ (x - loc)

    @classmethod
    def unit_normal_logcdf(cls, x):
        """bivariateskewnormal.py:46-49."""
        return torch.log(0.5 * (1 + torch.erf(x / math.sqrt(2))) + 1e-7)

    @classmethod
    def nll(cls, y, mu, cov, alpha):
        """bivariateskewnormal.py:51-61 (differentiable torch, training only)."""
        term1 = torch.log(torch.det(cov)).squeeze()
        term2 = (((mu - y).transpose(-1, -2) @ torch.inverse(cov)) @ (mu - y)).squeeze()
        term3 = cls.unit_normal_logcdf(cls.affine(y, mu, cov, alpha).squeeze()).squeeze()
        return 0.5 * term1 + 0.5 * term2 - term3, term1, term2, term3

    @classmethod
    def mode(cls, mu, cov, alpha):
        """bivariateskewnormal.py:73-81 (+ m0 / skewness / delta :195-215, cov2corr distributions/utils.py:38-73): closed-form
        approximation of the mode, a handful of scalar operations on (2,) / (2,2) tensors as in the reference."""
        mu, cov, alpha = (torch.as_tensor(t, dtype=torch.float32) for t in (mu, cov, alpha))
        c = cov.reshape(-1, 2, 2)
        std = torch.sqrt(torch.diagonal(c, dim1=1, dim2=2))
        corr = c / torch.bmm(std.unsqueeze(2), std.unsqueeze(1))
        alpha_star = torch.sqrt(alpha.T @ corr @ alpha)
        x = torch.tensor(2 / torch.pi)
        delta = alpha_star / torch.sqrt(1 + torch.square(alpha_star))
        mu_z = torch.sqrt(x) * delta
        sigma_z = torch.sqrt(1 - torch.square(mu_z))
        skew = ((4 - torch.pi) / 2) * torch.pow(delta * torch.sqrt(x), 3) / torch.pow(1 - 2 * torch.square(delta) / torch.pi, 3 / 2)
        m0 = mu_z - skew * sigma_z / 2 - torch.sign(alpha_star) / 2 * torch.exp(-2 * torch.pi / torch.abs(alpha_star))
        mode = mu + m0 / alpha_star * std @ corr @ alpha
        return mode.squeeze()

    @classmethod
    def marginal(cls, mu, cov, alpha, axis: int, angle=torch.tensor(0), *args, **kwargs):
        """bivariateskewnormal.py:92-135: (mu[axis], rotated variance, projected skewness)."""
        assert axis == 0 or axis == 1
        c, s = torch.cos(-angle), torch.sin(-angle)
        R = torch.tensor([[c, -s], [s, c]]).float()
        cov = R @ torch.as_tensor(cov, dtype=torch.float32) @ R.T
        alpha = torch.as_tensor(alpha, dtype=torch.float32).clone()
        alpha[1] = -alpha[1]
        alpha = R @ alpha
        std = torch.sqrt(torch.diagonal(cov))
        corr = cov / torch.outer(std, std)
        na = 1 - axis
        corr_22_1 = corr[na, na] - corr[0, 1] * corr[0, 1] / corr[axis, axis]
        alpha_1_2 = (alpha[axis] + (1 / corr[axis, axis]) * corr[0, 1] * alpha[na]) / torch.sqrt(1 + alpha[na] * corr_22_1 * alpha[na])
        return mu[axis], cov[axis, axis], alpha_1_2

    @classmethod
    def rvs_fast(cls, mu, cov, alpha, size=1, z=None, seed: int = 0):
        """bivariateskewnormal.py:160-191: ``size`` draws (n, 2).  ``z`` (n, 3) replays the standard normals."""
        n = int(size[0]) if isinstance(size, (tuple, list, torch.Size)) else int(size)
        d = _dev(mu)
        f = lambda t: torch.as_tensor(t, dtype=torch.float32).to(d).contiguous()
        out = E.skew_rvs_fast(f(mu).reshape(2), f(cov).reshape(2, 2), f(alpha).reshape(2), n, z=None if z is None else f(z), seed=seed)
        return out.to(mu.device)


__all__ = ["BivariateSkewNormal", "BivariateNormal"]
