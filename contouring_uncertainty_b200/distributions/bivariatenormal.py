"""Mirror of contour_uncertainty/distributions/bivariatenormal.py::BivariateNormal (classmethods, torch in / out).

``logpdf`` / ``pdf`` run the sm_100a kernel (``cub200_bivariate_pdf``); ``nll`` is the training loss and stays plain,
differentiable torch exactly as in the reference (training is outside the accelerated path)."""
import torch

from .. import engine as E


def _dev(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


class BivariateNormal:
    log2pi = torch.log(torch.tensor(2) * torch.pi)

    @classmethod
    def logpdf(cls, x, loc, cov, *args, **kwargs):
        """bivariatenormal.py:15-32: x (..., 2); loc (2,) / (N,2); cov (2,2) / (N,2,2) -> x.shape[:-1]."""
        d = _dev(x)
        out = E.bivariate_pdf(x.to(d, torch.float32), loc.to(d, torch.float32), cov.to(d, torch.float32), None, log=True)
        return out.to(x.device)

    @classmethod
    def pdf(cls, x, loc, cov, *args, **kwargs):
        d = _dev(x)
        out = E.bivariate_pdf(x.to(d, torch.float32), loc.to(d, torch.float32), cov.to(d, torch.float32), None, log=False)
        return out.to(x.device)

    @classmethod
    def nll(cls, y, mu, cov, *args, **kwargs):
        """bivariatenormal.py:38-43 (differentiable torch, training only)."""
        term1 = torch.log(torch.det(cov))
        term2 = ((mu - y).transpose(-1, -2) @ torch.inverse(cov)) @ (mu - y)
        return term1 + term2, term1, term2

    @classmethod
    def mode(cls, mu, cov, *args, **kwargs):
        return mu

    @classmethod
    def conditional_variance(cls, mu, cov, angle: torch.Tensor, *args, **kwargs):
        """bivariatenormal.py:57-65."""
        xx, xy, yy = cov[0, 0], cov[0, 1], cov[1, 1]
        c, s = torch.cos(angle), torch.sin(angle)
        return torch.det(cov) / (xx * s ** 2 + yy * c ** 2 - 2 * xy * s * c)

    @classmethod
    def marginal(cls, mu, cov, axis: int, angle=torch.tensor(0), *args, **kwargs):
        """bivariatenormal.py:67-86: (mu[axis], (R cov R^T)[axis, axis]) with R the rotation by -angle."""
        assert axis == 0 or axis == 1
        c, s = torch.cos(-angle), torch.sin(-angle)
        R = torch.tensor([[c, -s], [s, c]]).float()
        cov = R @ torch.as_tensor(cov, dtype=torch.float32) @ R.T
        return mu[axis], cov[axis, axis]

    @classmethod
    def rvs(cls, mu, cov, size=(1,), z=None, seed: int = 0):
        """bivariatenormal.py:88-90: ``MultivariateNormal(mu, cov).sample(size)`` = mu + chol(cov) z -> (n, 2).  The draw runs
        through the sampling kernel (a one-landmark "contour" whose factor is the 2x2 Cholesky factor); ``z`` (n, 2) replays
        the standard normals, otherwise Philox with ``seed``."""
        n = int(size[0]) if isinstance(size, (tuple, list, torch.Size)) else int(size)
        d = _dev(mu)
        m = torch.as_tensor(mu, dtype=torch.float32).to(d).reshape(2)
        c = torch.as_tensor(cov, dtype=torch.float32).to(d).reshape(2, 2)
        l00 = torch.sqrt(c[0, 0])
        l10 = c[1, 0] / l00
        l11 = torch.sqrt(c[1, 1] - l10 * l10)
        mu21 = torch.zeros((1, E.K, 2), dtype=torch.float32, device=d)
        mu21[0, 0] = m
        a_t = torch.zeros((1, 2, E.P), dtype=torch.float32, device=d)       # A^T: draw e -> coordinates of landmark 0
        a_t[0, 0, 0], a_t[0, 0, 1], a_t[0, 1, 1] = l00, l10, l11
        zt = None if z is None else torch.as_tensor(z, dtype=torch.float32).to(d).reshape(1, n, 2).contiguous()
        out = E.sample_contours(mu21, a_t, n, z=zt, seed=seed)[0, :, 0]
        return out.to(mu.device if isinstance(mu, torch.Tensor) else d)
