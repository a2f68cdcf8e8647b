"""``FunctionalMapping`` with the reference's surface (densematcher/pyFM/functional.py:19-831) for the
correspondence hot path: preprocess -> fit -> get_p2p -> icp_refine / zoomout_refine, plus
project / decode / transport / transfer and the ``FM`` / ``FM_type`` / ``k1`` / ``k2`` / ``eta`` /
``mapped_indicator`` attributes that ``compute_surface_map`` reads (functional_map.py:44-77).

What differs, deliberately:
* ``fit`` returns the float64 closed-form minimiser of the descriptor + Laplacian energy instead of running
  L-BFGS-B on a float32 energy (functional.py:381,477): same optimum, without the optimiser noise.  Energy terms
  With any of the dense-map terms (w_p2p, w_stochastic, w_ent, w_range01, w_sumto1 -- the notebook's default
  ``fit_params`` use w_ent and w_sumto1) the reference's L-BFGS-B loop is kept and the energy is evaluated on the
  GPU (``dm_dense_energy``).  The remaining terms (w_dcomm, w_orient, w_area, w_conformal, ...) raise
  ``NotImplementedError`` when given a non-zero weight -- never silently ignored.
* ``mapped_indicator`` (n2 x n1 float64, 32 MB at N = 2000) is materialised lazily, only if somebody reads it;
  ``get_p2p(dense=True)`` returns the dense-argmax override of functional_map.py:49-50 straight from the fused
  pass.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from .. import fm as _fm
from ._dev import to_dev
from . import refine as _refine
from . import spectral as _spectral

_UNSUPPORTED_WEIGHTS = ("w_dcomm", "w_orient", "w_area", "w_conformal", "w_area_difference", "w_mumford_shah",
                        "w_eta_entropy")
_DENSE_WEIGHTS = {"w_p2p": "p2p", "w_stochastic": "stochastic", "w_ent": "ent", "w_range01": "range01",
                  "w_sumto1": "sumto1"}


class FunctionalMapping:
    #: flags passed to the projection kernel: 0 = tcgen05 split-bf16 engine (fp32-grade, like the reference's float32
    #: fit), ``_lib.DM_F64_GEMM`` = float64 contraction
    projection_flags = 0

    def __init__(self, mesh1, mesh2, partial=False, optimizer="fmin_l_bfgs_b"):
        self.mesh1 = copy.deepcopy(mesh1)          # functional.py:58-59: inputs are never mutated
        self.mesh2 = copy.deepcopy(mesh2)
        self.descr1 = None
        self.descr2 = None
        self._FM_type = "classic"
        self._FM_base = None
        self._FM_icp = None
        self._FM_zo = None
        self._k1, self._k2 = None, None
        self.optimizer = optimizer                 # accepted for parity; the solve is closed form
        self.partial = partial
        self.eta = None
        self._mi = None                            # lazily materialised mapped_indicator
        self._mi_for = None

    # ------------------------------------------------------------------ dimensions / map switch
    @property
    def k1(self):
        if self._k1 is None and not self.preprocessed and not self.fitted:
            raise ValueError("No information known about dimensions")
        return self.FM.shape[1] if self.fitted else self._k1

    @k1.setter
    def k1(self, v):
        self._k1 = v

    @property
    def k2(self):
        if self._k2 is None and not self.preprocessed and not self.fitted:
            raise ValueError("No information known about dimensions")
        return self.FM.shape[0] if self.fitted else self._k2

    @k2.setter
    def k2(self, v):
        self._k2 = v

    @property
    def FM_type(self):
        return self._FM_type

    @FM_type.setter
    def FM_type(self, FM_type):
        if FM_type.lower() not in ["classic", "icp", "zoomout"]:
            raise ValueError(f'FM_type can only be set to "classic", "icp" or "zoomout", not {FM_type}')
        self._FM_type = FM_type

    def change_FM_type(self, FM_type):
        self.FM_type = FM_type

    @property
    def FM(self):
        return {"classic": self._FM_base, "icp": self._FM_icp, "zoomout": self._FM_zo}[self.FM_type.lower()]

    @FM.setter
    def FM(self, FM):
        self._FM_base = FM

    @property
    def preprocessed(self):
        return (self.descr1 is not None and self.descr2 is not None
                and self.mesh1.eigenvalues is not None and self.mesh2.eigenvalues is not None
                and self.mesh1.eigenvectors is not None and self.mesh2.eigenvectors is not None)

    @property
    def fitted(self):
        return self.FM is not None

    # ------------------------------------------------------------------ pipeline
    def preprocess(self, n_ev=(50, 50), n_descr=100, descr_type="WKS", landmarks=None, subsample_step=1,
                   k_process=None, verbose=False, descr1=None, descr2=None):
        """functional.py:264-350 for given ("neural") descriptors; HKS / WKS signatures and landmarks are outside
        the hot path."""
        self.k1, self.k2 = n_ev
        k_process = 1 if k_process is None else k_process
        if landmarks is not None and len(landmarks) > 0:
            raise NotImplementedError("landmark descriptors are outside the hot path")
        self.mesh1.process(max(self.k1, k_process), verbose=verbose, robust=True, intrinsic=False)
        self.mesh2.process(max(self.k2, k_process), verbose=verbose, robust=True, intrinsic=False)
        if descr1 is None or descr2 is None:
            raise NotImplementedError(f'descr_type "{descr_type}": only precomputed descriptors (descr1, descr2) '
                                      "are on the hot path")
        d1, d2 = np.asarray(descr1), np.asarray(descr2)
        self.descr1 = d1[:, np.arange(0, d1.shape[1], subsample_step)]
        self.descr2 = d2[:, np.arange(0, d2.shape[1], subsample_step)]
        return self

    def get_x0(self, optinit="zeros"):
        """functional.py:629-660 (only the pinned first column matters for the closed form)."""
        if optinit == "random":
            x0 = np.random.random((self.k2, self.k1))
            x0 = x0 / x0.sum()
        elif optinit == "identity":
            x0 = np.eye(self.k2, self.k1)
        else:
            x0 = np.zeros((self.k2, self.k1))
        ev_sign = np.sign(self.mesh1.eigenvectors[0, 0] * self.mesh2.eigenvectors[0, 0])
        x0[:, 0] = 0.0
        x0[0, 0] = ev_sign * np.sqrt(self.mesh2.area / self.mesh1.area)
        return x0

    def fit(self, w_descr=1e-1, w_lap=1e-3, w_dcomm=1, w_orient=0, w_area=0, w_conformal=0, w_p2p=0, w_stochastic=0,
            w_ent=0, w_range01=0, w_sumto1=0, w_area_difference=0, w_mumford_shah=0, mumford_shah_var=0.1,
            w_eta_entropy=0, orient_reversing=False, optinit="zeros", verbose=False, maxiter=1000000, device=None):
        """Minimiser of  w_descr/2 |C A - B|^2 + w_lap/2 sum C^2 Delta  with column 0 pinned (functional.py:352-487,
        base_functions.py:31-56, :79-102, :759).  The defaults are the reference's (functional.py:352-356), including
        ``w_dcomm=1`` -- a term that is not implemented here, so a call relying on the defaults raises
        ``NotImplementedError`` instead of silently minimising a different energy; pass ``w_dcomm=0`` as the DenseMatcher
        notebook does (example.ipynb cell 11)."""
        given = dict(w_dcomm=w_dcomm, w_orient=w_orient, w_area=w_area, w_conformal=w_conformal, w_p2p=w_p2p,
                     w_stochastic=w_stochastic, w_ent=w_ent, w_range01=w_range01, w_sumto1=w_sumto1,
                     w_area_difference=w_area_difference, w_mumford_shah=w_mumford_shah, w_eta_entropy=w_eta_entropy)
        bad = [n for n in _UNSUPPORTED_WEIGHTS if given[n] != 0]
        if bad:
            raise NotImplementedError(f"energy terms {bad} are not implemented (SURVEY.md 8f); supported: w_descr, "
                                      "w_lap and the dense-map terms w_p2p, w_stochastic, w_ent, w_range01, w_sumto1")
        dense = {t: float(given[n]) for n, t in _DENSE_WEIGHTS.items() if given[n] != 0}
        if self.partial:
            raise NotImplementedError()                                   # functional.py:479-480
        if not self.preprocessed:
            self.preprocess()
        k1, k2 = self._k1, self._k2
        P1 = to_dev(self.mesh1.eigenvectors[:, :k1], torch.float64)
        P2 = to_dev(self.mesh2.eigenvectors[:, :k2], torch.float64)
        a1, a2 = to_dev(self.mesh1.vertex_areas, torch.float64), to_dev(self.mesh2.vertex_areas, torch.float64)
        A = _fm.project(P1, a1, to_dev(self.descr1, torch.float32), flags=self.projection_flags)
        B = _fm.project(P2, a2, to_dev(self.descr2, torch.float32), flags=self.projection_flags)
        c00 = float(self.get_x0(optinit)[0, 0])
        ev1 = to_dev(self.mesh1.eigenvalues[:k1], torch.float64)[None]
        ev2 = to_dev(self.mesh2.eigenvalues[:k2], torch.float64)[None]
        if not dense:
            C = _fm.fmap_solve(A, B, ev1, ev2, torch.tensor([c00], dtype=torch.float64, device=A.device), w_descr, w_lap)
            self.FM = C[0].cpu().numpy()
        elif self.optimizer == "scipy":
            # the reference's host loop (scipy L-BFGS-B, one host round trip per callback) around dm_dense_energy
            self.FM = self._fit_lbfgs(A[0], B[0], c00, P1, P2, a1, w_descr, w_lap, dense, maxiter)
        else:
            # default: the batched on-device L-BFGS (fm.fit_dense) -- same energy, same pinned column, no host round
            # trip per iteration
            C, info = _fm.fit_dense(A, B, ev1, ev2, torch.tensor([c00], dtype=torch.float64, device=A.device), P1, P2, a1,
                                    dense, w_descr, w_lap, maxiter=min(int(maxiter), 2000), return_info=True)
            self.FM = C[0].cpu().numpy()
            self.fit_result = type("FitResult", (), {"nit": info[0], "nfev": info[1], "x": self.FM.ravel()})()
        self.eta = np.ones(self.mesh2.eigenvectors.shape[0])              # functional.py:483
        self._mi = None
        return self

    def _fit_lbfgs(self, A, B, c00, P1, P2, a1, w_descr, w_lap, dense, maxiter):
        """The reference's optimiser loop (functional.py:477: scipy L-BFGS-B, x0 = c00 e_00, gradient of column 0
        zeroed, base_functions.py:759) with the energy evaluated on the GPU in float64: descriptor and Laplacian
        terms as two small products, the dense-map terms by ``dm_dense_energy`` (M is never materialised)."""
        import scipy.optimize
        k1, k2 = self._k1, self._k2
        l1, l2 = np.asarray(self.mesh1.eigenvalues[:k1], np.float64), np.asarray(self.mesh2.eigenvalues[:k2], np.float64)
        scale = max(l1.max(), l2.max())
        Delta = torch.from_numpy(np.square(l1[None, :] / scale - l2[:, None] / scale)).to(A.device)
        At = A.T.contiguous()

        def fun(x):
            C = torch.from_numpy(np.ascontiguousarray(x.reshape(k2, k1))).to(A.device)
            R = C @ A - B
            e = 0.5 * w_descr * (R * R).sum() + 0.5 * w_lap * (C * C * Delta).sum()
            g = w_descr * (R @ At) + w_lap * (C * Delta)
            ed, gd = _fm.dense_energy(C, P1, P2, a1, dense)
            wvec = torch.tensor([dense.get(t, 0.0) for t in _fm.DENSE_TERMS], dtype=torch.float64, device=A.device)
            e = e + (ed[0] * wvec).sum()
            g = g + gd[0]
            g[:, 0] = 0.0
            return float(e), g.reshape(-1).cpu().numpy()

        x0 = np.zeros((k2, k1))
        x0[0, 0] = c00
        method = "L-BFGS-B" if self.optimizer in ("fmin_l_bfgs_b", "L-BFGS-B", "scipy") else self.optimizer
        res = scipy.optimize.minimize(fun, x0.ravel(), jac=True, method=method, options={"maxiter": int(maxiter)})
        self.fit_result = res
        return res.x.reshape(k2, k1)

    def _dev_bases(self):
        k2, k1 = self.FM.shape
        return (to_dev(self.mesh1.eigenvectors[:, :k1], torch.float64),
                to_dev(self.mesh2.eigenvectors[:, :k2], torch.float64),
                to_dev(self.mesh1.vertex_areas, torch.float64))

    def get_p2p(self, use_adj=False, n_jobs=1, dense=False):
        """(p2p_21, p2p_12) of the current map, like functional.py:201-219 (the kd-tree-equivalent searches of
        convert.py:134-140).  ``dense=True`` (extension) returns ``(p2p_21, p2p_12, dense_21, dense_12)`` where the
        last two are the argmax override of functional_map.py:49-50, all from one fused pass."""
        if not self.fitted:
            raise ValueError("Model should be fit before computing a point to point map")
        P1, P2, a1 = self._dev_bases()
        want = ("p2p_21", "p2p_12", "dense_21", "dense_12") if dense else ("p2p_21", "p2p_12")
        out = _fm.fm_to_p2p(to_dev(self.FM, torch.float64), P1, P2, a1, want=want)
        self._mi, self._mi_for = None, (self.FM_type, id(self.FM))
        res = tuple(out[n].cpu().numpy() for n in want)
        return res

    @property
    def mapped_indicator(self):
        """Phi2 C Phi1^T A1 (convert.py:144), materialised on first access for the current map."""
        if not self.fitted:
            raise ValueError("Model should be fit first")
        key = (self.FM_type, id(self.FM))
        if self._mi is None or self._mi_for != key:
            P1, P2, a1 = self._dev_bases()
            self._mi = _fm.mapped_indicator(to_dev(self.FM, torch.float64), P1, P2, a1).cpu().numpy()
            self._mi_for = key
        return self._mi

    def hungarian(self, indicator=None):
        """``linear_sum_assignment(MI * eta - 1000 (1 - eta), maximize=True)`` (functional_map.py:57,66,78) solved in
        HBM by ``dm_lap_solve``: the same (row_ind, col_ind) as scipy.  ``indicator``: a dense (n2, n1) map to use
        instead of the current ``mapped_indicator`` (the precise map, functional_map.py:62-66)."""
        if not self.fitted:
            raise ValueError("Model should be fit first")
        if indicator is None:
            P1, P2, a1 = self._dev_bases()
            mi = _fm.mapped_indicator(to_dev(self.FM, torch.float64), P1, P2, a1)
        else:
            mi = to_dev(indicator, torch.float64)
        eta = to_dev(self.eta, torch.float64)[:, None]
        cost = mi * eta - 1000 * (1 - eta)
        return _fm.lap_solve(cost, maximize=True)

    def icp_refine(self, nit=10, tol=None, use_adj=False, overwrite=True, verbose=False, n_jobs=1):
        """functional.py:564-586."""
        if not self.fitted:
            raise ValueError("The Functional map must be fit before refining it")
        self._FM_icp = _refine.mesh_icp_refine(self.FM, self.mesh1, self.mesh2, nit=nit, tol=tol, return_p2p=False,
                                               use_adj=use_adj, n_jobs=n_jobs, verbose=verbose)
        if overwrite:
            self.FM_type = "icp"

    def zoomout_refine(self, nit=10, step=1, subsample=None, overwrite=True, verbose=False):
        """functional.py:588-617 (upstream semantics, see refine/zoomout.py; farthest-point subsampling by count is
        outside the hot path, pass ``subsample=None`` or a pair of index arrays)."""
        if not self.fitted:
            raise ValueError("The Functional map must be fit before refining it")
        sub = None if subsample is None or (np.isscalar(subsample) and subsample == 0) else subsample
        self._FM_zo = _refine.mesh_zoomout_refine(self.FM, self.mesh1, self.mesh2, nit, step=step, subsample=sub,
                                                  verbose=verbose)
        if overwrite:
            self.FM_type = "zoomout"

    def compute_SD(self):
        """functional.py:619-627: shape-difference operators are outside the accelerated path."""
        raise NotImplementedError("shape-difference operators are outside the accelerated path (SURVEY.md section 2 row 9)")

    def get_precise_map(self, precompute_dmin=True, use_adj=True, batch_size=None, n_jobs=1, verbose=False):
        """functional.py:221-251: (n2, n1) sparse barycentric map of mesh 2 onto mesh 1."""
        if not self.fitted:
            raise ValueError("Model should be fit and fit to obtain p2p map")
        return _spectral.mesh_FM_to_p2p_precise(self.FM, self.mesh1, self.mesh2, precompute_dmin=precompute_dmin,
                                                use_adj=use_adj, batch_size=batch_size, n_jobs=n_jobs, verbose=verbose)

    # ------------------------------------------------------------------ function transfer
    def project(self, func, k=None, mesh_ind=1):
        """functional.py:730-752."""
        if k is None:
            k = self.k1 if mesh_ind == 1 else self.k2
        if mesh_ind == 1:
            return self.mesh1.project(func, k=k)
        if mesh_ind == 2:
            return self.mesh2.project(func, k=k)
        raise ValueError(f"Only indices 1 or 2 are accepted, not {mesh_ind}")

    def decode(self, encoded_func, mesh_ind=2):
        """functional.py:754-776."""
        if mesh_ind == 1:
            return self.mesh1.decode(encoded_func)
        if mesh_ind == 2:
            return self.mesh2.decode(encoded_func)
        raise ValueError(f"Only indices 1 or 2 are accepted, not {mesh_ind}")

    def transport(self, encoded_func, reverse=False):
        """functional.py:778-804."""
        if not self.preprocessed:
            raise ValueError("The Functional map must be fit before transporting a function")
        return (self.FM.T if reverse else self.FM) @ encoded_func

    def transfer(self, func, reverse=False):
        """functional.py:806-831."""
        if not reverse:
            return self.decode(self.transport(self.project(func)))
        return self.decode(self.transport(self.project(func, mesh_ind=2), reverse=True), mesh_ind=1)
